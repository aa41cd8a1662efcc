"""Minimal Hydra-compatible composer for the reference's ``config/auto_aggl.yaml``.

hydra-core / omegaconf are not installed in this image; the hot path only needs: the ``defaults:``
list (group files land under their directory's package, ``_self_`` first so group files win),
dotted ``key=value`` overrides exactly as scripts/inference.sh passes them, and the interpolations
``${hydra:runtime.cwd}``, ``${project_root_path}``, ``${experiment_name}``.  The YAML files are the
reference's, byte-for-byte (copied under config/ of this repo for convenience, or any checkout).
"""
import os
import re

import yaml


class Cfg(dict):
    """Attribute-style dict (what the modules read as ``cfg.denoiser.model.embed_dim``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(d):
    if isinstance(d, dict):
        return Cfg({k: _wrap(v) for k, v in d.items()})
    if isinstance(d, list):
        return [_wrap(v) for v in d]
    return d


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _set_dotted(cfg, key, value):
    parts = key.lstrip("+").split(".")
    d = cfg
    for p in parts[:-1]:
        d = d.setdefault(p, {})
    d[parts[-1]] = yaml.safe_load(value) if value != "" else None


def _resolve(node, root, cwd):
    def sub(s):
        def rep(m):
            key = m.group(1)
            if key == "hydra:runtime.cwd":
                return cwd
            v = root
            for p in key.split("."):
                v = v[p]
            return str(sub(v) if isinstance(v, str) else v)
        return re.sub(r"\$\{([^}]+)\}", rep, s)

    if isinstance(node, dict):
        return {k: _resolve(v, root, cwd) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, cwd) for v in node]
    if isinstance(node, str) and "${" in node:
        return sub(node)
    return node


def compose(config_dir, config_name="auto_aggl", overrides=(), cwd=None):
    """Compose ``config_dir/config_name.yaml`` the way ``@hydra.main`` does for test.py:9."""
    cwd = cwd or os.getcwd()
    with open(os.path.join(config_dir, config_name + ".yaml")) as f:
        primary = yaml.safe_load(f)
    defaults = primary.pop("defaults", [])
    primary.pop("hydra", None)
    cfg = {}
    for item in defaults:
        if item == "_self_":
            _merge(cfg, primary)
        elif isinstance(item, str):
            with open(os.path.join(config_dir, item + ".yaml")) as f:
                content = yaml.safe_load(f) or {}
            node = cfg
            for p in os.path.dirname(item).split("/"):
                if p:
                    node = node.setdefault(p, {})
            _merge(node, content)
        # dict items are hydra-internal overrides (logging) -- nothing to load
    if "_self_" not in defaults:
        _merge(cfg, primary)
    for ov in overrides:
        k, _, v = ov.partition("=")
        _set_dotted(cfg, k, v)
    return _wrap(_resolve(cfg, cfg, cwd))
