"""sys.modules shims that let the reference's OWN first-party modules import.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

Used only where /root/reference exists (the build container): by
``oracle/gen_golden.py`` to produce tests/golden/*.npz from the reference's
verbatim first-party code, and by the optional cross-check tests.  The shims
supply ONLY the packages that are missing here (lightning, hydra, diffusers,
torch_cluster, pytorch3d, chamferdist) using oracle/third_party.py; nothing
from the reference is copied or modified.
"""
import contextlib
import os
import sys
import types

import torch

from . import third_party as tp

REFERENCE_ROOT = os.environ.get("PFPP_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "puzzlefusion_plusplus"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _LightningModule(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self._logged = {}

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, name, value, **k):
        self._logged[name] = value


def install():
    """Idempotently install the shims and put the reference on sys.path."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "pfpp_shims_installed" in sys.modules:
        return
    _mod("pfpp_shims_installed")
    # lightning
    plm = _mod("lightning.pytorch", LightningModule=_LightningModule,
               seed_everything=lambda s, workers=False: torch.manual_seed(s))
    _mod("lightning", pytorch=plm)
    # hydra
    hu = _mod("hydra.utils", instantiate=lambda *a, **k: None)
    _mod("hydra", utils=hu, main=lambda *a, **k: (lambda f: f))
    # diffusers
    att = _mod("diffusers.models.attention", Attention=tp.Attention, FeedForward=tp.FeedForward)
    dm = _mod("diffusers.models", attention=att)
    _mod("diffusers", DDPMScheduler=tp.DDPMScheduler, models=dm)
    # torch_cluster
    _mod("torch_cluster", fps=tp.fps)
    # pytorch3d
    tr = _mod("pytorch3d.transforms",
              quaternion_apply=tp.quaternion_apply,
              quaternion_raw_multiply=tp.quaternion_raw_multiply,
              quaternion_invert=tp.quaternion_invert,
              quaternion_to_matrix=tp.quaternion_to_matrix,
              matrix_to_quaternion=tp.matrix_to_quaternion,
              matrix_to_euler_angles=lambda m, convention="XYZ": tp.matrix_to_euler_angles_xyz(m))
    ops = _mod("pytorch3d.ops", estimate_pointcloud_normals=tp.estimate_pointcloud_normals)
    _mod("pytorch3d", transforms=tr, ops=ops)
    # chamferdist
    _mod("chamferdist", ChamferDistance=tp.ChamferDistance)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


@contextlib.contextmanager
def cpu_cuda_noop():
    """The reference calls ``.cuda()`` (auto_aggl.py:138); make it a no-op on CPU."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


class AttrDict(dict):
    """Minimal OmegaConf-like attribute dict for constructing reference modules."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        if isinstance(d, list):
            return [AttrDict.wrap(v) for v in d]
        return d
