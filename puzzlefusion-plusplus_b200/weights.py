"""Weight packing: reference ``state_dict``s -> flat device buffers in the layouts the kernels read.

Input keys/shapes are exactly the reference checkpoints' (test.py:24-38
prefix-stripped).  Packing does, once per checkpoint:
  * BatchNorm (eval) folded into the 1x1-conv weight/bias in float64 (utils/pn2_utils.py:209-212);
  * K dimension zero-padded to the GEMM engines' alignment (4 fp32 / 8 bf16 elements);
  * to_q/to_k/to_v concatenated into one [3C, C] projection (attention.py:46-72 via diffusers);
  * GEGLU projection rows interleaved (value_j, gate_j) so the activation fuses into the epilogue;
  * the two output heads' first layers concatenated;
  * AdaLN modulation rows Linear(SiLU(Embedding[t])) tabulated for the scheduler's T timesteps
    (attention.py:22-24; they depend only on (layer, t));
  * bf16 copies for the tcgen05 path.
"""
import torch

from . import _lib
from ._lib import EPI_NONE


def _pad_k(w, mult):
    n, k = w.shape
    kp = (k + mult - 1) // mult * mult
    if kp == k:
        return w.contiguous()
    out = w.new_zeros(n, kp)
    out[:, :k] = w
    return out


class Linear:
    """One packed projection: fp32 [N, K4] (+ optional bf16 [N, K8]) and fp32 bias."""

    def __init__(self, w, b, device, bf16, split=False):
        w = w.detach().to(torch.float32)
        self.n, self.k = w.shape
        self.w32 = _pad_k(w, 4).to(device)
        self.k32 = self.w32.shape[1]
        self.b = None if b is None else b.detach().to(torch.float32).contiguous().to(device)
        self.w16 = self.w16s = None
        if bf16 or split:
            wp = _pad_k(w, 8).to(device)
            self.k16 = wp.shape[1]
            if bf16:
                self.w16 = wp.to(torch.bfloat16).contiguous()
            if split:
                # bf16 hi/lo split rows [N, 2*K8] (hi | lo): the weight operand of pfpp_gemm_bf16x3
                hi = wp.to(torch.bfloat16)
                lo = (wp - hi.float()).to(torch.bfloat16)
                self.w16s = torch.cat([hi, lo], 1).contiguous()


def fold_bn(sd, prefix, i):
    w = sd[f"{prefix}.mlp_convs.{i}.weight"].double().flatten(1)
    b = sd[f"{prefix}.mlp_convs.{i}.bias"].double()
    g = sd[f"{prefix}.mlp_bns.{i}.weight"].double()
    beta = sd[f"{prefix}.mlp_bns.{i}.bias"].double()
    mean = sd[f"{prefix}.mlp_bns.{i}.running_mean"].double()
    var = sd[f"{prefix}.mlp_bns.{i}.running_var"].double()
    s = g / torch.sqrt(var + 1e-5)
    return (w * s[:, None]).float(), ((b - mean) * s + beta).float()


class EncoderWeights:
    def __init__(self, sd, device, bf16, split=False):
        self.sa = []
        for li in range(3):
            layers = []
            for i in range(3):
                w, b = fold_bn(sd, f"pn2.sa{li + 1}", i)
                layers.append(Linear(w, b, device, bf16, split))
            self.sa.append(layers)
            if bf16:
                # layer 0 of the fused kernel: feature columns on the tensor cores (bf16), the three
                # centroid-offset columns as fp32 epilogue FMAs
                w0 = layers[0]
                w = fold_bn(sd, f"pn2.sa{li + 1}", 0)[0]
                w0.w16_feat = w[:, 3:].contiguous().to(device).to(torch.bfloat16) if w.shape[1] > 3 else None
                wx = torch.zeros(w0.n, 4)
                wx[:, :3] = w[:, :3]
                w0.wxyz = wx.contiguous().to(device)
        self.conv6 = Linear(sd["pn2.conv6.weight"].flatten(1), sd["pn2.conv6.bias"], device, bf16, split)
        self.codebook = sd["vector_quantization.embedding.weight"].detach().float().contiguous().to(device)


class DenoiserWeights:
    def __init__(self, sd, device, bf16, num_layers, timesteps, split=False):
        C = sd["param_fc.weight"].shape[0]
        self.C = C
        self.layers = []
        ts = torch.as_tensor(timesteps, dtype=torch.long)
        for i in range(num_layers):
            p = f"transformer_layers.{i}"
            L = {}
            for a in ("self_attn", "global_attn"):
                wqkv = torch.cat([sd[f"{p}.{a}.to_q.weight"], sd[f"{p}.{a}.to_k.weight"], sd[f"{p}.{a}.to_v.weight"]], 0)
                L[a + ".qkv"] = Linear(wqkv, None, device, bf16, split)
                L[a + ".out"] = Linear(sd[f"{p}.{a}.to_out.0.weight"], sd[f"{p}.{a}.to_out.0.bias"], device, bf16, split)
            w1, b1 = sd[f"{p}.ff.net.0.proj.weight"], sd[f"{p}.ff.net.0.proj.bias"]
            half = w1.shape[0] // 2
            wi = torch.stack([w1[:half], w1[half:]], 1).reshape(2 * half, -1)
            bi = torch.stack([b1[:half], b1[half:]], 1).reshape(2 * half)
            L["ff1"] = Linear(wi, bi, device, bf16, split)
            L["ff2"] = Linear(sd[f"{p}.ff.net.2.weight"], sd[f"{p}.ff.net.2.bias"], device, bf16, split)
            L["norm3.w"] = sd[f"{p}.norm3.weight"].detach().float().contiguous().to(device)
            L["norm3.b"] = sd[f"{p}.norm3.bias"].detach().float().contiguous().to(device)
            self.layers.append(L)
        # mod[(layer*2 + which)] : [T, 2C] -- rows selected per fragment by the current step index
        self._ada = [{k: sd[f"transformer_layers.{i}.{n}.{k}"].detach() for k in ("emb.weight", "linear.weight", "linear.bias")}
                     for i in range(num_layers) for n in ("norm1", "norm2")]
        self._device = device
        self.mod = self.modulation_table(ts)
        self._mod_all = None
        self.shape_embedding = Linear(sd["shape_embedding.weight"], sd["shape_embedding.bias"], device, bf16, split)
        self.param_fc = Linear(sd["param_fc.weight"], sd["param_fc.bias"], device, bf16, split)
        self.ref_emb = sd["ref_part_emb.weight"].detach().float().contiguous().to(device)
        self.pe = sd["pos_encoding.pe"][0].detach().float().contiguous().to(device)  # [P, C]
        # output heads: fp32 SIMT in fp32 mode, split-operand (fp32-grade) tensor-core GEMMs in the other modes
        hs = bf16 or split
        self.head0 = Linear(torch.cat([sd["mlp_out_trans.0.weight"], sd["mlp_out_rot.0.weight"]], 0),
                            torch.cat([sd["mlp_out_trans.0.bias"], sd["mlp_out_rot.0.bias"]], 0), device, False, hs)
        self.head_t2 = Linear(sd["mlp_out_trans.2.weight"], sd["mlp_out_trans.2.bias"], device, False, hs)
        self.head_r2 = Linear(sd["mlp_out_rot.2.weight"], sd["mlp_out_rot.2.bias"], device, False, hs)
        self.head_t4 = Linear(sd["mlp_out_trans.4.weight"], sd["mlp_out_trans.4.bias"], device, False, hs)
        self.head_r4 = Linear(sd["mlp_out_rot.4.weight"], sd["mlp_out_rot.4.bias"], device, False, hs)


    def modulation_table(self, timesteps):
        """MyAdaLayerNorm's Linear(SiLU(Embedding[t])) (attention.py:22-24) for the given timesteps ->
        [2 * layers, len(timesteps), 2C] fp32 (row order: layer-major, norm1 then norm2)."""
        ts = torch.as_tensor(timesteps, dtype=torch.long)
        C, device = self.C, self._device
        mods = []
        for ada in self._ada:
            emb = ada["emb.weight"][ts].float().to(device)
            a = torch.nn.functional.silu(emb).contiguous()
            lin = Linear(ada["linear.weight"], ada["linear.bias"], device, False)
            out = torch.empty(len(ts), 2 * C, device=device, dtype=torch.float32)
            _lib.call("pfpp_gemm_f32", a.data_ptr(), C, lin.w32.data_ptr(), lin.k32, lin.b.data_ptr(), None, 0,
                      out.data_ptr(), 2 * C, len(ts), 2 * C, C, EPI_NONE)
            mods.append(out)
        return torch.stack(mods, 0).contiguous()

    def mod_all(self, num_train_timesteps):
        """the table for EVERY training timestep (row = t), built on first use: DenoiserTransformer.forward accepts
        any timestep (training / validation forward, denoiser.py:80-113), not only the inference schedule's"""
        if self._mod_all is None or self._mod_all.shape[1] != num_train_timesteps:
            self._mod_all = self.modulation_table(range(num_train_timesteps))
        return self._mod_all


class VerifierWeights:
    def __init__(self, sd, device, num_layers, split=False):
        self.C = sd["edge_feature_emb.weight"].shape[0]
        f = lambda k: sd[k].detach().float().contiguous().to(device)  # noqa: E731
        self.emb_w, self.emb_b = f("edge_feature_emb.weight"), f("edge_feature_emb.bias")
        self.pe = sd["edge_indices_pe.pe"][0].detach().float().contiguous().to(device)
        self.out_w, self.out_b = f("mlp_out.weight").reshape(-1), f("mlp_out.bias")
        self.layers = []
        for i in range(num_layers):
            p = f"transformer_encoder.layers.{i}"
            self.layers.append({
                "qkv": Linear(sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"], device, False, split),
                "out": Linear(sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"], device, False, split),
                "l1": Linear(sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"], device, False, split),
                "l2": Linear(sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"], device, False, split),
                "n1w": f(f"{p}.norm1.weight"), "n1b": f(f"{p}.norm1.bias"),
                "n2w": f(f"{p}.norm2.weight"), "n2b": f(f"{p}.norm2.bias"),
            })
