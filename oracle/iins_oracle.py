"""CPU oracle for the IIns-VAE hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch CPU ops on a flat ``dict`` of
named parameters, no nn.Module) of the algorithm in the reference repository
JadeLilyx/IIns-VAE.  It exists so the CUDA path can be checked for parity; it is
never imported by the product package ``iins_vae_b200`` (only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it).

Pinned?  The reference ships no tests and no golden vectors.
The oracle is pinned instead against OUTPUTS OF THE REFERENCE ITSELF: the live
``/root/reference/models.py`` modules are run in the authoring container by
``tests/golden/make_golden.py`` and the results are committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.

Where the arithmetic lives: the reference calls torch (2.11.0+cu128 in this image)
``nn.Conv1d / nn.Linear / nn.InstanceNorm1d / F.batch_norm / nn.Upsample /
nn.AdaptiveAvgPool1d``.  Here the convolutions and matmuls are ``F.conv1d`` /
``F.linear``; every other operator (adaptive pooling windows, reflection padding,
instance / adaptive-instance / custom layer norm, nearest upsampling, the losses,
Adam) is restated explicitly from its published definition so that the semantics
the CUDA kernels must reproduce are written down in one place.

All ``file:line`` citations are into ``/root/reference``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass(frozen=True)
class PathConfig:
    """Shape options of the 1-D path (models.py:33,68,97,118; train_semi.py:77-82)."""
    cir_len: int = 157        # README_diverse.md:8  (zenodo: 157 taps)
    dim: int = 4              # models.py:39 comment "dim=4"
    n_residual: int = 3
    n_downsample: int = 4
    env_dim: int = 16         # style_dim; utils.py:39
    range_dim: int = 2        # models.py:169 comment
    num_classes: int = 5      # train_semi.py:46 (room_full)
    pooled_len: int = 128     # models.py:146, 264

    @property
    def trunk_dim(self) -> int:          # channels after the downsampling stack
        return self.dim * 2 ** self.n_downsample

    @property
    def code_len(self) -> int:           # train_semi.py:70
        return self.pooled_len // (2 ** self.n_downsample)

    @property
    def n_adain(self) -> int:            # models.py:442-449
        return 2 * self.n_residual * 2 * self.trunk_dim


# ------------------------------------------------------------------ parameter inventory
def encoder_param_shapes(cfg: PathConfig) -> "OrderedDict[str, tuple]":
    """state_dict keys/shapes of ``Encoder`` in registration order
    (models.py:140-173 range encoder, :258-281 env encoder)."""
    sh = OrderedDict()
    d = cfg.dim
    idx = 2                                            # [pool, pad, conv(2), IN, ReLU]
    sh[f"range_encoder.model.{idx}.weight"] = (d, 1, 7)
    sh[f"range_encoder.model.{idx}.bias"] = (d,)
    idx += 3
    c = d
    for _ in range(cfg.n_downsample):                  # conv, IN, ReLU
        sh[f"range_encoder.model.{idx}.weight"] = (2 * c, c, 4)
        sh[f"range_encoder.model.{idx}.bias"] = (2 * c,)
        c *= 2
        idx += 3
    for _ in range(cfg.n_residual):
        for j in (1, 5):                               # models.py:994-1002
            sh[f"range_encoder.model.{idx}.block.{j}.weight"] = (c, c, 3)
            sh[f"range_encoder.model.{idx}.block.{j}.bias"] = (c,)
        idx += 1
    sh[f"range_encoder.model.{idx}.weight"] = (cfg.range_dim, c, 1)
    sh[f"range_encoder.model.{idx}.bias"] = (cfg.range_dim,)
    # env encoder: EnvEncoder1d(dim*4, n_downsample-2, style_dim)  (models.py:40)
    e = 4 * d
    idx = 2                                            # [pool, pad, conv(2), ReLU]
    sh[f"env_encoder.model.{idx}.weight"] = (e, 1, 7)
    sh[f"env_encoder.model.{idx}.bias"] = (e,)
    idx += 2
    for _ in range(2):
        sh[f"env_encoder.model.{idx}.weight"] = (2 * e, e, 4)
        sh[f"env_encoder.model.{idx}.bias"] = (2 * e,)
        e *= 2
        idx += 2
    for _ in range(cfg.n_downsample - 2 - 2):          # models.py:275 (0 iterations by default)
        sh[f"env_encoder.model.{idx}.weight"] = (e, e, 4)
        sh[f"env_encoder.model.{idx}.bias"] = (e,)
        idx += 2
    idx += 1                                           # AdaptiveAvgPool1d(1)
    sh[f"env_encoder.model.{idx}.weight"] = (cfg.env_dim, e, 1)
    sh[f"env_encoder.model.{idx}.bias"] = (cfg.env_dim,)
    return sh


def decoder_param_shapes(cfg: PathConfig, buffers: bool = True) -> "OrderedDict[str, tuple]":
    """state_dict keys of ``Decoder`` (models.py:405-439 + MLP :951-959).  ``buffers``
    adds the dummy AdaIN running_mean / running_var entries (models.py:1057-1059)."""
    sh = OrderedDict()
    D = cfg.trunk_dim
    sh["decoder.model.0.weight"] = (D, cfg.range_dim, 1)
    sh["decoder.model.0.bias"] = (D,)
    idx = 2
    for _ in range(cfg.n_residual):
        for j in (1, 5):
            sh[f"decoder.model.{idx}.block.{j}.weight"] = (D, D, 3)
            sh[f"decoder.model.{idx}.block.{j}.bias"] = (D,)
            if buffers:
                sh[f"decoder.model.{idx}.block.{j + 1}.running_mean"] = (D,)
                sh[f"decoder.model.{idx}.block.{j + 1}.running_var"] = (D,)
        idx += 1
    c = D
    for _ in range(cfg.n_downsample):                  # up, conv, LN, ReLU
        sh[f"decoder.model.{idx + 1}.weight"] = (c // 2, c, 5)
        sh[f"decoder.model.{idx + 1}.bias"] = (c // 2,)
        sh[f"decoder.model.{idx + 2}.gamma"] = (c // 2,)
        sh[f"decoder.model.{idx + 2}.beta"] = (c // 2,)
        c //= 2
        idx += 4
    sh[f"decoder.model.{idx + 1}.weight"] = (1, c, 7)
    sh[f"decoder.model.{idx + 1}.bias"] = (1,)
    sh["decoder.mlp.model.0.weight"] = (256, cfg.env_dim)
    sh["decoder.mlp.model.0.bias"] = (256,)
    sh["decoder.mlp.model.2.weight"] = (256, 256)
    sh["decoder.mlp.model.2.bias"] = (256,)
    sh["decoder.mlp.model.4.weight"] = (cfg.n_adain, 256)
    sh["decoder.mlp.model.4.bias"] = (cfg.n_adain,)
    return sh


def restorer_param_shapes(cfg: PathConfig) -> "OrderedDict[str, tuple]":
    """``RestorerLinear`` (models.py:619-633)."""
    n_in = cfg.range_dim * cfg.code_len
    sh = OrderedDict()
    for i, (o, k) in zip((0, 2, 4), ((512, n_in), (256, 512), (256, 256))):
        sh[f"restorer.layers.{i}.weight"] = (o, k)
        sh[f"restorer.layers.{i}.bias"] = (o,)
    sh["restorer.linear_layer1.weight"] = (1, 256)
    sh["restorer.linear_layer1.bias"] = (1,)
    sh["restorer.linear_layer2.weight"] = (2, 256)     # unused when soft=False (:632, :656)
    sh["restorer.linear_layer2.bias"] = (2,)
    return sh


def classifier_param_shapes(cfg: PathConfig, filters: int = 16) -> "OrderedDict[str, tuple]":
    """``ClassifierLinear`` (models.py:846-856)."""
    dims = (cfg.env_dim, filters, 2 * filters, filters, cfg.num_classes)
    sh = OrderedDict()
    for i in range(4):
        sh[f"classifier.layers.{2 * i}.weight"] = (dims[i + 1], dims[i])
        sh[f"classifier.layers.{2 * i}.bias"] = (dims[i + 1],)
    return sh


def grad_is_structurally_zero(name: str, cfg: "PathConfig" = None) -> bool:
    """Conv biases that feed an InstanceNorm / AdaIN (mean subtraction over L) receive an exactly
    zero gradient in exact arithmetic: range-encoder stem + downsampling + residual convs
    (models.py:151-160, 994-1002) and the decoder's AdaIN residual convs.  In fp32 they hold
    rounding noise only, so parity checks treat them with an absolute tolerance."""
    cfg = cfg or PathConfig()
    if not name.endswith(".bias"):
        return False
    if "range_encoder.model." in name:
        if ".block." in name:
            return True
        idx = int(name.split("range_encoder.model.")[1].split(".")[0])
        last = 2 + 3 * (1 + cfg.n_downsample) + cfg.n_residual      # 1x1 conv + ReLU (models.py:171)
        return idx != last
    return "decoder.model." in name and ".block." in name


def is_buffer(name: str) -> bool:
    return name.endswith("running_mean") or name.endswith("running_var") or name.endswith("num_batches_tracked")


def init_params(shapes: "OrderedDict[str, tuple]", gen: torch.Generator) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic parameter values with the reference's init DISTRIBUTIONS:
    conv weights N(0, 0.02) (models.py:8-11); conv biases and Linear layers keep torch's
    default U(-1/sqrt(fan_in), 1/sqrt(fan_in)); custom LayerNorm gamma U(0,1), beta 0
    (models.py:973-974).  The draw order is this file's own, not torch's module RNG order:
    golden fixtures record the seed and are regenerated through this function."""
    out = OrderedDict()
    last_fan_in = 1
    for name, shp in shapes.items():
        if name.endswith("weight"):
            last_fan_in = int(np.prod(shp[1:]))
        if name.endswith("running_mean"):
            t = torch.zeros(shp)
        elif name.endswith("running_var"):
            t = torch.ones(shp)
        elif name.endswith("gamma"):
            t = torch.rand(shp, generator=gen)
        elif name.endswith("beta"):
            t = torch.zeros(shp)
        elif len(shp) == 3 and name.endswith("weight"):
            t = torch.randn(shp, generator=gen) * 0.02
        else:
            bound = 1.0 / math.sqrt(last_fan_in)
            t = (torch.rand(shp, generator=gen) * 2 - 1) * bound
        out[name] = t.float()
    return out


def init_all(cfg: PathConfig, seed: int):
    """Four parameter dicts (Enc, Dec, Res, Cls) from one seed."""
    gen = torch.Generator().manual_seed(seed)
    return (init_params(encoder_param_shapes(cfg), gen),
            init_params(decoder_param_shapes(cfg), gen),
            init_params(restorer_param_shapes(cfg), gen),
            init_params(classifier_param_shapes(cfg), gen))


def synthetic_batch(cfg: PathConfig, batch: int, seed: int):
    """Synthetic (CIR, Err, Label) of the shape/dtype ``UWBDataset.__getitem__`` yields
    (dataset.py:118-133): CIR (B,157) f32 ~ N(0,1) (StandardScaler'd, dataset.py:73-76),
    Err (B,1) f32 = clip(|N(0,0.15)|,0,1), Label (B,1) f32 holding integers in [0,NC)."""
    gen = torch.Generator().manual_seed(seed)
    cir = torch.randn(batch, cfg.cir_len, generator=gen)
    err = (torch.randn(batch, 1, generator=gen) * 0.15).abs().clamp_(0, 1)
    label = torch.randint(0, cfg.num_classes, (batch, 1), generator=gen).float()
    return cir, err, label


# ------------------------------------------------------------------------ primitives
# ---------------------------------------------------------------------------------------------------
# Operand rounding (test infrastructure for the bf16 compute mode, BASELINE configs[2]).
# With ``operand_rounding("bf16")`` active every Conv1d / Linear of the restatement rounds its two GEMM operands to
# bfloat16 (round-to-nearest-even) and accumulates in the working precision -- forward (x, W), data gradient (dy, W) and
# weight gradient (x, dy) -- which is the arithmetic the reference would see with bf16 tensor-core operands and fp32
# accumulation / statistics.  Its distance from the fp64 evaluation is the yardstick the bf16 parity test uses
# (tests/test_gpu_parity.py::test_bf16_mode_full_gradient_parity): the CUDA path must not be further from exact
# arithmetic than a small multiple of what ANY bf16-operand evaluation of the reference is.
_OPERAND_ROUNDING = None


class operand_rounding:
    def __init__(self, kind):
        assert kind in (None, "bf16")
        self.kind = kind

    def __enter__(self):
        global _OPERAND_ROUNDING
        self.prev, _OPERAND_ROUNDING = _OPERAND_ROUNDING, self.kind
        return self

    def __exit__(self, *exc):
        global _OPERAND_ROUNDING
        _OPERAND_ROUNDING = self.prev


def _rnd(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundedConv1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        xr, wr = _rnd(x), _rnd(w)
        ctx.save_for_backward(xr, wr)
        ctx.conf = (stride, padding, b is not None)
        return F.conv1d(xr, wr, b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        stride, padding, has_bias = ctx.conf
        gr = _rnd(gy)
        gx = torch.nn.grad.conv1d_input(xr.shape, wr, gr, stride=stride, padding=padding)
        gw = torch.nn.grad.conv1d_weight(xr, wr.shape, gr, stride=stride, padding=padding)
        return gx, gw, (gy.sum((0, 2)) if has_bias else None), None, None


class _RoundedLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = _rnd(x), _rnd(w)
        ctx.save_for_backward(xr, wr)
        ctx.has_bias = b is not None
        return F.linear(xr, wr, b)

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        gr = _rnd(gy)
        return gr @ wr, gr.t() @ xr, (gy.sum(0) if ctx.has_bias else None)


def _conv1d(x, w, b=None, stride=1, padding=0):
    if _OPERAND_ROUNDING is None:
        return F.conv1d(x, w, b, stride=stride, padding=padding)
    return _RoundedConv1d.apply(x, w, b, stride, padding)


def _linear(x, w, b=None):
    if _OPERAND_ROUNDING is None:
        return F.linear(x, w, b)
    return _RoundedLinear.apply(x, w, b)


def adaptive_pool_windows(lin: int, lout: int):
    """AdaptiveAvgPool1d window table: start=floor(i*lin/lout), end=ceil((i+1)*lin/lout)
    (torch semantics; used at models.py:146, 264, 279, 436)."""
    return [((i * lin) // lout, -((-(i + 1) * lin) // lout)) for i in range(lout)]


_POOL_CACHE = {}


def adaptive_pool_matrix(lin: int, lout: int, dtype=torch.float32) -> torch.Tensor:
    """(lout, lin) averaging matrix of the window table above (row i = 1/len on window i)."""
    key = (lin, lout, dtype)
    if key not in _POOL_CACHE:
        m = torch.zeros(lout, lin, dtype=dtype)
        for i, (s, e) in enumerate(adaptive_pool_windows(lin, lout)):
            m[i, s:e] = 1.0 / (e - s)
        _POOL_CACHE[key] = m
    return _POOL_CACHE[key]


def adaptive_avg_pool1d(x: torch.Tensor, lout: int) -> torch.Tensor:
    return x @ adaptive_pool_matrix(x.shape[-1], lout, x.dtype).t()


def reflect_index(u: int, n: int) -> int:
    """ReflectionPad1d source index (no edge repeat)."""
    if u < 0:
        return -u
    if u >= n:
        return 2 * (n - 1) - u
    return u


def reflection_pad1d(x: torch.Tensor, p: int) -> torch.Tensor:
    """out[q] = x[reflect_index(q - p)]: the p samples next to each edge mirrored, edge not repeated."""
    left = x[..., 1:p + 1].flip(-1)
    right = x[..., -p - 1:-1].flip(-1)
    return torch.cat([left, x, right], dim=-1)


def instance_norm1d(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm1d defaults (models.py:152,160): per (b,c) over L, biased variance,
    no affine, batch statistics always."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def adaptive_instance_norm1d(x, weight_bc, bias_bc, eps: float = 1e-5):
    """AdaptiveInstanceNorm1d.forward (models.py:1061-1076): F.batch_norm in training mode on
    a (1, B*C, L) view == instance norm with a per-(b,c) scale and shift."""
    b, c, _ = x.shape
    return instance_norm1d(x, eps) * weight_bc.view(b, c, 1) + bias_bc.view(b, c, 1)


def custom_layer_norm(x, gamma, beta, eps: float = 1e-5):
    """models.py:976-985: per-sample mean and UNBIASED std over (C*L); eps added to std."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    mean = flat.mean(1).view(b, 1, 1)
    std = flat.std(1).view(b, 1, 1)            # unbiased (n-1)
    y = (x - mean) / (std + eps)
    return y * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def upsample_nearest2(x: torch.Tensor) -> torch.Tensor:
    """nn.Upsample(scale_factor=2) default mode 'nearest' (models.py:422): out[j]=in[j//2]."""
    return x.repeat_interleave(2, dim=-1)


# --------------------------------------------------------------------------- modules
def range_encoder(p, x1, cfg: PathConfig, taps=None):
    """RangeEncoder1d (models.py:140-176).  x1: (B,1,L) -> (B,range_dim,code_len)."""
    pre = "range_encoder.model."
    h = adaptive_avg_pool1d(x1, cfg.pooled_len)
    idx = 2
    h = _conv1d(reflection_pad1d(h, 3), p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"])
    h = torch.relu(instance_norm1d(h))
    if taps is not None:
        taps["r0"] = h
    idx += 3
    for i in range(cfg.n_downsample):
        h = _conv1d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"], stride=2, padding=1)
        h = torch.relu(instance_norm1d(h))
        if taps is not None:
            taps[f"r{i + 1}"] = h
        idx += 3
    for i in range(cfg.n_residual):                                  # models.py:988-1005
        q = f"{pre}{idx}.block."
        t = _conv1d(reflection_pad1d(h, 1), p[q + "1.weight"], p[q + "1.bias"])
        t = torch.relu(instance_norm1d(t))
        t = _conv1d(reflection_pad1d(t, 1), p[q + "5.weight"], p[q + "5.bias"])
        h = h + instance_norm1d(t)
        if taps is not None:
            taps[f"rres{i}"] = h
        idx += 1
    h = torch.relu(_conv1d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"]))
    return h


def env_encoder(p, x1, cfg: PathConfig, noise=None, taps=None):
    """EnvEncoder1d (models.py:258-298).  Returns cat (B,E,1), latent (B,E/2,1), kl ()."""
    pre = "env_encoder.model."
    h = adaptive_avg_pool1d(x1, cfg.pooled_len)
    idx = 2
    h = torch.relu(_conv1d(reflection_pad1d(h, 3), p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"]))
    if taps is not None:
        taps["e0"] = h
    idx += 2
    n_conv = 2 + max(0, cfg.n_downsample - 2 - 2)
    for i in range(n_conv):
        h = torch.relu(_conv1d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"], stride=2, padding=1))
        if taps is not None:
            taps[f"e{i + 1}"] = h
        idx += 2
    h = h.mean(dim=-1, keepdim=True)                                  # AdaptiveAvgPool1d(1)
    idx += 1
    cat = _conv1d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"])
    half = cat.shape[1] // 2
    mu, log_sigma = cat[:, :half], cat[:, half:]
    if noise is None:
        noise = torch.randn_like(mu)
    latent = noise * log_sigma.exp() + mu                             # models.py:288
    kl = 0.5 * torch.sum((2 * log_sigma).exp() + mu ** 2 - 1 - 2 * log_sigma, dim=1)
    return cat, latent, kl.mean()                                     # models.py:294-298


def encoder(p, x, cfg: PathConfig, noise=None, taps=None):
    """Encoder.forward (models.py:49-61)."""
    x1 = x.view(x.size(0), 1, x.size(1))
    rc = range_encoder(p, x1, cfg, taps)
    cat, latent, kl = env_encoder(p, x1, cfg, noise, taps)
    return rc, cat, latent, kl


def decoder(p, range_code, env_code, cfg: PathConfig, taps=None):
    """Decoder.forward -> Decoder1d.forward (models.py:81-91, 405-471)."""
    b = range_code.shape[0]
    D = cfg.trunk_dim
    s = env_code.reshape(b, -1)                                       # MLP.forward :961
    s = torch.relu(_linear(s, p["decoder.mlp.model.0.weight"], p["decoder.mlp.model.0.bias"]))
    s = torch.relu(_linear(s, p["decoder.mlp.model.2.weight"], p["decoder.mlp.model.2.bias"]))
    adain = _linear(s, p["decoder.mlp.model.4.weight"], p["decoder.mlp.model.4.bias"])
    if taps is not None:
        taps["adain"] = adain
    pre = "decoder.model."
    h = torch.relu(_conv1d(range_code, p[pre + "0.weight"], p[pre + "0.bias"]))
    idx = 2
    off = 0
    for i in range(cfg.n_residual):
        q = f"{pre}{idx}.block."
        # assign_adain_params (models.py:452-464): per AdaIN layer, bias first then weight.
        b1, w1 = adain[:, off:off + D], adain[:, off + D:off + 2 * D]
        b2, w2 = adain[:, off + 2 * D:off + 3 * D], adain[:, off + 3 * D:off + 4 * D]
        off += 4 * D
        t = _conv1d(reflection_pad1d(h, 1), p[q + "1.weight"], p[q + "1.bias"])
        t = torch.relu(adaptive_instance_norm1d(t, w1, b1))
        t = _conv1d(reflection_pad1d(t, 1), p[q + "5.weight"], p[q + "5.bias"])
        h = h + adaptive_instance_norm1d(t, w2, b2)
        if taps is not None:
            taps[f"dres{i}"] = h
        idx += 1
    for i in range(cfg.n_downsample):
        h = _conv1d(upsample_nearest2(h), p[f"{pre}{idx + 1}.weight"], p[f"{pre}{idx + 1}.bias"], padding=2)
        h = torch.relu(custom_layer_norm(h, p[f"{pre}{idx + 2}.gamma"], p[f"{pre}{idx + 2}.beta"]))
        if taps is not None:
            taps[f"u{i + 1}"] = h
        idx += 4
    h = torch.tanh(_conv1d(reflection_pad1d(h, 3), p[f"{pre}{idx + 1}.weight"], p[f"{pre}{idx + 1}.bias"]))
    h = adaptive_avg_pool1d(h, cfg.cir_len)
    return h.squeeze()                                                # models.py:90


def restorer(p, range_code):
    """RestorerLinear.forward, soft=False branch (models.py:642-658)."""
    h = range_code.reshape(range_code.size(0), -1)
    for i in (0, 2, 4):
        h = F.leaky_relu(_linear(h, p[f"restorer.layers.{i}.weight"], p[f"restorer.layers.{i}.bias"]), 0.2)
    return _linear(h, p["restorer.linear_layer1.weight"], p["restorer.linear_layer1.bias"])


def classifier(p, env_code):
    """ClassifierLinear.forward (models.py:858-862); note LeakyReLU(0.2) on the logits (:854)."""
    h = env_code.reshape(env_code.size(0), -1)
    for i, slope in zip((0, 2, 4, 6), (0.01, 0.01, 0.01, 0.2)):
        h = F.leaky_relu(_linear(h, p[f"classifier.layers.{i}.weight"], p[f"classifier.layers.{i}.bias"]), slope)
    return h



# ------------------------------------------------------------------ Conv1d heads
BN_EPS = 0.8                 # ``nn.BatchNorm1d(out_filters, 0.8)``: the second positional argument is eps (models.py:676, 881)
BN_MOMENTUM = 0.1
DROPOUT_P = 0.25


def restorer_conv1d_param_shapes(cfg: PathConfig) -> "OrderedDict[str, tuple]":
    """``RestorerConv1d`` (models.py:661-693): state_dict order incl. the BatchNorm buffers."""
    sh = OrderedDict()
    sh["restorer.conv_blocks.0.weight"] = (16, cfg.range_dim, 4)
    sh["restorer.conv_blocks.0.bias"] = (16,)
    sh["restorer.conv_blocks.3.weight"] = (32, 16, 4)
    sh["restorer.conv_blocks.3.bias"] = (32,)
    sh["restorer.conv_blocks.6.weight"] = (32,)
    sh["restorer.conv_blocks.6.bias"] = (32,)
    sh["restorer.conv_blocks.6.running_mean"] = (32,)
    sh["restorer.conv_blocks.6.running_var"] = (32,)
    sh["restorer.conv_blocks.6.num_batches_tracked"] = ()
    sh["restorer.linear_layer1.weight"] = (1, 64)
    sh["restorer.linear_layer1.bias"] = (1,)
    sh["restorer.linear_layer2.0.weight"] = (2, 64)      # unused when soft=False
    sh["restorer.linear_layer2.0.bias"] = (2,)
    return sh


def classifier_conv1d_param_shapes(cfg: PathConfig, filters: int = 16) -> "OrderedDict[str, tuple]":
    """``ClassifierConv1d`` (models.py:865-891)."""
    sh = OrderedDict()
    sh["classifier.conv_blocks.0.weight"] = (filters, cfg.env_dim, 1)
    sh["classifier.conv_blocks.0.bias"] = (filters,)
    sh["classifier.conv_blocks.3.weight"] = (filters, filters, 1)
    sh["classifier.conv_blocks.3.bias"] = (filters,)
    sh["classifier.conv_blocks.6.weight"] = (filters,)
    sh["classifier.conv_blocks.6.bias"] = (filters,)
    sh["classifier.conv_blocks.6.running_mean"] = (filters,)
    sh["classifier.conv_blocks.6.running_var"] = (filters,)
    sh["classifier.conv_blocks.6.num_batches_tracked"] = ()
    sh["classifier.linear.0.weight"] = (cfg.num_classes, filters)
    sh["classifier.linear.0.bias"] = (cfg.num_classes,)
    return sh


def init_conv_head_params(shapes, gen):
    """Reference init distributions (models.py:8-14 via .apply(weights_init_normal)): Conv weights N(0, 0.02), BatchNorm weight
    N(1, 0.02) and bias 0, conv biases / Linear torch defaults; buffers at their PyTorch initial values."""
    out = OrderedDict()
    fan_in = 1
    for name, shp in shapes.items():
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros((), dtype=torch.long)
            continue
        if name.endswith("running_mean"):
            out[name] = torch.zeros(shp)
            continue
        if name.endswith("running_var"):
            out[name] = torch.ones(shp)
            continue
        if len(shp) >= 2 and name.endswith("weight"):
            fan_in = int(np.prod(shp[1:]))
        if ".6.weight" in name:
            t = 1.0 + torch.randn(shp, generator=gen) * 0.02
        elif ".6.bias" in name:
            t = torch.zeros(shp)
        elif len(shp) == 3:
            t = torch.randn(shp, generator=gen) * 0.02
        else:
            t = (torch.rand(shp, generator=gen) * 2 - 1) / math.sqrt(fan_in)
        out[name] = t.float()
    return out


def dropout_apply(x, mask, training: bool):
    """nn.Dropout(0.25) (models.py:672, 877) with an explicit keep-mask (0 / 1, same shape as x): kept values are scaled by
    1 / (1 - p); identity in eval mode."""
    if not training:
        return x
    return x * mask.to(x.dtype) / (1.0 - DROPOUT_P)


def batch_norm1d(x, weight, bias, running_mean, running_var, training: bool, eps: float = BN_EPS):
    """nn.BatchNorm1d(C, eps=0.8) on (B, C, L): training -> batch statistics over (B, L), biased variance; returns
    (y, new_running_mean, new_running_var) with the momentum-0.1 update (unbiased variance) PyTorch applies; eval -> running
    statistics."""
    if training:
        n = x.shape[0] * x.shape[2]
        mean = x.mean(dim=(0, 2))
        var = x.var(dim=(0, 2), unbiased=False)
        new_rm = (1 - BN_MOMENTUM) * running_mean.to(x.dtype) + BN_MOMENTUM * mean.detach()
        new_rv = (1 - BN_MOMENTUM) * running_var.to(x.dtype) + BN_MOMENTUM * var.detach() * (n / max(n - 1, 1))
    else:
        mean, var, new_rm, new_rv = running_mean.to(x.dtype), running_var.to(x.dtype), running_mean, running_var
    y = (x - mean.view(1, -1, 1)) / torch.sqrt(var.view(1, -1, 1) + eps) * weight.view(1, -1, 1) + bias.view(1, -1, 1)
    return y, new_rm, new_rv


def restorer_conv1d(p, range_code, masks=None, training: bool = True):
    """RestorerConv1d.forward (models.py:702-716), soft=False.  masks = (m1 (B,16,4), m2 (B,32,2)) dropout keep-masks.
    Returns (err_est (B,1), (new_running_mean, new_running_var))."""
    q = "restorer.conv_blocks."
    m1, m2 = masks if masks is not None else (None, None)
    h = F.leaky_relu(_conv1d(range_code, p[q + "0.weight"], p[q + "0.bias"], stride=2, padding=1), 0.2)
    h = dropout_apply(h, m1, training)
    h = F.leaky_relu(_conv1d(h, p[q + "3.weight"], p[q + "3.bias"], stride=2, padding=1), 0.2)
    h = dropout_apply(h, m2, training)
    h, rm, rv = batch_norm1d(h, p[q + "6.weight"], p[q + "6.bias"], p[q + "6.running_mean"], p[q + "6.running_var"], training)
    flat = h.reshape(h.shape[0], -1)
    return _linear(flat, p["restorer.linear_layer1.weight"], p["restorer.linear_layer1.bias"]), (rm, rv)


def classifier_conv1d(p, env_code, masks=None, training: bool = True):
    """ClassifierConv1d.forward (models.py:893-902).  masks = (m1 (B,F,1), m2 (B,F,1))."""
    q = "classifier.conv_blocks."
    m1, m2 = masks if masks is not None else (None, None)
    h = env_code.reshape(env_code.shape[0], -1).unsqueeze(2)
    h = F.leaky_relu(_conv1d(h, p[q + "0.weight"], p[q + "0.bias"]), 0.2)
    h = dropout_apply(h, m1, training)
    h = F.leaky_relu(_conv1d(h, p[q + "3.weight"], p[q + "3.bias"]), 0.2)
    h = dropout_apply(h, m2, training)
    h, rm, rv = batch_norm1d(h, p[q + "6.weight"], p[q + "6.bias"], p[q + "6.running_mean"], p[q + "6.running_var"], training)
    logits = F.leaky_relu(_linear(h.reshape(h.shape[0], -1), p["classifier.linear.0.weight"], p["classifier.linear.0.bias"]), 0.2)
    return logits, (rm, rv)


def emnet(pe, pr, pc, cir, cfg: PathConfig, noise=None):
    """The composite ``network(cir) -> (label_est, env_latent, err_est)`` used by
    train.py:82 / test.py:73.  ``EMNet`` is missing from the reference (run.py:59-62 is the
    only trace); per SURVEY.md section 8(b) it is Encoder -> (Classifier(env_code), env_code,
    Restorer(range_code))."""
    rc, cat, _, _ = encoder(pe, cir, cfg, noise)
    return classifier(pc, cat), cat, restorer(pr, rc)


# ------------------------------------------------------------------------------ losses
LAMBDA_AE, LAMBDA_RES, LAMBDA_RANGE, LAMBDA_ENV = 1.0, 10.0, 1.0, 1.0     # train_semi.py:111-114


def l1_mean(a, b):
    return (a - b).abs().mean()                                       # torch.nn.L1Loss()


def cross_entropy_mean(logits, target):
    """torch.nn.CrossEntropyLoss() default: mean over batch of logsumexp(z) - z[target]."""
    lse = torch.logsumexp(logits, dim=1)
    return (lse - logits.gather(1, target.view(-1, 1)).squeeze(1)).mean()


def semi_forward(pe, pd, pr, pc, cir, err, label, cfg: PathConfig, supervised: bool, noise=None, taps=None, head_fns=None):
    """One forward of the semi-supervised step (train_semi.py:186-225).  ``head_fns`` = dict(res=f(p, range_code) -> err_fake,
    cls=f(p, env_code) -> logits) selects other heads than the Linear ones (net_type='Conv1d': restorer_conv1d / classifier_conv1d
    with their dropout masks bound)."""
    rc, cat, latent, kl = encoder(pe, cir, cfg, noise, taps)
    cir_gen = decoder(pd, rc, cat, cfg, taps)
    err_fake = restorer(pr, rc) if head_fns is None else head_fns["res"](pr, rc)
    label_fake = classifier(pc, cat) if head_fns is None else head_fns["cls"](pc, cat)
    out = dict(range_code=rc, env_code=cat, env_code_rv=latent, kl=kl,
               cir_gen=cir_gen, err_fake=err_fake, label_fake=label_fake)
    out["loss_ae"] = LAMBDA_AE * l1_mean(cir, cir_gen.reshape(cir.shape))        # :199
    out["loss_range"] = LAMBDA_RANGE * kl                                        # :200
    if not supervised:                                                           # :204-206
        out["loss"] = out["loss_ae"] + out["loss_range"]
        return out
    tgt = label.to(torch.int64).squeeze(-1) if label.dim() > 1 else label.to(torch.int64)
    out["loss_res"] = LAMBDA_RES * l1_mean(err, err_fake)                        # :218
    out["loss_env"] = LAMBDA_ENV * cross_entropy_mean(label_fake, tgt)           # :220 (room_full)
    out["loss"] = out["loss_ae"] + out["loss_range"] + out["loss_res"] + out["loss_env"]   # :225
    return out


def supervised_forward(pe, pr, pc, cir, err, label, cfg: PathConfig, noise=None):
    """train.py:82-91 (lambda_idy = lambda_reg = 1, :51-52)."""
    label_est, env_latent, err_est = emnet(pe, pr, pc, cir, cfg, noise)
    tgt = label.to(torch.int64).squeeze(-1) if label.dim() > 1 else label.to(torch.int64)
    loss_idy = cross_entropy_mean(label_est, tgt)
    loss_reg = l1_mean(err_est, err)
    return dict(label_est=label_est, env_latent=env_latent, err_est=err_est,
                loss_idy=loss_idy, loss_reg=loss_reg, loss=loss_idy + loss_reg)


def batch_metrics(err_est, err_gt, logits, label):
    """Per-batch RMSE / MAE / accuracy (train.py:104-115, test.py:76-85)."""
    tgt = label.to(torch.int64).reshape(-1)
    rmse = torch.mean((err_est - err_gt) ** 2) ** 0.5
    mae = torch.mean(torch.abs(err_est - err_gt))
    pred = torch.argmax(logits, dim=1)
    acc = torch.sum(pred == tgt).float() / tgt.shape[0]
    return rmse, mae, acc, pred


def supervision_mask(rng: np.random.RandomState, rate: float) -> int:
    """train_semi.py:203: ``0 if np.random.randn(1) > rate else 1`` (a NORMAL draw)."""
    return 0 if rng.randn(1)[0] > rate else 1


def lambda_lr(epoch: int, n_epochs: int = 500, offset: int = 0, decay_start: int = 100) -> float:
    """models.py:24-25."""
    return 1.0 - max(0, epoch + offset - decay_start) / (n_epochs - decay_start)


# -------------------------------------------------------------------------------- Adam
class AdamState:
    """torch.optim.Adam defaults restated (train_semi.py:118-122): eps 1e-8, no weight decay,
    per-parameter step counter, parameters whose grad is None are skipped entirely."""

    def __init__(self, params: "dict[str, torch.Tensor]", lr=1e-4, betas=(0.5, 0.999), eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}

    def step(self, params, grads, lr_scale: float = 1.0):
        for k, p in params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.t[k] += 1
            t = self.t[k]
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            denom = self.v[k].sqrt() / math.sqrt(bc2) + self.eps
            params[k] = p - (self.lr * lr_scale / bc1) * self.m[k] / denom
        return params


# --------------------------------------------------------------------- one train step
def trainable(params: dict) -> dict:
    return {k: v for k, v in params.items() if not is_buffer(k)}


def semi_step_with_grads(pe, pd, pr, pc, cir, err, label, cfg, supervised, noise=None, head_fns=None):
    """Forward + autograd backward of the semi step; returns (out, grads) where grads maps
    'enc.<key>' / 'dec.<key>' / 'res.<key>' / 'cls.<key>' -> tensor or None (None == the
    parameter received no gradient, e.g. restorer.linear_layer2 always, Res/Cls when the
    batch is unsupervised; train_semi.py:204-214)."""
    groups = dict(enc=pe, dec=pd, res=pr, cls=pc)
    leaves = {}
    work = {}
    for g, pdict in groups.items():
        work[g] = {}
        for k, v in pdict.items():
            t = v.detach().clone()
            if not is_buffer(k):
                t.requires_grad_(True)
                leaves[f"{g}.{k}"] = t
            work[g][k] = t
    out = semi_forward(work["enc"], work["dec"], work["res"], work["cls"], cir, err, label, cfg, supervised, noise, head_fns=head_fns)
    out["loss"].backward()
    grads = {k: (t.grad.detach() if t.grad is not None else None) for k, t in leaves.items()}
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    return out, grads
