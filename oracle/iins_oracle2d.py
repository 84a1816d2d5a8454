"""CPU oracle for the 2-D variant of the IIns-VAE path (conv_type = 2, expand = True)  --  TEST INFRASTRUCTURE ONLY.

Functional restatement (plain torch CPU ops on flat dicts of named parameters) of ``RangeEncoder2d`` (models.py:179-215),
``EnvEncoder2d`` (:304-346), ``Decoder2d`` (:474-539), ``ResidualBlock2d`` (:1008-1025), ``AdaptiveInstanceNorm2d`` (:1082-1113)
and the ``Encoder`` / ``Decoder`` glue for conv_type = 2 with expand = True (:49-61, :81-91) of JadeLilyx/IIns-VAE.  Same rules as
``oracle/iins_oracle.py``: never imported by the product package; only ``tests/`` (and the fixture generator) use it.

Pinned against outputs of the reference itself: ``tests/golden/make_golden2d.py`` runs the live ``/root/reference/models.py``
modules with these parameters and commits outputs + gradients (``tests/golden/iins_golden2d.npz``);
``tests/test_oracle_golden.py::test_oracle2d_matches_reference_fixture`` checks this file against them.

Convolutions are ``F.conv2d``; every other operator (2-D adaptive pooling as the outer product of the 1-D window tables,
reflection padding, instance / adaptive-instance norm over (H, W), the custom LayerNorm over (C, H, W), nearest upsampling)
is restated from its published definition.  All ``file:line`` citations are into ``/root/reference``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from oracle import iins_oracle as orc


# --------------------------------------------------------------------------- parameter shapes
def _to2d(shapes):
    """Conv1d weight (co, ci, k) -> Conv2d weight (co, ci, k, k); everything else (biases, Linear, LayerNorm, buffers) unchanged."""
    return OrderedDict((k, tuple(v) + (v[-1],) if (k.endswith("weight") and len(v) == 3) else tuple(v)) for k, v in shapes.items())


def encoder_param_shapes(cfg):
    return _to2d(orc.encoder_param_shapes(cfg))


def decoder_param_shapes(cfg, buffers=True):
    return _to2d(orc.decoder_param_shapes(cfg, buffers))


def restorer_param_shapes(cfg):
    """RestorerLinear on the flattened (range_dim, 8, 8) code (models.py:619-633 with code_shape of three entries)."""
    sh = orc.restorer_param_shapes(cfg)
    sh["restorer.layers.0.weight"] = (512, cfg.range_dim * cfg.code_len * cfg.code_len)
    return sh


def init_params(shapes, gen):
    """The 1-D oracle's init rules for this variant's shapes: conv weights N(0, 0.02) (models.py:8-11), conv biases and Linear
    layers U(-1/sqrt(fan_in), 1/sqrt(fan_in)), LayerNorm gamma U(0, 1), beta 0 (:973-974), AdaIN dummy buffers 0 / 1."""
    res = OrderedDict()
    last_fan_in = 1
    for name, shp in shapes.items():
        if name.endswith("running_mean") or name.endswith("beta"):
            res[name] = torch.zeros(shp)
        elif name.endswith("running_var"):
            res[name] = torch.ones(shp)
        elif name.endswith("gamma"):
            res[name] = torch.rand(shp, generator=gen)
        elif len(shp) == 4:
            res[name] = (torch.randn(shp, generator=gen) * 0.02).float()
            last_fan_in = shp[1] * shp[2] * shp[3]
        else:
            if name.endswith("weight"):
                last_fan_in = shp[1]
            res[name] = ((torch.rand(shp, generator=gen) * 2 - 1) / last_fan_in ** 0.5).float()
    return res


def init_all(cfg, seed):
    gen = torch.Generator().manual_seed(seed)
    return (init_params(encoder_param_shapes(cfg), gen), init_params(decoder_param_shapes(cfg), gen),
            init_params(restorer_param_shapes(cfg), gen))


# --------------------------------------------------------------------------- primitives
def adaptive_avg_pool2d(x, out):
    """nn.AdaptiveAvgPool2d(out) (models.py:185, 311, 509): windows are the 1-D table in each direction."""
    mh = orc.adaptive_pool_matrix(x.shape[-2], out, x.dtype)
    mw = orc.adaptive_pool_matrix(x.shape[-1], out, x.dtype)
    return mh @ x @ mw.t()


def reflection_pad2d(x, p):
    """nn.ReflectionPad2d(p): the 1-D rule (edge not repeated) along W, then along H."""
    x = orc.reflection_pad1d(x, p)
    return orc.reflection_pad1d(x.transpose(-1, -2), p).transpose(-1, -2)


def instance_norm2d(x, eps=1e-5):
    """nn.InstanceNorm2d defaults (models.py:191, 199): per (b, c) over (H, W), biased variance, no affine."""
    mean = x.mean(dim=(-2, -1), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(-2, -1), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def adaptive_instance_norm2d(x, weight_bc, bias_bc, eps=1e-5):
    """AdaptiveInstanceNorm2d.forward (models.py:1095-1110): F.batch_norm in training mode on a (1, B*C, H, W) view."""
    b, c = x.shape[:2]
    return instance_norm2d(x, eps) * weight_bc.view(b, c, 1, 1) + bias_bc.view(b, c, 1, 1)


def custom_layer_norm(x, gamma, beta, eps=1e-5):
    """models.py:976-985 on a 4-D tensor: per-sample mean and UNBIASED std over (C, H, W); eps added to std."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    mean = flat.mean(1).view(b, 1, 1, 1)
    std = flat.std(1).view(b, 1, 1, 1)
    return (x - mean) / (std + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def upsample_nearest2(x):
    """nn.Upsample(scale_factor=2), mode 'nearest', on (B, C, H, W) (models.py:494)."""
    return x.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)


# --------------------------------------------------------------------------- modules
def expand_input(x):
    """Encoder.forward, conv_type != 1 and expand = True (models.py:55): (B, L) -> (B, 1, L, L) with x[h] at every column."""
    return x.view(x.size(0), 1, x.size(1), 1).expand(x.size(0), 1, x.size(1), x.size(1))


def range_encoder(p, x2, cfg, taps=None):
    """RangeEncoder2d (models.py:179-215).  x2: (B, 1, L, L) -> (B, range_dim, 8, 8)."""
    pre = "range_encoder.model."
    h = adaptive_avg_pool2d(x2, cfg.pooled_len)
    idx = 2
    h = F.conv2d(reflection_pad2d(h, 3), p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"])
    h = torch.relu(instance_norm2d(h))
    if taps is not None:
        taps["r0"] = h
    idx += 3
    for i in range(cfg.n_downsample):
        h = torch.relu(instance_norm2d(F.conv2d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"], stride=2, padding=1)))
        if taps is not None:
            taps[f"r{i + 1}"] = h
        idx += 3
    for i in range(cfg.n_residual):                                   # ResidualBlock2d (models.py:1008-1025)
        q = f"{pre}{idx}.block."
        t = torch.relu(instance_norm2d(F.conv2d(reflection_pad2d(h, 1), p[q + "1.weight"], p[q + "1.bias"])))
        t = F.conv2d(reflection_pad2d(t, 1), p[q + "5.weight"], p[q + "5.bias"])
        h = h + instance_norm2d(t)
        idx += 1
    return torch.relu(F.conv2d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"]))


def env_encoder(p, x2, cfg, noise=None, taps=None):
    """EnvEncoder2d (models.py:304-346).  Returns cat (B, E, 1, 1), latent (B, E/2, 1, 1), kl ()."""
    pre = "env_encoder.model."
    h = adaptive_avg_pool2d(x2, cfg.pooled_len)
    idx = 2
    h = torch.relu(F.conv2d(reflection_pad2d(h, 3), p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"]))
    idx += 2
    for i in range(2 + max(0, cfg.n_downsample - 2 - 2)):
        h = torch.relu(F.conv2d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"], stride=2, padding=1))
        if taps is not None:
            taps[f"e{i + 1}"] = h
        idx += 2
    h = h.mean(dim=(-2, -1), keepdim=True)                             # AdaptiveAvgPool2d(1)
    idx += 1
    cat = F.conv2d(h, p[f"{pre}{idx}.weight"], p[f"{pre}{idx}.bias"])
    half = cat.shape[1] // 2
    mu, log_sigma = cat[:, :half], cat[:, half:]
    if noise is None:
        noise = torch.randn_like(mu)
    latent = noise * log_sigma.exp() + mu                              # models.py:336
    kl = 0.5 * torch.sum((2 * log_sigma).exp() + mu ** 2 - 1 - 2 * log_sigma, dim=1)
    return cat, latent, kl.mean()                                      # models.py:341-345


def encoder(p, x, cfg, noise=None, taps=None):
    """Encoder.forward with conv_type = 2, expand = True (models.py:49-61)."""
    x2 = expand_input(x)
    rc = range_encoder(p, x2, cfg, taps)
    cat, latent, kl = env_encoder(p, x2, cfg, noise, taps)
    return rc, cat, latent, kl


def decoder(p, range_code, env_code, cfg, taps=None):
    """Decoder.forward -> Decoder2d.forward (models.py:81-91, 474-539); expand = True keeps x_recon[:, :, :, 0]."""
    b = range_code.shape[0]
    D = cfg.trunk_dim
    s = env_code.reshape(b, -1)                                        # MLP.forward (models.py:961)
    s = torch.relu(F.linear(s, p["decoder.mlp.model.0.weight"], p["decoder.mlp.model.0.bias"]))
    s = torch.relu(F.linear(s, p["decoder.mlp.model.2.weight"], p["decoder.mlp.model.2.bias"]))
    adain = F.linear(s, p["decoder.mlp.model.4.weight"], p["decoder.mlp.model.4.bias"])
    pre = "decoder.model."
    h = torch.relu(F.conv2d(range_code, p[pre + "0.weight"], p[pre + "0.bias"]))
    idx, off = 2, 0
    for _ in range(cfg.n_residual):
        q = f"{pre}{idx}.block."
        # assign_adain_params (models.py:520-533): per AdaIN layer, bias (mean) first then weight (std)
        b1, w1 = adain[:, off:off + D], adain[:, off + D:off + 2 * D]
        b2, w2 = adain[:, off + 2 * D:off + 3 * D], adain[:, off + 3 * D:off + 4 * D]
        off += 4 * D
        t = torch.relu(adaptive_instance_norm2d(F.conv2d(reflection_pad2d(h, 1), p[q + "1.weight"], p[q + "1.bias"]), w1, b1))
        t = F.conv2d(reflection_pad2d(t, 1), p[q + "5.weight"], p[q + "5.bias"])
        h = h + adaptive_instance_norm2d(t, w2, b2)
        idx += 1
    for i in range(cfg.n_downsample):
        h = F.conv2d(upsample_nearest2(h), p[f"{pre}{idx + 1}.weight"], p[f"{pre}{idx + 1}.bias"], padding=2)
        h = torch.relu(custom_layer_norm(h, p[f"{pre}{idx + 2}.gamma"], p[f"{pre}{idx + 2}.beta"]))
        if taps is not None:
            taps[f"u{i + 1}"] = h
        idx += 4
    h = torch.tanh(F.conv2d(reflection_pad2d(h, 3), p[f"{pre}{idx + 1}.weight"], p[f"{pre}{idx + 1}.bias"]))
    h = adaptive_avg_pool2d(h, cfg.cir_len)
    return h[:, :, :, 0].squeeze()                                     # models.py:90


def restorer(p, range_code):
    """RestorerLinear on the flattened 2-D code (models.py:642-658)."""
    return orc.restorer(p, range_code)


def step_loss(pe, pd, pr, cir, err, cfg, noise):
    """The unsupervised + range-error part of the semi-supervised loss (train_semi.py:199-225 without the classifier term, whose
    head is unchanged by conv_type):  L1(x, x_recon) + KL + 10 * L1(err, err_est)."""
    rc, cat, latent, kl = encoder(pe, cir, cfg, noise)
    xrec = decoder(pd, rc, cat, cfg)
    err_est = restorer(pr, rc)
    loss = orc.l1_mean(xrec, cir) + kl + 10.0 * orc.l1_mean(err_est, err)
    return loss, dict(rc=rc, cat=cat, latent=latent, kl=kl, xrec=xrec, err_est=err_est)
