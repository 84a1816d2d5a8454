"""Drop-in replacements for the nn.Module classes of the reference's ``models.py`` (1-D path).

Same constructor signatures, same ``forward`` signatures / return shapes, same ``state_dict`` keys:

    Encoder    models.py:32-64    Decoder     models.py:67-91
    Restorer   models.py:94-112   Classifier  models.py:115-132
    weights_init_normal :8-14     LambdaLR    :17-25

Underneath, every ``forward`` / ``backward`` is ONE call into the C ABI of ``libiins_b200.so``
(include/iins_b200.h) which sequences the hand-written sm_100a kernels.  The leaf modules
(``nn.Conv1d`` / ``nn.Linear`` instances, ``LayerNorm``, ``AdaptiveInstanceNorm1d``) exist only to own
the parameters under the reference's names and init distributions -- their own ``forward`` is never
used.  There is no CPU / eager fallback: non-CUDA inputs or a missing library raise.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from ._capi import IinsConfig, IinsHeadState, get_lib, ptr, ptr_array

__all__ = ["Encoder", "Decoder", "Restorer", "Classifier", "EMNet", "weights_init_normal", "LambdaLR",
           "LayerNorm", "AdaptiveInstanceNorm1d"]


def weights_init_normal(m):
    """models.py:8-14 (identical behaviour: Conv weights ~ N(0, 0.02); BatchNorm untouched here
    because the Linear heads have none)."""
    if isinstance(m, (nn.Conv1d, nn.Conv2d)):
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


class LambdaLR:
    """models.py:17-25."""

    def __init__(self, n_epochs, offset, decay_start_epoch):
        assert (n_epochs - decay_start_epoch) > 0, "Decay must start before the training session ends!"
        self.n_epochs = n_epochs
        self.offset = offset
        self.decay_start_epoch = decay_start_epoch

    def step(self, epoch):
        return 1.0 - max(0, epoch + self.offset - self.decay_start_epoch) / (self.n_epochs - self.decay_start_epoch)


# ------------------------------------------------------------------ parameter-holding leaves
class LayerNorm(nn.Module):
    """Parameter holder for the reference's custom LayerNorm (models.py:965-985): gamma ~ U(0,1), beta 0."""

    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__()
        self.num_features, self.eps, self.affine = num_features, eps, affine
        self.gamma = nn.Parameter(torch.Tensor(num_features).uniform_())
        self.beta = nn.Parameter(torch.zeros(num_features))


class AdaptiveInstanceNorm1d(nn.Module):
    """Holder of the dummy running_mean / running_var buffers (models.py:1057-1059) so checkpoints
    keep the reference's keys; the per-sample weight/bias come from the MLP inside the kernels."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))


class _Slots(nn.Module):
    """A bare container whose children are registered under explicit names ("2", "block", ...)."""

    def put(self, name, module):
        self.add_module(str(name), module)
        return module


def _res_block(features, adain, conv=nn.Conv1d):
    """ResidualBlock1d (models.py:988-1005) / ResidualBlock2d (:1008-1025): the same Sequential indices in both."""
    blk = _Slots()
    inner = blk.put("block", _Slots())
    inner.put(1, conv(features, features, 3))
    if adain:
        inner.put(2, AdaptiveInstanceNorm1d(features))      # AdaptiveInstanceNorm2d (:1082-1113) holds the same two buffers
    inner.put(5, conv(features, features, 3))
    if adain:
        inner.put(6, AdaptiveInstanceNorm1d(features))
    return blk


def _check_input(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"iins_vae_b200: {name} must be a CUDA tensor (this package has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"iins_vae_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _params_of(module):
    ps = list(module.parameters())
    for p in ps:
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("iins_vae_b200: parameters must be contiguous float32 CUDA tensors (call .cuda())")
    return ps


def _cfg(batch, cir_len=157, dim=4, n_residual=3, n_downsample=4, env_dim=16, range_dim=2, num_classes=2, filters=16, conv_type=1):
    return IinsConfig(int(batch), int(cir_len), int(dim), int(n_residual), int(n_downsample), int(env_dim),
                      int(range_dim), int(num_classes), int(filters), int(conv_type))


def _empty(n, dev):
    return torch.empty(int(n) + 16, dtype=torch.float32, device=dev)


# ----------------------------------------------------------------------------- autograd glue
class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, noise, opts, *params):
        lib = get_lib()
        B = x.shape[0]
        two_d = opts.get("conv_type", 1) == 2
        mod = "encoder2d" if two_d else "encoder"
        cfg = _cfg(B, x.shape[1], opts["dim"], opts["n_residual"], opts["n_downsample"], opts["env_dim"], opts["range_dim"],
                   conv_type=2 if two_d else 1)
        lib.check(lib.iins_validate_config(cfg), "Encoder config")
        dev = x.device
        E, R = opts["env_dim"], opts["range_dim"]
        code = 128 >> opts["n_downsample"]
        tail = (1, 1) if two_d else (1,)                            # models.py:60: (B, 8, 1, 1) / (B, 8, 1)
        rc = torch.empty((B, R, code, code) if two_d else (B, R, code), device=dev)
        cat = torch.empty((B, E) + tail, device=dev)
        lat = torch.empty((B, E // 2) + tail, device=dev)
        kl = torch.empty((), device=dev)
        ws = _empty(getattr(lib, f"iins_{mod}_ws_floats")(cfg), dev)
        ctx.mod = mod
        lib.check(getattr(lib, f"iins_{mod}_forward")(cfg, ptr_array(params), ptr(x), ptr(noise), opts["seed"], opts["offset"],
                                           ptr(rc), ptr(cat), ptr(lat), ptr(kl), ptr(ws), _stream()), "Encoder forward")
        ctx.cfg, ctx.opts, ctx.params = cfg, opts, params
        ctx.save_for_backward(rc, cat, ws, noise if noise is not None else torch.empty(0, device=dev))
        ctx.has_noise = noise is not None
        ctx.set_materialize_grads(False)
        return rc, cat, lat, kl

    @staticmethod
    def backward(ctx, d_rc, d_cat, d_lat, d_kl):
        lib = get_lib()
        rc, cat, ws, noise = ctx.saved_tensors
        dev = rc.device
        grads = [torch.zeros_like(p) for p in ctx.params]
        scratch = _empty(getattr(lib, f"iins_{ctx.mod}_scratch_floats")(ctx.cfg), dev)
        fix = lambda g: None if g is None else g.contiguous().float()
        d_rc, d_cat, d_lat, d_kl = fix(d_rc), fix(d_cat), fix(d_lat), fix(d_kl)
        lib.check(getattr(lib, f"iins_{ctx.mod}_backward")(ctx.cfg, ptr_array(ctx.params), ptr(noise) if ctx.has_noise else None,
                                            ctx.opts["seed"], ctx.opts["offset"], ptr(rc), ptr(cat), ptr(ws),
                                            ptr(d_rc), ptr(d_cat), ptr(d_lat), ptr(d_kl), ptr_array(grads),
                                            ptr(scratch), _stream()), "Encoder backward")
        return (None, None, None) + tuple(grads)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rc, cat, opts, *params):
        lib = get_lib()
        B = rc.shape[0]
        two_d = opts.get("conv_type", 1) == 2
        mod = "decoder2d" if two_d else "decoder"
        cfg = _cfg(B, opts["in_dim"], opts["dim"], opts["n_residual"], opts["n_downsample"], opts["env_dim"], opts["range_dim"],
                   conv_type=2 if two_d else 1)
        lib.check(lib.iins_validate_config(cfg), "Decoder config")
        dev = rc.device
        xrec = torch.empty(B, opts["in_dim"], device=dev)
        ws = _empty(getattr(lib, f"iins_{mod}_ws_floats")(cfg), dev)
        ctx.mod = mod
        lib.check(getattr(lib, f"iins_{mod}_forward")(cfg, ptr_array(params), ptr(rc), ptr(cat), ptr(xrec), ptr(ws), _stream()),
                  "Decoder forward")
        ctx.cfg, ctx.params = cfg, params
        ctx.save_for_backward(rc, cat, ws)
        return xrec

    @staticmethod
    def backward(ctx, d_xrec):
        lib = get_lib()
        rc, cat, ws = ctx.saved_tensors
        grads = [torch.zeros_like(p) for p in ctx.params]
        d_rc, d_cat = torch.empty_like(rc), torch.empty_like(cat)
        scratch = _empty(getattr(lib, f"iins_{ctx.mod}_scratch_floats")(ctx.cfg), rc.device)
        lib.check(getattr(lib, f"iins_{ctx.mod}_backward")(ctx.cfg, ptr_array(ctx.params), ptr(rc), ptr(cat), ptr(ws),
                                            ptr(d_xrec.contiguous().float()), ptr_array(grads), ptr(d_rc), ptr(d_cat), 0,
                                            ptr(scratch), _stream()), "Decoder backward")
        return (d_rc, d_cat, None) + tuple(grads)


class _HeadFn(torch.autograd.Function):
    """Restorer / Classifier: a stack of Linear(+LeakyReLU) layers."""

    @staticmethod
    def forward(ctx, inp, kind, cfg, out_dim, *params):
        lib = get_lib()
        lib.check(lib.iins_validate_config(cfg), f"{kind} config")
        dev = inp.device
        out = torch.empty(inp.shape[0], out_dim, device=dev)
        ws = _empty(getattr(lib, f"iins_{kind}_ws_floats")(cfg), dev)
        lib.check(getattr(lib, f"iins_{kind}_forward")(cfg, ptr_array(params), ptr(inp), ptr(out), ptr(ws), _stream()),
                  f"{kind} forward")
        ctx.kind, ctx.cfg, ctx.params = kind, cfg, params
        ctx.save_for_backward(inp, ws)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = get_lib()
        inp, ws = ctx.saved_tensors
        # parameters the reference never touches (restorer.linear_layer2) get grad None
        n_used = 8
        grads = [torch.zeros_like(p) for p in ctx.params[:n_used]]
        d_in = torch.empty_like(inp)
        scratch = _empty(getattr(lib, f"iins_{ctx.kind}_scratch_floats")(ctx.cfg), inp.device)
        lib.check(getattr(lib, f"iins_{ctx.kind}_backward")(
            ctx.cfg, ptr_array(ctx.params), ptr(inp), ptr(ws), ptr(d_out.contiguous().float()),
            ptr_array(grads + [None] * (len(ctx.params) - n_used)), ptr(d_in), 0, ptr(scratch), _stream()),
            f"{ctx.kind} backward")
        return (d_in, None, None, None) + tuple(grads) + (None,) * (len(ctx.params) - n_used)


def _addr(t):
    return None if t is None else t.data_ptr()


class _ConvHeadFn(torch.autograd.Function):
    """Restorer / Classifier with net_type='Conv1d' (models.py:661-716, :865-902): Conv1d + LeakyReLU + Dropout blocks, BatchNorm1d
    (eps 0.8), Linear.  ``hs`` carries the dropout source (explicit masks or a Philox seed / offset) and the BatchNorm buffers."""

    @staticmethod
    def forward(ctx, inp, kind, cfg, out_dim, hs, *params):
        lib = get_lib()
        lib.check(lib.iins_validate_config(cfg), f"{kind} config")
        dev = inp.device
        out = torch.empty(inp.shape[0], out_dim, device=dev)
        ws = _empty(getattr(lib, f"iins_{kind}_conv_ws_floats")(cfg), dev)
        st = IinsHeadState(int(hs["training"]), _addr(hs.get("mask1")), _addr(hs.get("mask2")), int(hs["seed"]), int(hs["offset"]),
                           _addr(hs["running_mean"]), _addr(hs["running_var"]), _addr(hs["num_batches_tracked"]),
                           _addr(hs["bn_stats"]), 0, 1.0, 0, None)
        lib.check(getattr(lib, f"iins_{kind}_conv_forward")(cfg, ptr_array(params), ptr(inp), ptr(out), ptr(ws), C.byref(st), _stream()),
                  f"{kind} (Conv1d) forward")
        ctx.kind, ctx.cfg, ctx.params, ctx.hs, ctx.st = kind, cfg, params, hs, st
        ctx.save_for_backward(inp, ws)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = get_lib()
        inp, ws = ctx.saved_tensors
        n_used = 8                                                   # restorer.linear_layer2 is never touched (grad None)
        grads = [torch.zeros_like(p) for p in ctx.params[:n_used]]
        d_in = torch.empty_like(inp)
        scratch = _empty(getattr(lib, f"iins_{ctx.kind}_conv_scratch_floats")(ctx.cfg), inp.device)
        lib.check(getattr(lib, f"iins_{ctx.kind}_conv_backward")(
            ctx.cfg, ptr_array(ctx.params), ptr(inp), ptr(ws), ptr(d_out.contiguous().float()),
            ptr_array(grads + [None] * (len(ctx.params) - n_used)), ptr(d_in), 0, ptr(scratch), C.byref(ctx.st), _stream()),
            f"{ctx.kind} (Conv1d) backward")
        return (d_in, None, None, None, None) + tuple(grads) + (None,) * (len(ctx.params) - n_used)


class _SoftRestorerFn(torch.autograd.Function):
    """RestorerLinear with soft=True (models.py:634-655): trunk -> linear_layer2 (mu, logvar) -> the reference's (B, B)-broadcast
    reparameterisation z[i][j] = noise[i] * exp(logvar[j] / 2) + mu[j]."""

    @staticmethod
    def forward(ctx, rc, noise, cfg, *params):
        lib = get_lib()
        lib.check(lib.iins_validate_config(cfg), "restorer config")
        B = rc.shape[0]
        z = torch.empty(B, B, device=rc.device)
        ws = _empty(lib.iins_restorer_soft_ws_floats(cfg), rc.device)
        lib.check(lib.iins_restorer_soft_forward(cfg, ptr_array(params), ptr(rc), ptr(noise), ptr(z), ptr(ws), _stream()), "restorer (soft) forward")
        ctx.cfg, ctx.params = cfg, params
        ctx.save_for_backward(rc, noise, ws)
        return z

    @staticmethod
    def backward(ctx, d_z):
        lib = get_lib()
        rc, noise, ws = ctx.saved_tensors
        grads = [torch.zeros_like(p) for p in ctx.params]
        d_rc = torch.empty_like(rc)
        scratch = _empty(lib.iins_restorer_soft_scratch_floats(ctx.cfg), rc.device)
        lib.check(lib.iins_restorer_soft_backward(ctx.cfg, ptr_array(ctx.params), ptr(rc), ptr(noise), ptr(ws), ptr(d_z.contiguous().float()),
                                                  ptr_array(grads), ptr(d_rc), 0, ptr(scratch), _stream()), "restorer (soft) backward")
        grads[6] = grads[7] = None                                   # linear_layer1 is not on the soft path
        return (d_rc, None, None) + tuple(grads)


def _conv_block(holder, first_index, cin, cout, ks, stride, pad, bn):
    """Parameter holders of one ``conv_block`` of the reference (Conv1d, LeakyReLU, Dropout[, BatchNorm1d(c, 0.8)]) under the
    reference's Sequential indices."""
    holder.put(first_index, nn.Conv1d(cin, cout, ks, stride, pad))
    if bn:
        holder.put(first_index + 3, nn.BatchNorm1d(cout, 0.8))      # second positional argument = eps (models.py:676, 881)


class _ConvHeadMixin:
    """Shared forward plumbing of the two Conv1d heads."""

    def _head_state(self, bn, masks, dev):
        if not hasattr(self, "_bn_stats") or self._bn_stats.device != dev:
            self._bn_stats = torch.zeros(4 * bn.num_features, dtype=torch.float64, device=dev)
        self._offset += 1
        hs = dict(training=self.training, seed=self._seed, offset=self._offset, running_mean=bn.running_mean,
                  running_var=bn.running_var, num_batches_tracked=bn.num_batches_tracked, bn_stats=self._bn_stats)
        if masks is not None:
            hs["mask1"], hs["mask2"] = (m.contiguous().float() for m in masks)
        return hs


# ------------------------------------------------------------------------------------ modules
def _check_conv_type(conv_type, expand):
    """conv_type 1: the Conv1d path; conv_type 2 with expand=True: the Conv2d variant (models.py:179-346, 474-539).  The reference's
    conv_type 2 with expand=False squeezes the decoder output to (B, L, L) and cannot be trained against a (B, L) CIR; conv_type 3
    ("NoExpand") is marked "not available yet" in the reference itself (models.py:45-47)."""
    if conv_type == 1:
        return
    if conv_type == 2 and expand:
        return
    raise NotImplementedError("iins_vae_b200: conv_type=1, or conv_type=2 with expand=True (the reference's other combinations are not runnable)")


class Encoder(nn.Module):
    """models.py:32-64.  ``forward(x:(B,L)) -> (range_code (B,out_dim,8), env_code (B,style_dim,1),
    env_code_rv (B,style_dim/2,1), kl_div ())``.

    ``noise``: "torch" (default) draws ``torch.randn`` exactly where the reference calls
    ``torch.randn_like(mu)`` (models.py:287), so a seeded run consumes the same generator stream;
    "philox" generates it inside the fused reparameterisation kernel (seed, per-call offset)."""

    def __init__(self, conv_type=1, dim=4, n_residual=3, n_downsample=4, style_dim=8, out_dim=2, expand=False,
                 noise="torch", seed=0):
        super().__init__()
        _check_conv_type(conv_type, expand)
        conv = nn.Conv2d if conv_type == 2 else nn.Conv1d
        self.conv_type, self.expand, self.latent_dim = conv_type, expand, style_dim
        self.opts = dict(dim=dim, n_residual=n_residual, n_downsample=n_downsample, env_dim=style_dim, range_dim=out_dim,
                         conv_type=conv_type)
        self.noise_mode, self.seed, self._calls = noise, seed, 0
        # ---- RangeEncoder1d (models.py:140-173) / RangeEncoder2d (:179-215): indices into the reference's nn.Sequential
        self.range_encoder = _Slots()
        m = self.range_encoder.put("model", _Slots())
        idx = 2
        m.put(idx, conv(1, dim, 7))
        idx += 3
        c = dim
        for _ in range(n_downsample):
            m.put(idx, conv(c, 2 * c, 4, stride=2, padding=1))
            c *= 2
            idx += 3
        for _ in range(n_residual):
            m.put(idx, _res_block(c, adain=False, conv=conv))
            idx += 1
        m.put(idx, conv(c, out_dim, 1, 1, 0))
        # ---- EnvEncoder1d(dim*4, n_downsample-2, style_dim) (models.py:258-281) / EnvEncoder2d (:304-329)
        self.env_encoder = _Slots()
        m = self.env_encoder.put("model", _Slots())
        e = 4 * dim
        idx = 2
        m.put(idx, conv(1, e, 7))
        idx += 2
        for _ in range(2):
            m.put(idx, conv(e, 2 * e, 4, stride=2, padding=1))
            e *= 2
            idx += 2
        for _ in range(n_downsample - 2 - 2):
            m.put(idx, conv(e, e, 4, stride=2, padding=1))
            idx += 2
        idx += 1
        m.put(idx, conv(e, style_dim, 1, 1, 0))

    def forward(self, x, noise=None):
        """``noise`` (extension, optional): explicit (B, style_dim/2[, 1]) standard normals to use instead of
        drawing them -- lets a test pin ``env_code_rv`` to a reference run."""
        x = _check_input(x, "x")
        x = x.view(x.size(0), -1)
        opts = dict(self.opts, seed=int(self.seed), offset=int(self._calls) * 64)
        self._calls += 1
        nshape = (x.size(0), self.latent_dim // 2, 1, 1) if self.conv_type == 2 else (x.size(0), self.latent_dim // 2, 1)
        if noise is not None:
            noise = _check_input(noise, "noise").view(nshape)
        elif self.noise_mode == "torch":
            noise = torch.randn(nshape, device=x.device)
        return _EncoderFn.apply(x, noise, opts, *_params_of(self))

    def sample(self, n):
        return torch.randn(n, self.latent_dim)


class Decoder(nn.Module):
    """models.py:67-91.  ``forward(range_code, env_code) -> (B, in_dim)`` (squeezed like the reference)."""

    def __init__(self, conv_type=1, dim=4, n_residual=3, n_upsample=4, style_dim=8, in_dim=152, out_dim=2, expand=False):
        super().__init__()
        _check_conv_type(conv_type, expand)
        conv = nn.Conv2d if conv_type == 2 else nn.Conv1d           # Decoder1d (models.py:405-471) / Decoder2d (:474-539)
        self.conv_type, self.expand = conv_type, expand
        self.opts = dict(dim=dim, n_residual=n_residual, n_downsample=n_upsample, env_dim=style_dim, range_dim=out_dim,
                         in_dim=in_dim, conv_type=conv_type)
        D = dim * 2 ** n_upsample
        self.decoder = _Slots()
        m = self.decoder.put("model", _Slots())
        m.put(0, conv(out_dim, D, 1, 1, 0))
        idx = 2
        for _ in range(n_residual):
            m.put(idx, _res_block(D, adain=True, conv=conv))
            idx += 1
        c = D
        for _ in range(n_upsample):
            m.put(idx + 1, conv(c, c // 2, 5, stride=1, padding=2))
            m.put(idx + 2, LayerNorm(c // 2))
            c //= 2
            idx += 4
        m.put(idx + 1, conv(c, 1, 7))
        mlp = self.decoder.put("mlp", _Slots())
        mm = mlp.put("model", _Slots())
        n_adain = 2 * n_residual * 2 * D
        mm.put(0, nn.Linear(style_dim, 256))
        mm.put(2, nn.Linear(256, 256))
        mm.put(4, nn.Linear(256, n_adain))

    def forward(self, range_code, env_code):
        rc = _check_input(range_code, "range_code")
        cat = _check_input(env_code, "env_code").view(env_code.size(0), -1)
        x_recon = _DecoderFn.apply(rc, cat, self.opts, *_params_of(self))
        return x_recon.squeeze()                      # models.py:90


class Restorer(_ConvHeadMixin, nn.Module):
    """models.py:94-112.  net_type='Linear' -> RestorerLinear (:615-658); net_type='Conv1d' -> RestorerConv1d (:661-716):
    two Conv1d(k4,s2,p1) + LeakyReLU(0.2) + Dropout(0.25) blocks, BatchNorm1d(32, eps=0.8) behind the second, Linear(64, 1).
    ``forward(range_code, masks=None)``: ``masks`` = (m1 (B,16,4), m2 (B,32,2)) explicit dropout keep-masks (tests replay the
    reference's); by default Philox4x32-10 decides inside the kernel (``seed``, one offset per call)."""

    def __init__(self, code_shape, soft=False, filters=64, conv_type=1, expand=False, net_type="Linear", seed=0):
        super().__init__()
        if net_type not in ("Linear", "Conv1d"):
            raise NotImplementedError("iins_vae_b200: net_type 'Linear' and 'Conv1d' are on the B200 path (Conv2d: SURVEY 8f row 3)")
        if soft and net_type != "Linear":
            raise NotImplementedError("iins_vae_b200: soft=True is on the path for the Linear restorer only")
        self.soft, self.net_type = soft, net_type
        self.code_shape = tuple(int(v) for v in code_shape)
        n_in = int(np.prod(code_shape))
        self.restorer = _Slots()
        self._seed, self._offset = int(seed), 0
        if net_type == "Linear":
            layers = self.restorer.put("layers", _Slots())
            layers.put(0, nn.Linear(n_in, 512))
            layers.put(2, nn.Linear(512, 256))
            layers.put(4, nn.Linear(256, 256))
            self.restorer.put("linear_layer1", nn.Linear(256, 1))
            self.restorer.put("linear_layer2", nn.Linear(256, 2))       # present, unused (models.py:632)
        else:
            blocks = self.restorer.put("conv_blocks", _Slots())
            _conv_block(blocks, 0, self.code_shape[0], 16, 4, 2, 1, bn=False)
            _conv_block(blocks, 3, 16, 32, 4, 2, 1, bn=True)
            self.restorer.put("linear_layer1", nn.Linear(64, 1))
            l2 = self.restorer.put("linear_layer2", _Slots())           # nn.Sequential(nn.Linear(64, 2)): present, unused
            l2.put(0, nn.Linear(64, 2))

    def forward(self, range_code, masks=None):
        rc = _check_input(range_code, "range_code")
        if tuple(rc.shape[1:]) != self.code_shape or any(v != 8 for v in self.code_shape[1:]):
            raise RuntimeError(f"Restorer expects range_code (B,{self.code_shape}) with code length 8")
        two_d = len(self.code_shape) == 3                       # (R, 8, 8): the 2-D variant's code, flattened like the reference's view(B, -1)
        if two_d and (self.net_type != "Linear" or self.soft):
            raise NotImplementedError("iins_vae_b200: the 2-D range code goes through the Linear restorer (soft=False)")
        cfg = _cfg(rc.shape[0], range_dim=self.code_shape[0], conv_type=2 if two_d else 1)
        if self.net_type == "Linear" and self.soft:
            # the reference draws np.random.normal(0, 1, (B, 1)) on the host (models.py:637): same call, same generator stream
            noise = torch.from_numpy(np.random.normal(0, 1, (rc.shape[0], 1)).astype(np.float32)).to(rc.device).view(-1)
            return _SoftRestorerFn.apply(rc, noise, cfg, *_params_of(self))
        if self.net_type == "Linear":
            return _HeadFn.apply(rc, "restorer", cfg, 1, *_params_of(self))
        hs = self._head_state(getattr(self.restorer.conv_blocks, "6"), masks, rc.device)
        return _ConvHeadFn.apply(rc, "restorer", cfg, 1, hs, *_params_of(self))


class Classifier(_ConvHeadMixin, nn.Module):
    """models.py:115-132.  net_type='Linear' -> ClassifierLinear (:838-862); net_type='Conv1d' -> ClassifierConv1d (:865-902):
    two Conv1d(k1) + LeakyReLU(0.2) + Dropout(0.25) blocks on (B, env_dim, 1), BatchNorm1d(filters, eps=0.8) behind the second,
    Linear(filters, num_classes) + LeakyReLU(0.2)."""

    def __init__(self, env_dim, num_classes, filters=16, net_type="Linear", seed=0):
        super().__init__()
        if net_type not in ("Linear", "Conv1d"):
            raise NotImplementedError("iins_vae_b200: net_type 'Linear' and 'Conv1d' are on the B200 path (Conv2d: SURVEY 8f row 3)")
        self.env_dim, self.num_classes, self.filters, self.net_type = env_dim, num_classes, filters, net_type
        self.classifier = _Slots()
        self._seed, self._offset = int(seed) + 1, 0
        if net_type == "Linear":
            layers = self.classifier.put("layers", _Slots())
            layers.put(0, nn.Linear(env_dim, filters))
            layers.put(2, nn.Linear(filters, filters * 2))
            layers.put(4, nn.Linear(filters * 2, filters))
            layers.put(6, nn.Linear(filters, num_classes))
        else:
            blocks = self.classifier.put("conv_blocks", _Slots())
            _conv_block(blocks, 0, env_dim, filters, 1, 1, 0, bn=False)
            _conv_block(blocks, 3, filters, filters, 1, 1, 0, bn=True)
            lin = self.classifier.put("linear", _Slots())
            lin.put(0, nn.Linear(filters, num_classes))

    def forward(self, env_code, masks=None):
        cat = _check_input(env_code, "env_code").view(env_code.size(0), -1)
        cfg = _cfg(cat.shape[0], env_dim=self.env_dim, num_classes=self.num_classes, filters=self.filters)
        if self.net_type == "Linear":
            return _HeadFn.apply(cat, "classifier", cfg, self.num_classes, *_params_of(self))
        hs = self._head_state(getattr(self.classifier.conv_blocks, "6"), masks, cat.device)
        return _ConvHeadFn.apply(cat, "classifier", cfg, self.num_classes, hs, *_params_of(self))


class EMNet(nn.Module):
    """The composite ``network(cir) -> (label_est, env_latent, err_est)`` that train.py:82 / test.py:73 call.
    The class is missing from the reference (run.py:59-62 is its only trace); following SURVEY.md 8(b) it is
    Encoder -> (Classifier(env_code), env_code, Restorer(range_code)).  ``filters`` maps to Encoder ``dim``
    only through ``dim`` (default 4); enet_type / mnet_type select the Linear heads (1)."""

    def __init__(self, cir_len=157, num_classes=2, env_dim=16, filters=16, enet_type=1, mnet_type=1, dim=4,
                 n_residual=3, n_downsample=4, range_dim=2):
        super().__init__()
        if enet_type not in (1, 2) or mnet_type not in (1, 2):
            raise NotImplementedError("iins_vae_b200: identifier / regressor types 1 (Linear) and 2 (Conv1d); 3 (Conv2d) is SURVEY 8f row 3")
        self.cir_len = cir_len
        self.encoder = Encoder(1, dim, n_residual, n_downsample, env_dim, range_dim)
        kinds = {1: "Linear", 2: "Conv1d"}                  # utils.py:43-44: 1 for linear, 2 for conv1d
        self.classifier = Classifier(env_dim, num_classes, filters=16, net_type=kinds[enet_type])
        self.restorer = Restorer((range_dim, 128 // 2 ** n_downsample), net_type=kinds[mnet_type])

    def forward(self, cir):
        range_code, env_code, _, _ = self.encoder(cir)
        return self.classifier(env_code), env_code, self.restorer(range_code)
