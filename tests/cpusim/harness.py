"""Drive the C ABI of the CPU logic simulator build (tests/cpusim/build.sh) with host tensors.

TEST INFRASTRUCTURE ONLY: lets the SIMT kernels' indexing / barrier / reduction logic and the host-side
launch plans be checked against the oracle in the authoring container (no GPU).  The product package
never loads this library.
"""
import os
import subprocess

import torch

from iins_vae_b200._capi import IinsConfig, IinsLib, ptr, ptr_array
from oracle import iins_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_PATH = os.path.join(HERE, "_build", "libiins_cpusim.so")


def build_sim() -> IinsLib:
    srcs = [os.path.join(HERE, "cuda_sim.h")] + [
        os.path.join(HERE, "..", "..", "iins_vae_b200", "csrc", f)
        for f in os.listdir(os.path.join(HERE, "..", "..", "iins_vae_b200", "csrc"))]
    if not os.path.exists(SIM_PATH) or any(os.path.getmtime(s) > os.path.getmtime(SIM_PATH) for s in srcs):
        subprocess.run(["bash", os.path.join(HERE, "build.sh")], check=True, capture_output=True)
    return IinsLib(SIM_PATH)


def make_cfg(cfg: orc.PathConfig, batch: int) -> IinsConfig:
    return IinsConfig(batch, cfg.cir_len, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim,
                      cfg.num_classes, 16)


def plist(pdict):
    """Parameter tensors in named_parameters() order (buffers dropped)."""
    return [v.contiguous() for k, v in pdict.items() if not orc.is_buffer(k)]


def pnames(pdict):
    return [k for k in pdict if not orc.is_buffer(k)]


class SimStep:
    """One semi-supervised step (forward, loss, backward) through the C ABI on host tensors."""

    def __init__(self, lib: IinsLib, cfg: orc.PathConfig, batch: int, pe, pd, pr, pc):
        self.lib, self.cfg, self.B = lib, cfg, batch
        self.c = make_cfg(cfg, batch)
        lib.check(lib.iins_validate_config(self.c), "validate")
        self.P = dict(enc=plist(pe), dec=plist(pd), res=plist(pr), cls=plist(pc))
        self.names = dict(enc=pnames(pe), dec=pnames(pd), res=pnames(pr), cls=pnames(pc))
        z = lambda n: torch.zeros(int(n) + 16)
        self.ws = {m: z(getattr(lib, f"iins_{m}_ws_floats")(self.c)) for m in ("encoder", "decoder", "restorer", "classifier")}
        self.scratch = {m: z(getattr(lib, f"iins_{m}_scratch_floats")(self.c)) for m in ("encoder", "decoder", "restorer", "classifier")}

    def forward(self, cir, noise):
        lib, c, B, cfg = self.lib, self.c, self.B, self.cfg
        self.cir, self.noise = cir.contiguous(), (None if noise is None else noise.contiguous())
        self.rc = torch.zeros(B, cfg.range_dim, cfg.code_len)
        self.cat = torch.zeros(B, cfg.env_dim)
        self.lat = torch.zeros(B, cfg.env_dim // 2)
        self.kl = torch.zeros(1)
        self.xrec = torch.zeros(B, cfg.cir_len)
        self.err_est = torch.zeros(B, 1)
        self.logits = torch.zeros(B, cfg.num_classes)
        lib.check(lib.iins_encoder_forward(c, ptr_array(self.P["enc"]), ptr(self.cir), ptr(self.noise), 1234, 0,
                                           ptr(self.rc), ptr(self.cat), ptr(self.lat), ptr(self.kl),
                                           ptr(self.ws["encoder"]), None), "enc fwd")
        lib.check(lib.iins_decoder_forward(c, ptr_array(self.P["dec"]), ptr(self.rc), ptr(self.cat), ptr(self.xrec),
                                           ptr(self.ws["decoder"]), None), "dec fwd")
        lib.check(lib.iins_restorer_forward(c, ptr_array(self.P["res"]), ptr(self.rc), ptr(self.err_est),
                                            ptr(self.ws["restorer"]), None), "res fwd")
        lib.check(lib.iins_classifier_forward(c, ptr_array(self.P["cls"]), ptr(self.cat), ptr(self.logits),
                                              ptr(self.ws["classifier"]), None), "cls fwd")

    def loss_backward(self, err, label, supervised: bool):
        lib, c, B, cfg = self.lib, self.c, self.B, self.cfg
        self.G = {m: [torch.zeros_like(p) for p in ps] for m, ps in self.P.items()}
        self.out = torch.zeros(8)
        d_xrec = torch.zeros(B, cfg.cir_len)
        d_err = torch.zeros(B, 1)
        d_logits = torch.zeros(B, cfg.num_classes)
        err = err.contiguous()
        label = label.contiguous()
        lib.check(lib.iins_loss_forward_backward(
            B, cfg.cir_len, cfg.num_classes, ptr(self.cir), ptr(self.xrec),
            ptr(err) if supervised else None, ptr(self.err_est) if supervised else None,
            ptr(self.logits) if supervised else None, ptr(label) if supervised else None, None, getattr(self, "label_offset", 0),
            orc.LAMBDA_AE, orc.LAMBDA_RES, orc.LAMBDA_ENV, ptr(self.out), ptr(d_xrec),
            ptr(d_err) if supervised else None, ptr(d_logits) if supervised else None, None, None), "loss")
        d_rc = torch.zeros_like(self.rc)
        d_cat = torch.zeros_like(self.cat)
        lib.check(lib.iins_decoder_backward(c, ptr_array(self.P["dec"]), ptr(self.rc), ptr(self.cat), ptr(self.ws["decoder"]),
                                            ptr(d_xrec), ptr_array(self.G["dec"]), ptr(d_rc), ptr(d_cat), 0,
                                            ptr(self.scratch["decoder"]), None), "dec bwd")
        if supervised:
            lib.check(lib.iins_restorer_backward(c, ptr_array(self.P["res"]), ptr(self.rc), ptr(self.ws["restorer"]), ptr(d_err),
                                                 ptr_array(self.G["res"]), ptr(d_rc), 1, ptr(self.scratch["restorer"]), None), "res bwd")
            lib.check(lib.iins_classifier_backward(c, ptr_array(self.P["cls"]), ptr(self.cat), ptr(self.ws["classifier"]),
                                                   ptr(d_logits), ptr_array(self.G["cls"]), ptr(d_cat), 1,
                                                   ptr(self.scratch["classifier"]), None), "cls bwd")
        d_kl = torch.full((1,), orc.LAMBDA_RANGE)
        lib.check(lib.iins_encoder_backward(c, ptr_array(self.P["enc"]), ptr(self.noise), 1234, 0, ptr(self.rc), ptr(self.cat),
                                            ptr(self.ws["encoder"]), ptr(d_rc), ptr(d_cat), None, ptr(d_kl),
                                            ptr_array(self.G["enc"]), ptr(self.scratch["encoder"]), None), "enc bwd")
        self.d_rc, self.d_cat = d_rc, d_cat

    def grads(self, supervised: bool):
        out = {}
        for m in ("enc", "dec", "res", "cls"):
            for n, g in zip(self.names[m], self.G[m]):
                none = (not supervised and m in ("res", "cls")) or "linear_layer2" in n
                out[f"{m}.{n}"] = None if none else g
        return out
