"""Golden fixture for the 2-D variant (conv_type = 2, expand = True) from the LIVE reference, run in the
authoring container:

    python tests/golden/make_golden2d.py

The reference's own ``Encoder(conv_type=2, expand=True)`` / ``Decoder(conv_type=2, expand=True)`` / ``Restorer((2, 8, 8))``
(models.py:33-112 -> RangeEncoder2d :179-215, EnvEncoder2d :304-346, Decoder2d :474-539) are loaded with the oracle's seeded
parameters (``oracle/iins_oracle2d.init_all`` -- the tests regenerate them from the seed, only a checksum is stored), run on a
seeded batch, and   L1(x, x_recon) + KL + 10 * L1(err, err_est)   is back-propagated.  Stored: inputs, the latent noise torch drew
inside the reference (recovered from its outputs), every output, every parameter gradient.  The script also asserts that the
restatement in ``oracle/iins_oracle2d.py`` reproduces these numbers (outputs 1e-5, gradients 2e-4 rel-L2)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import models as ref                        # noqa: E402
from oracle import iins_oracle as orc        # noqa: E402
from oracle import iins_oracle2d as orc2      # noqa: E402
from tests.golden.make_golden_common import BIG, sample_positions_2d   # noqa: E402


def digest(store, key, t):
    """Small tensors in full; large ones as their L2 norm + 1024 sampled entries (positions: sample_positions_2d)."""
    a = t.detach().numpy().ravel()
    if a.size <= BIG:
        store[key + "|full"] = a.astype(np.float32)
    else:
        store[key + "|norm"] = np.float64(np.linalg.norm(a.astype(np.float64)))
        store[key + "|samples"] = a[sample_positions_2d(a.size)].astype(np.float32)

CASES = [(0, 2), (1, 3)]                     # (seed, batch)


def run_case(store, seed, batch):
    cfg = orc.PathConfig()
    pe, pd, pr = orc2.init_all(cfg, seed)
    E = ref.Encoder(conv_type=2, dim=cfg.dim, n_residual=cfg.n_residual, n_downsample=cfg.n_downsample, style_dim=cfg.env_dim,
                    out_dim=cfg.range_dim, expand=True)
    D = ref.Decoder(conv_type=2, dim=cfg.dim, n_residual=cfg.n_residual, n_upsample=cfg.n_downsample, style_dim=cfg.env_dim,
                    in_dim=cfg.cir_len, out_dim=cfg.range_dim, expand=True)
    R = ref.Restorer((cfg.range_dim, cfg.code_len, cfg.code_len), net_type="Linear")
    for m, p in ((E, pe), (D, pd), (R, pr)):
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in p.items()]
        m.load_state_dict(p)
        m.train()
    cir, err, _ = orc.synthetic_batch(cfg, batch, seed + 100)
    torch.manual_seed(seed + 7)
    rc, cat, lat, kl = E(cir)
    xrec = D(rc, cat)
    err_est = R(rc)
    loss = torch.nn.L1Loss()(xrec, cir) + kl + 10.0 * torch.nn.L1Loss()(err_est, err)
    loss.backward()
    half = cat.shape[1] // 2
    noise = ((lat - cat[:, :half]) / cat[:, half:].exp()).detach()          # what torch.randn_like(mu) returned (models.py:335)
    pre = f"c2d.s{seed}.b{batch}."
    store[pre + "meta"] = np.array([seed, batch], dtype=np.int64)
    store[pre + "cir"] = cir.numpy(); store[pre + "err"] = err.numpy(); store[pre + "noise"] = noise.numpy()
    for k, v in (("rc", rc), ("cat", cat), ("latent", lat), ("kl", kl), ("xrec", xrec), ("err_est", err_est), ("loss", loss)):
        store[pre + "out." + k] = v.detach().numpy()
    grads = {}
    for g, m in (("enc", E), ("dec", D), ("res", R)):
        for k, v in m.named_parameters():
            if v.grad is not None:
                grads[f"{g}.{k}"] = v.grad.detach().clone()
                digest(store, pre + f"g.{g}.{k}", v.grad)
    store[pre + "param_checksum"] = np.array([float(sum(v.double().abs().sum() for v in p.values())) for p in (pe, pd, pr)])
    # ---- the restatement must reproduce the reference
    tp = [{k: v.clone().requires_grad_(not orc.is_buffer(k)) for k, v in p.items()} for p in (pe, pd, pr)]
    l2, outs = orc2.step_loss(tp[0], tp[1], tp[2], cir, err, cfg, noise)
    l2.backward()
    assert abs(float(l2) - float(loss)) < 1e-5 * max(1.0, abs(float(loss))), (float(l2), float(loss))
    for k, v in (("rc", rc), ("cat", cat), ("latent", lat), ("xrec", xrec), ("err_est", err_est)):
        assert torch.allclose(outs[k], v, rtol=1e-4, atol=1e-5), k
    worst = 0.0
    for g, p in zip(("enc", "dec", "res"), tp):
        for k, v in p.items():
            key = f"{g}.{k}"
            if key not in grads:
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, key
                continue
            n = float(grads[key].norm())
            if orc.grad_is_structurally_zero(k) or n == 0.0:
                continue
            worst = max(worst, float((v.grad - grads[key]).norm()) / n)
    assert worst < 2e-4, worst
    print(f"case seed {seed} batch {batch}: loss {float(loss):.6f}, oracle vs reference worst gradient rel-L2 {worst:.2e}")


if __name__ == "__main__":
    store = {}
    for seed, batch in CASES:
        run_case(store, seed, batch)
    out = os.path.join(HERE, "iins_golden2d.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")
