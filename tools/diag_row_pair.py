"""Diagnostics: which class of paired-row launches moves which gradient tensor (two contexts per mask, fp32-grade mode)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iins_vae_b200
from iins_vae_b200._capi import get_lib
from iins_vae_b200.engine import SemiTrainEngine
from oracle import iins_oracle as orc
from tests.test_gpu_parity import _mods

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = get_lib().dll
cfg = orc.PathConfig()
cir, err, label = orc.synthetic_batch(cfg, B, 977)


def run(pair, mask, eps=0.0):
    os.environ["IINS_ROW_PAIR"] = str(pair)
    os.environ["IINS_ROW_PAIR_MASK"] = str(mask)
    ctx = d.iins_ctx_create()
    d.iins_ctx_make_current(ctx)
    iins_vae_b200.set_compute_mode("fp32")
    mods, _ = _mods(cfg, 41)
    eng = SemiTrainEngine(*mods, batch_size=B, cir_len=cfg.cir_len, use_graph=False)
    x = cir if eps == 0.0 else cir * (1.0 + eps * torch.randn(cir.shape, generator=torch.Generator().manual_seed(1)))
    eng.step(x, err, label, supervised=True, update=False)
    torch.cuda.synchronize()
    out = {k: v.clone() for k, v in eng.named_grads().items()}
    out["_rc"], out["_xrec"] = eng.rc.clone(), eng.xrec.clone()
    d.iins_ctx_make_current(None)
    d.iins_ctx_destroy(ctx)
    return out


base = run(0, 31)
again = run(0, 31)
for name, other in [("run-to-run", again), ("input * (1 + 1e-7 N(0,1)), unpaired", run(0, 31, 1e-7)), ("input * (1 + 1e-6 N(0,1)), unpaired", run(0, 31, 1e-6))] + [(f"mask {m}", run(1, m)) for m in (1, 2, 4, 8, 16)]:
    rows = []
    for k, e in base.items():
        if not k.startswith("_") and orc.grad_is_structurally_zero(k):
            continue
        n = float(e.norm())
        if n > 0:
            rows.append((float((other[k] - e).norm()) / n, k))
    rows.sort(reverse=True)
    print(name, " | ".join(f"{k} {v:.1e}" for v, k in rows[:6]))
