"""Which (parameter seed s, data seed s + 1) pairs give a kink-free B = 32 smoke step (every gradient tensor strict against the
fp64 oracle in tensor-core mode)?  Run under several library settings: a seed that is clean in all of them has no ReLU /
LeakyReLU / sign() input within rounding of zero.   python tools/smoke_seed_scan.py 0:16"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import iins_oracle as orc
from tests import parity
from tests.test_gpu_parity import _mods
from iins_vae_b200.engine import SemiTrainEngine

lo, hi = (int(v) for v in sys.argv[1].split(":"))
cfg = orc.PathConfig()
B = 32
for seed in range(lo, hi):
    mods, pdicts = _mods(cfg, seed)
    cir, err, label = orc.synthetic_batch(cfg, B, seed + 1)
    zero = torch.zeros(B, cfg.env_dim // 2, 1)
    _, ref32 = orc.semi_step_with_grads(*pdicts, cir, err, label, cfg, True, zero)
    dbl = lambda d: {k: v.double() for k, v in d.items()}
    _, truth = orc.semi_step_with_grads(*(dbl(p) for p in pdicts), cir.double(), err.double(), label.double(), cfg, True, zero.double())
    gscale = max(float(g.abs().max()) for g in ref32.values() if g is not None)
    eng = SemiTrainEngine(*mods, batch_size=B, cir_len=cfg.cir_len, use_graph=False)
    eng.step(cir, err, label, supervised=True, update=False)
    torch.cuda.synchronize()
    rows = parity.grad_report(eng.named_grads(), truth, ref32, gscale, parity.REF_FACTOR_TC)
    bad = [r for r in rows if not r[3]]
    worst = max(r[1] for r in rows if not orc.grad_is_structurally_zero(r[0]))
    print(f"seed {seed}: {len(bad)} tensors beyond strict, worst rel-L2 {worst:.2e}", flush=True)
