"""Which gradient tensors of a CUDA-graph replay differ from the eager pass of the same engine (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import iins_oracle as orc
from tests.test_gpu_parity import _mods
from iins_vae_b200.engine import SemiTrainEngine

cfg = orc.PathConfig()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for supervised in (True, False):
    mods, _ = _mods(cfg, 23)
    cir, err, label = orc.synthetic_batch(cfg, batch, 623)
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=False)
    eng.step(cir, err, label, supervised=supervised, update=False)
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in eng.named_grads().items()}
    eng.step(cir, err, label, supervised=supervised, update=False)
    torch.cuda.synchronize()
    eager2 = {k: v.clone() for k, v in eng.named_grads().items()}
    eng.use_graph = True
    for rep in range(3):
        eng.step(cir, err, label, supervised=supervised, update=False)
        torch.cuda.synchronize()
        for k, e in eager.items():
            n = float(e.norm())
            if n == 0:
                continue
            d = float((eng.named_grads()[k] - e).norm()) / n
            d2 = float((eager2[k] - e).norm()) / n
            if d > 1e-5 or d2 > 1e-5:
                print(f"sup={supervised} replay {rep}: {k:60s} replay-vs-eager {d:.3e}  eager-vs-eager {d2:.3e}")
print("REPLAY_DIAG done")
