"""Per-launch timing of one train step (CUDA events around every kernel launch, in-process)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iins_vae_b200
from oracle import iins_oracle as orc
from iins_vae_b200 import models as M
from iins_vae_b200.engine import SemiTrainEngine
from iins_vae_b200._capi import get_lib

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
DIM = int(sys.argv[3]) if len(sys.argv) > 3 else 4
SUP = (sys.argv[4] != "unsup") if len(sys.argv) > 4 else True
iins_vae_b200.set_compute_mode(mode)
cfg = orc.PathConfig(dim=DIM)
pe, pd, pr, pc = orc.init_all(cfg, 0)
Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
Res = M.Restorer((cfg.range_dim, cfg.code_len)); Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
    m.load_state_dict(p); m.cuda()
cir, err, label = orc.synthetic_batch(cfg, B, 1)
eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, use_graph=False)
eng.set_concurrency(False)          # serial launches: one event pair = one kernel
for _ in range(3):
    eng.step(cir, err, label, supervised=SUP)
torch.cuda.synchronize()
lib = get_lib()
prof = lib.profile(lambda: eng.step(cir, err, label, supervised=SUP))
sh = lib.last_shapes
tot = sum(p[1] for p in prof)
print(f"mode {mode} B {B}: {len(prof)} launches, sum of kernel times {tot:.3f} ms")
for i, (name, ms, fl) in enumerate(prof):
    m, n, k = sh[3 * i], sh[3 * i + 1], sh[3 * i + 2]
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 and fl > 0 else 0
    print(f"{i:4d} {name:28s} M={m:8d} N={n:4d} K={k:4d} {ms * 1e3:9.1f} us {tf:8.2f} TFLOP/s")
