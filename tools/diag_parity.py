"""Diagnostic: per-tensor error of the GPU step vs an fp64 oracle run, next to the fp32 oracle's own error."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import iins_oracle as orc
from iins_vae_b200 import models as M
from iins_vae_b200.engine import SemiTrainEngine

def run(B, seed, sup=True):
    cfg = orc.PathConfig()
    pe, pd, pr, pc = orc.init_all(cfg, seed)
    Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
    Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
    Res = M.Restorer((cfg.range_dim, cfg.code_len)); Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
    for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
        m.load_state_dict(p); m.cuda()
    cir, err, label = orc.synthetic_batch(cfg, B, 500 + B)
    eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, use_graph=False)
    eng.step(cir, err, label, supervised=sup, update=False)
    torch.cuda.synchronize()
    z = torch.zeros(B, 8, 1)
    o32, g32 = orc.semi_step_with_grads(pe, pd, pr, pc, cir, err, label, cfg, sup, z)
    d = lambda x: {k: v.double() for k, v in x.items()}
    o64, g64 = orc.semi_step_with_grads(d(pe), d(pd), d(pr), d(pc), cir.double(), err.double(), label.double(), cfg, sup, z.double())
    print(f"==== B={B} seed={seed} sup={sup}")
    for k, t in (("range_code", eng.rc), ("env_code", eng.cat), ("cir_gen", eng.xrec), ("err_fake", eng.err_est), ("label_fake", eng.logits)):
        ref = o64[k].reshape(t.shape)
        e_gpu = float((t.cpu().double() - ref).abs().max()); e_cpu = float((o32[k].reshape(t.shape).double() - ref).abs().max())
        print(f"  fwd {k:12s} max|err| gpu {e_gpu:.2e} cpu32 {e_cpu:.2e} (scale {float(ref.abs().max()):.2e})")
    got = eng.named_grads()
    for k, t in g64.items():
        if t is None or orc.grad_is_structurally_zero(k): continue
        n = float(t.norm()) + 1e-30
        e_gpu = float((got[k].cpu().double() - t).norm()) / n
        e_cpu = float((g32[k].double() - t).norm()) / n
        flag = " <<<" if e_gpu > 1e-4 else ""
        print(f"  grad {k:46s} gpu {e_gpu:.2e} cpu32 {e_cpu:.2e} |g| {n:.2e}{flag}")

if __name__ == "__main__":
    from iins_vae_b200._capi import get_lib
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cases = [(int(b), 100 + int(b)) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [(2, 0), (64, 75), (4096, 4107)]
    get_lib().check(get_lib().iins_set_compute_mode(mode), "mode")
    print("compute mode", mode)
    sup = not (len(sys.argv) > 3 and sys.argv[3] == "unsup")
    if len(sys.argv) > 4:                       # explicit seed (e.g. the pytest case: seed = 11 + batch)
        cases = [(b, int(sys.argv[4])) for b, _ in cases]
    for B, seed in cases:
        run(B, seed, sup)
