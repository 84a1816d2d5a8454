import sys; sys.path.insert(0,'/root/repo')
import torch
from oracle import iins_oracle as orc
from iins_vae_b200 import models as M
from iins_vae_b200.engine import SemiTrainEngine
import iins_vae_b200
cfg = orc.PathConfig(); B = 32
for mode in ("fp32", "simt"):
    iins_vae_b200.set_compute_mode(mode)
    pe, pd, pr, pc = orc.init_all(cfg, 0)
    Enc = M.Encoder(1, 4, 3, 4, 16, 2); Dec = M.Decoder(1, 4, 3, 4, 16, 157, 2); Res = M.Restorer((2, 8)); Cls = M.Classifier(16, 5)
    for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)): m.load_state_dict(p); m.cuda()
    cir, err, label = orc.synthetic_batch(cfg, B, 1)
    eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, use_graph=False)
    eng.step(cir, err, label, supervised=True, update=False); torch.cuda.synchronize()
    d = lambda x: {k: v.double() for k, v in x.items()}
    z = torch.zeros(B, 8, 1)
    _, g32 = orc.semi_step_with_grads(pe, pd, pr, pc, cir, err, label, cfg, True, z)
    _, g64 = orc.semi_step_with_grads(d(pe), d(pd), d(pr), d(pc), cir.double(), err.double(), label.double(), cfg, True, z.double())
    got = eng.named_grads()
    print("mode", mode)
    for k, t in g64.items():
        if t is None or orc.grad_is_structurally_zero(k): continue
        n = float(t.norm()) + 1e-30
        e = float((got[k].cpu().double() - t).norm()) / n; e2 = float((g32[k].double() - t).norm()) / n
        if e > 1e-4: print(f"   {k:46s} gpu {e:.2e} cpu32 {e2:.2e}")
