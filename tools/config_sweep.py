"""Robustness sweep: one train step through the fused engine for unusual shape options, checked against the CPU oracle
(loss terms rtol 1e-4, every gradient tensor finite and within the kink band; tensors whose gradient is ~0 relative to
the largest one -- e.g. a conv bias cancelled by the following norm -- are checked absolutely).  dim=3 must be REFUSED
(non-power-of-two channel counts), not mis-computed.  python tools/config_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import iins_oracle as orc
from iins_vae_b200 import models as M
from iins_vae_b200.engine import SemiTrainEngine

CASES = [dict(dim=2), dict(dim=1), dict(dim=3), dict(dim=8), dict(dim=16), dict(dim=16, n_residual=1, env_dim=8), dict(n_residual=1), dict(n_residual=0), dict(num_classes=2), dict(num_classes=10),
         dict(env_dim=8), dict(range_dim=4), dict(dim=2, n_residual=2, env_dim=32, num_classes=3)]


def run_sweep(cases=CASES, batches=(3, 130)):
  bad = 0
  for kw in cases:
    for B in batches:
          cfg = orc.PathConfig(**kw)
          try:
              pe, pd, pr, pc = orc.init_all(cfg, 5)
              Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
              Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
              Res = M.Restorer((cfg.range_dim, cfg.code_len)); Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
              for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
                  m.load_state_dict(p); m.cuda()
              cir, err, label = orc.synthetic_batch(cfg, B, 77)
              eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, use_graph=False)
              eng.step(cir, err, label, supervised=True, update=False)
              torch.cuda.synchronize()
              got = eng.loss_terms()
              ref, grads = orc.semi_step_with_grads(pe, pd, pr, pc, cir, err, label, cfg, True, torch.zeros(B, cfg.env_dim // 2, 1))
              rel_loss = abs(got["loss"] - float(ref["loss"])) / abs(float(ref["loss"]))
              named = eng.named_grads()
              worst = 0.0
              gmax = max(float(r.norm()) for r in grads.values() if r is not None)
              for name, r in grads.items():
                  if r is None or orc.grad_is_structurally_zero(name):
                      continue
                  g = named[name].cpu()
                  assert torch.isfinite(g).all(), name
                  if float(r.norm()) < 1e-6 * gmax:                 # effectively zero gradient: absolute check
                      assert float((g - r).norm()) < 1e-6 * gmax, name
                      continue
                  worst = max(worst, float((g - r).norm() / (r.norm() + 1e-30)))
              ok = rel_loss < 1e-4 and worst < 8.0 / B
              bad += not ok
              print(f"{'ok ' if ok else 'BAD'} {kw} B={B}: loss rel err {rel_loss:.1e}, worst gradient rel-L2 {worst:.1e}")
          except Exception as e:
              refused = kw.get("dim") == 3 and "power of two" in str(e)
              bad += not refused
              print(f"{'ok  (refused)' if refused else 'ERR'} {kw} B={B}: {type(e).__name__}: {e}")
  print("failures:", bad)
  return bad


if __name__ == "__main__":
    sys.exit(1 if run_sweep() else 0)
