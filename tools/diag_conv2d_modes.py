import sys; sys.path.insert(0,'/root/repo')
import torch, iins_vae_b200
from oracle import iins_oracle as orc, iins_oracle2d as orc2
from tests.test_gpu_conv2d import _mods2d, _step
cfg = orc.PathConfig(); batch, seed = 4, 11
cir, err, _ = orc.synthetic_batch(cfg, batch, seed + 300)
noise = torch.randn(batch, cfg.env_dim // 2, 1, 1, generator=torch.Generator().manual_seed(seed + 5))
_, pdicts = _mods2d(cfg, seed)
tp = [{k: v.clone().double() for k, v in p.items()} for p in pdicts]
l64, o64 = orc2.step_loss(tp[0], tp[1], tp[2], cir.double(), err.double(), cfg, noise.double())
for mode in ("fp32", "bf16", "simt"):
    iins_vae_b200.set_compute_mode(mode)
    mods, _ = _mods2d(cfg, seed)
    loss, outs = _step(mods, cir.cuda(), err.cuda(), noise.cuda())
    print(mode, "loss", float(loss), float(l64))
    for k in ("rc", "cat", "xrec", "err_est"):
        ref = o64[k].float(); d = (outs[k].detach().cpu() - ref)
        print(f"   {k:8s} scale {float(ref.abs().max()):.3e} max err {float(d.abs().max()):.3e} mean err {float(d.abs().mean()):.3e} mean |ref| {float(ref.abs().mean()):.3e}")
