"""Replay time of the step's phases under the real stream concurrency (CUDA graphs): forward only, forward + loss + backward,
the full step -- for the supervised and the unsupervised branch.  python tools/phase_times.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import iins_oracle as orc
from tests.test_gpu_parity import _mods
from iins_vae_b200.engine import SemiTrainEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = orc.PathConfig()
mods, _ = _mods(cfg, 3)
cir, err, label = orc.synthetic_batch(cfg, B, 5)
eng = SemiTrainEngine(*mods, batch_size=B, use_graph=False)
eng.load_batch(cir.cuda(), err.cuda(), label.cuda())


def timed(fn, n=30):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for sup in (True, False):
    def fwd():
        eng._forward(sup)

    def fwd_loss_bwd():
        eng.flat.grad.zero_()
        eng._forward(sup); eng._loss(sup)
        eng.lib.iins_set_deferred_join(1)
        eng._backward(sup)
        eng.lib.iins_set_deferred_join(0)
        eng._join_weight_gradients()

    def enc_only():
        lib, c = eng.lib, eng.cfg
        from iins_vae_b200._capi import ptr
        from iins_vae_b200.engine import _stream
        lib.iins_encoder_forward(c, eng.ptab["enc"], ptr(eng.cir), None, 0, 0, ptr(eng.rc), ptr(eng.cat), None, ptr(eng.kl), ptr(eng.ws["encoder"]), _stream())

    def full():
        eng._step_body(sup, True)

    t_enc, t_f, t_fb, t_all = timed(enc_only), timed(fwd), timed(fwd_loss_bwd), timed(full)
    print(f"supervised={sup} B={B}: encoder fwd {t_enc:.0f} us | forward {t_f:.0f} us | forward+loss+backward {t_fb:.0f} us (backward ~{t_fb - t_f:.0f}) | full step {t_all:.0f} us")
