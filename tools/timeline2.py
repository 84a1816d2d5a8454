"""Clock trace (CTA 0, thread 0) of chosen tensor-core forward launches inside an encoder forward."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iins_vae_b200
from iins_vae_b200._capi import get_lib, IinsConfig, ptr, ptr_array
from oracle import iins_oracle as orc
lib = get_lib(); iins_vae_b200.set_compute_mode("fp32")
cfg = orc.PathConfig(); B = 4096
pe, pd, pr, pc = orc.init_all(cfg, 0)
c = IinsConfig(B, 157, 4, 3, 4, 16, 2, 5, 16)
P = [v.cuda() for k, v in pe.items()]
x = torch.randn(B, 157, device="cuda")
rc = torch.zeros(B, 2, 8, device="cuda"); cat = torch.zeros(B, 16, device="cuda"); kl = torch.zeros(1, device="cuda")
ws = torch.zeros(lib.iins_encoder_ws_floats(c) + 16, device="cuda")
tl = torch.zeros(1024, dtype=torch.int64, device="cuda")
names = {0: "start", 1: "stage-free wait", 2: "split+store+arrive", 3: "prefetch next", 13: "drain (last MMAs)", 14: "tmem->smem", 15: "epilogue"}
def run():
    lib.check(lib.iins_encoder_forward(c, ptr_array(P), ptr(x), None, 0, 0, ptr(rc), ptr(cat), None, ptr(kl), ptr(ws), None), "enc")
for _ in range(3): run()
torch.cuda.synchronize()
for which in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4]:
    tl.zero_()
    lib.dll.iins_debug_set_timeline(C.c_void_p(tl.data_ptr()), which)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); run(); t1.record(); torch.cuda.synchronize()
    lib.dll.iins_debug_set_timeline(None, -1)
    t = tl.cpu().tolist(); n = t[1022]
    print(f"--- tensor-core forward launch #{which} of encoder_forward ({n} events; whole encoder fwd {t0.elapsed_time(t1)*1e3:.0f} us)")
    prev = t[1]; agg = {}
    for i in range(n):
        tag, clk = t[2 * i], t[2 * i + 1]
        agg[tag] = agg.get(tag, 0) + clk - prev; prev = clk
    for tag, v in agg.items(): print(f"   {names.get(tag, tag):22s} {v:8d} cycles")
    print("   total", t[2 * (n - 1) + 1] - t[1])
    print(f"   inside epilogue: stats {t[901] - t[900]} cycles, apply+store {t[902] - t[901]} cycles")
