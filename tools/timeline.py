"""Per-phase clock trace of CTA(0,0)/thread 0 of the tensor-core forward kernel for chosen layer shapes."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iins_vae_b200
from iins_vae_b200._capi import get_lib, IinsConfig, ptr, ptr_array
from oracle import iins_oracle as orc

lib = get_lib()
iins_vae_b200.set_compute_mode("fp32")
cfg = orc.PathConfig()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pe, pd, pr, pc = orc.init_all(cfg, 0)
c = IinsConfig(B, 157, 4, 3, 4, 16, 2, 5, 16)
P = [v.cuda() for k, v in pr.items()]
rc = torch.rand(B, 2, 8, device="cuda")
out = torch.zeros(B, 1, device="cuda")
ws = torch.zeros(lib.iins_restorer_ws_floats(c) + 16, device="cuda")
tl = torch.zeros(1024, dtype=torch.int64, device="cuda")
names = {0: "start", 1: "mma-wait", 2: "gather+store", 3: "fence", 4: "syncthreads", 5: "tma-wait", 6: "mma-issue+commit", 13: "drain", 14: "tmem->smem", 15: "epilogue"}
for it in range(3):
    lib.iins_restorer_forward(c, ptr_array(P), ptr(rc), ptr(out), ptr(ws), None)
torch.cuda.synchronize()
names[7] = "  mma"
for which, desc in ((1, "layer 2: M=%d N=256 (NT=64) K=512" % B), (3, "layer 4: M=%d N=1 (NT=16) K=256" % B)):
    tl.zero_()
    lib.dll.iins_debug_set_timeline(C.c_void_p(tl.data_ptr()), which)
    lib.iins_restorer_forward(c, ptr_array(P), ptr(rc), ptr(out), ptr(ws), None)
    torch.cuda.synchronize()
    lib.dll.iins_debug_set_timeline(None, -1)
    t = tl.cpu().tolist()
    n = t[1022]
    print("restorer forward,", desc, n, "events")
    prev = t[1]
    shown = 0
    for i in range(n):
        tag, clk = t[2 * i], t[2 * i + 1]
        if shown < 44 or tag >= 13:
            print(f"  {names.get(tag, tag):18s} +{clk - prev:7d} cycles")
            shown += 1
        prev = clk
    print("total", t[2 * (n - 1) + 1] - t[1])
