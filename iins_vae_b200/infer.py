"""Sharded inference driver (BASELINE configs[4]: test.py inference-only throughput -- range error + environment id -- on
10 M CIR windows over 8 B200s).

    python -m iins_vae_b200.infer --synthetic 10000000 --batch_size 32768                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m iins_vae_b200.infer --synthetic 10000000

The path of test.py:55-85 (no-grad ``network(cir)`` = Encoder -> (Classifier, Restorer), per-batch RMSE / MAE / argmax accuracy)
shards along the windows with no data-path collective: every rank takes a contiguous 1/N of the windows (parallel.shard_range),
streams it from PINNED HOST memory through a two-slot device ring -- the host-to-device copy of batch i+1 runs on a copy stream
while batch i computes (the reference does a synchronous pageable ``.cuda()`` per tensor, test.py:61-64) -- and runs the
captured inference graph per batch.  Range-error estimates (fp32) and argmax labels (int32) of every window stay in device
arrays; the three metric sums are all-reduced ONCE at the end.  The timed region (CUDA events, max over ranks) contains every
H2D copy.
"""
import argparse
import json
import os
import sys
import time

import torch

from .engine import InferenceEngine
from .parallel import init_distributed, shard_range, shutdown_distributed

BYTES_PER_WINDOW = 157 * 4 + 4 + 4          # CIR + err + label in, (err_est + argmax = 8 B out)


class ShardedInference:
    """One rank's share of a window set held in pinned host memory."""

    def __init__(self, network, cir, err, label, batch_size=32768, label_offset=0, device=None):
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.cir, self.err, self.label = cir, err, label
        self.n = int(cir.shape[0])
        self.B = int(min(batch_size, self.n))
        self.net = network
        L = int(cir.shape[1])
        self.engines = {}
        self.label_offset = label_offset
        self.copy_stream = torch.cuda.Stream(device=self.device)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        self.ring = [(f(self.B, L), f(self.B, 1), f(self.B, 1)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.err_est = f(self.n, 1)
        self.pred = torch.empty(self.n, dtype=torch.int32, device=self.device)
        self.sums = torch.zeros(4, dtype=torch.float64, device=self.device)     # sum |e|, sum e^2, #correct, #windows

    def _engine(self, b):
        if b not in self.engines:
            self.engines[b] = InferenceEngine(self.net.encoder, self.net.restorer, self.net.classifier, batch_size=b,
                                              cir_len=self.cir.shape[1], label_offset=self.label_offset, device=self.device)
        return self.engines[b]

    def _issue_copy(self, i, slot):
        lo, hi = i * self.B, min((i + 1) * self.B, self.n)
        self.copy_stream.wait_event(self.free[slot])
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.ring[slot], (self.cir, self.err, self.label)):
                dst[:hi - lo].copy_(src[lo:hi].view(hi - lo, -1), non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def run(self):
        """One pass over the shard.  Returns the device tensor of metric sums (no host synchronisation)."""
        main = torch.cuda.current_stream()
        nb = (self.n + self.B - 1) // self.B
        self.sums.zero_()
        for s in range(2):
            self.free[s].record(main)
        self._issue_copy(0, 0)
        for i in range(nb):
            slot = i & 1
            if i + 1 < nb:
                self._issue_copy(i + 1, slot ^ 1)
            lo, hi = i * self.B, min((i + 1) * self.B, self.n)
            b = hi - lo
            eng = self._engine(b)
            main.wait_event(self.ready[slot])
            cir, err, label = (t[:b] for t in self.ring[slot])
            e, pred, out = eng.run(cir, err, label)
            self.free[slot].record(main)
            self.err_est[lo:hi].copy_(e, non_blocking=True)
            self.pred[lo:hi].copy_(pred, non_blocking=True)
            # out[1] = mean |err - est|, out[4] = mean squared error, out[5] = number of correct argmax (iins_b200.h)
            self.sums[0] += out[1].double() * b
            self.sums[1] += out[4].double() * b
            self.sums[2] += out[5].double()
            self.sums[3] += b
        return self.sums


def synthetic_windows(n, cir_len=157, num_classes=5, seed=1234, label_base=0, chunk=1 << 20):
    """n windows of the zenodo loader's shape in PINNED host memory (generated on the device in chunks, copied back)."""
    cir = torch.empty(n, cir_len, dtype=torch.float32).pin_memory()
    err = torch.empty(n, 1, dtype=torch.float32).pin_memory()
    label = torch.empty(n, 1, dtype=torch.float32).pin_memory()
    g = torch.Generator(device="cuda").manual_seed(seed)
    for lo in range(0, n, chunk):
        hi = min(lo + chunk, n)
        cir[lo:hi].copy_(torch.randn(hi - lo, cir_len, device="cuda", generator=g))
        err[lo:hi].copy_((torch.randn(hi - lo, 1, device="cuda", generator=g) * 0.15).abs().clamp_(0, 1))
        label[lo:hi].copy_((torch.randint(0, num_classes, (hi - lo, 1), device="cuda", generator=g) + label_base).float())
    torch.cuda.synchronize()
    return cir, err, label


def run_sharded(network, n_windows, batch_size=32768, num_classes=5, passes=1, warmup=1, label_offset=0, seed=1234):
    """Shard n_windows over the ranks of the torchrun launch, run `warmup` untimed and `passes` timed passes; returns a dict
    with the aggregate windows/s (all ranks' windows / max-over-ranks device time) and the reduced metrics."""
    import torch.distributed as dist
    rank, local, world, pg = init_distributed()
    torch.cuda.set_device(local)
    lo, hi = shard_range(n_windows - n_windows % world, rank, world)
    cir, err, label = synthetic_windows(hi - lo, 157, num_classes, seed + rank, label_offset)
    sh = ShardedInference(network, cir, err, label, batch_size, label_offset)
    for _ in range(warmup):
        sh.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        sums = sh.run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    tot = sums.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)                                   # the ONE collective of the path: 4 scalars
    ms = float(ms)
    s = tot.tolist()
    n_total = (hi - lo) * world
    res = dict(windows=n_total, passes=passes, n_gpus=world, ms=ms, windows_per_s=n_total * passes / (ms * 1e-3),
               mae=s[0] / s[3], rmse=(s[1] / s[3]) ** 0.5, accuracy=s[2] / s[3], batch_size=sh.B,
               h2d_bytes_per_window=BYTES_PER_WINDOW, h2d_gbs=n_total * passes * BYTES_PER_WINDOW / (ms * 1e-3) / 1e9 / world)
    if world > 1:
        shutdown_distributed(list(sh.engines.values()))
    return res, rank


def main(argv=None):
    from . import models as M, set_compute_mode
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", type=int, default=10_000_000, help="number of synthetic CIR windows (BASELINE configs[4]: 10 M)")
    ap.add_argument("--batch_size", type=int, default=32768)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--num_classes", type=int, default=5)
    ap.add_argument("--compute_mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--checkpoint", default=None, help="Network_%%d.pth written by train.py (random init when absent)")
    opt = ap.parse_args(argv)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    set_compute_mode(opt.compute_mode)
    torch.manual_seed(1234)
    net = M.EMNet(cir_len=157, num_classes=opt.num_classes, env_dim=16)
    if opt.checkpoint:
        net.load_state_dict(torch.load(opt.checkpoint))
    else:
        for m in (net.encoder, net.restorer, net.classifier):
            m.apply(M.weights_init_normal)
    net.cuda().eval()
    res, rank = run_sharded(net, opt.synthetic, opt.batch_size, opt.num_classes, opt.passes)
    if rank == 0:
        peak = 6548.8
        try:
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        res["hbm_frac_algorithmic"] = res["windows_per_s"] / res["n_gpus"] * (BYTES_PER_WINDOW + 8) / 1e9 / peak
        res["metric"] = "inference windows/sec (range error + env id), H2D inside the timed region"
        print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    main()
