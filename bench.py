#!/usr/bin/env python
"""bench.py -- IIns-VAE train throughput (samples/s of 157-tap CIR windows) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the box's HOST cores (oracle port)

Workload (BASELINE.json configs[1]): the full IIns-VAE training step -- Encoder + Decoder + Restorer +
Classifier forward, reconstruction + KL (+ range-error L1 + env cross-entropy on supervised batches,
train_semi.py:203 mask with supervision_rate 0.1), backward, Adam(lr 1e-4, betas (0.5, 0.999)) -- at batch
4096 per GPU, fp32, dim=4 / env_dim=16 / range_dim=2 / NC=5, synthetic CIR tensors of the zenodo loader's
shape.  One "step" = one batch.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train samples/sec (CIR windows)"
UNIT = "samples/s"
N_BATCHES = 8            # distinct synthetic batches cycled through (no step sees the previous step's inputs)


# ----------------------------------------------------------------------------------------------------
class PathShape:
    """Shape options of the benchmarked 1-D path (models.py:33,68,97,118; train_semi.py:77-82 with the SURVEY 8(d)
    defaults).  The product arm takes nothing from oracle/: parameters come from the modules' own reference-style
    initialisation (train_semi.py:104-107), batches from iins_vae_b200.data.SyntheticCIR, the supervision mask from
    iins_vae_b200.parallel.SupervisionMask."""
    cir_len, dim, n_residual, n_downsample, env_dim, range_dim, num_classes, pooled_len = 157, 4, 3, 4, 16, 2, 5, 128

    @property
    def trunk_dim(self):
        return self.dim * 2 ** self.n_downsample

    @property
    def code_len(self):
        return self.pooled_len // (2 ** self.n_downsample)

    @property
    def n_adain(self):
        return 2 * self.n_residual * 2 * self.trunk_dim


def algorithmic_flops(cfg, supervised: bool) -> float:
    """2*MACs of every Conv1d / Linear executed per SAMPLE in one train step:
    forward + dgrad + wgrad, no data gradient into the two stem convs (their input is the data)."""
    d, D, E, R, NC = cfg.dim, cfg.trunk_dim, cfg.env_dim, cfg.range_dim, cfg.num_classes
    fwd = 0.0
    no_dgrad = 0.0

    def conv(L, cin, cout, k, first=False):
        nonlocal fwd, no_dgrad
        f = 2.0 * L * cin * cout * k
        fwd += f
        if first:
            no_dgrad += f

    conv(128, 1, d, 7, first=True)
    L, c = 128, d
    for _ in range(cfg.n_downsample):
        conv(L // 2, c, 2 * c, 4)
        L //= 2
        c *= 2
    for _ in range(2 * cfg.n_residual):
        conv(L, c, c, 3)
    conv(L, c, R, 1)
    conv(128, 1, 4 * d, 7, first=True)
    conv(64, 4 * d, 8 * d, 4)
    conv(32, 8 * d, 16 * d, 4)
    conv(1, 16 * d, E, 1)
    # decoder
    conv(1, E, 256, 1); conv(1, 256, 256, 1); conv(1, 256, cfg.n_adain, 1)
    conv(L, R, D, 1)
    for _ in range(2 * cfg.n_residual):
        conv(L, D, D, 3)
    c = D
    for _ in range(cfg.n_downsample):
        conv(2 * L, c, c // 2, 5)
        L *= 2
        c //= 2
    conv(L, c, 1, 7)
    if supervised:
        conv(1, R * cfg.code_len, 512, 1); conv(1, 512, 256, 1); conv(1, 256, 256, 1); conv(1, 256, 1, 1)
        conv(1, E, 16, 1); conv(1, 16, 32, 1); conv(1, 32, 16, 1); conv(1, 16, NC, 1)
    return 3.0 * fwd - no_dgrad


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Index of the next sample: brackets the timed region inside a sampler that was started well before it
        (nvidia-smi needs ~100 ms to produce its first line; the timed region of the default run is ~50 ms)."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)                            # let the last in-region sample arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        last = len(self.rows) if last is None else min(len(self.rows), last + 2)   # + the samples straddling the end
        window = self.rows[first:last]
        where = "timed region"
        if not window:                              # region shorter than one sampling period: nearest samples under the same load
            window, where = self.rows[max(0, first - 3):last + 3], "timed region +/- 3 samples (same load: warm-up / e2e steps)"
        self.where = where
        for r in window:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": where, "period_ms": 20}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
def hbm_kernel_rooflines(lib, peaks):
    """The path's pure-streaming kernels (fused loss + seed gradients, fused Adam) timed ALONE at a size that leaves
    the 126 MB L2 (at the bench batch they move 8-18 MB and are launch-bound): achieved = algorithmic bytes / CUDA-event
    time, peak = measured HBM copy bandwidth (burst figure: kernel timed alone).  SURVEY.md 8(d): loss 2.07 KB/sample,
    Adam 28 B/parameter."""
    import ctypes as C
    from iins_vae_b200._capi import ptr
    out = []
    peak = float(peaks.get("hbm_gbs", 6548.8))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    Bb, L, NC = 1 << 19, 157, 5                      # 524288 windows: x + xrec + d_xrec = 3 x 329 MB
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(Bb, L, device="cuda", generator=g); xr = torch.randn(Bb, L, device="cuda", generator=g)
    err = torch.rand(Bb, 1, device="cuda", generator=g); ee = torch.rand(Bb, 1, device="cuda", generator=g)
    lg = torch.randn(Bb, NC, device="cuda", generator=g); lab = torch.randint(0, NC, (Bb, 1), device="cuda", generator=g).float()
    o = torch.zeros(8, device="cuda"); dx = torch.empty_like(x); de = torch.empty_like(ee); dl = torch.empty_like(lg)
    pred = torch.zeros(Bb, dtype=torch.int32, device="cuda")
    ms = timed(lambda: lib.check(lib.iins_loss_forward_backward(Bb, L, NC, ptr(x), ptr(xr), ptr(err), ptr(ee), ptr(lg), ptr(lab), None, 0,
                                                                1.0, 10.0, 1.0, ptr(o), ptr(dx), ptr(de), ptr(dl), ptr(pred), st), "loss"))
    nbytes = Bb * (3 * L * 4 + 2 * 4 + 4 + 4 + NC * 4 + 4 + 4 + NC * 4 + 4)
    out.append({"kernel": "iins_loss_kernel", "workload": f"{Bb} windows (fused L1 recon + L1 err + CE + seed gradients + metrics)",
                "bytes": nbytes, "ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": nbytes / (ms * 1e-3) / 1e9 / peak})
    del x, xr, dx
    P = 1 << 26                                      # 64 Mi parameters: 7 x 4 B x P = 1.9 GB per update
    w = torch.randn(P, device="cuda", generator=g); gr = torch.randn(P, device="cuda", generator=g)
    m = torch.zeros(P, device="cuda"); v = torch.zeros(P, device="cuda")
    steps = torch.zeros(8, dtype=torch.int32, device="cuda"); lr = torch.full((1,), 1e-4, device="cuda")
    gb, ge, ga = (C.c_int64 * 1)(0), (C.c_int64 * 1)(P), (C.c_int32 * 1)(1)
    ms = timed(lambda: lib.check(lib.iins_adam_step(ptr(w), ptr(gr), ptr(m), ptr(v), gb, ge, ga, 1, ptr(steps), ptr(lr), 0.5, 0.999,
                                                    1e-8, 1.0, 1, st), "adam"))
    nbytes = 28 * P
    out.append({"kernel": "iins_adam_kernel", "workload": f"{P} parameters (read p,g,m,v; write p,m,v)", "bytes": nbytes, "ms": ms,
                "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak})
    return out


def cpu_reference_run(batch, steps, warmup, threads):
    """The reference's own implementation of the step on the host cores; returns (samples/s, ms/step, kind, note).
    kind "reference": the UNMODIFIED reference modules (baseline/_ref/models.py, installed by baseline/install_ref.py) driven
    through the restated train_semi.py:183-228 loop body with torch.optim.Adam (baseline/ref_step.py) -- none of this
    repository's code on the path.  kind "port": only if that install is absent -- the oracle port (oracle/iins_oracle.py);
    this function and the smoke test are the only non-test code that may touch oracle/."""
    torch.set_num_threads(threads)
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_step
    if ref_step.available():
        tr = ref_step.ReferenceTrainer("cpu")
        batches = ref_step.synthetic_batches(2, batch)
        masks = ref_step.mask_stream()
        t0 = None
        for step in range(warmup + steps):
            if step == warmup:
                t0 = time.perf_counter()
            loss = tr.step(*batches[step % 2], next(masks))
            float(loss)                                   # the reference reads the loss every step (train_semi.py:240-268)
        dt = time.perf_counter() - t0
        sha = ref_step.manifest().get("sha256", "?")[:16]
        return batch * steps / dt, dt / steps * 1e3, "reference", f"unmodified reference models.py (sha256 {sha}) + train_semi.py:183-228 loop body, torch CPU"
    from oracle import iins_oracle as orc
    cfg = orc.PathConfig()
    pe, pd, pr, pc = orc.init_all(cfg, 1234)
    flat = {f"{g}.{k}": v for g, d in zip(("enc", "dec", "res", "cls"), (pe, pd, pr, pc)) for k, v in d.items()
            if not orc.is_buffer(k)}
    adam = orc.AdamState(flat)
    batches = [orc.synthetic_batch(cfg, batch, 1234 + j) for j in range(2)]
    rng = np.random.RandomState(1234)
    t0 = None
    for step in range(warmup + steps):
        if step == warmup:
            t0 = time.perf_counter()
        cir, err, label = batches[step % 2]
        mask = orc.supervision_mask(rng, 0.1)
        groups = {g: {} for g in ("enc", "dec", "res", "cls")}
        for g, d in zip(("enc", "dec", "res", "cls"), (pe, pd, pr, pc)):
            for k, v in d.items():
                groups[g][k] = v if orc.is_buffer(k) else flat[f"{g}.{k}"]
        _, grads = orc.semi_step_with_grads(groups["enc"], groups["dec"], groups["res"], groups["cls"], cir, err, label,
                                            cfg, bool(mask))
        flat = adam.step(flat, grads)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, "port", "oracle port of models.py + train_semi.py:183-228, torch CPU"


def gpu_eager_reference(batch, steps=10, warmup=3, dim=4):
    """SURVEY.md 2.2 / BASELINE.md section 4: "the kernel to beat on the same box" = the UNMODIFIED reference modules in eager
    PyTorch-CUDA (cuDNN / cuBLAS / ATen) on this B200, same step, same batch.  Timed with CUDA events outside the headline
    timed region.  `value`: inputs resident in HBM, no host read-back; `e2e`: the reference's own loop shape -- pageable
    host batch -> .cuda() per tensor and loss.item() every step (train_semi.py:174-180, :240-268)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_step
    if not ref_step.available():
        return {"unavailable": "baseline/_ref/models.py not installed (run baseline/install_ref.py where /root/reference exists)"}
    tr = ref_step.ReferenceTrainer("cuda", dim=dim)
    host = ref_step.synthetic_batches(4, batch)
    dev = [tuple(t.cuda() for t in b) for b in host]
    masks = ref_step.mask_stream()
    for i in range(warmup):
        tr.step(*dev[i % 4], next(masks))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        tr.step(*dev[i % 4], next(masks))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        float(tr.step(*host[i % 4], next(masks)))
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / steps
    return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "e2e_value": batch / (ms_e2e * 1e-3),
            "e2e_ms_per_step": ms_e2e, "batch": batch, "steps": steps, "dtype": "fp32 (torch default: TF32 off)",
            "what": "unmodified reference models.py (baseline/_ref) + train_semi.py:183-228 loop body + torch.optim.Adam, eager "
                    "PyTorch-CUDA on the same GPU, CUDA-event timing"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    val, ms, kind, note = cpu_reference_run(args.batch, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} timed steps of batch {args.batch} after {args.warmup} warm-up "
                                   f"({note}, {threads} threads)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def workload_config(batch, gpus, dim=4):
    return {"workload": "IIns-VAE semi-supervised train step (Enc+Dec+Res+Cls, recon+KL+L1+CE, Adam), "
                        f"BASELINE configs[1]: batch {batch} per GPU, fp32" + ("" if dim == 4 else f", dim = {dim} (--filters {dim})"),
            "batch_per_gpu": batch, "global_batch": batch * gpus, "cir_len": 157, "dim": dim, "env_dim": 16,
            "range_dim": 2, "num_classes": 5, "supervision_rate": 0.1, "parallelism": f"dp{gpus}",
            "l2_policy": f"{N_BATCHES} distinct input batches cycled; per-step working set (saved activations "
                         "~0.5 MB/sample) is far larger than the 126 MB L2"}


# ----------------------------------------------------------------------------------------------------
_STDOUT_FD = None


def _capture_stdout():
    """stdout carries exactly ONE line (the JSON): native libraries (NCCL prints its version banner there) and anything
    else that writes to fd 1 during the run go to stderr; the original stdout is restored by _emit() only."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _STDOUT_FD is not None:
        os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the eager PyTorch-CUDA reference leg")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (BASELINE configs[3]): fixed GLOBAL batch, per-GPU batch = global / world")
    ap.add_argument("--compute-mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-overlap", action="store_true", help="one all-reduce after the backward instead of overlapped buckets")
    ap.add_argument("--dim", type=int, default=4,
                    help="base channel count (models.py:33 `dim`; 4 = the benchmarked configuration, 16 = utils.py:38 --filters 16)")
    args = ap.parse_args()
    _capture_stdout()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    from iins_vae_b200 import models as M
    from iins_vae_b200.data import SyntheticCIR
    from iins_vae_b200.engine import SemiTrainEngine
    from iins_vae_b200.parallel import SupervisionMask
    from iins_vae_b200._capi import get_lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL banners off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    lib = get_lib()
    cfg = PathShape()
    cfg.dim = args.dim
    scaling = "weak"
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must divide by the number of ranks")
        args.batch, scaling = args.global_batch // world, "strong"
    B, K, W = args.batch, args.steps, args.warmup
    import iins_vae_b200
    iins_vae_b200.set_compute_mode(args.compute_mode)

    torch.manual_seed(1234)                                     # random-init weights of the named architecture, same on every rank
    Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
    Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
    Res = M.Restorer((cfg.range_dim, cfg.code_len))
    Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
    for m in (Enc, Dec, Res, Cls):
        m.apply(M.weights_init_normal)                          # train_semi.py:104-107
        m.cuda()
    eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, cir_len=cfg.cir_len, lr=1e-4, betas=(0.5, 0.999),
                          use_graph=True, process_group=pg, overlap_allreduce=not args.no_overlap)

    data = SyntheticCIR(N_BATCHES * B, B, cfg.cir_len, cfg.num_classes, seed=1234 + 100 * rank, pin=True)
    host = [(b["CIR"], b["Err"], b["Label"]) for b in data]     # pinned host batches of the loader's shape (dataset.py:118-133)
    dev = [tuple(t.cuda() for t in b) for b in host]
    mask_stream = SupervisionMask(0.1, seed=1234)               # identical mask sequence on every rank (train_semi.py:203)
    masks = [mask_stream() for _ in range(W + 2 * K + 8)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                             # long before the timed region: nvidia-smi needs ~100 ms to start
    # launches per step, counted from eager passes of both branches (also captures both CUDA graphs)
    launches = {}
    for sup in (True, False):
        eng.step(*dev[0], supervised=sup)            # eager warm-up + capture inside
        torch.cuda.synchronize()
    eng_eager = eng.use_graph
    eng.use_graph = False
    for sup in (True, False):
        c0 = lib.iins_launch_count()
        eng.step(*dev[0], supervised=sup)
        torch.cuda.synchronize()
        launches[sup] = lib.iins_launch_count() - c0
    eng.use_graph = eng_eager

    step_i = 0
    for _ in range(W):
        eng.step(*dev[step_i % N_BATCHES], supervised=bool(masks[step_i]))
        step_i += 1

    # ---------------- timed region 1: inputs resident in HBM
    barrier()
    mark0 = sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_launch = 0
    flops = 0.0
    ev0.record()
    for _ in range(K):
        sup = bool(masks[step_i])
        eng.step(*dev[step_i % N_BATCHES], supervised=sup)
        n_launch += launches[sup]
        flops += algorithmic_flops(cfg, sup) * B
        step_i += 1
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    mark1 = sampler.mark()
    barrier()

    # ---------------- timed region 2 (e2e): pinned host inputs, H2D inside, loss read back every step
    # Every step: one H2D copy of a pinned host batch (issued by the engine's copy stream while the previous step computes:
    # SemiTrainEngine.prefetch, the repo's input pipeline) and one D2H read of the step's loss vector.  The read-back is
    # pipelined by one step -- the host waits for the loss of step i after it has queued step i+1 -- so the device never
    # idles on the host; every step's loss is read (and checked finite) inside the timed region.
    out_hosts = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)]
    read_evs = [torch.cuda.Event() for _ in range(2)]
    seen = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.prefetch(*host[step_i % N_BATCHES])
    for k in range(K):
        sup = bool(masks[step_i])
        out = eng.step(supervised=sup, prefetched=True)
        if k + 1 < K:
            eng.prefetch(*host[(step_i + 1) % N_BATCHES])
        out_hosts[k % 2].copy_(out, non_blocking=True)
        read_evs[k % 2].record()
        if k >= 1:
            read_evs[(k - 1) % 2].synchronize()
            seen.append(float(out_hosts[(k - 1) % 2][0]))
        step_i += 1
    read_evs[(K - 1) % 2].synchronize()
    seen.append(float(out_hosts[(K - 1) % 2][0]))
    out_host = out_hosts[(K - 1) % 2]
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    barrier()
    assert len(seen) == K and np.isfinite(seen).all() and np.isfinite(out_host.numpy()).all(), "non-finite loss"
    clocks = sampler.stop(mark0, mark1) if rank == 0 else None      # stopped after the e2e region (same load)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # ---------------- per-kernel timing inside a real step (CUDA events around every launch, eager pass)
    roofline = None
    kernel_table = None
    # every rank runs the instrumented steps (they contain the gradient all-reduce); rank 0 reports
    eng.use_graph = False
    eng.set_concurrency(False)              # serial launches: an event pair then brackets exactly one kernel
    prof = []
    for sup in (True, False):
        rows = lib.profile(lambda: eng.step(*dev[0], supervised=sup))
        prof += [(n_, ms_, fl_, float(lib.last_bytes[i])) for i, (n_, ms_, fl_) in enumerate(rows)]
    eng.set_concurrency(True)
    eng.use_graph = eng_eager
    barrier()
    if rank == 0:
        # Roofline of a kernel family: each launch is bounded by the SLOWER of its two rooflines,
        #   t_roof = max(algorithmic flops / measured bf16 tensor peak, algorithmic HBM bytes / measured HBM bandwidth)
        # (flops = 2*M*N*K of the layer; bytes = every operand tensor read once + every result tensor written once, fp32:
        # iins_runtime.cu nt_bytes / dz_bytes).  frac = sum(t_roof) / sum(measured time); `bound` says
        # which roofline supplies most of sum(t_roof); `achieved` / `peak` are quoted in that roofline's unit.
        peaks, peak_src = measured_peaks()
        peak_t = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        peak_h = float(peaks.get("hbm_gbs", 6548.8))
        agg = {}
        for name, kms, fl, by in prof:
            a = agg.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0, "roof_ms": 0.0, "roof_hbm_ms": 0.0, "roof_tensor_ms": 0.0})
            t_t, t_h = fl / (peak_t * 1e12) * 1e3, by / (peak_h * 1e9) * 1e3
            a["ms"] += kms; a["flops"] += fl; a["bytes"] += by; a["n"] += 1
            a["roof_ms"] += max(t_t, t_h)
            if t_h >= t_t:
                a["roof_hbm_ms"] += t_h
            else:
                a["roof_tensor_ms"] += t_t
        total_ms = sum(a["ms"] for a in agg.values())
        top_name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        # DRAM traffic per launch of that kernel from the committed ncu --set full capture (profiles/), if present
        traffic = None
        for tname in ("r02f_ncu_traffic.json", "r02e_ncu_traffic.json", "r02a_ncu_traffic.json", "r01b_ncu_traffic.json"):   # newest capture that holds this family
            tpath = os.path.join(ROOT, "profiles", tname)
            if traffic is None and os.path.exists(tpath):
                with open(tpath) as f:
                    rows = [r for r in json.load(f) if top_name.rstrip("_") in r["kernel"]]
                if rows:
                    traffic = sum(r["dram_bytes"] for r in rows) / len(rows)

        def family_roofline(a):
            hbm = a["roof_hbm_ms"] >= a["roof_tensor_ms"]
            t = max(a["ms"], 1e-9) * 1e-3
            return {"bound": "hbm" if hbm else "tensor",
                    "achieved": a["bytes"] / t / 1e9 if hbm else a["flops"] / t / 1e12,
                    "peak": peak_h if hbm else peak_t, "unit": "GB/s" if hbm else "TFLOP/s",
                    "frac": a["roof_ms"] / max(a["ms"], 1e-9),
                    "tensor_tflops": a["flops"] / t / 1e12, "tensor_frac": a["flops"] / t / 1e12 / peak_t,
                    "hbm_gbs": a["bytes"] / t / 1e9, "hbm_frac": a["bytes"] / t / 1e9 / peak_h}

        roofline = family_roofline(top)
        roofline.update({"kernel": top_name, "traffic": traffic, "share_of_step": top["ms"] / total_ms,
                         "avg_launch_ms": top["ms"] / top["n"],
                         "algorithmic_bytes_per_launch": top["bytes"] / top["n"], "algorithmic_flops_per_launch": top["flops"] / top["n"],
                         "peak_source": peak_src + ": HBM copy bandwidth / bf16 dense sustained (kernels timed inside a long step)",
                         "step": {"frac": sum(a["roof_ms"] for a in agg.values()) / total_ms,
                                  "roof_ms": sum(a["roof_ms"] for a in agg.values()), "kernel_ms": total_ms,
                                  "note": "sum over ALL launches of one supervised + one unsupervised step; launches without a byte / "
                                          "flop model (pooling, reparameterisation, small reductions) count as time with no roofline"},
                         "note": "every launch is bounded by the slower of its two rooflines, max(2*M*N*K / tensor peak, algorithmic "
                                 "bytes / HBM peak); frac = sum of those bounds / CUDA-event time of the family's launches inside an "
                                 "instrumented step (launches serialised); tensor_* / hbm_* give both rooflines separately; the "
                                 "fp32-grade mode issues 8 bf16 piece products per fp32 product (bf16x3 split), so the tensor pipe "
                                 "does ~2.7x the algorithmic flops; traffic = mean DRAM bytes per launch from the committed ncu "
                                 f"capture; whole-step algorithmic TFLOP/s = {flops / (ms * 1e-3) / 1e12:.2f}"})
        kernel_table = {}
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            r = family_roofline(a)
            kernel_table[k] = {"ms": round(a["ms"], 4), "launches": a["n"], "tflops": round(r["tensor_tflops"], 3),
                               "gbs": round(r["hbm_gbs"], 1), "bound": r["bound"], "frac": round(r["frac"], 4)}

    hbm_rooflines = None
    if rank == 0 and world == 1:
        try:
            hbm_rooflines = hbm_kernel_rooflines(lib, measured_peaks()[0])
        except Exception as e:                      # never let the side measurement take the headline down
            hbm_rooflines = [{"error": repr(e)}]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        val, _, kind, note = cpu_reference_run(B, 8, 2, threads)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"8 timed steps of batch {B} after 2 warm-up ({note}, {threads} threads)"}

    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        try:
            gpu_ref = gpu_eager_reference(B, dim=cfg.dim)
        except Exception as e:
            gpu_ref = {"error": repr(e)}

    if rank == 0:
        value = world * B * K / (ms * 1e-3)
        e2e = world * B * K / (ms_e2e * 1e-3)
        h2d = sum(t.numel() * t.element_size() for t in host[0])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "fp32" if args.compute_mode == "fp32" else "bf16",
            "data": "synthetic", "config": workload_config(B, world, cfg.dim),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                    "ms_per_step": ms_e2e / K,
                    "pipeline": "H2D of batch i+1 on a copy stream during step i; loss of step i read back after step i+1 is queued"},
            "gpu_launches": int(n_launch), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernels": kernel_table, "roofline_hbm": hbm_rooflines, "gpu_eager_reference": gpu_ref,
            "final_loss_terms": {k: float(v) for k, v in zip(("l1_recon", "l1_err", "ce", "weighted_sum"), out_host[:4].tolist())},
        }
        _emit(line)
    if world > 1:
        from iins_vae_b200.parallel import shutdown_distributed
        shutdown_distributed([eng])


if __name__ == "__main__":
    main()
