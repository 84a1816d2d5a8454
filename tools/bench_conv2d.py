"""Throughput of the 2-D variant (conv_type = 2, expand = True) on one B200, with the per-kernel roofline
and the UNMODIFIED reference's 2-D modules beside it (eager PyTorch-CUDA on the same GPU, and on the host cores).

    python tools/bench_conv2d.py [--batch 64] [--steps 10] [--mode fp32|bf16]

One step = the reference's loop shape on the drop-in modules (autograd path): Encoder(2, expand=True) -> Decoder -> Restorer ->
Classifier, L1 recon + KL + 10 L1 err + CE, backward, torch.optim.Adam (the optimizer the reference builds, train_semi.py:118).
Prints ONE JSON line."""
import argparse
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import numpy as np
import torch


def build(M, dev, cls_kw):
    Enc = M.Encoder(conv_type=2, dim=4, n_residual=3, n_downsample=4, style_dim=16, out_dim=2, expand=True)
    Dec = M.Decoder(conv_type=2, dim=4, n_residual=3, n_upsample=4, style_dim=16, in_dim=157, out_dim=2, expand=True)
    Res = M.Restorer((2, 8, 8), net_type="Linear")
    Cls = M.Classifier(16, 5, **cls_kw)
    mods = (Enc, Dec, Res, Cls)
    for m in mods:
        m.apply(M.weights_init_normal)
        m.to(dev)
    opt = torch.optim.Adam(itertools.chain(*(m.parameters() for m in mods)), lr=1e-4, betas=(0.5, 0.999))
    return mods, opt


def step(mods, opt, cir, err, label):
    Enc, Dec, Res, Cls = mods
    opt.zero_grad()
    rc, cat, lat, kl = Enc(cir)
    gen = Dec(rc, cat)
    loss = torch.nn.functional.l1_loss(cir, gen) + kl + 10 * torch.nn.functional.l1_loss(err, Res(rc)) + \
        torch.nn.functional.cross_entropy(Cls(cat), label)
    loss.backward()
    opt.step()
    return loss


def timed(fn, steps, warmup, cuda=True):
    for _ in range(warmup):
        fn()
    if not cuda:
        t0 = time.perf_counter()
        for _ in range(steps):
            float(fn())
        return (time.perf_counter() - t0) / steps * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    import iins_vae_b200
    from iins_vae_b200 import models as M
    from iins_vae_b200._capi import get_lib
    import bench
    iins_vae_b200.set_compute_mode(args.mode)
    lib = get_lib()
    B = args.batch
    g = torch.Generator().manual_seed(1234)
    cir = torch.randn(B, 157, generator=g).cuda()
    err = (torch.randn(B, 1, generator=g) * 0.15).abs().clamp_(0, 1).cuda()
    label = torch.randint(0, 5, (B,), generator=g).cuda()
    torch.manual_seed(1234)
    mods, opt = build(M, "cuda", {})
    ms = timed(lambda: step(mods, opt, cir, err, label), args.steps, args.warmup)
    # per-kernel roofline (launches serialised by the profile's event pairs)
    rows = lib.profile(lambda: step(mods, opt, cir, err, label))
    prof = [(n, t, f, float(lib.last_bytes[i])) for i, (n, t, f) in enumerate(rows)]
    peaks, src = bench.measured_peaks()
    pt, ph = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))), float(peaks.get("hbm_gbs"))
    agg = {}
    for n, t, f, by in prof:
        a = agg.setdefault(n, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0, "roof_ms": 0.0})
        a["ms"] += t; a["flops"] += f; a["bytes"] += by; a["n"] += 1
        a["roof_ms"] += max(f / (pt * 1e12), by / (ph * 1e9)) * 1e3
    tot = sum(a["ms"] for a in agg.values())
    kernels = {k: {"ms": round(a["ms"], 3), "launches": a["n"], "tflops": round(a["flops"] / max(a["ms"], 1e-9) / 1e9, 2),
                   "gbs": round(a["bytes"] / max(a["ms"], 1e-9) / 1e6, 1), "frac": round(a["roof_ms"] / max(a["ms"], 1e-9), 4)}
               for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    flops = sum(a["flops"] for a in agg.values())
    line = {"metric": "train samples/sec (CIR windows), 2-D variant", "value": B / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1,
            "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "dtype": args.mode, "data": "synthetic",
            "config": {"workload": f"IIns-VAE conv_type=2 expand=True train step (autograd path, torch.optim.Adam), batch {B}",
                       "dim": 4, "env_dim": 16, "range_dim": 2, "num_classes": 5, "cir_len": 157},
            "gpu_launches": len(prof), "kernel_ms_serial": tot, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "roofline_step_frac": sum(a["roof_ms"] for a in agg.values()) / max(tot, 1e-9), "peak_source": src, "kernels": kernels}
    if not args.no_reference:
        import ref_step
        if ref_step.available():
            ref = ref_step.load_reference_models()
            torch.manual_seed(1234)
            rm, ropt = build(ref, "cuda", {"net_type": "Linear"})
            rms = timed(lambda: step(rm, ropt, cir, err, label), args.steps, args.warmup)
            line["gpu_eager_reference"] = {"value": B / (rms * 1e-3), "ms_per_step": rms,
                                           "what": "unmodified reference models.py (conv_type=2, expand=True), eager PyTorch-CUDA, same GPU"}
            del rm, ropt
            torch.set_num_threads(os.cpu_count() or 1)
            cm, copt = build(ref, "cpu", {"net_type": "Linear"})
            cb = min(B, 16)
            cms = timed(lambda: step(cm, copt, cir[:cb].cpu(), err[:cb].cpu(), label[:cb].cpu()), 3, 1, cuda=False)
            line["cpu_baseline"] = {"value": cb / (cms * 1e-3), "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"3 timed steps of batch {cb} after 1 warm-up, unmodified reference 2-D modules, torch CPU"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
