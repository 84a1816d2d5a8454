"""Worker of tests/test_gpu_multi.py (launched under torchrun, one process per GPU, NCCL):

    N ranks x (B/N) samples must reproduce 1 rank x B samples -- gradients (after the all-reduce) and parameters after
    the fused Adam update -- for both supervision branches, with the step replayed from a CUDA graph and the gradient
    buckets reduced on the communication stream next to the encoder backward (train_semi.py:207,227 have no collective:
    this is the data-parallel exchange the B200 build adds).

Bounds (stated): per-sample arithmetic is identical on both sides (every normalisation is per sample), so only the order
of the fp32 sums over the batch differs: every gradient tensor within 1e-4 rel-L2 (the fp32 bound of north_star; measured
values are printed), parameters after one Adam step within parity.assert_traj_close (2e-5 relative, a few entries whose
gradient is rounding noise may move by up to 2*lr).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    from oracle import iins_oracle as orc
    from tests import parity
    from iins_vae_b200 import models as M
    from iins_vae_b200.engine import SemiTrainEngine
    from iins_vae_b200.parallel import init_distributed, shard_range, shutdown_distributed

    rank, local, world, pg = init_distributed()
    torch.cuda.set_device(local)
    B = int(os.environ.get("IINS_DP_BATCH", "4096"))
    cfg = orc.PathConfig()
    cir, err, label = orc.synthetic_batch(cfg, B, 4242)
    lo, hi = shard_range(B, rank, world)

    def modules(seed=77):
        pe, pd, pr, pc = orc.init_all(cfg, seed)
        Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
        Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
        Res = M.Restorer((cfg.range_dim, cfg.code_len))
        Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
        for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
            m.load_state_dict(p)
            m.cuda()
        return Enc, Dec, Res, Cls

    failures = []
    engines = []
    for overlap in (True, False):
        for supervised in (True, False):
            dp = SemiTrainEngine(*modules(), batch_size=hi - lo, use_graph=True, process_group=pg, overlap_allreduce=overlap)
            engines.append(dp)
            # gradients only (SUM over ranks: the 1/world factor lives in the fused Adam), replayed twice from the graph
            for _ in range(2):
                dp.step(cir[lo:hi], err[lo:hi], label[lo:hi], supervised=supervised, update=False)
            torch.cuda.synchronize()
            g_dp = {k: (v / world).cpu() for k, v in dp.named_grads().items()}
            loss_dp = torch.tensor([dp.loss_terms()["loss"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(loss_dp)
            dp.step(cir[lo:hi], err[lo:hi], label[lo:hi], supervised=supervised, update=True)
            torch.cuda.synchronize()
            p_dp = torch.cat([p.detach().reshape(-1) for p in dp.flat.params]).cpu()
            # every rank must hold the same parameters after the update
            chk = torch.stack([dp.flat.flat.double().sum(), dp.flat.flat.double().abs().sum()])
            lo_chk, hi_chk = chk.clone(), chk.clone()
            dist.all_reduce(lo_chk, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_chk, op=dist.ReduceOp.MAX)
            if not torch.equal(lo_chk, hi_chk):
                failures.append(f"overlap={overlap} sup={supervised}: ranks hold different parameters after the update")
            if rank == 0:
                one = SemiTrainEngine(*modules(), batch_size=B, use_graph=True)
                one.step(cir, err, label, supervised=supervised, update=False)
                torch.cuda.synchronize()
                g_one = {k: v.cpu() for k, v in one.named_grads().items()}
                loss_one = one.loss_terms()["loss"]
                one.step(cir, err, label, supervised=supervised, update=True)
                torch.cuda.synchronize()
                names = [n for n in g_one]
                worst = ("", 0.0)
                for k in names:
                    n1 = float(g_one[k].norm())
                    if n1 == 0.0:
                        if float(g_dp[k].abs().max()) != 0.0:
                            failures.append(f"{k}: gradient must be absent (None in the reference)")
                        continue
                    if orc.grad_is_structurally_zero(k):
                        continue
                    rel = float((g_dp[k] - g_one[k]).norm()) / n1
                    if rel > worst[1]:
                        worst = (k, rel)
                    if rel > parity.RTOL_FP32:
                        failures.append(f"overlap={overlap} sup={supervised} {k}: rel-L2 {rel:.2e} > {parity.RTOL_FP32}")
                if abs(float(loss_dp) / world - loss_one) > 1e-5 * abs(loss_one):
                    failures.append(f"loss: mean over ranks {float(loss_dp) / world} vs single rank {loss_one}")
                o = 0
                for (k, p) in zip(names, one.flat.params):
                    n = p.numel()
                    if not orc.grad_is_structurally_zero(k):
                        try:
                            parity.assert_traj_close(k, p_dp[o:o + n], p.detach().cpu().reshape(-1), 1)
                        except AssertionError as e:
                            failures.append(f"overlap={overlap} sup={supervised} post-Adam {e}")
                    o += n
                print(f"[dp] world={world} B={B} ({hi - lo}/rank) overlap={overlap} supervised={supervised}: worst gradient rel-L2 vs "
                      f"1 rank x {B}: {worst[1]:.2e} ({worst[0]}); loss {float(loss_dp) / world:.7f} vs {loss_one:.7f}", flush=True)
    # ---- Conv1d heads: BatchNorm over the batch -> SyncBN (2 * C double sums all-reduced between the two
    # phases of the head's pass), dropout masks drawn from Philox counters over GLOBAL sample indices: N ranks x B/N == 1 rank x B
    def conv_modules(seed=88):
        pe, pd, _, _ = orc.init_all(cfg, seed)
        gen = torch.Generator().manual_seed(seed + 99)
        pr = orc.init_conv_head_params(orc.restorer_conv1d_param_shapes(cfg), gen)
        pc = orc.init_conv_head_params(orc.classifier_conv1d_param_shapes(cfg), gen)
        Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
        Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
        Res = M.Restorer((cfg.range_dim, cfg.code_len), net_type="Conv1d")
        Cls = M.Classifier(cfg.env_dim, cfg.num_classes, net_type="Conv1d")
        for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
            m.load_state_dict(p)
            m.cuda()
        return Enc, Dec, Res, Cls

    mods_dp = conv_modules()
    dp = SemiTrainEngine(*mods_dp, batch_size=hi - lo, use_graph=True, process_group=pg)
    engines.append(dp)
    for _ in range(2):
        dp.step(cir[lo:hi], err[lo:hi], label[lo:hi], supervised=True, update=False)
    torch.cuda.synchronize()
    g_dp = {k: (v / world).cpu() for k, v in dp.named_grads().items()}
    bn_dp = getattr(mods_dp[2].restorer.conv_blocks, "6").running_var.clone()
    if rank == 0:
        mods_one = conv_modules()
        one = SemiTrainEngine(*mods_one, batch_size=B, use_graph=True)
        for _ in range(2):
            one.step(cir, err, label, supervised=True, update=False)
        torch.cuda.synchronize()
        worst = ("", 0.0)
        for k, v in one.named_grads().items():
            n1 = float(v.norm())
            if n1 == 0.0 or orc.grad_is_structurally_zero(k):
                continue
            rel = float((g_dp[k] - v.cpu()).norm()) / n1
            if rel > worst[1]:
                worst = (k, rel)
            if rel > parity.RTOL_FP32:
                failures.append(f"conv heads {k}: rel-L2 {rel:.2e} > {parity.RTOL_FP32}")
        bn_one = getattr(mods_one[2].restorer.conv_blocks, "6").running_var
        if not torch.allclose(bn_dp, bn_one, rtol=1e-5, atol=1e-7):
            failures.append("conv heads: BatchNorm running_var differs between 2 ranks (SyncBN) and 1 rank")
        print(f"[dp] world={world} Conv1d heads (SyncBN, global-index Philox dropout): worst gradient rel-L2 vs 1 rank x {B}: "
              f"{worst[1]:.2e} ({worst[0]})", flush=True)
    ok = torch.tensor([len(failures)], device="cuda")
    dist.broadcast(ok, src=0)
    if rank == 0:
        for f in failures:
            print("FAIL:", f, flush=True)
        print("DP_CHECK", "FAILED" if failures else "PASSED", flush=True)
    shutdown_distributed(engines)
    sys.exit(1 if int(ok) else 0)


if __name__ == "__main__":
    main()
