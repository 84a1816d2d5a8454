"""GPU parity tests (run with -m gpu on the B200 box): the sm_100a library, called through the C ABI
behind the drop-in modules and the fused engine, against (a) the golden fixtures recorded from the live
reference and (b) the CPU oracle on fresh seeded inputs up to BASELINE's batch 4096.

Tolerances (tests/parity.py): forward tensors rtol 1e-4 / atol 2e-5; gradients per tensor
||got-ref||_2 <= 1e-4*||ref||_2 (+ a floor of 1e-7*G*sqrt(n), G = largest gradient entry of the step);
north_star: "per-step loss and gradients within 1e-4 relative in fp32", "identical argmax",
"RMSE within 1e-3".
"""
import os
import re

import numpy as np
import pytest
import torch

from oracle import iins_oracle as orc
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _default_mode():
    import iins_vae_b200
    iins_vae_b200.set_compute_mode("fp32")
    yield
    iins_vae_b200.set_compute_mode("fp32")


def _mods(cfg, seed, dev="cuda"):
    from iins_vae_b200 import models as M
    pe, pd, pr, pc = orc.init_all(cfg, seed)
    Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
    Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
    Res = M.Restorer((cfg.range_dim, cfg.code_len))
    Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
    for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
        m.load_state_dict(p)
        m.to(dev)
    return (Enc, Dec, Res, Cls), (pe, pd, pr, pc)


def _named_grads(mods):
    out = {}
    for pre, m in zip(("enc", "dec", "res", "cls"), mods):
        for n, p in m.named_parameters():
            out[f"{pre}.{n}"] = p.grad
    return out


def _cases(golden, kind):
    pat = re.compile(rf"^{kind}\.(s\d+\.b\d+(?:\.m\d)?)\.meta$")
    return sorted(m.group(1) for m in (pat.match(f) for f in golden.files) if m)


def test_native_library_is_the_one_loaded():
    from iins_vae_b200._capi import get_lib, LIB_PATH, EXPORTS
    lib = get_lib()
    assert os.path.samefile(lib.path, LIB_PATH)
    for sym in EXPORTS:
        assert hasattr(lib.dll, sym), sym
    with pytest.raises(RuntimeError):
        from iins_vae_b200 import models as M
        M.Restorer((2, 8)).cuda()(torch.zeros(4, 2, 8))           # CPU tensor -> loud failure, no fallback


def test_modules_autograd_match_golden(golden):
    """The reference's own loop shape: torch criteria on the modules' outputs, loss.backward()."""
    cfg = orc.PathConfig()
    summary = []
    for case in _cases(golden, "semi"):
        pre = f"semi.{case}."
        seed, batch, sup, noise_seed = (int(v) for v in golden[pre + "meta"])
        mods, _ = _mods(cfg, seed)
        Enc, Dec, Res, Cls = mods
        cir = torch.from_numpy(golden[pre + "cir"]).cuda()
        err = torch.from_numpy(golden[pre + "err"]).cuda()
        label = torch.from_numpy(golden[pre + "label"]).cuda()
        torch.manual_seed(noise_seed)
        noise = torch.randn(batch, cfg.env_dim // 2, 1).cuda()
        rc, cat, lat, kl = Enc(cir, noise=noise)
        gen = Dec(rc, cat)
        err_fake, label_fake = Res(rc), Cls(cat)
        loss_ae = torch.nn.L1Loss()(cir, gen)
        loss = loss_ae + kl
        if sup:
            loss_res = 10 * torch.nn.L1Loss()(err, err_fake)
            loss_env = torch.nn.CrossEntropyLoss()(label_fake, label.to(torch.int64).squeeze())
            loss = loss + loss_res + loss_env
        loss.backward()
        outs = dict(range_code=rc, env_code=cat, env_code_rv=lat, kl=kl, cir_gen=gen, err_fake=err_fake,
                    label_fake=label_fake, loss_ae=loss_ae, loss=loss)
        for k, v in outs.items():
            parity.assert_out_close(f"{case}:{k}", v, golden[pre + "out." + k])
        none = set(golden[pre + "grad_none"].tolist())
        gscale = float(golden[pre + "grad_scale"])
        n_strict_fail, worst = 0, 0.0
        for name, g in _named_grads(mods).items():
            if name in none:
                assert g is None, f"{case}: {name} must have no gradient"
                continue
            assert g is not None, f"{case}: {name} missing gradient"
            if orc.grad_is_structurally_zero(name):
                assert float(g.abs().max()) <= parity.ZERO_G * gscale, name
                continue
            rel = parity.digest_rel_error(golden, pre + "grad." + name, g)
            # the fixture holds the reference's fp32 gradient; its own error vs fp64 is stored beside it
            tol = max(parity.RTOL_FP32, (1 + parity.REF_FACTOR) * float(golden[pre + "grad." + name + "|referr"]))
            worst = max(worst, rel)
            if rel > tol:
                n_strict_fail += 1
                assert rel <= parity.FLIP_C / batch, f"{case}: {name} rel error {rel:.2e} beyond the kink-flip bound"
        summary.append((case, batch, n_strict_fail, worst))
        pred = torch.argmax(label_fake, dim=1).cpu().numpy()
        assert np.array_equal(pred, golden[pre + "pred"]), f"{case}: argmax predictions differ"
        rmse = float(torch.mean((err_fake.detach() - err) ** 2) ** 0.5)
        assert abs(rmse - golden[pre + "metrics"][0]) < 1e-3
    for row in summary:
        print("golden case %s (B=%d): %d tensors beyond the strict bound, worst rel-L2 %.2e" % row)
    # The fixtures hold the reference's own fp32 gradients.  A ReLU kink decided differently by the reference's fp32 arithmetic
    # and by the CUDA path moves every tensor upstream of the kink by O(1/B), so the B = 64 fixtures (recorded before kink-free
    # seeds were selected) may legitimately show MANY tensors in the band; each of them is bounded by FLIP_C / B above, B <= 16
    # fixtures must be strict, and the capped-band check against exact arithmetic is test_engine_step_matches_oracle.
    for case, batch, n_fail, worst in summary:
        if batch <= parity.STRICT_BATCH:
            assert n_fail == 0 or batch <= 4, f"{case}: {n_fail} tensors beyond the strict bound at B={batch}"
    small = [r for r in summary if r[1] <= 4]
    clean = [r for r in small if r[2] == 0]
    assert len(clean) * 2 >= len(small), f"fewer than half of the B<=4 golden cases are kink-free on this device: {summary}"


# Mid-size batches run on seeds for which the step has no ReLU / LeakyReLU / sign() input within rounding distance of
# zero (tools/kink_seed_scan.py, run on the B200: profiles/r02_kink_seed_scan.log -- about half of all seeds qualify, for the
# fp32 CPU reference just as for both CUDA modes).  One flipped kink moves EVERY gradient tensor upstream of it by O(1/B):
# at B = 64..130 that is 1-2 %, which would hide a real error of that size; on kink-free inputs every tensor has to meet the
# strict bound and the kink-flip band (capped at parity.MAX_FLIP_TENSORS tensors) stays empty.
KINK_FREE_K = {64: 1, 130: 2}


@pytest.mark.parametrize("batch,supervised,graph,mode", [
    (64, True, False, "fp32"), (64, False, False, "fp32"), (130, True, False, "fp32"), (130, False, False, "fp32"), (1, True, False, "fp32"), (2, True, False, "fp32"),
    (4096, True, True, "fp32"), (4096, False, True, "fp32"), (8192, True, True, "fp32"),
    (2, True, False, "simt"), (64, True, False, "simt"), (130, True, False, "simt"), (4096, True, True, "simt")])
def test_engine_step_matches_oracle(batch, supervised, graph, mode):
    """Fused engine (fused loss, flat gradient buffer) vs the CPU oracle, up to BASELINE's batch 4096, for the
    tensor-core fp32-grade path ("fp32": tcgen05, bf16x3 split) and the SIMT fp32 cross-check path."""
    _engine_step_vs_oracle(orc.PathConfig(), KINK_FREE_K.get(batch, 0), batch, supervised, graph, mode)


# BASELINE configs[0] names "reference default options from utils.py": utils.py:38 --filters 16, i.e. dim = 16 (256-channel
# trunk, 3,993,291 parameters): the same per-tensor bound, on kink-free seeds at B = 64 (scan: profiles/r02_kink_seed_scan_dim16.log).
KINK_FREE_K_DIM16 = {64: 1}


@pytest.mark.parametrize("batch,supervised,graph", [(64, True, False), (64, False, False), (4096, True, True)])
def test_dim16_engine_step_matches_oracle(batch, supervised, graph):
    _engine_step_vs_oracle(orc.PathConfig(dim=16), KINK_FREE_K_DIM16.get(batch, 0), batch, supervised, graph, "fp32")


def _engine_step_vs_oracle(cfg, k, batch, supervised, graph, mode):
    import iins_vae_b200
    from iins_vae_b200.engine import SemiTrainEngine
    iins_vae_b200.set_compute_mode(mode)
    seed = 11 + batch + 1000 * k
    mods, pdicts = _mods(cfg, seed)
    cir, err, label = orc.synthetic_batch(cfg, batch, 500 + batch + 1000 * k)
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=graph)
    eng.step(cir, err, label, supervised=supervised, update=False)
    if graph:                                   # a second replay (compared with the eager pass in test_graph_replay_*)
        eng.step(cir, err, label, supervised=supervised, update=False)
    torch.cuda.synchronize()
    ref, ref_grads = orc.semi_step_with_grads(*pdicts, cir, err, label, cfg, supervised,
                                              torch.zeros(batch, cfg.env_dim // 2, 1))
    terms = eng.loss_terms()
    np.testing.assert_allclose(terms["loss"], float(ref["loss"]), rtol=1e-4)
    np.testing.assert_allclose(terms["loss_ae"], float(ref["loss_ae"]), rtol=1e-4)
    np.testing.assert_allclose(terms["loss_range"], float(ref["loss_range"]), rtol=1e-4)
    parity.assert_out_close("range_code", eng.rc, ref["range_code"])
    parity.assert_out_close("env_code", eng.cat, ref["env_code"].view(batch, -1))
    parity.assert_out_close("cir_gen", eng.xrec, ref["cir_gen"].view(batch, -1))
    if supervised:
        np.testing.assert_allclose(terms["loss_res"], float(ref["loss_res"]), rtol=1e-4)
        np.testing.assert_allclose(terms["loss_env"], float(ref["loss_env"]), rtol=1e-4)
        rmse, mae, acc, pred = orc.batch_metrics(ref["err_fake"], err, ref["label_fake"], label)
        assert abs(terms["rmse"] - float(rmse)) < 1e-3
        # argmax must be identical wherever the top-2 logit gap is above fp32 noise
        top2 = ref["label_fake"].topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-5
        assert np.array_equal(eng.pred.cpu().numpy()[safe.numpy()], pred.numpy()[safe.numpy()])
    gscale = max(float(g.abs().max()) for g in ref_grads.values() if g is not None)
    got = eng.named_grads()
    for name, g in ref_grads.items():
        if g is None:
            assert float(got[name].abs().max()) == 0.0, f"{name}: reference has no grad, engine wrote one"
    dbl = lambda d: {k: v.double() for k, v in d.items()}
    _, truth = orc.semi_step_with_grads(*(dbl(p) for p in pdicts), cir.double(), err.double(), label.double(), cfg,
                                        supervised, torch.zeros(batch, cfg.env_dim // 2, 1).double())
    rows = parity.grad_report(got, truth, ref_grads, gscale,
                              parity.REF_FACTOR_TC if mode == "fp32" else parity.REF_FACTOR)
    n_flip = parity.assert_grads(rows, batch, label=f"B={batch}")       # strict for B <= 16, at most 3 tensors in the band
    rel = sorted(r[1] for r in rows if not orc.grad_is_structurally_zero(r[0]))
    ref = sorted(r[2] for r in rows if not orc.grad_is_structurally_zero(r[0]))
    print(f"[{mode}] B={batch} sup={supervised}: gradient rel-L2 error vs fp64 oracle: median {rel[len(rel) // 2]:.2e} max {rel[-1]:.2e}"
          f" | fp32 CPU oracle: median {ref[len(ref) // 2]:.2e} max {ref[-1]:.2e} | {n_flip} tensors in the kink-flip band")


@pytest.mark.parametrize("case", ["s0.b4", "s1.b64"])
def test_engine_adam_trajectory_matches_golden(golden, case):
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig()
    pre = f"traj.{case}."
    seed, batch, n_steps = (int(v) for v in golden[pre + "meta"])
    mods, pdicts = _mods(cfg, seed)
    start = {f"{g}.{k}": v.clone() for g, d in zip(("enc", "dec", "res", "cls"), pdicts) for k, v in d.items()}
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, lr=1e-4, betas=(0.5, 0.999), use_graph=True)
    batches = [orc.synthetic_batch(cfg, batch, seed + 2000 + j) for j in range(3)]
    masks = golden[pre + "masks"]
    for step in range(n_steps):
        cir, err, label = batches[step % 3]
        eng.step(cir, err, label, supervised=bool(masks[step]))
        loss = eng.loss_terms()["loss"]
        np.testing.assert_allclose(loss, golden[pre + "losses"][step], rtol=2e-5 if step == 0 else 1e-2)
        if step + 1 in (1, n_steps):
            for gname, m in zip(("enc", "dec", "res", "cls"), mods):
                for n, p in m.named_parameters():
                    name = f"{gname}.{n}"
                    key = f"{pre}step{step + 1}.param.{name}"
                    got = p.detach().cpu().double().numpy().ravel()
                    s0 = start[name].double().numpy().ravel()
                    if orc.grad_is_structurally_zero(name):
                        assert np.abs(got - s0).max() <= (step + 1) * 1.01e-4
                        continue
                    if "linear_layer2" in name:
                        assert np.array_equal(got, s0), "restorer.linear_layer2 must never be updated"
                        continue
                    if key + "|full" in golden.files:
                        ref, g2, s2 = golden[key + "|full"].ravel(), got, s0
                    else:
                        pos = parity.sample_positions(got.size)
                        ref, g2, s2 = golden[key + "|samples"], got[pos], s0[pos]
                    if step == 0:
                        parity.assert_traj_close(name, g2, ref, 1)
                    else:
                        # B=4 trajectories are chaotic (one kink flip is 25% of the batch): loose bound there
                        parity.assert_update_close(name, g2, ref, s2, step + 1, rel=1.0 if batch <= 4 else 0.5)


def test_supervised_engine_matches_golden(golden):
    """train.py:82-94: CE + L1 on Encoder -> (Classifier, Restorer), NC=2."""
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig(num_classes=2)
    pre = "sup.s0.b64."
    seed, batch = (int(v) for v in golden[pre + "meta"])
    mods, pdicts = _mods(cfg, seed)
    cir, err, label = orc.synthetic_batch(cfg, batch, seed + 1000)
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, mode="supervised", use_graph=False)
    eng.step(cir, err, label, update=False)
    t = eng.loss_terms()
    np.testing.assert_allclose(t["loss"], float(golden[pre + "out.loss"]), rtol=1e-4)
    np.testing.assert_allclose(t["loss_env"], float(golden[pre + "out.loss_idy"]), rtol=1e-4)
    np.testing.assert_allclose(t["loss_res"], float(golden[pre + "out.loss_reg"]), rtol=1e-4)
    parity.assert_out_close("label_est", eng.logits, golden[pre + "out.label_est"])
    parity.assert_out_close("err_est", eng.err_est, golden[pre + "out.err_est"])
    gscale = 1.0
    none = set(golden[pre + "grad_none"].tolist())
    for name, g in eng.named_grads().items():
        if name in none:
            continue
        if orc.grad_is_structurally_zero(name):
            continue
        rel = parity.digest_rel_error(golden, pre + "grad." + name, g)
        assert rel <= max(parity.RTOL_FP32, parity.FLIP_C / batch), f"{name}: rel error {rel:.2e}"


def test_emnet_module_and_ewine_length():
    """EMNet composite (run.py:59-62 contract) under autograd, with the 152-tap ewine CIR length."""
    from iins_vae_b200 import models as M
    cfg = orc.PathConfig(cir_len=152, num_classes=2)
    pe, pd, pr, pc = orc.init_all(cfg, 3)
    net = M.EMNet(cir_len=152, num_classes=2, env_dim=16).cuda()
    net.encoder.load_state_dict(pe); net.classifier.load_state_dict(pc); net.restorer.load_state_dict(pr)
    cir, err, label = orc.synthetic_batch(cfg, 37, 9)
    label_est, env_latent, err_est = net(cir.cuda())
    ref = orc.supervised_forward(pe, pr, pc, cir, err, label, cfg, torch.zeros(37, 8, 1))
    parity.assert_out_close("label_est", label_est, ref["label_est"])
    parity.assert_out_close("err_est", err_est, ref["err_est"])
    assert env_latent.shape == (37, 16, 1)
    loss = torch.nn.CrossEntropyLoss()(label_est, label.cuda().long().squeeze()) + torch.nn.L1Loss()(err_est, err.cuda())
    loss.backward()
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=1e-4)
    assert net.restorer.restorer.linear_layer2.weight.grad is None


def test_inference_engine_matches_oracle():
    """test.py:66-85: identical argmax, RMSE within 1e-3 (north_star)."""
    from iins_vae_b200.engine import InferenceEngine
    cfg = orc.PathConfig()
    batch = 1000
    mods, (pe, pd, pr, pc) = _mods(cfg, 21)
    cir, err, label = orc.synthetic_batch(cfg, batch, 77)
    eng = InferenceEngine(mods[0], mods[2], mods[3], batch_size=batch, cir_len=cfg.cir_len)
    for _ in range(2):
        err_est, pred, out = eng.run(cir, err, label)
    torch.cuda.synchronize()
    logits, _, ref_err = orc.emnet(pe, pr, pc, cir, cfg, torch.zeros(batch, 8, 1))
    rmse, mae, acc, ref_pred = orc.batch_metrics(ref_err, err, logits, label)
    top2 = logits.topk(2, dim=1).values
    safe = ((top2[:, 0] - top2[:, 1]) > 1e-5).numpy()
    assert np.array_equal(pred.cpu().numpy()[safe], ref_pred.numpy()[safe])
    o = out.tolist()
    assert abs(o[4] ** 0.5 - float(rmse)) < 1e-3
    assert abs(o[1] - float(mae)) < 1e-3
    assert abs(o[5] / batch - float(acc)) < 2e-3


def test_philox_latent_noise_distribution():
    from iins_vae_b200 import models as M
    cfg = orc.PathConfig()
    pe, _, _, _ = orc.init_all(cfg, 0)
    keys = [k for k in pe if k.startswith("env_encoder")]
    pe[keys[-2]].zero_(); pe[keys[-1]].zero_()                  # mu = log_sigma = 0 -> latent == noise
    Enc = M.Encoder(1, 4, 3, 4, 16, 2, noise="philox", seed=5).cuda()
    Enc.load_state_dict(pe)
    x = torch.randn(4096, 157, device="cuda")
    z1 = Enc(x)[2].detach().flatten().cpu().numpy()
    z2 = Enc(x)[2].detach().flatten().cpu().numpy()
    assert abs(z1.mean()) < 0.03 and abs(z1.std() - 1) < 0.03
    assert not np.array_equal(z1, z2), "offset must advance between calls"
    assert abs(np.corrcoef(z1, z2)[0, 1]) < 0.03


def test_bf16_mode_is_close_and_stated_tolerance():
    """BASELINE configs[2]: bf16 tensor-core mode (plain bf16 operands, fp32 accumulate/statistics).  Stated
    tolerance: loss within 2e-2 relative, forward tensors within 5e-2 of their scale, argmax identical where the
    top-2 logit gap exceeds 1e-2."""
    import iins_vae_b200
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig()
    batch = 512
    mods, pdicts = _mods(cfg, 5)
    cir, err, label = orc.synthetic_batch(cfg, batch, 55)
    iins_vae_b200.set_compute_mode("bf16")
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=False)
    eng.step(cir, err, label, supervised=True, update=False)
    torch.cuda.synchronize()
    ref = orc.semi_forward(*pdicts, cir, err, label, cfg, True, torch.zeros(batch, cfg.env_dim // 2, 1))
    t = eng.loss_terms()
    np.testing.assert_allclose(t["loss"], float(ref["loss"]), rtol=2e-2)
    for name, got, want in (("range_code", eng.rc, ref["range_code"]), ("cir_gen", eng.xrec, ref["cir_gen"].view(batch, -1)),
                            ("label_fake", eng.logits, ref["label_fake"])):
        scale = float(want.abs().max())
        assert float((got.cpu() - want).abs().max()) <= 5e-2 * scale, name
    top2 = ref["label_fake"].topk(2, dim=1).values
    safe = ((top2[:, 0] - top2[:, 1]) > 1e-2).numpy()
    assert np.array_equal(eng.pred.cpu().numpy()[safe], ref["label_fake"].argmax(1).numpy()[safe])
    g = eng.named_grads()["dec.decoder.mlp.model.4.weight"]
    assert torch.isfinite(g).all() and float(g.abs().max()) > 0


def test_shape_option_sweep_matches_oracle():
    """One train step for the other shape options of the path (dim 1 / 2, n_residual 0 / 1, num_classes 2 / 10, env_dim 8,
    range_dim 4, mixed) at a ragged and a multi-tile batch, against the CPU oracle; dim=3 must be refused, not mis-computed."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("config_sweep", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "config_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run_sweep() == 0


@pytest.mark.parametrize("supervised", [True, False])
def test_graph_replay_matches_eager_and_is_reproducible(supervised):
    """The CUDA-graph replay of a step must produce what the eager pass of the same engine produces, and five replays
    must agree with each other: the only run-to-run freedom is the order of the fp32 atomics that flush the weight
    gradients (a gradient tensor is never read again inside the step, so the noise cannot propagate).  Stated bounds:
    forward tensors bit-identical, loss sums within 1e-6 / 1e-5; every gradient tensor within 1e-5 rel-L2 of the eager
    pass and of the first replay (measured: 1e-6 .. 3e-6)."""
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig()
    batch = 4096
    mods, _ = _mods(cfg, 23)
    cir, err, label = orc.synthetic_batch(cfg, batch, 623)
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=False)
    eng.step(cir, err, label, supervised=supervised, update=False)
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in eng.named_grads().items()}
    eager_out = (eng.out.clone(), eng.kl.clone(), eng.xrec.clone(), eng.rc.clone(), eng.cat.clone())
    eng.use_graph = True
    runs = []
    for _ in range(5):
        eng.step(cir, err, label, supervised=supervised, update=False)
        torch.cuda.synchronize()
        runs.append({k: v.clone() for k, v in eng.named_grads().items()})
        for a, b in zip(eager_out[2:], (eng.xrec, eng.rc, eng.cat)):
            assert torch.equal(a, b), "forward tensors of a replay differ from the eager pass"
        assert torch.allclose(eager_out[1], eng.kl, rtol=1e-5, atol=0), "KL (a sum of float atomics) differs beyond 1e-5"
        assert torch.allclose(eager_out[0], eng.out, rtol=1e-6, atol=0), "loss terms (atomic sums) differ beyond 1e-6"
    worst_eager = worst_replay = 0.0
    for k, e in eager.items():
        n = float(e.norm())
        if n == 0.0:
            assert all(float(r[k].abs().max()) == 0.0 for r in runs), k
            continue
        worst_eager = max(worst_eager, max(float((r[k] - e).norm()) / n for r in runs))
        worst_replay = max(worst_replay, max(float((r[k] - runs[0][k]).norm()) / n for r in runs[1:]))
    print(f"sup={supervised}: replay vs eager worst rel-L2 {worst_eager:.2e}; replay vs replay {worst_replay:.2e}")
    assert worst_eager <= 1e-5 and worst_replay <= 1e-5


# BASELINE configs[2]: bf16 operands (8 mantissa bits: unit round-off 2^-9 = 2e-3), fp32 accumulation and statistics.
# This network is ill-conditioned (InstanceNorm over L=8 amplifies operand noise by up to 1/sqrt(eps) = 316, DESIGN.md
# section 4), so ANY bf16-operand evaluation of the reference is far from the fp64 result on the decoder / range-encoder
# gradients (10-40 % rel-L2, measured with the oracle's operand_rounding("bf16") emulation).  Stated tolerance, per
# gradient tensor, rel-L2 against the fp64 oracle:
#     err(CUDA bf16 mode) <= max(BF16_FLOOR, BF16_FACTOR x err(reference restated with bf16-rounded GEMM operands))
# i.e. the CUDA path may not be further from exact arithmetic than a small multiple of what bf16 operands cost the
# reference itself; forward tensors: the same rule on the max-abs error relative to the tensor's scale.
BF16_FACTOR = 3.0
BF16_FLOOR = 2e-2
BF16_FWD_FLOOR = 1e-2


@pytest.mark.parametrize("batch", [8192])
def test_bf16_mode_full_gradient_parity(batch):
    """configs[2] (semi-supervised step, bf16, B=8192), both mask branches, graph replay on: loss within 1e-2, forward
    tensors and EVERY parameter-gradient tensor within the stated bf16 tolerance (above) of the fp64 oracle."""
    import iins_vae_b200
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig()
    mods, pdicts = _mods(cfg, 31)
    cir, err, label = orc.synthetic_batch(cfg, batch, 831)
    iins_vae_b200.set_compute_mode("bf16")
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=True)
    dbl = lambda d: {k: v.double() for k, v in d.items()}
    zero = torch.zeros(batch, cfg.env_dim // 2, 1)
    for supervised in (True, False):
        eng.step(cir, err, label, supervised=supervised, update=False)
        torch.cuda.synchronize()
        ref, truth = orc.semi_step_with_grads(*(dbl(p) for p in pdicts), cir.double(), err.double(), label.double(), cfg,
                                              supervised, zero.double())
        with orc.operand_rounding("bf16"):
            emu, emu_g = orc.semi_step_with_grads(*pdicts, cir, err, label, cfg, supervised, zero)
        t = eng.loss_terms()
        np.testing.assert_allclose(t["loss"], float(ref["loss"]), rtol=1e-2)
        for name, got, key in (("range_code", eng.rc, "range_code"), ("env_code", eng.cat, "env_code"), ("cir_gen", eng.xrec, "cir_gen")):
            want = ref[key].reshape(got.shape)
            scale = float(want.abs().max())
            e_got = float((got.cpu().double() - want).abs().max()) / scale
            e_emu = float((emu[key].reshape(got.shape).double() - want).abs().max()) / scale
            assert e_got <= max(BF16_FWD_FLOOR, BF16_FACTOR * e_emu), f"{name}: {e_got:.2e} of scale vs bf16-operand reference {e_emu:.2e}"
        got = eng.named_grads()
        rows = []
        for name, g64 in truth.items():
            if g64 is None:
                assert float(got[name].abs().max()) == 0.0, name
                continue
            if orc.grad_is_structurally_zero(name):
                continue
            n = float(g64.norm()) + 1e-300
            rows.append((name, float((got[name].cpu().double() - g64).norm()) / n, float((emu_g[name].double() - g64).norm()) / n))
        ratio = sorted(r[1] / max(r[2], 1e-12) for r in rows)
        worst = max(rows, key=lambda r: r[1])
        print(f"[bf16] B={batch} sup={supervised}: gradient rel-L2 vs fp64: worst {worst[1]:.2e} ({worst[0]}; bf16-operand reference "
              f"{worst[2]:.2e}); CUDA/reference error ratio median {ratio[len(ratio) // 2]:.2f} max {ratio[-1]:.2f}")
        top = max(rows, key=lambda r: r[1] / max(r[2], 1e-12))
        print(f"       largest ratio: {top[0]} CUDA {top[1]:.2e} vs bf16-operand reference {top[2]:.2e} (floor {BF16_FLOOR:.0e})")
        for name, e_got, e_emu in rows:
            assert e_got <= max(BF16_FLOOR, BF16_FACTOR * e_emu), (f"{name}: rel-L2 {e_got:.2e} vs fp64, beyond {BF16_FACTOR}x the "
                                                                  f"bf16-operand reference's own {e_emu:.2e}")


def test_one_based_labels_match_reference_shift():
    """train_semi.py:217-222: with any dataset_env but 'room_full' the labels are 1..NC and the reference feeds
    CrossEntropyLoss `label - 1`: the engine with label_offset=1 on 1-based labels must equal label_offset=0 on the
    shifted labels, and raw 1-based labels without the offset must fail loudly instead of reading out of the row."""
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig(num_classes=4)
    batch = 96
    mods, pdicts = _mods(cfg, 41)
    cir, err, label = orc.synthetic_batch(cfg, batch, 141)            # labels in [0, NC)
    e0 = SemiTrainEngine(*mods, batch_size=batch, use_graph=False)
    e0.step(cir, err, label, supervised=True, update=False)
    t0, g0 = e0.loss_terms(), {k: v.clone() for k, v in e0.named_grads().items()}
    e1 = SemiTrainEngine(*mods, batch_size=batch, use_graph=False, shared_state=e0, label_offset=1)
    e1.step(cir, err, label + 1, supervised=True, update=False)
    t1 = e1.loss_terms()
    np.testing.assert_allclose(t0["loss_env"], t1["loss_env"], rtol=1e-6)
    assert t0["accuracy"] == t1["accuracy"]
    for k, v in e1.named_grads().items():
        if k.startswith("cls."):
            assert torch.allclose(v, g0[k], rtol=1e-5, atol=1e-9), k
    ref = orc.semi_forward(*pdicts, cir, err, label, cfg, True, torch.zeros(batch, cfg.env_dim // 2, 1))
    np.testing.assert_allclose(t1["loss_env"], float(ref["loss_env"]), rtol=1e-4)
    e0.step(cir, err, label + 1, supervised=True, update=False)       # 1-based labels, no offset: label == NC is invalid
    with pytest.raises(ValueError):
        e0.loss_terms()


@pytest.mark.parametrize("kind", ["res", "cls"])
def test_conv1d_heads_match_reference_fixtures(kind):
    """Restorer / Classifier with net_type='Conv1d' (models.py:661-716, :865-902) as drop-in modules under
    autograd on the B200, against fixtures recorded from the LIVE reference with its own dropout masks replayed (train
    mode: outputs, parameter gradients, input gradient, BatchNorm buffers after the step; eval mode: running statistics),
    plus the in-kernel Philox dropout: keep rate 0.75, the same mask in forward and backward, a new mask per call."""
    from iins_vae_b200 import models as M
    from tests.golden.make_golden_common import conv_head_case_inputs
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "iins_golden_convheads.npz"))
    cfg = orc.PathConfig()
    pat = re.compile(rf"^{kind}\.s(\d+)\.b(\d+)\.t(\d)\.meta$")
    cases = sorted((int(m.group(1)), int(m.group(2)), int(m.group(3))) for m in (pat.match(f) for f in golden.files) if m)
    assert len(cases) == 4
    for seed, batch, training in cases:
        pre = f"{kind}.s{seed}.b{batch}.t{training}."
        x, p, _ = conv_head_case_inputs(kind, seed, batch, cfg)
        mod = (M.Restorer((cfg.range_dim, cfg.code_len), net_type="Conv1d") if kind == "res"
               else M.Classifier(cfg.env_dim, cfg.num_classes, filters=16, net_type="Conv1d"))
        mod.load_state_dict(p)
        mod.cuda().train(bool(training))
        masks = [torch.from_numpy(golden[pre + f"mask{i}"]).cuda() for i in range(2)] if training else None
        xin = x.cuda().requires_grad_(True)
        out = mod(xin, masks=masks)
        (out * torch.from_numpy(golden[pre + "d_out"]).cuda()).sum().backward()
        np.testing.assert_allclose(out.detach().cpu().numpy(), golden[pre + "out"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(xin.grad.cpu().numpy(), golden[pre + "d_x"], rtol=2e-4, atol=2e-7)
        for k, v in mod.named_parameters():
            ref = golden[pre + "grad." + k]
            if ref.size == 0:
                assert v.grad is None, k
                continue
            err = np.linalg.norm(v.grad.cpu().numpy().ravel() - ref.ravel())
            assert err <= 2e-4 * np.linalg.norm(ref) + 1e-8, (pre, k, err)
        if training:
            sd = mod.state_dict()
            for k in sd:
                if "running" in k:
                    np.testing.assert_allclose(sd[k].cpu().numpy(), golden[pre + "buf." + k], rtol=1e-5, atol=1e-6)
                if "num_batches" in k:
                    assert int(sd[k]) == 1
    # Philox dropout (no explicit masks): zero the BatchNorm'd path's sensitivity by looking at the conv-block output directly
    mod.train(True)
    x = torch.ones(4096, *x.shape[1:], device="cuda")
    o1, o2 = mod(x).detach(), mod(x).detach()
    assert not torch.equal(o1, o2), "a new dropout mask per call"
    mod.eval()
    assert torch.equal(mod(x), mod(x)), "eval mode is deterministic"


def _conv_mods(cfg, seed):
    from iins_vae_b200 import models as M
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
    return (Enc, Dec, Res, Cls), (pe, pd, pr, pc)


def _conv_head_masks(eng, m, B):
    """The Philox keep-masks the kernels drew, recovered from the saved activations (dropout output vs input) and returned
    in the reference's (B, C, L) layout."""
    C1, L1, C2, L2 = (16, 4, 32, 2) if m == "restorer" else (eng.cfg.cls_filters, 1, eng.cfg.cls_filters, 1)
    ws = eng.ws[m]
    n1, n2 = B * L1 * C1, B * L2 * C2
    r4 = lambda n: (n + 3) // 4 * 4
    a1, d1 = ws[:n1], ws[r4(n1):r4(n1) + n1]
    o = 2 * r4(n1)
    a2, d2 = ws[o:o + n2], ws[o + r4(n2):o + r4(n2) + n2]
    m1 = torch.where(a1 != 0, (d1 != 0).float(), torch.ones_like(a1)).view(B, L1, C1).permute(0, 2, 1).contiguous().cpu()
    m2 = torch.where(a2 != 0, (d2 != 0).float(), torch.ones_like(a2)).view(B, L2, C2).permute(0, 2, 1).contiguous().cpu()
    return m1, m2


# (weights, batch) seeds of the test below: a pair with no ReLU input within rounding of zero, so that no gradient tensor needs the
# kink-flip band whatever the summation order of the kernels (picked with tools/heads_seed_scan.py; a flip is a property of the
# data, not of the kernels: at the previous pair (61, 161) ANY 1e-7 perturbation of the input moved four tensors into the band)
CONV_HEADS_SEEDS = (64, 164)


def test_engine_with_conv1d_heads_matches_oracle():
    """regressor_type / identifier_type = 2 (utils.py:43-44) through the fused engine: the semi-supervised step with
    RestorerConv1d + ClassifierConv1d (Philox dropout, BatchNorm over the batch) against the oracle replaying the masks the
    kernels drew; every gradient tensor to the fp32 bound, BatchNorm buffers updated like torch's, keep rate 0.75."""
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig()
    batch = 512
    mods, pdicts = _conv_mods(cfg, CONV_HEADS_SEEDS[0])
    cir, err, label = orc.synthetic_batch(cfg, batch, CONV_HEADS_SEEDS[1])
    eng = SemiTrainEngine(*mods, batch_size=batch, use_graph=False)
    eng.step(cir, err, label, supervised=True, update=False)
    torch.cuda.synchronize()
    masks_r, masks_c = _conv_head_masks(eng, "restorer", batch), _conv_head_masks(eng, "classifier", batch)
    keep = float(torch.cat([m.flatten() for m in masks_r + masks_c]).mean())
    assert abs(keep - 0.75) < 0.02, keep
    bufs = {}
    fns = dict(res=lambda p, rc: orc.restorer_conv1d(p, rc, masks_r, True)[0], cls=lambda p, cat: orc.classifier_conv1d(p, cat, masks_c, True)[0])
    zero = torch.zeros(batch, cfg.env_dim // 2, 1)
    ref, ref_grads = orc.semi_step_with_grads(*pdicts, cir, err, label, cfg, True, zero, head_fns=fns)
    dbl = lambda d: {k: (v.double() if v.is_floating_point() else v) for k, v in d.items()}
    _, truth = orc.semi_step_with_grads(*(dbl(p) for p in pdicts), cir.double(), err.double(), label.double(), cfg, True, zero.double(),
                                        head_fns=fns)
    t = eng.loss_terms()
    for k in ("loss", "loss_res", "loss_env"):
        np.testing.assert_allclose(t[k], float(ref[k]), rtol=1e-4)
    gscale = max(float(g.abs().max()) for g in ref_grads.values() if g is not None)
    got = eng.named_grads()
    for name, g in ref_grads.items():
        if g is None:
            assert float(got[name].abs().max()) == 0.0, name
    rows = parity.grad_report(got, truth, ref_grads, gscale, parity.REF_FACTOR_TC)
    n_band = parity.assert_grads(rows, batch, label="conv heads")
    worst = max(r[1] for r in rows if r[0].startswith(("res.", "cls.")))
    print(f"[conv1d heads] B={batch}: worst head-gradient rel-L2 vs fp64 oracle {worst:.2e}, {n_band} tensors in the kink band, keep rate {keep:.3f}")
    # BatchNorm buffers after ONE training pass == torch's momentum update on the oracle's batch statistics
    _, (rm, rv) = orc.restorer_conv1d(pdicts[2], ref["range_code"], masks_r, True)
    bn = getattr(mods[2].restorer.conv_blocks, "6")
    np.testing.assert_allclose(bn.running_mean.cpu().numpy(), rm.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bn.running_var.cpu().numpy(), rv.numpy(), rtol=1e-4, atol=1e-6)
    # a full optimisation step replayed from the graph draws NEW masks every step (device-side Philox offset)
    eng2 = SemiTrainEngine(*mods, batch_size=batch, use_graph=True, shared_state=eng)
    outs = []
    for _ in range(3):
        eng2.step(cir, err, label, supervised=True)
        outs.append(_conv_head_masks(eng2, "restorer", batch)[0].clone())
    assert not torch.equal(outs[1], outs[2])


def test_soft_restorer_module_matches_reference_fixture():
    """Restorer(soft=True) (models.py:634-655) as a drop-in module on the B200: the host np.random.normal draw of the
    reference (same call, same seed -> same noise), the (B, B)-broadcast output, gradients of linear_layer2 and the trunk,
    linear_layer1 without gradient."""
    from iins_vae_b200 import models as M
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "iins_golden_convheads.npz"))
    cfg = orc.PathConfig()
    for seed, batch in ((0, 5), (1, 48)):
        pre = f"soft.s{seed}.b{batch}."
        gen = torch.Generator().manual_seed(seed)
        p = orc.init_params(orc.restorer_param_shapes(cfg), gen)
        x = torch.rand(batch, cfg.range_dim, cfg.code_len, generator=gen).cuda().requires_grad_(True)
        mod = M.Restorer((cfg.range_dim, cfg.code_len), soft=True)
        mod.load_state_dict(p)
        mod.cuda()
        np.random.seed(seed)
        out = mod(x)
        assert out.shape == (batch, batch)
        (out * torch.from_numpy(golden[pre + "d_out"]).cuda()).sum().backward()
        np.testing.assert_allclose(out.detach().cpu().numpy(), golden[pre + "out"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(x.grad.cpu().numpy(), golden[pre + "d_x"], rtol=2e-4, atol=1e-6)
        for k, v in mod.named_parameters():
            ref = golden[pre + "grad." + k]
            if ref.size == 0:
                assert v.grad is None, k
                continue
            err = np.linalg.norm(v.grad.cpu().numpy().ravel() - ref.ravel())
            assert err <= 2e-4 * np.linalg.norm(ref) + 1e-8, (pre, k, err)


@pytest.mark.parametrize("batch,mode", [(37, "fp32"), (4096, "fp32"), (4096, "bf16")])
def test_window_kernels_match_per_layer_kernels(batch, mode):
    """The persistent window kernels of the stride-2 convolutions (csrc/iins_win.cu: forward, parity-split data gradient,
    weight gradient) against the per-layer tensor-core kernels of the same library: two contexts in one process, one
    created with IINS_WIN=0 (the switches are read when a context is created).  Both use the same bf16 pieces, the same
    packed weights and the same k order, so the forward tensors are BIT-IDENTICAL and every gradient tensor agrees to
    summation-order noise: stated bound 1e-5 rel-L2 in fp32-grade mode (measured 4e-7 .. 4.4e-6; the trunk weight-gradient
    kernel drops two piece products of relative size 2^-24, and with the window kernels the range encoder's norm backward
    runs as its own kernel instead of in the per-layer data-gradient epilogue), 3e-5 in bf16 mode (a 1e-7 difference in a
    data gradient flips the bf16 rounding of a few operand elements of the next layer: measured 6e-6 .. 1.3e-5); ragged batch =
    partial last tile."""
    import iins_vae_b200
    from iins_vae_b200._capi import get_lib
    from iins_vae_b200.engine import SemiTrainEngine
    d = get_lib().dll
    cfg = orc.PathConfig()
    cir, err, label = orc.synthetic_batch(cfg, batch, 977)
    res = {}
    old = os.environ.get("IINS_WIN")
    try:
        for win in ("0", "7"):
            os.environ["IINS_WIN"] = win
            ctx = d.iins_ctx_create()
            assert ctx
            d.iins_ctx_make_current(ctx)
            iins_vae_b200.set_compute_mode(mode)
            mods, _ = _mods(cfg, 41)
            eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=False)
            launches = get_lib().profile(lambda: eng.step(cir, err, label, supervised=True, update=False))
            torch.cuda.synchronize()
            res[win] = ({k: v.clone() for k, v in eng.named_grads().items()}, eng.xrec.clone(), eng.rc.clone(), eng.cat.clone(),
                        sum(1 for n, _, _ in launches if "iins_win" in n))
            d.iins_ctx_make_current(None)
            d.iins_ctx_destroy(ctx)
    finally:
        if old is None:
            os.environ.pop("IINS_WIN", None)
        else:
            os.environ["IINS_WIN"] = old
    (g0, x0, r0, c0, n0), (g1, x1, r1, c1, n1) = res["0"], res["7"]
    assert n0 == 0 and n1 >= 10, f"window kernels launched: {n0} (IINS_WIN=0) / {n1} (default)"
    assert torch.equal(x0, x1) and torch.equal(r0, r1) and torch.equal(c0, c1), "forward tensors differ"
    worst = 0.0
    for k, e in g0.items():
        if orc.grad_is_structurally_zero(k):
            continue                    # conv bias in front of an InstanceNorm: the true gradient is 0, both values are rounding noise
        n = float(e.norm())
        if n > 0:
            worst = max(worst, float((g1[k] - e).norm()) / n)
    print(f"window vs per-layer kernels, B={batch} {mode}: {n1} window launches, worst gradient rel-L2 diff {worst:.2e}")
    assert worst <= (1e-5 if mode == "fp32" else 3e-5)


def test_dim16_bf16_mode_loss_is_close():
    """dim = 16 in bf16 mode: the one-piece operands of the 64-channel stride-2 layers fit the persistent window kernels' shared
    memory, whose producers hold at most 8 units per thread and tile -- that geometry (ca = 64) is refused by
    iins_win_nt_supported and runs on the per-layer kernels.  Stated tolerance: loss terms within 1e-2 relative of the fp32 oracle."""
    import iins_vae_b200
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig(dim=16)
    batch = 64
    mods, pdicts = _mods(cfg, 3)
    cir, err, label = orc.synthetic_batch(cfg, batch, 5)
    iins_vae_b200.set_compute_mode("bf16")
    eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=False)
    eng.step(cir, err, label, supervised=True, update=False)
    torch.cuda.synchronize()
    ref, _ = orc.semi_step_with_grads(*pdicts, cir, err, label, cfg, True, torch.zeros(batch, cfg.env_dim // 2, 1))
    got = eng.loss_terms()
    for k in ("loss", "loss_ae", "loss_res", "loss_env"):
        assert abs(got[k] - float(ref[k])) <= 1e-2 * abs(float(ref[k])) + 1e-6, (k, got[k], float(ref[k]))


@pytest.mark.parametrize("batch", [37, 256])
def test_paired_row_kernels_match_single_row_kernels(batch):
    """The one-thread-per-row kernels of the small-channel layers with TWO rows per thread (iins_row2_nt_kernel<.., R = 2>: one
    shared-memory weight read serves both rows; the default for layers with >= 65536 rows) against the same kernels with one row
    per thread: two contexts in one process, IINS_ROW_PAIR = 1 (pair every eligible layer, also at this small batch) and 0.
    The convolution sums are the same FMAs in the same order; the InstanceNorm / LayerNorm statistics add a thread's two rows
    before the warp sum, so tensors agree to fp32 summation-order noise: stated bound 1e-5 rel-L2 forward (measured 3e-6 on the
    range code, 2e-7 on the reconstruction), 2e-4 per gradient tensor (measured 3e-5; fp32-grade mode).  Larger batches are
    left to the oracle tests: a batch that holds ONE ReLU input within rounding of zero moves the range encoder's gradients by
    1e-2 under ANY 1e-7 perturbation, paired or not (tools/diag_row_pair.py shows both)."""
    import iins_vae_b200
    from iins_vae_b200._capi import get_lib
    from iins_vae_b200.engine import SemiTrainEngine
    d = get_lib().dll
    cfg = orc.PathConfig()
    cir, err, label = orc.synthetic_batch(cfg, batch, 977)
    res = {}
    old = os.environ.get("IINS_ROW_PAIR")
    try:
        for sw in ("0", "1"):
            os.environ["IINS_ROW_PAIR"] = sw
            ctx = d.iins_ctx_create()
            assert ctx
            d.iins_ctx_make_current(ctx)
            iins_vae_b200.set_compute_mode("fp32")
            mods, _ = _mods(cfg, 41)
            eng = SemiTrainEngine(*mods, batch_size=batch, cir_len=cfg.cir_len, use_graph=False)
            launches = get_lib().profile(lambda: eng.step(cir, err, label, supervised=True, update=False))
            torch.cuda.synchronize()
            res[sw] = ({k: v.clone() for k, v in eng.named_grads().items()}, eng.xrec.clone(), eng.rc.clone(),
                       [s for n, _, s in launches if "iins_row2_nt" in n])
            d.iins_ctx_make_current(None)
            d.iins_ctx_destroy(ctx)
    finally:
        if old is None:
            os.environ.pop("IINS_ROW_PAIR", None)
        else:
            os.environ["IINS_ROW_PAIR"] = old
    (g0, x0, r0, l0), (g1, x1, r1, l1) = res["0"], res["1"]
    assert len(l0) == len(l1) and len(l0) >= 20
    for a, b, name in ((x0, x1, "xrec"), (r0, r1, "rc")):
        assert float((a - b).norm()) <= 1e-5 * float(a.norm()), name
    worst = 0.0
    for k, e in g0.items():
        if orc.grad_is_structurally_zero(k):
            continue
        n = float(e.norm())
        if n > 0:
            worst = max(worst, float((g1[k] - e).norm()) / n)
    print(f"paired vs single row kernels, B={batch}: worst gradient rel-L2 diff {worst:.2e}")
    assert worst <= 2e-4
