"""GPU: the drop-in entry points (train_semi.run, train.train_gem, test.test_gem) drive the fused engines and
match the oracle's training dynamics."""
import os

import numpy as np
import pytest
import torch

from oracle import iins_oracle as orc

pytestmark = pytest.mark.gpu


def test_train_semi_entry_point_reduces_loss(tmp_path, monkeypatch):
    from iins_vae_b200 import train_semi
    from iins_vae_b200.utils import get_args
    monkeypatch.chdir(tmp_path)
    parser = get_args(None)
    parser.add_argument("--supervision_rate", type=float, default=0.1)
    opt = parser.parse_args(["--dataset_env", "room_full", "--batch_size", "256", "--synthetic", "2048", "--n_epochs", "200",
                             "--log_every", "1", "--lr", "0.001", "--checkpoint_interval", "1"])
    torch.manual_seed(0)
    first = train_semi.run(opt, max_steps=2, quiet=True)
    torch.manual_seed(0)
    last = train_semi.run(opt, max_steps=120, quiet=True)
    assert np.isfinite(last["loss"]) and last["loss"] < first["loss"], (first, last)
    assert os.path.exists(os.path.join("saved_models_semi", "room_full_mode_full"))


def test_train_gem_and_test_gem_signatures(tmp_path):
    """train.py:26 / test.py:26 signatures around EMNet; after training the torch optimizer's parameters (the
    network's) hold the trained weights and inference matches the oracle on them."""
    from iins_vae_b200 import models as M
    from iins_vae_b200.data import SyntheticCIR
    from iins_vae_b200.test import test_gem
    from iins_vae_b200.train import train_gem
    from iins_vae_b200.utils import get_args
    opt = get_args(None).parse_args(["--n_epochs", "2", "--checkpoint_interval", "1", "--dataset_env", "nlos", "--log_every", "1"])
    torch.manual_seed(1)
    net = M.EMNet(cir_len=157, num_classes=2, env_dim=16).cuda()
    optim = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.5, 0.999))
    train = SyntheticCIR(1024, 128, 157, 2, seed=3)
    val = SyntheticCIR(500, 250, 157, 2, seed=4)
    w0 = net.restorer.restorer.linear_layer1.weight.detach().clone()
    hist = train_gem(opt, torch.device("cuda"), torch.cuda.FloatTensor, str(tmp_path), str(tmp_path), train, None, optim, net, None)
    assert len(hist) == 16 and hist[-1]["loss"] < hist[0]["loss"]
    assert not torch.equal(w0, net.restorer.restorer.linear_layer1.weight.detach())
    assert os.path.exists(tmp_path / "Network_1.pth")
    res = test_gem(opt, torch.device("cuda"), torch.cuda.FloatTensor, str(tmp_path), str(tmp_path), val, net, 1, None)
    # oracle on the SAME (trained, reloaded) weights
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    pe = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    pr = {k[len("restorer."):]: v for k, v in sd.items() if k.startswith("restorer.")}
    pc = {k[len("classifier."):]: v for k, v in sd.items() if k.startswith("classifier.")}
    cfg = orc.PathConfig(num_classes=2)
    rm, ab, ac = [], [], []
    for batch in val:
        logits, _, err_est = orc.emnet(pe, pr, pc, batch["CIR"], cfg, torch.zeros(250, 8, 1))
        rmse, mae, acc, _ = orc.batch_metrics(err_est, batch["Err"], logits, batch["Label"])
        rm.append(float(rmse)); ab.append(float(mae)); ac.append(float(acc))
    assert abs(res["rmse"] - np.mean(rm)) < 1e-3 and abs(res["abs"] - np.mean(ab)) < 1e-3
    assert abs(res["accuracy"] - np.mean(ac)) < 5e-3
    assert res["err_est"].shape == (500, 1) and res["env_latent"].shape == (500, 16)


def test_checkpoint_roundtrip_with_reference_key_names(tmp_path):
    from iins_vae_b200 import models as M
    cfg = orc.PathConfig()
    pe, pd, pr, pc = orc.init_all(cfg, 9)
    Dec = M.Decoder(1, 4, 3, 4, 16, 157, 2)
    Dec.load_state_dict(pd)
    torch.save(Dec.state_dict(), tmp_path / "Dec_0.pth")
    sd = torch.load(tmp_path / "Dec_0.pth")
    assert list(sd.keys()) == list(orc.decoder_param_shapes(cfg).keys())
    assert torch.equal(sd["decoder.model.2.block.2.running_var"], torch.ones(64))


def _engine(seed, batch):
    from iins_vae_b200 import models as M
    from iins_vae_b200.engine import SemiTrainEngine
    cfg = orc.PathConfig()
    pe, pd, pr, pc = orc.init_all(cfg, seed)
    Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
    Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
    Res = M.Restorer((cfg.range_dim, cfg.code_len))
    Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
    for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
        m.load_state_dict(p)
        m.cuda()
    return cfg, (Enc, Dec, Res, Cls), SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=batch, lr=1e-3, use_graph=False)


def test_optimizer_state_roundtrip_and_torch_adam_compat(tmp_path):
    """the engine's Adam state saves in torch.optim.Adam's own layout over the reference's parameter
    chain (train_semi.py:118-122), loads into a stock torch optimizer, and one stock torch step from it equals one fused
    step on the same gradients; a fresh engine that loads it continues exactly where the first one stopped."""
    import itertools
    B = 64
    cfg, mods, eng = _engine(21, B)
    batches = [orc.synthetic_batch(cfg, B, 900 + j) for j in range(4)]
    for j, sup in enumerate((True, False, True)):
        eng.step(*batches[j], supervised=sup)
    sd = eng.optimizer_state_dict()
    torch.save(sd, tmp_path / "Opt_3.pth")
    sd = torch.load(tmp_path / "Opt_3.pth")
    params = list(itertools.chain(*(m.parameters() for m in mods)))
    assert sd["param_groups"][0]["params"] == list(range(len(params)))
    # restorer.linear_layer2 never gets a gradient -> no state, like torch's lazily created state
    names = [n for m in mods for n, _ in m.named_parameters()]
    missing = [names[i] for i in range(len(params)) if i not in sd["state"]]
    assert missing == ["restorer.linear_layer2.weight", "restorer.linear_layer2.bias"], missing
    enc_idx, res_idx = 0, names.index("restorer.layers.0.weight")
    assert int(sd["state"][enc_idx]["step"]) == 3 and int(sd["state"][res_idx]["step"]) == 2      # Res/Cls skip the unsupervised step

    # (1) a stock torch Adam over detached copies of the parameters, fed the engine's next gradients
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    topt = torch.optim.Adam(ref_params, lr=1e-3, betas=(0.5, 0.999))
    topt.load_state_dict(sd)
    eng.step(*batches[3], supervised=True, update=False)            # gradients only
    grads = [g.clone() for g in eng.flat.grad_views]
    for p, g, n in zip(ref_params, grads, names):
        p.grad = None if "linear_layer2" in n else g
    topt.step()
    eng._adam(True)                                                   # the fused update on the same gradients
    torch.cuda.synchronize()
    for p_ref, p, n in zip(ref_params, params, names):
        assert torch.allclose(p_ref, p.detach(), rtol=0, atol=2e-7), n

    # (2) exact resume: a fresh engine on the same weights + loaded state reproduces the moments and counters
    cfg2, mods2, eng2 = _engine(21, B)
    for m2, m in zip(mods2, mods):
        m2.load_state_dict(m.state_dict())
    eng2.load_optimizer_state_dict(eng.optimizer_state_dict())
    assert torch.equal(eng2.flat.exp_avg, eng.flat.exp_avg) and torch.equal(eng2.flat.exp_avg_sq, eng.flat.exp_avg_sq)
    assert eng2.steps.tolist()[:3] == eng.steps.tolist()[:3] == [4, 3, 3]
    assert abs(float(eng2.lr) - 1e-3) < 1e-9          # the learning rate lives in an fp32 device scalar


def test_test_gem_device_side_accumulation_matches_per_batch_path():
    """test.py:55-107 tail on the device: one host sync, preallocated output arrays -- same numbers as summing per batch."""
    from iins_vae_b200 import models as M
    from iins_vae_b200.data import SyntheticCIR
    from iins_vae_b200.engine import InferenceEngine
    from iins_vae_b200.test import test_gem
    from iins_vae_b200.utils import get_args
    opt = get_args(None).parse_args(["--dataset_env", "nlos"])
    torch.manual_seed(5)
    net = M.EMNet(cir_len=157, num_classes=2, env_dim=16).cuda()
    val = SyntheticCIR(1000, 250, 157, 2, seed=8)
    res = test_gem(opt, torch.device("cuda"), torch.cuda.FloatTensor, "/tmp", "/tmp", val, net, 0, None)
    eng = InferenceEngine(net.encoder, net.restorer, net.classifier, batch_size=250)
    rm, ab, ac, errs = [], [], [], []
    for batch in val:
        err_est, pred, out = eng.run(batch["CIR"], batch["Err"], batch["Label"])
        o = out.tolist()
        rm.append(max(o[4], 0.0) ** 0.5); ab.append(o[1]); ac.append(o[5] / 250)
        errs.append(err_est.clone())
    assert abs(res["rmse"] - np.mean(rm)) < 1e-6 and abs(res["abs"] - np.mean(ab)) < 1e-6 and abs(res["accuracy"] - np.mean(ac)) < 1e-6
    assert res["err_est"].shape == (1000, 1) and torch.equal(res["err_est"], torch.cat(errs))
    assert res["pred"].shape == (1000,) and res["env_latent"].shape == (1000, 16)


def test_prefetched_batch_gives_the_same_step():
    """the copy-stream input pipeline (prefetch + step(prefetched=True)) feeds the step the same
    tensors as the direct path."""
    B = 128
    cfg, mods, eng = _engine(33, B)
    a, b = (tuple(t.pin_memory() for t in orc.synthetic_batch(cfg, B, 70 + j)) for j in range(2))
    eng.step(*a, supervised=True, update=False)
    ref_a = eng.out.clone()
    eng.step(*b, supervised=True, update=False)
    ref_b = eng.out.clone()
    eng.prefetch(*a)
    eng.step(supervised=True, update=False, prefetched=True)
    eng.prefetch(*b)                                   # overlaps the step above
    got_a = eng.out.clone()
    eng.step(supervised=True, update=False, prefetched=True)
    got_b = eng.out.clone()
    torch.cuda.synchronize()
    assert torch.allclose(got_a, ref_a, rtol=1e-6, atol=1e-7) and torch.allclose(got_b, ref_b, rtol=1e-6, atol=1e-7)
    assert not torch.allclose(ref_a, ref_b)


def test_ring_fed_training_sees_exactly_the_ring_batches(tmp_path, monkeypatch):
    """train_semi.run fed by the pinned batch ring (StandardScaler'd arrays -> shuffled pinned batches ->
    prefetch on the copy stream, ragged last batch included, several epochs so that ring slots are recycled): the tensors
    every step actually computes on must be exactly the ring's batches, in order, with the seeded supervision mask -- no
    stale or half-overwritten pinned buffer.  (Comparing trained parameters would not do: the float atomics of the weight
    gradients make two identical runs drift by +-lr on noise-dominated entries.)"""
    from iins_vae_b200 import dataset as D, train_semi
    from iins_vae_b200.engine import SemiTrainEngine
    from iins_vae_b200.parallel import SupervisionMask
    from iins_vae_b200.utils import get_args
    monkeypatch.chdir(tmp_path)
    rng = np.random.RandomState(0)
    cir = rng.randn(1100, 157) * 2 + 0.5
    data = (cir, np.abs(rng.randn(1100, 1)) * 0.15, rng.randint(0, 5, (1100, 1)).astype(float))
    train, _, _, _ = D.err_mitigation_dataset(None, data=data, split_factor=0.8, scaling=True, mode="full")     # 880 windows
    ds = D.UWBDataset(train)
    parser = get_args(None)
    parser.add_argument("--supervision_rate", type=float, default=0.1)
    opt = parser.parse_args(["--dataset_env", "room_full", "--batch_size", "256", "--n_epochs", "4", "--decay_epoch", "3", "--lr", "0.001",
                             "--checkpoint_interval", "-1"])
    seen = []
    orig = SemiTrainEngine.step

    def spy(self, *a, **kw):
        out = orig(self, *a, **kw)
        seen.append((self.B, bool(kw.get("supervised", True)), self.cir.double().sum().item(), self.cir[:, 3].double().sum().item(),
                     self.err.double().sum().item(), self.label.double().sum().item()))
        return out

    monkeypatch.setattr(SemiTrainEngine, "step", spy)
    torch.manual_seed(0)
    last = train_semi.run(opt, dataloader=D.PinnedBatchRing(ds, 256, shuffle=True, seed=7), quiet=True)
    assert np.isfinite(last["loss"])
    mask = SupervisionMask(opt.supervision_rate, seed=1234)
    ring = D.PinnedBatchRing(ds, 256, shuffle=True, seed=7, pin=False)
    want = []
    for epoch in range(4):
        for batch in ring:
            c, e, l = batch["CIR"].double(), batch["Err"].double(), batch["Label"].double()
            want.append((c.shape[0], bool(mask()), float(c.sum()), float(c[:, 3].sum()), float(e.sum()), float(l.sum())))
    assert len(seen) == len(want) == 16
    for i, (g, w) in enumerate(zip(seen, want)):
        assert g[:2] == w[:2], (i, g, w)
        np.testing.assert_allclose(g[2:], w[2:], rtol=1e-9, atol=1e-6, err_msg=f"step {i}: the engine computed on other data than ring batch {i}")


def test_sharded_inference_driver_single_rank():
    """BASELINE configs[4] driver (iins_vae_b200/infer.py) on one rank: pinned host windows -> two-slot device ring -> captured
    inference graph; outputs and metric sums equal a direct InferenceEngine pass over the same windows (ragged tail included)."""
    from iins_vae_b200 import models as M
    from iins_vae_b200.engine import InferenceEngine
    from iins_vae_b200.infer import ShardedInference, synthetic_windows
    torch.manual_seed(2)
    net = M.EMNet(cir_len=157, num_classes=5, env_dim=16).cuda()
    cir, err, label = synthetic_windows(10_000, 157, 5, seed=9)
    assert cir.is_pinned()
    sh = ShardedInference(net, cir, err, label, batch_size=4096)
    for _ in range(2):
        sums = sh.run()
    torch.cuda.synchronize()
    s = sums.tolist()
    ref_e, ref_p = [], []
    for lo in range(0, 10_000, 4096):
        hi = min(lo + 4096, 10_000)
        eng = InferenceEngine(net.encoder, net.restorer, net.classifier, batch_size=hi - lo)
        e, p, _ = eng.run(cir[lo:hi].cuda(), err[lo:hi].cuda(), label[lo:hi].cuda())
        ref_e.append(e.clone()); ref_p.append(p.clone())
    ref_e, ref_p = torch.cat(ref_e), torch.cat(ref_p)
    assert torch.equal(sh.err_est, ref_e) and torch.equal(sh.pred, ref_p)
    assert s[3] == 10_000
    np.testing.assert_allclose(s[0] / s[3], float((ref_e.cpu() - err).abs().mean()), rtol=1e-5)
    np.testing.assert_allclose(s[1] / s[3], float(((ref_e.cpu() - err) ** 2).mean()), rtol=1e-5)
    assert s[2] == float((ref_p.cpu() == label.view(-1).int()).sum())
