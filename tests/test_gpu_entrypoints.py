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
