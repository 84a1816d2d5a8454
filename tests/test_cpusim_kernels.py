"""CPU (no GPU): the SIMT kernels + host launch plans compiled for the logic simulator
(tests/cpusim) must reproduce the oracle: forward tensors, the four loss terms, every parameter
gradient, for both supervision branches, whole and ragged tiles.  This checks LOGIC only (indexing,
barriers, reductions, sequencing); the parity tests proper are the `-m gpu` tests, which call the
nvcc-built sm_100a library through the same C ABI."""
import shutil

import numpy as np
import pytest
import torch

from oracle import iins_oracle as orc
from tests import parity

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ needed for the simulator build")


@pytest.fixture(scope="module")
def sim():
    from tests.cpusim import harness
    return harness, harness.build_sim()


@pytest.mark.parametrize("batch,supervised,seed", [(2, True, 0), (5, False, 1), (19, True, 2)])
def test_semi_step_matches_oracle(sim, batch, supervised, seed):
    H, lib = sim
    cfg = orc.PathConfig()
    pe, pd, pr, pc = orc.init_all(cfg, seed)
    cir, err, label = orc.synthetic_batch(cfg, batch, 1000 + seed)
    torch.manual_seed(7 + seed)
    noise = torch.randn(batch, cfg.env_dim // 2)
    ref, ref_grads = orc.semi_step_with_grads(pe, pd, pr, pc, cir, err, label, cfg, supervised, noise.view(batch, -1, 1))
    st = H.SimStep(lib, cfg, batch, pe, pd, pr, pc)
    st.forward(cir, noise)
    parity.assert_out_close("range_code", st.rc, ref["range_code"])
    parity.assert_out_close("env_code", st.cat, ref["env_code"].view(batch, -1))
    parity.assert_out_close("env_code_rv", st.lat, ref["env_code_rv"].view(batch, -1))
    parity.assert_out_close("kl", st.kl, ref["kl"].view(1))
    parity.assert_out_close("cir_gen", st.xrec, ref["cir_gen"].view(batch, -1))
    parity.assert_out_close("err_fake", st.err_est, ref["err_fake"])
    parity.assert_out_close("label_fake", st.logits, ref["label_fake"])
    st.loss_backward(err, label.view(-1), supervised)
    np.testing.assert_allclose(float(st.out[0]), float(ref["loss_ae"]), rtol=1e-5)
    if supervised:
        np.testing.assert_allclose(float(st.out[1]) * orc.LAMBDA_RES, float(ref["loss_res"]), rtol=1e-5)
        np.testing.assert_allclose(float(st.out[2]), float(ref["loss_env"]), rtol=1e-5)
        rmse, mae, acc, _ = orc.batch_metrics(ref["err_fake"], err, ref["label_fake"], label)
        np.testing.assert_allclose(float(st.out[4]) ** 0.5, float(rmse), rtol=1e-5)
        np.testing.assert_allclose(float(st.out[5]) / batch, float(acc), rtol=1e-6)
    total = float(st.out[3]) + orc.LAMBDA_RANGE * float(st.kl)
    np.testing.assert_allclose(total, float(ref["loss"]), rtol=1e-5)
    gscale = max(float(g.abs().max()) for g in ref_grads.values() if g is not None)
    got = st.grads(supervised)
    for name, g in ref_grads.items():
        if g is None:
            assert got[name] is None
            continue
        ok, msg = parity.grad_error(name, got[name], g, gscale)
        assert ok, msg


def test_config_validation_fails_loudly(sim):
    H, lib = sim
    c = H.make_cfg(orc.PathConfig(dim=16), 4)          # trunk 256 channels: not supported by this build
    assert lib.iins_validate_config(c) != 0
    assert b"dim" in lib.dll.iins_last_error()


def test_philox_noise_is_standard_normal(sim):
    """noise=None -> Philox4x32-10 Box-Muller normals (the latent is never consumed by a
    loss, so only the distribution matters)."""
    H, lib = sim
    cfg = orc.PathConfig()
    B = 256
    pe, pd, pr, pc = orc.init_all(cfg, 0)
    # force mu = 0, log_sigma = 0 so the latent IS the noise: zero the last conv of the env encoder
    key_w = [k for k in pe if k.startswith("env_encoder")][-2]
    key_b = [k for k in pe if k.startswith("env_encoder")][-1]
    pe[key_w].zero_(); pe[key_b].zero_()
    st = H.SimStep(lib, cfg, B, pe, pd, pr, pc)
    st.forward(orc.synthetic_batch(cfg, B, 5)[0], None)
    z = st.lat.numpy().ravel()
    assert abs(z.mean()) < 0.1 and abs(z.std() - 1.0) < 0.1 and np.abs(z).max() < 6
    assert len(np.unique(z)) == z.size
