"""CPU tests of the host-side logic: C-ABI library exports, option parsing, LR schedule, data-parallel helpers
over gloo with world_size 2, state_dict surface of the drop-in modules."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    """include/iins_b200.h <-> libiins_b200.so (no compute calls: there is no GPU here)."""
    from iins_vae_b200 import build
    path = build.build()
    dll = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "iins_b200.h")).read()
    declared = set(re.findall(r"\b(iins_[a-z0-9_]+)\s*\(", header))
    declared -= {"iins_config", "iins_status", "iins_stream_t"}
    assert len(declared) >= 30
    for sym in sorted(declared):
        assert hasattr(dll, sym), f"{sym} declared in the header but not exported"
    from iins_vae_b200._capi import EXPORTS
    assert set(EXPORTS) <= declared
    dll.iins_abi_version.restype = ctypes.c_int
    assert dll.iins_abi_version() == 3


def test_config_validation_and_sizes_without_gpu():
    from iins_vae_b200._capi import IinsConfig, IinsLib
    from iins_vae_b200 import build
    lib = IinsLib(build.build())
    ok = IinsConfig(4096, 157, 4, 3, 4, 16, 2, 5, 16)
    assert lib.iins_validate_config(ok) == 0
    assert lib.iins_encoder_num_params(ok) == 32 and lib.iins_decoder_num_params(ok) == 38
    assert lib.iins_restorer_num_params(ok) == 10 and lib.iins_classifier_num_params(ok) == 8
    assert lib.iins_encoder_ws_floats(ok) > 4096 * 15000          # saved activations, ~18k floats / sample
    for bad in (IinsConfig(0, 157, 4, 3, 4, 16, 2, 5, 16), IinsConfig(8, 157, 32, 3, 4, 16, 2, 5, 16),
                IinsConfig(8, 157, 4, 3, 3, 16, 2, 5, 16), IinsConfig(8, 157, 4, 3, 4, 15, 2, 5, 16)):
        assert lib.iins_validate_config(bad) != 0
        assert len(lib.dll.iins_last_error()) > 0
    assert lib.iins_set_compute_mode(7) != 0 and lib.iins_set_compute_mode(0) == 0


def test_contexts_hold_the_state():
    """SURVEY 8(b) "no global state": compute mode and concurrency live in iins_ctx handles; a thread's current context is
    what the plain entry points use, other contexts (and other threads) are unaffected."""
    import threading
    from iins_vae_b200._capi import IinsLib
    from iins_vae_b200 import build
    lib = IinsLib(build.build())
    d = lib.dll
    assert d.iins_set_compute_mode(0) == 0
    a, b = d.iins_ctx_create(), d.iins_ctx_create()
    assert a and b and a != b
    assert d.iins_ctx_set_compute_mode(a, 1) == 0 and d.iins_ctx_set_compute_mode(b, 2) == 0
    assert d.iins_ctx_set_compute_mode(a, 9) != 0
    assert d.iins_get_compute_mode() == 0                      # the default context is untouched
    assert d.iins_ctx_make_current(a) == 0 and d.iins_get_compute_mode() == 1 and d.iins_ctx_get_current() == a
    seen = []
    t = threading.Thread(target=lambda: seen.append((d.iins_ctx_get_current(), d.iins_get_compute_mode())))
    t.start(); t.join()
    assert seen == [(None, 0)]                                 # another thread: no current context -> default context
    assert d.iins_ctx_make_current(b) == 0 and d.iins_get_compute_mode() == 2
    assert d.iins_set_compute_mode(0) == 0 and d.iins_ctx_get_compute_mode(b) == 0 and d.iins_ctx_get_compute_mode(a) == 1
    assert d.iins_ctx_make_current(None) == 0 and d.iins_get_compute_mode() == 0
    d.iins_ctx_destroy(a); d.iins_ctx_destroy(b)


def test_modules_fail_loudly_on_cpu_tensors():
    from iins_vae_b200 import models as M
    with pytest.raises(RuntimeError, match="CUDA"):
        M.Encoder(1, 4, 3, 4, 16, 2)(torch.zeros(2, 157))
    with pytest.raises(NotImplementedError):
        M.Encoder(conv_type=2)
    with pytest.raises(NotImplementedError):
        M.Restorer((2, 8), net_type="Conv2d")
    with pytest.raises(RuntimeError, match="CUDA"):
        M.Restorer((2, 8), net_type="Conv1d")(torch.zeros(4, 2, 8))
    # the Conv1d heads keep the reference's state_dict keys, BatchNorm buffers included (models.py:661-693, 865-891)
    from oracle import iins_oracle as orc
    cfg = orc.PathConfig()
    assert list(M.Restorer((2, 8), net_type="Conv1d").state_dict()) == list(orc.restorer_conv1d_param_shapes(cfg))
    assert list(M.Classifier(16, 5, net_type="Conv1d").state_dict()) == list(orc.classifier_conv1d_param_shapes(cfg))


def test_state_dict_surface_matches_oracle_inventory():
    from iins_vae_b200 import models as M, model as M2
    from oracle import iins_oracle as orc
    cfg = orc.PathConfig()
    pairs = [(M.Encoder(1, 4, 3, 4, 16, 2), orc.encoder_param_shapes(cfg)),
             (M.Decoder(1, 4, 3, 4, 16, 157, 2), orc.decoder_param_shapes(cfg)),
             (M.Restorer((2, 8)), orc.restorer_param_shapes(cfg)), (M.Classifier(16, 5), orc.classifier_param_shapes(cfg)),
             (M2.Encoder(1, 4, 3, 4, 16, 2), orc.encoder_param_shapes(cfg)), (M2.Restorer(False, 1, 1, 2, 4), orc.restorer_param_shapes(cfg))]
    for mod, shapes in pairs:
        sd = mod.state_dict()
        assert [(k, tuple(v.shape)) for k, v in sd.items()] == list(shapes.items())
    net = M.EMNet(cir_len=157, num_classes=2, env_dim=16)
    assert any(k.startswith("encoder.range_encoder.model.2.") for k in net.state_dict())


def test_options_and_lr_schedule():
    from iins_vae_b200.utils import get_args, num_classes_for
    from iins_vae_b200.models import LambdaLR
    opt = get_args(None).parse_args([])
    assert (opt.batch_size, opt.lr, opt.b1, opt.b2, opt.n_residual, opt.n_downsample, opt.env_dim) == (500, 1e-4, 0.5, 0.999, 3, 4, 16)
    assert (opt.conv_type, opt.dim, opt.range_dim, opt.restorer_type, opt.classifier_type) == (1, 4, 2, "Linear", "Linear")
    assert num_classes_for("room_full") == 5 and num_classes_for("nlos") == 2
    with pytest.raises(ValueError):
        num_classes_for("nowhere")
    s = LambdaLR(500, 0, 100)
    assert s.step(0) == 1.0 and s.step(100) == 1.0 and abs(s.step(300) - 0.5) < 1e-12 and s.step(500) == 0.0


def test_supervision_mask_stream_matches_reference_draw():
    from iins_vae_b200.parallel import SupervisionMask
    from oracle import iins_oracle as orc
    a, rng = SupervisionMask(0.1, seed=99), np.random.RandomState(99)
    seq = [a() for _ in range(200)]
    assert seq == [orc.supervision_mask(rng, 0.1) for _ in range(200)]
    assert 0.4 < np.mean(seq) < 0.7            # P(mask=1) = Phi(0.1) ~ 0.54 (a NORMAL draw, train_semi.py:203)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from iins_vae_b200.parallel import init_distributed, shard_range, allreduce_mean_, SupervisionMask, broadcast_parameters
rank, local, world, pg = init_distributed("gloo")
assert world == 2 and pg is not None
b, e = shard_range(8192, rank, world)
assert (b, e) == (rank * 4096, (rank + 1) * 4096)
# gradient buckets: only the first n_active entries are reduced on an unsupervised step
flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
allreduce_mean_(flat, 6, pg)
want = torch.arange(10, dtype=torch.float32) * (rank + 1)
want[:6] = torch.arange(6, dtype=torch.float32) * 1.5
assert torch.equal(flat, want), (rank, flat)
# identical supervision decisions on every rank
m = SupervisionMask(0.1, seed=1234)
seq = torch.tensor([m() for _ in range(64)])
other = [torch.zeros_like(seq) for _ in range(world)]
dist.all_gather(other, seq, group=pg)
assert all(torch.equal(o, seq) for o in other)
lin = torch.nn.Linear(3, 2)
broadcast_parameters([lin], pg)
w = [torch.zeros_like(lin.weight) for _ in range(world)]
dist.all_gather(w, lin.weight.data, group=pg)
assert torch.equal(w[0], w[1])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_data_parallel_helpers_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + os.getpid() % 400
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_conv2d_variant_config_sizes_and_state_dict_surface():
    """2-D variant (conv_type = 2, expand = True): the config is accepted only with conv_type 2 by the 2-D entry points' sizing
    functions, the modules hold the reference's 2-D keys / shapes (oracle2d.*_param_shapes, pinned to the live reference by
    tests/golden/make_golden2d.py), unrunnable combinations raise."""
    from iins_vae_b200._capi import IinsConfig, IinsLib
    from iins_vae_b200 import build, models as M
    from oracle import iins_oracle as orc, iins_oracle2d as orc2
    lib = IinsLib(build.build())
    c1, c2 = IinsConfig(8, 157, 4, 3, 4, 16, 2, 5, 16, 1), IinsConfig(8, 157, 4, 3, 4, 16, 2, 5, 16, 2)
    assert lib.iins_validate_config(c2) == 0 and lib.iins_validate_config(IinsConfig(8, 157, 4, 3, 4, 16, 2, 5, 16, 3)) != 0
    assert lib.iins_validate_config(IinsConfig(8, 157, 2, 3, 4, 16, 2, 5, 16, 2)) != 0          # dim < 4: norm kernels need >= 4 channels
    assert lib.iins_encoder2d_ws_floats(c1) == 0 and lib.iins_decoder2d_scratch_floats(c1) == 0  # wrong conv_type: refused
    assert lib.iins_encoder2d_ws_floats(c2) > 8 * 128 * 128 * 4 * 2 and lib.iins_decoder2d_ws_floats(c2) > 8 * 128 * 128 * 4 * 2
    assert lib.iins_restorer_ws_floats(c2) == lib.iins_restorer_ws_floats(c1)                   # hidden layers only; input is 128 wide
    cfg = orc.PathConfig()
    Enc = M.Encoder(2, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim, expand=True)
    Dec = M.Decoder(2, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim, expand=True)
    Res = M.Restorer((cfg.range_dim, cfg.code_len, cfg.code_len))
    for m, shapes in ((Enc, orc2.encoder_param_shapes(cfg)), (Dec, orc2.decoder_param_shapes(cfg)), (Res, orc2.restorer_param_shapes(cfg))):
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v)) for k, v in shapes.items()]
    for bad in (lambda: M.Encoder(2, expand=False), lambda: M.Decoder(3, expand=True), lambda: M.Restorer((2, 8, 8), net_type="Conv1d")(torch.zeros(1))):
        with pytest.raises((NotImplementedError, RuntimeError)):
            bad()
