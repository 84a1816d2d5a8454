"""Digest layout shared by make_golden.py (writer) and tests/parity.py (reader)."""
import numpy as np

BIG = 4096          # tensors with more elements are stored as a digest
N_SAMPLES = 64


def sample_positions(n: int) -> np.ndarray:
    rs = np.random.RandomState(n % 65521)
    return np.sort(rs.choice(n, N_SAMPLES, replace=False))


N_SAMPLES_2D = 1024


def sample_positions_2d(n: int) -> np.ndarray:
    """Sampled entries of a large gradient tensor of the 2-D fixture (tests/golden/make_golden2d.py)."""
    rs = np.random.RandomState((n * 31 + 7) % 65521)
    return np.sort(rs.choice(n, N_SAMPLES_2D, replace=False))


def conv_head_case_inputs(kind, seed, batch, cfg):
    """Inputs and parameters of one Conv1d-head fixture case (tests/golden/make_golden_convheads.py records the reference's
    outputs for exactly these): x, the parameter / buffer dict in state_dict order."""
    import torch
    from oracle import iins_oracle as orc
    gen = torch.Generator().manual_seed(seed)
    if kind == "res":
        shapes = orc.restorer_conv1d_param_shapes(cfg)
        x = torch.rand(batch, cfg.range_dim, cfg.code_len, generator=gen)
    else:
        shapes = orc.classifier_conv1d_param_shapes(cfg)
        x = torch.randn(batch, cfg.env_dim, 1, generator=gen) * 0.5
    p = orc.init_conv_head_params(shapes, gen)
    p[[k for k in p if k.endswith("running_mean")][0]] += 0.05          # non-trivial buffers (eval mode uses them)
    p[[k for k in p if k.endswith("running_var")][0]] *= 1.3
    return x, p, gen
