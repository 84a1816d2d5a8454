"""Digest layout shared by make_golden.py (writer) and tests/parity.py (reader)."""
import numpy as np

BIG = 4096          # tensors with more elements are stored as a digest
N_SAMPLES = 64


def sample_positions(n: int) -> np.ndarray:
    rs = np.random.RandomState(n % 65521)
    return np.sort(rs.choice(n, N_SAMPLES, replace=False))
