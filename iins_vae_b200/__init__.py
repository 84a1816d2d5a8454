"""iins_vae_b200 -- B200-native (sm_100a) implementation of the IIns-VAE training / inference hot path.

Drop-in surface (same names as the reference's flat scripts): ``models`` (Encoder, Decoder, Restorer,
Classifier, weights_init_normal, LambdaLR), ``utils.get_args``, ``train_semi``, ``train``, ``test``.
Fast path: ``engine.SemiTrainEngine`` / ``engine.InferenceEngine``.
"""
__version__ = "0.1.0"

COMPUTE_MODES = {"fp32": 0, "bf16": 1, "simt": 2}


def set_compute_mode(mode="fp32"):
    """Arithmetic of the conv / linear GEMMs (process-wide, see include/iins_b200.h):
    "fp32" (default) tcgen05 tensor cores with the 3-piece bf16 split (fp32-grade results);
    "bf16" tcgen05 tensor cores, plain bf16 operands, fp32 accumulation;
    "simt" the fp32 FFMA bring-up kernels."""
    from ._capi import get_lib
    lib = get_lib()
    lib.check(lib.iins_set_compute_mode(COMPUTE_MODES[mode] if isinstance(mode, str) else int(mode)), "set_compute_mode")


def get_compute_mode():
    from ._capi import get_lib
    inv = {v: k for k, v in COMPUTE_MODES.items()}
    return inv[get_lib().iins_get_compute_mode()]
