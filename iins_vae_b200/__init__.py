"""iins_vae_b200 -- B200-native (sm_100a) implementation of the IIns-VAE training / inference hot path.

Drop-in surface (same names as the reference's flat scripts): ``models`` (Encoder, Decoder, Restorer,
Classifier, weights_init_normal, LambdaLR), ``utils.get_args``, ``train_semi``, ``train``, ``test``.
Fast path: ``engine.SemiTrainEngine`` / ``engine.InferenceEngine``.
"""
__version__ = "0.1.0"
