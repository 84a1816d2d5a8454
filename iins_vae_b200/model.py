"""The ``model.py`` dialect of the reference (a broken rewrite of models.py with different keyword names,
SURVEY.md 8(b)): same modules, constructor signatures of model.py:31, :61, :85, :111 mapped onto models.py's."""
from . import models as _m
from .models import EMNet, LambdaLR, weights_init_normal  # noqa: F401


class Encoder(_m.Encoder):
    def __init__(self, conv_type=1, filters=4, n_residual=3, n_downsample=4, env_dim=16, range_dim=2):
        super().__init__(conv_type=conv_type, dim=filters, n_residual=n_residual, n_downsample=n_downsample,
                         style_dim=env_dim, out_dim=range_dim)


class Decoder(_m.Decoder):
    def __init__(self, conv_type=1, filters=4, n_residual=3, n_upsample=4, env_dim=16, range_dim=2, out_dim=157):
        super().__init__(conv_type=conv_type, dim=filters, n_residual=n_residual, n_upsample=n_upsample,
                         style_dim=env_dim, in_dim=out_dim, out_dim=range_dim)


class Restorer(_m.Restorer):
    def __init__(self, use_soft=False, layer_type=1, conv_type=1, range_dim=2, n_downsample=4):
        if layer_type != 1:
            raise NotImplementedError("iins_vae_b200: only the Linear restorer (layer_type=1) is on the B200 path")
        super().__init__((range_dim, 128 // 2 ** n_downsample), soft=use_soft, conv_type=conv_type)


class Classifier(_m.Classifier):
    def __init__(self, env_dim=16, num_classes=2, filters=16, layer_type=1):
        if layer_type != 1:
            raise NotImplementedError("iins_vae_b200: only the Linear classifier (layer_type=1) is on the B200 path")
        super().__init__(env_dim, num_classes, filters=filters)
