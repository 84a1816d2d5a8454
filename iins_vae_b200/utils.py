"""Options of the reference (utils.py:17-54) without its plotting imports (matplotlib / umap are not part of
the hot path and are absent on the GPU box).  ``get_args`` keeps the reference's quirk of ignoring the parser
passed in and returning a fresh one, and adds the five options ``train_semi.py`` reads but ``utils.get_args``
never defines (train_semi.py:67-82: conv_type, dim, range_dim, restorer_type, classifier_type)."""
import argparse


def get_args(parser=None):
    parser = argparse.ArgumentParser()
    # learning setting
    parser.add_argument("--epoch", type=int, default=0, help="epoch to start training from")
    parser.add_argument("--n_epochs", type=int, default=500, help="number of epochs of training")
    parser.add_argument("--test_epoch", type=int, default=500, help="epoch to test model performance")
    # optimization parameters
    parser.add_argument("--batch_size", type=int, default=500, help="size of the batch size for training")
    parser.add_argument("--lr", type=float, default=0.0001, help="adam: learning rate")
    parser.add_argument("--b1", type=float, default=0.5, help="adam: decay of first order moment of gradient")
    parser.add_argument("--b2", type=float, default=0.999, help="adam: decay of first order moment of gradient")
    parser.add_argument("--decay_epoch", type=int, default=100, help="epoch from which to decay")
    parser.add_argument("--n_cpu", type=int, default=8, help="number of cpu threads to load data")
    # network choice
    parser.add_argument("--net_ablation", type=str, default="loop", help="choices: loop, loops")
    # network structure
    parser.add_argument("--n_residual", type=int, default=3, help="number of residual blocks")
    parser.add_argument("--n_downsample", type=int, default=4, help="number of downsampling layers")
    parser.add_argument("--filters", type=int, default=16, help="number of filters in first encoder layer")
    parser.add_argument("--env_dim", type=int, default=16, help="dimension of environment code")
    parser.add_argument("--use_soft", type=bool, default=False, help="estimate soft range information")
    parser.add_argument("--identifier_type", type=int, default=1, help="1 for linear, 2 for conv1d, 3 for conv2d")
    parser.add_argument("--regressor_type", type=int, default=1, help="1 for linear, 2 for conv1d, 3 for conv2d")
    # data choices
    parser.add_argument("--dataset_name", type=str, default="zenodo", help="name of the dataset for usage")
    parser.add_argument("--dataset_env", type=str, default="nlos", help="environment options for zenodo dataset")
    parser.add_argument("--mode", type=str, default="full", help="mode to assign train and test data")
    parser.add_argument("--split_factor", type=float, default=0.8, help="factor to split train and test data")
    # check intervals
    parser.add_argument("--sample_interval", type=int, default=20, help="epoch interval between saving samples")
    parser.add_argument("--checkpoint_interval", type=int, default=50, help="epoch interval between checkpoints")
    # --- read by train_semi.py:67-82 but never declared by the reference
    parser.add_argument("--conv_type", type=int, default=1, help="1: Conv1d path (the only one on the B200 path)")
    parser.add_argument("--dim", type=int, default=4, help="Encoder/Decoder dim (models.py:33 default)")
    parser.add_argument("--range_dim", type=int, default=2, help="channels of the range code")
    parser.add_argument("--restorer_type", type=str, default="Linear")
    parser.add_argument("--classifier_type", type=str, default="Linear")
    # --- B200 runtime options (new)
    parser.add_argument("--synthetic", type=int, default=0, help="number of synthetic samples to train on (0 = load the dataset)")
    parser.add_argument("--compute_mode", type=str, default="fp32", choices=["fp32", "bf16", "simt"])
    parser.add_argument("--log_every", type=int, default=20, help="read the device-side loss/metric accumulators every N steps")
    return parser


NUM_CLASSES = {  # run.py:40-55 and train_semi.py:46-63
    "room_full": 5, "obstacle_full": 10, "nlos": 2, "room_part": 3, "room_full_rough": 3, "obstacle_part": 4,
    "obstacle_part2": 2, "room_full_rough2": 2, "paper": 4,
}


def num_classes_for(dataset_env: str) -> int:
    if dataset_env not in NUM_CLASSES:
        raise ValueError(f"Unknown environment {dataset_env!r}")
    return NUM_CLASSES[dataset_env]
