"""The streaming kernels (fused loss, fused Adam) at L2-exceeding sizes, alone: the launches bench.py's roofline_hbm times
(run under ncu to capture them: ncu -k regex:iins_loss_kernel|iins_adam_kernel ...)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from iins_vae_b200._capi import get_lib
torch.cuda.set_device(0)
print(json.dumps(bench.hbm_kernel_rooflines(get_lib(), bench.measured_peaks()[0])))
