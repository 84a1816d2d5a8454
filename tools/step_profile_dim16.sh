#!/bin/bash
# dim = 16 (utils.py:38 --filters 16): ncu --set full of the widest per-layer tensor-core launches of one step (256-channel trunk convolutions)
ncu --set full --clock-control none --import-source on -k regex:"iins_tc_(nt|tn)" --launch-skip 60 --launch-count 14 -o gpurun_out/r02e_dim16_tc python tools/step_profile.py fp32 1024 16 > gpurun_out/r02e_dim16_ncu.log 2>&1
tail -2 gpurun_out/r02e_dim16_ncu.log
