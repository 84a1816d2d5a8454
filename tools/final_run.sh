set -x
O=gpurun_out
TAG=${TAG:-r02f}
timeout 900 python -m pytest tests -m gpu -q -s > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err
python bench.py --impl reference --steps 8 --warmup 3 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_ref.err
python bench.py --batch 8192 --no-cpu-baseline --no-gpu-reference > $O/${TAG}_bench_1gpu_b8192.json 2> /dev/null
python bench.py --batch 8192 --compute-mode bf16 --no-cpu-baseline --no-gpu-reference > $O/${TAG}_bench_1gpu_b8192_bf16.json 2> /dev/null
python bench.py --dim 16 --no-cpu-baseline > $O/${TAG}_bench_1gpu_dim16.json 2> /dev/null
python tools/bench_conv2d.py --batch 64 > $O/${TAG}_bench_conv2d_b64.json 2> /dev/null
python tools/step_profile.py fp32 4096 > $O/${TAG}_step_profile_serial.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"iins_(win|trunk)" --launch-skip 30 --launch-count 18 -o $O/${TAG}_win_trunk python tools/step_profile.py fp32 4096 > $O/${TAG}_ncu_full.log 2>&1
ls -la $O/${TAG}_*
python -m iins_vae_b200.infer --synthetic 10000000 --batch_size 32768 > $O/${TAG}_infer_10M_1gpu.json 2> $O/${TAG}_infer.err
IINS_PDL=1 python -m iins_vae_b200.infer --synthetic 10000000 --batch_size 32768 > $O/${TAG}_infer_10M_1gpu_pdl1.json 2>> $O/${TAG}_infer.err
tail -c 400 $O/${TAG}_infer_10M_1gpu.json; tail -c 400 $O/${TAG}_infer_10M_1gpu_pdl1.json
