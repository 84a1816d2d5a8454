set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > $O/r02e_pytest_gpu.log 2>&1; tail -3 $O/r02e_pytest_gpu.log
python __graft_entry__.py --smoke > $O/r02e_smoke.log 2>&1; tail -1 $O/r02e_smoke.log
python bench.py > $O/r02e_bench_1gpu.json 2> $O/r02e_bench_1gpu.err
python bench.py --impl reference --steps 8 --warmup 3 > $O/r02e_bench_reference_arm.json 2> $O/r02e_ref.err
python bench.py --batch 8192 --no-cpu-baseline --no-gpu-reference > $O/r02e_bench_1gpu_b8192.json 2> /dev/null
python bench.py --batch 8192 --compute-mode bf16 --no-cpu-baseline --no-gpu-reference > $O/r02e_bench_1gpu_b8192_bf16.json 2> /dev/null
python bench.py --dim 16 --no-cpu-baseline > $O/r02e_bench_1gpu_dim16.json 2> /dev/null
python tools/bench_conv2d.py --batch 64 > $O/r02e_bench_conv2d_b64.json 2> /dev/null
python tools/step_profile.py fp32 4096 > $O/r02e_step_profile_serial.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/r02e_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"iins_(win|trunk)" --launch-skip 30 --launch-count 18 -o $O/r02e_win_trunk python tools/step_profile.py fp32 4096 > $O/r02e_ncu_full.log 2>&1
ls -la $O/r02e_*
python -m iins_vae_b200.infer --synthetic 10000000 --batch_size 32768 > $O/r02e_infer_10M_1gpu.json 2> $O/r02e_infer.err
IINS_PDL=1 python -m iins_vae_b200.infer --synthetic 10000000 --batch_size 32768 > $O/r02e_infer_10M_1gpu_pdl1.json 2>> $O/r02e_infer.err
tail -c 400 $O/r02e_infer_10M_1gpu.json; tail -c 400 $O/r02e_infer_10M_1gpu_pdl1.json
