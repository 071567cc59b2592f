#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ per-layer times), the ncu launch list of the bench command, and
# ncu --set full captures of the conv kernels of one chunk and of the HBM-bound kernels.   tools/gpu_round.sh [noncu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 --layers gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/bench.log
tail -c 3000 gpurun_out/bench.log
if [ "$1" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_halo|conv_wide|conv_tc|first_conv' -s 24 -c 24 -o gpurun_out/prof_conv -f python tools/prof_forward.py 14 > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'aggregate_kernel|eval_fused' -c 3 -o gpurun_out/prof_hbm -f python tools/prof_forward.py 14 > gpurun_out/ncu_hbm.log 2>&1
fi
tail -n 3 gpurun_out/pytest.log; tail -n 3 gpurun_out/smoke.log
