#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list and one --set full capture of the hot kernels of the same command.
# usage (build container): tools/gpurun_retry.sh 1500 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v ref-kernel gpurun_out/pytest_gpu.log | tail -4
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --only main --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:k_shade_tc8|k_color_tc2|k_knn|k_march|k_pack_quads|k_composite" -s 18 -c 6 \
    -f -o gpurun_out/full python bench.py --steps 1 --warmup 3 --only main --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/train_profile.py > gpurun_out/train_profile.log 2>&1; echo "train profile rc=$?"; head -4 gpurun_out/train_profile.log
timeout 300 python tools/shard_latency.py > gpurun_out/shard_latency_lego.log 2>&1; head -5 gpurun_out/shard_latency_lego.log
PNB_PROF=1 timeout 200 python tools/tc_profile.py 2>&1 | grep -v "per-CTA kernel cycles" > gpurun_out/tc_profile.log; grep "cycles per 128" gpurun_out/tc_profile.log
