mkdir -p gpurun_out
PNB_TRAIN_PROFILE=1 timeout 400 python bench.py --only train --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/train_bench.json 2> gpurun_out/train_cpu_profile.log; tail -32 gpurun_out/train_cpu_profile.log
echo "=== no weights"; PNB_NO_WEIGHTS=1 timeout 300 python tools/tc_profile.py 2>&1 | grep "cycles per 128-row"
