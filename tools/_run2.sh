mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shade.py -x -q -k "deferred or lego_scale or matches_oracle or silhouette" 2>&1 | tail -5
echo "=== DEFER (default)"; PNB_PROF=1 timeout 300 python tools/tc_profile.py 2>&1 | grep -v "per-CTA kernel cycles" | tee gpurun_out/tc_profile_defer.log
echo "=== non-deferred"; PNB_PROF=1 PNB_DBG_FLAGS=8 timeout 300 python tools/tc_profile.py 2>&1 | grep -v "per-CTA kernel cycles" | tee gpurun_out/tc_profile_nodefer.log
timeout 300 python bench.py --only main --no-cpu-baseline --steps 5 2>&1 | tail -1 | cut -c1-300
timeout 300 python tools/shard_latency.py 2>&1 | tee gpurun_out/shard_latency_lego.log | head -6
