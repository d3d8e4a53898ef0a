mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dist.py -q -rs -s > gpurun_out/pytest_gpu_dist2.log 2>&1; echo "dist pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_dist2.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_2gpu.json | cut -c1-300; tail -3 gpurun_out/bench_2gpu.err
