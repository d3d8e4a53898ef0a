mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "ref-kernel" gpurun_out/pytest_gpu.log | tail -25
