mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_last.log
tail -3 gpurun_out/pytest_gpu_last.log
