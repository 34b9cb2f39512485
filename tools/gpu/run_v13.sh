set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v13.log
tail -3 gpurun_out/pytest_gpu_v13.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v13.log 2>&1; tail -1 gpurun_out/smoke_v13.log
