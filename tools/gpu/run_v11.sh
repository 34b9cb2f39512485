set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v11.log
tail -3 gpurun_out/pytest_gpu_v11.log
NMSM_ROWS=7,8,9 timeout 600 python tests/bench_configs.py --fixed-base > gpurun_out/configs_ntt_v2.jsonl 2> gpurun_out/configs_ntt_v2.err
cat gpurun_out/configs_ntt_v2.jsonl | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v11.log 2>&1; tail -2 gpurun_out/smoke_v11.log
