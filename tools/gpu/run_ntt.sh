set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ntt" > gpurun_out/pytest_ntt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ntt.log
tail -15 gpurun_out/pytest_ntt.log
NMSM_ROWS=7,8,9 timeout 600 python tests/bench_configs.py --fixed-base > gpurun_out/configs_ntt.jsonl 2> gpurun_out/configs_ntt.err
tail -3 gpurun_out/configs_ntt.err
cat gpurun_out/configs_ntt.jsonl
