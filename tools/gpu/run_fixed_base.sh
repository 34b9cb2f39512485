set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fixed or point_table" > gpurun_out/pytest_fixed.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fixed.log
tail -5 gpurun_out/pytest_fixed.log
for K in 2 4; do NMSM_TK=$K NMSM_ROWS=0,1,3 timeout 300 python tools/bench_configs.py --fixed-base > gpurun_out/configs_fixed_base_TK$K.jsonl 2>/dev/null; done
timeout 900 python tools/bench_configs.py --fixed-base > gpurun_out/configs_fixed_base.jsonl 2> gpurun_out/configs_fixed_base.err
tail -3 gpurun_out/configs_fixed_base.err
