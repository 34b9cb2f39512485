set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fixed" > gpurun_out/pytest_fixed.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fixed.log
tail -5 gpurun_out/pytest_fixed.log
timeout 600 python tools/bench_configs.py --fixed-base > gpurun_out/configs_fixed_base.jsonl 2> gpurun_out/configs_fixed_base.err
tail -3 gpurun_out/configs_fixed_base.err
