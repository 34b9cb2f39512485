set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v8.log
tail -5 gpurun_out/pytest_gpu_v8.log
timeout 900 python tests/bench_configs.py --fixed-base > gpurun_out/configs_fixed_base.jsonl 2> gpurun_out/configs_fixed_base.err
tail -3 gpurun_out/configs_fixed_base.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
NMSM_K=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v8_K2.json 2>/dev/null
NMSM_K=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v8_K4.json 2>/dev/null
