set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v10.log
tail -3 gpurun_out/pytest_gpu_v10.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err
tail -2 gpurun_out/bench_v10.err
timeout 600 python tests/bench_configs.py > gpurun_out/configs_v10.jsonl 2> gpurun_out/configs_v10.err
NMSM_ROWS=0,1,2,3 timeout 600 python tests/bench_configs.py --fixed-base > gpurun_out/configs_fixed_base_v10.jsonl 2> gpurun_out/configs_fixed_base_v10.err
NMSM_TK=4 NMSM_ROWS=0,1 timeout 600 python tests/bench_configs.py --fixed-base > gpurun_out/configs_fixed_base_v10_TK4.jsonl 2>/dev/null
