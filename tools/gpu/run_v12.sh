set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mul or golden or object_api or config_secp" > gpurun_out/pytest_mul_v12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mul_v12.log
tail -3 gpurun_out/pytest_mul_v12.log
NMSM_ROWS=0 timeout 300 python tests/bench_configs.py > gpurun_out/configs_v12_row0.jsonl 2>/dev/null; cat gpurun_out/configs_v12_row0.jsonl | cut -c1-300
NMSM_ROWS=4,6 timeout 300 python tests/bench_configs.py --fixed-base > gpurun_out/configs_v12_mul.jsonl 2>/dev/null; cat gpurun_out/configs_v12_mul.jsonl | cut -c1-600
NMSM_K=4 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed-base > gpurun_out/bench_v12_K4.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed-base > gpurun_out/bench_v12.json 2>/dev/null
