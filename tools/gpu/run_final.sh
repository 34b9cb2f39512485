set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
NMSM_ROWS=7,8,9 timeout 600 python tests/bench_configs.py --fixed-base > gpurun_out/configs_ntt_v3.jsonl 2> gpurun_out/configs_ntt_v3.err
timeout 600 python tests/bench_configs.py > gpurun_out/configs_final.jsonl 2> gpurun_out/configs_final.err
