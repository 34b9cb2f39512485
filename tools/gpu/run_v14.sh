set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14.err
tail -2 gpurun_out/bench_v14.err
timeout 600 python -m pytest tests -m gpu -x -q -k "soak or sweep or giant or config_bls12_381_g1_2p16 or fixed_base_table_small" > gpurun_out/pytest_v14.log 2>&1; tail -2 gpurun_out/pytest_v14.log
