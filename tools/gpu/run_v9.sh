set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err
tail -2 gpurun_out/bench_v9.err
NMSM_ROWS=4,5,6 timeout 600 python tests/bench_configs.py --fixed-base > gpurun_out/configs_point_table.jsonl 2> gpurun_out/configs_point_table.err
tail -2 gpurun_out/configs_point_table.err
