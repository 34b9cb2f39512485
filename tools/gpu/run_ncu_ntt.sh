set -x
mkdir -p gpurun_out
NMSM_ROWS=7 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_ntt.csv python tests/bench_configs.py --fixed-base > gpurun_out/ncu_ntt_launch.log 2>&1
NMSM_ROWS=7 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 2 -c 2 -o gpurun_out/prof_ntt_pass -f python tests/bench_configs.py --fixed-base > gpurun_out/ncu_ntt_full.log 2>&1
NMSM_ROWS=4 ncu --set full --clock-control none --import-source on -k regex:k_table_mul -s 1 -c 1 -o gpurun_out/prof_table_mul -f python tests/bench_configs.py --fixed-base > gpurun_out/ncu_table_mul.log 2>&1
ls -la gpurun_out/*.ncu-rep
