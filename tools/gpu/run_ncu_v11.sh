set -x
mkdir -p gpurun_out
# launch list of one serial MSM step (per-launch device times; numbers under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_v11.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fixed-base --in-flight 1 > gpurun_out/ncu_launch_bench.log 2>&1
# full captures: k_accumulate, k_reduce1 (serial loop, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 3 -c 1 -o gpurun_out/prof_acc_v11 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fixed-base --in-flight 1 > gpurun_out/ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reduce1 -s 3 -c 1 -o gpurun_out/prof_reduce1_v11 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fixed-base --in-flight 1 > gpurun_out/ncu_r1.log 2>&1
ls -la gpurun_out/*.ncu-rep
