set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fixed-base > gpurun_out/bench_stdout_1gpu.json 2> gpurun_out/bench_stdout_1gpu.err
wc -l gpurun_out/bench_stdout_1gpu.json; head -c 60 gpurun_out/bench_stdout_1gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_final_2gpu.json 2> gpurun_out/bench_final_2gpu.err
wc -l gpurun_out/bench_final_2gpu.json; head -c 60 gpurun_out/bench_final_2gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_final_2gpu_ref.json 2>/dev/null
wc -l gpurun_out/bench_final_2gpu_ref.json
