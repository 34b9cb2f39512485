set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_final_2gpu.json 2> gpurun_out/bench_final_2gpu.err
wc -l gpurun_out/bench_final_2gpu.json; head -c 60 gpurun_out/bench_final_2gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/bench_final_2gpu_strong.json 2> gpurun_out/bench_final_2gpu_strong.err
tail -c 300 gpurun_out/bench_final_2gpu_strong.err
