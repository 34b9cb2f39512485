set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_v10_2gpu.json 2> gpurun_out/bench_v10_2gpu.err
tail -3 gpurun_out/bench_v10_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_v10_2gpu_ref.json 2> gpurun_out/bench_v10_2gpu_ref.err
tail -2 gpurun_out/bench_v10_2gpu_ref.err
