set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_v15_4gpu.json 2> gpurun_out/bench_v15_4gpu.err
tail -3 gpurun_out/bench_v15_4gpu.err
head -c 400 gpurun_out/bench_v15_4gpu.json
