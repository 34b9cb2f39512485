set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v7.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
for K in 4 16; do NMSM_K=$K python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_K$K.json 2>/dev/null; done
for L in 16 24 48; do NMSM_L=$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_L$L.json 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v7.json 2>/dev/null
tail -3 gpurun_out/pytest_gpu_v7.log
