#!/bin/bash
# Parametrised GPU-side runner (replaces the round-1 one-off run_v*.sh scripts).
#   tools/gpu/run.sh <task> [args...]      — always invoked through gpurun from the repo root; writes under gpurun_out/
# tasks: micro | tests [pytest-args] | bench [bench-args] | configs | ncu-launches [bench-args] | ncu-full <kernel-regex> [bench-args]
#        sanitizer <tool> | dist <nproc> [bench-args]
set -u
mkdir -p gpurun_out
task=${1:-tests}; shift || true
case "$task" in
  micro)
    for b in tools/gpu/micro/*_nl tools/gpu/micro/*_inl; do [ -x "$b" ] && { echo "== $b"; timeout 300 "$b"; } ; done > gpurun_out/micro.jsonl 2>&1 ;;
  tests)
    timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests.log; tail -5 gpurun_out/gpu_tests.log ;;
  bench)
    timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json ;;
  configs)
    timeout 900 python bench.py --configs "$@" > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "rc=$?"; cat gpurun_out/configs.json ;;
  ncu-launches)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fixed-base "$@" > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?" ;;
  ncu-full)
    pat=$1; shift   # one capture of the 2nd launch whose name matches (the first is a cold start)
    timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 1 -c 1 -f -o gpurun_out/prof_$pat \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fixed-base "$@" > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" ;;
  sanitizer)
    tool=${1:-memcheck}; shift || true
    timeout 1500 compute-sanitizer --tool "$tool" python -m pytest tests -m gpu -x -q -k "${1:-boundary_soak or degenerate}" > gpurun_out/sanitizer_$tool.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/sanitizer_$tool.log ;;
  dist)
    np=$1; shift
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$np" --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus "$np" "$@" > gpurun_out/bench_g$np.json 2> gpurun_out/bench_g$np.err; echo "rc=$?"; cat gpurun_out/bench_g$np.json ;;
  distworker)
    np=${1:-2}
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$np" --master-addr 127.0.0.1 --master-port 29519 \
      tests/dist_worker.py > gpurun_out/dist_worker_g$np.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/dist_worker_g$np.log ;;
  sweep)
    timeout 1500 python tools/gpu/sweep_lk.py "$@" > gpurun_out/sweep_lk.jsonl 2> gpurun_out/sweep_lk.err; echo "rc=$?"; cat gpurun_out/sweep_lk.jsonl ;;
  tracedist)
    np=${1:-2}
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$np" --master-addr 127.0.0.1 --master-port 29521 \
      tools/gpu/trace_dist.py > gpurun_out/trace_dist_g$np.log 2>&1; echo "rc=$?"; grep "nmsm trace" gpurun_out/trace_dist_g$np.log | tail -$((2*np)) ;;
  ncu-shard1)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_shard1.csv \
      python tools/gpu/shard1_profile.py "$@" > gpurun_out/ncu_shard1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_shard1.log ;;
  sweepc)
    timeout 1500 python tools/gpu/sweep_c.py "$@" > gpurun_out/sweep_c.jsonl 2> gpurun_out/sweep_c.err; echo "rc=$?"; cat gpurun_out/sweep_c.jsonl ;;
  sweeptail)
    timeout 1500 python tools/gpu/sweep_tail.py "$@" > gpurun_out/sweep_tail.jsonl 2> gpurun_out/sweep_tail.err; echo "rc=$?"; cat gpurun_out/sweep_tail.jsonl ;;
  explore)
    timeout 900 python tools/gpu/explore_groups.py "$@" > gpurun_out/explore.jsonl 2> gpurun_out/explore.err; echo "rc=$?"; cat gpurun_out/explore.jsonl ;;
  *) echo "unknown task $task"; exit 2 ;;
esac
