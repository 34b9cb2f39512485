set -x
mkdir -p gpurun_out
SEL="ntt_matches_oracle or ntt_errors or fixed_base_table_small or point_table_multiply or outside_the_subgroup or test_mul_batch_vs_oracle or ntt_output_feeds"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_v15.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_v15.log
tail -6 gpurun_out/sanitizer_memcheck_v15.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "ntt_errors or ntt_output_feeds or point_table_getpublickey" > gpurun_out/sanitizer_racecheck_v15.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_v15.log
tail -6 gpurun_out/sanitizer_racecheck_v15.log
