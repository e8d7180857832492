#!/bin/bash
# First GPU call of round 2: the state round 1 ended in, measured once more so that every later change has a baseline
# taken on this round's boxes.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/r2_first_call.sh'
mkdir -p gpurun_out
CFP_TEST_EXTRA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -k z6 -q > gpurun_out/r2a_z6.log 2>&1
echo "z6 rc=$?" | tee -a gpurun_out/r2a_z6.log
timeout 700 python -m pytest tests -m gpu -q -x -rxX > gpurun_out/r2a_gpu_tests.log 2>&1
echo "suite rc=$?" | tee -a gpurun_out/r2a_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
CFP_SEQUENTIAL_LEVELS=1 timeout 300 python bench.py --no-cpu > gpurun_out/r2a_bench_seq.json 2> gpurun_out/r2a_bench_seq.err
timeout 300 python bench.py --workload baseline_b16 > gpurun_out/r2a_bench_baseline_b16.json 2> gpurun_out/r2a_bench_baseline_b16.err
timeout 300 python bench.py --workload latency_480 > gpurun_out/r2a_bench_latency_480.json 2> gpurun_out/r2a_bench_latency_480.err
tail -3 gpurun_out/r2a_z6.log gpurun_out/r2a_gpu_tests.log
head -c 400 gpurun_out/r2a_bench.json; echo
head -c 300 gpurun_out/r2a_bench_seq.json; echo
cat gpurun_out/r2a_bench_baseline_b16.json gpurun_out/r2a_bench_latency_480.json
