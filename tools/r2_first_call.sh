#!/bin/bash
# First GPU call of round 2 (nothing of this could be run in round 1: the GPU budget was spent).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# 1. the 6x6-zone cases alone (fusion cases confirmed by tools/z6_quick.py at the end of round 1; the mask-export
#    cases of that layout run here for the first time);
# 2. the whole GPU suite;
# 3. smoke() (fp32 + bf16 engines against the oracle) and the default bench line;
# Everything lands in gpurun_out/ so it comes back.
mkdir -p gpurun_out
CFP_TEST_EXTRA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -k z6 -q > gpurun_out/r2_z6.log 2>&1
echo "z6 rc=$?" | tee -a gpurun_out/r2_z6.log
timeout 600 python -m pytest tests -m gpu -q -x -rxX > gpurun_out/r2_gpu_tests.log 2>&1
echo "suite rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2_smoke.log
timeout 300 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -3 gpurun_out/r2_z6.log gpurun_out/r2_gpu_tests.log
head -c 600 gpurun_out/r2_bench.json
