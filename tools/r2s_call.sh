#!/bin/bash
mkdir -p gpurun_out
export CFP_BENCH_NO_SAMPLER=1
K='regex:_tc_kernel|layout_|dw_plane_pack|dwconv_bn_relu|sr_bias_ln|loftr_query|kv_state|hist_encoder|lkpm_mlp|canvas_resize'
timeout 900 ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section SourceCounters --section SchedulerStats \
    --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --import-source on --clock-control none \
    -k "$K" --launch-skip 95 --launch-count 95 -o gpurun_out/r2s_step -f python tools/prof_step.py 2 > gpurun_out/r2s_step.log 2>&1; echo "step capture rc=$?"
python tools/ncu_export.py gpurun_out/r2s_step.ncu-rep --keep-rep-below-mb 0
python __graft_entry__.py --smoke > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 12 gpurun_out/r2s_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r2s_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -n 3 gpurun_out/r2s_gpu_tests.log
