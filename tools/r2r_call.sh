#!/bin/bash
# profiling pass of round 2: launch list of the bench command, one headline step under ncu (sections + DRAM bytes), the
# decoder / head kernels under ncu --set full
mkdir -p gpurun_out
export CFP_BENCH_NO_SAMPLER=1
CFP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2r_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2r_ncu_bench.log 2>&1; echo "launch list rc=$?"
N=$(python tools/prof_step.py 1 | tail -n 1 | sed 's/.*: \([0-9]*\) libcfp.*/\1/'); echo "launches per step: $N"
timeout 900 ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section SourceCounters --section SchedulerStats \
    --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --import-source on --clock-control none \
    --launch-skip $N --launch-count $N -o gpurun_out/r2r_step -f python tools/prof_step.py 2 > gpurun_out/r2r_step.log 2>&1; echo "step capture rc=$?"
python tools/ncu_export.py gpurun_out/r2r_step.ncu-rep --keep-rep-below-mb 0
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_gen_tc|head_expect|upsample_concat|channel_mean|head_regressor|posenc_tokens_nhwc|copy_channels" \
    --launch-skip 22 --launch-count 22 -o gpurun_out/r2r_tail -f python tools/prof_tail.py 2 16 > gpurun_out/r2r_tail.log 2>&1; echo "tail capture rc=$?"
python tools/ncu_export.py gpurun_out/r2r_tail.ncu-rep --keep-rep-below-mb 0
ls -la gpurun_out/r2r_*
