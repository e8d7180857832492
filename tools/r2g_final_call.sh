#!/bin/bash
# final artefacts of round 2: GPU suite, smoke, every bench workload, launch list + one step under ncu
mkdir -p gpurun_out
P=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q -rxXs > ${P}_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -n 3 ${P}_gpu_tests.log
python __graft_entry__.py --smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 ${P}_smoke.log
timeout 600 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
CFP_SEQUENTIAL_LEVELS=1 timeout 300 python bench.py --no-cpu > ${P}_bench_seq.json 2> ${P}_bench_seq.err; echo "bench seq rc=$?"
for w in baseline_b16 latency_480 train_b32 tail_b16; do timeout 400 python bench.py --workload $w > ${P}_bench_$w.json 2> ${P}_bench_$w.err; echo "bench $w rc=$?"; done
timeout 400 python bench.py --dtype f32 --no-cpu > ${P}_bench_f32.json 2> ${P}_bench_f32.err; echo "bench f32 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d.get("metric"), round(d.get("value",0),1), d.get("unit"), "ms/step", round(d.get("ms_per_step",0),3), "e2e", round((d.get("e2e") or {}).get("value",0),1), "roofline", (d.get("roofline") or {}).get("kernel"), round((d.get("roofline") or {}).get("frac") or 0,3))
    except Exception as e: print(f, "failed", e)
PY
export CFP_BENCH_NO_SAMPLER=1
CFP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${P}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > ${P}_ncu_bench.log 2>&1; echo "launch list rc=$?"
N=$(python tools/prof_step.py 1 | tail -n 1 | sed 's/.*: \([0-9]*\) libcfp.*/\1/'); echo "launches per step: $N"
K='regex:_tc_kernel|_tma_kernel|layout_|dw_plane_pack|dwconv_bn_relu|sr_bias_ln|loftr_query|kv_state|hist_encoder|lkpm_mlp|canvas_resize'
timeout 900 ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section SourceCounters --section SchedulerStats \
    --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --import-source on --clock-control none \
    -k "$K" --launch-skip $N --launch-count $N -o ${P}_step -f python tools/prof_step.py 2 > ${P}_step.log 2>&1; echo "step capture rc=$?"
python tools/ncu_export.py ${P}_step.ncu-rep --keep-rep-below-mb 0
ls -la gpurun_out/r2g_* | head -40
