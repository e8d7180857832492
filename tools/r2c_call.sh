#!/bin/bash
# host-side packing + smoke order: GPU suite, smoke, smoke under an ncu launch list (what the driver's window shows), bench
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r2c_gpu_tests.log 2>&1; echo "suite rc=$?"
tail -n 4 gpurun_out/r2c_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 8 gpurun_out/r2c_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r2c_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/r2c_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"
grep -c "gpu__time_duration" gpurun_out/r2c_smoke_launches.csv
timeout 300 python bench.py --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; head -c 300 gpurun_out/r2c_bench.json; echo
