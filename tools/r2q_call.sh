#!/bin/bash
# CUDA-graph replay of the headline step: parity test, then graph on/off x levels concurrent/sequential x micro-batches
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k graph 2>&1 | tail -n 5
run() { tag=$1; shift; env "$@" timeout 240 python bench.py --no-cpu > gpurun_out/r2q_$tag.json 2> gpurun_out/r2q_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2q_$tag.json").read())
    print("$tag", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$tag", "failed", e); print(open("gpurun_out/r2q_$tag.err").read()[-600:])
PY
}
run eager_conc CFP_GRAPH=0
run graph_conc CFP_GRAPH=1
run graph_conc_m221 CFP_GRAPH=1 CFP_MICRO=128:2,64:2,32:1
run graph_conc_m421 CFP_GRAPH=1 CFP_MICRO=128:4,64:2,32:1
run graph_conc_m422 CFP_GRAPH=1 CFP_MICRO=128:4,64:2,32:2
run eager_seq CFP_GRAPH=0 CFP_SEQUENTIAL_LEVELS=1
run graph_seq CFP_GRAPH=1 CFP_SEQUENTIAL_LEVELS=1
run graph_seq_m221 CFP_GRAPH=1 CFP_SEQUENTIAL_LEVELS=1 CFP_MICRO=128:2,64:2,32:1
run graph_seq_m422 CFP_GRAPH=1 CFP_SEQUENTIAL_LEVELS=1 CFP_MICRO=128:4,64:2,32:2
run graph_seq_m842 CFP_GRAPH=1 CFP_SEQUENTIAL_LEVELS=1 CFP_MICRO=128:8,64:4,32:2
timeout 200 python bench.py --workload tail_b16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('tail', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'fps e2e', round(d['e2e']['value']))
for k,v in list(d['kernels'].items())[:8]: print(f'{k:32s} n={v[\"launches_per_step\"]:5.1f} {v[\"ms_per_step\"]:.3f} ms  {v.get(\"bound\")} {v.get(\"roofline_frac\")}')
"
