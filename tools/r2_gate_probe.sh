#!/bin/bash
# focused check of the rewritten NMS reduce / point-pool scan + an ncu capture of k_sir_gate in a frame
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nms or pool or refine or frame or roi" > gpurun_out/probe_pytest.txt 2>&1; tail -n 4 gpurun_out/probe_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sir_gate -s 6 -c 1 -o gpurun_out/prof_r2_sir_gate -f python tools/frame_once.py > gpurun_out/prof_r2_sir_gate.log 2>&1
tail -n 2 gpurun_out/prof_r2_sir_gate.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/probe_bench.json 2> gpurun_out/probe_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/probe_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "stage_ms", d["stage_ms"]); print("breach", d["parity"]["breach"], d["parity"]["index_mismatches"])
PY
