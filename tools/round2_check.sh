#!/bin/bash
# Round-2 hardware check: the whole GPU suite, the smoke entry, the default bench line (full scope, parity block, library bar).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.txt 2>&1; echo "exit $?" >> gpurun_out/r2_pytest.txt
tail -n 15 gpurun_out/r2_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE ENTRY OK')" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    print("stage_ms", d["stage_ms"])
    print("parity", d.get("parity"))
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "us_per_launch", "share_of_step", "traffic")})
    print("cpu", d.get("cpu_baseline"))
    print("library", json.dumps(d.get("library_baseline"), indent=0)[:1800])
    print("frame_stats", d.get("frame_stats"))
except Exception as e:
    print("bench failed", e, open("gpurun_out/r2_bench.err").read()[-3000:])
PY
