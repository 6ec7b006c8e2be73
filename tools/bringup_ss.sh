#!/bin/bash
# Bring-up of the shared-memory-operand gather-GEMM (csrc/gemm_ss.cu): smoke + parity tests first (each under its own timeout:
# a hang in a new kernel must not hold the box), then role timers, the per-shape op table and a short bench.
mkdir -p gpurun_out
timeout 90 python tools/ss_smoke.py > gpurun_out/ss_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/ss_smoke.txt
tail -n 4 gpurun_out/ss_smoke.txt
grep -q "SMOKE OK" gpurun_out/ss_smoke.txt || exit 0
timeout 120 python tools/gemm_role_timers.py > gpurun_out/ss_timers.txt 2>&1
cat gpurun_out/ss_timers.txt
if [ "$1" != "quick" ]; then
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/ss_gemm_test.txt 2>&1; echo "exit $?" >> gpurun_out/ss_gemm_test.txt
tail -n 5 gpurun_out/ss_gemm_test.txt
timeout 600 python -m pytest -m gpu tests/test_gpu_gemm_f16.py tests/test_gpu_conv.py tests/test_gpu_modules.py tests/test_gpu_frame.py -x -q > gpurun_out/ss_more_tests.txt 2>&1; echo "exit $?" >> gpurun_out/ss_more_tests.txt
tail -n 5 gpurun_out/ss_more_tests.txt
fi
timeout 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes_ss.txt 2>&1
head -n 40 gpurun_out/gemm_shapes_ss.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ss.json 2> gpurun_out/bench_ss.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ss.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["roofline"]["us_per_launch"], d["kernels"]["gather_gemm_conv"], d["kernels"]["gather_gemm_linear"])
except Exception as e:
    print("bench failed", e, open("gpurun_out/bench_ss.err").read()[-2000:])
PY
