mkdir -p gpurun_out
timeout 90 python tools/ss_smoke.py 2>&1 | tail -3
for hv in 1 0; do
  echo "=== FSFB_GEMM_HV=$hv"
  FSFB_GEMM_HV=$hv timeout 200 python tools/gemm_ss_timers.py 2>&1 | grep "== "
  FSFB_GEMM_HV=$hv timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_hv$hv.json 2> gpurun_out/bench_hv$hv.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_hv$hv.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["kernels"]["gather_gemm_conv"]["ms_per_frame"], d["kernels"]["gather_gemm_linear"]["ms_per_frame"], d["stage_ms"])
except Exception as e:
    print("bench failed", e, open("gpurun_out/bench_hv$hv.err").read()[-2000:])
PY
done
