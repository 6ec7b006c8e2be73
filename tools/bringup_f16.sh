#!/bin/bash
# First GPU call for the experimental fp16-split gather-GEMM and the gather micro-benchmark:
#   gpurun --timeout 1800 -- 'bash tools/bringup_f16.sh'
# Everything is wrapped in its own timeout: a hang in an untested kernel must not hold the box.
mkdir -p gpurun_out
[ -x tools/microbench/gather_paths ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/gather_paths tools/microbench/gather_paths.cu
timeout 120 tools/microbench/gather_paths > gpurun_out/gather_paths.txt 2>&1
FSFB_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_gemm_f16.py -x -q > gpurun_out/f16_test.txt 2>&1
echo "f16 test exit $?" >> gpurun_out/f16_test.txt
FSFB_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_loading.py -x -q -k hwc16_projection > gpurun_out/hwc16_test.txt 2>&1
# the four tests of validated paths that were written after the GPU budget ran out (un-gate them once they pass)
FSFB_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest -q -m gpu tests/test_loading.py tests/test_shim_autograd.py tests/test_gpu_frame.py \
  tests/test_gpu_modules.py tests/test_conv_autograd.py tests/test_plane_split.py -k "plane_split_on_device or disk_to_ids or backward_on_device or simple_test_entry or vote_seg_head_reference" > gpurun_out/new_tests.txt 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
FSFB_PLANE_SPLIT=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_planes.json 2> gpurun_out/bench_planes.err
FSFB_ROW_KEY=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tf32_oldkey.json 2> gpurun_out/bench_tf32_oldkey.err
if grep -q "F16 OK" gpurun_out/f16_test.txt || grep -q "1 passed" gpurun_out/f16_test.txt; then
  FSFB_GEMM_F16=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
fi
if grep -q "1 passed" gpurun_out/hwc16_test.txt; then timeout 200 python tools/op_bench_hwc.py > gpurun_out/op_bench_hwc.json 2>&1; fi
tail -n 30 gpurun_out/gather_paths.txt gpurun_out/f16_test.txt gpurun_out/hwc16_test.txt gpurun_out/new_tests.txt
cat gpurun_out/bench_tf32.json gpurun_out/bench_planes.json gpurun_out/bench_tf32_oldkey.json gpurun_out/bench_f16.json 2>/dev/null
