#!/bin/bash
# ncu evidence of round 2: (1) full capture of the dominant kernel (SubM 27x128->128 on the level-0 voxels + a Linear),
# (2) full captures of the scatter / projection kernels BASELINE's metric names, (3) the launch list of one steady-state frame.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gather_gemm_ss|k_linear_ss" -s 2 -c 2 -o gpurun_out/prof_r2_gemm_ss -f python tools/ss_profile_target.py > gpurun_out/prof_r2_gemm_ss.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_segreduce|k_project_sample_select" -s 4 -c 2 -o gpurun_out/prof_r2_scatter_proj -f python tools/profile_targets.py all > gpurun_out/prof_r2_scatter_proj.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "frame/" --csv --log-file gpurun_out/r2_launches_frame.csv python tools/frame_once.py > gpurun_out/r2_launches_frame.log 2>&1
tail -n 3 gpurun_out/prof_r2_gemm_ss.log gpurun_out/prof_r2_scatter_proj.log gpurun_out/r2_launches_frame.log
wc -l gpurun_out/r2_launches_frame.csv
