#!/bin/bash
# Round-2 profiles of the final build (run on the GPU box from the repo root): launch lists + ncu --set full of the top kernels.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c3_n16384.csv python scratch/profile_step.py C3 16384 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c3_n512.csv python scratch/profile_step.py C3 512 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c2_n4096.csv python scratch/profile_step.py C2 4096 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:gemm_tch2p -c 4 -o gpurun_out/r02_prof_tc -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"kdir_fwd_v4|kdir_bwd_v4|dA_half" -c 4 -o gpurun_out/r02_prof_kdir -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:kdir_fwd_v4 -s 1 -c 1 -o gpurun_out/r02_prof_kdir_konly -f python scratch/profile_kdir_only.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:potrf_cluster -s 8 -c 1 -o gpurun_out/r02_prof_potrf -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"gemm64_async" -s 30 -c 2 -o gpurun_out/r02_prof_tail -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
ls -la gpurun_out/r02_*
