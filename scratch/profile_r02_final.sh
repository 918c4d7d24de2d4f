#!/bin/bash
# Round-2 profiles of the FINAL build, summarised on the GPU box (the .ncu-rep files together exceed what gpurun copies back):
# launch lists + ncu --set full of the top kernels -> gpurun_out/final_*.txt (copied to profiles/r02_*.txt afterwards).
set -x
O=gpurun_out; T=/tmp/prof; mkdir -p $O $T
NCU="ncu --clock-control none --profile-from-start off"
H="ncu --set full --clock-control none, one C3 training step (n=16384, fp32, 3xFP16 path), round-2 final build (end of round); per launch --"
for w in "C3 16384" "C3 512" "C2 4096"; do
  set -- $w; tag=$(echo $1 | tr A-Z a-z)_n$2
  $NCU --metrics gpu__time_duration.sum --csv --log-file $T/launches_$tag.csv python scratch/profile_step.py $1 $2 > /dev/null 2>&1
  python scratch/summarize_launches.py $T/launches_$tag.csv > $O/final_launches_$tag.txt
done
$NCU --set full -k regex:gemm_tch2p -c 4 -o $T/tc -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
python scratch/ncu_summary.py $T/tc.ncu-rep "$H the four persistent 3xFP16 products in step order: A = W K_zx, C = (S-I) A, G = A_g A^T (stream-K), dK_zx = W^T dA" > $O/final_ncu_prof_tc.txt
$NCU --set full -k regex:"kdir_fwd_v4|kdir_bwd_v4|dA_half" -c 4 -o $T/kdir -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
python scratch/ncu_summary.py $T/kdir.ncu-rep "$H kdir_fwd_v4 (generic + canonical variant, half-split output), dA_half_kernel, kdir_bwd_v4 (cp.async-staged upstream rows, exact d = 10)" > $O/final_ncu_prof_kdir.txt
$NCU --set full -k regex:kdir_fwd_v4 -s 1 -c 1 -o $T/konly -f python scratch/profile_kdir_only.py > /dev/null 2>&1
python scratch/ncu_summary.py $T/konly.ncu-rep "ncu --set full --clock-control none, kdir_fwd_v4 writing K only (C3 shapes, canonical directions: the <2,2,0,1> variant): what RBFKernelDirectionalGrad.forward returns; round-2 final build (end of round)" > $O/final_ncu_prof_kdir_konly.txt
$NCU --set full -k regex:potrf_cluster -s 8 -c 1 -o $T/potrf -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
python scratch/ncu_summary.py $T/potrf.ncu-rep "$H potrf_cluster (diagonal block 9 of 32)" > $O/final_ncu_prof_potrf.txt
$NCU --set full -k regex:"gemm64_async" -s 30 -c 2 -o $T/tail -f python scratch/profile_step.py C3 16384 > /dev/null 2>&1
python scratch/ncu_summary.py $T/tail.ncu-rep "$H gemm64_async_kernel: the two fp64 products of the Cholesky-backward tail" > $O/final_ncu_prof_tail.txt
ls -la $O/final_*
