#!/bin/bash
# round 2, call 1: ncu --set full captures of the kernels that had no committed summary in round 1
# (ICAR solve, probit site pass + its GEMMs + solve, tiled large-q solve, FP64 DMMA Gram); each after a plain run exited 0.
mkdir -p gpurun_out
P="python benchmarks/configs.py --steps 3 --warmup 2"
NCU="ncu --set full --clock-control none --import-source on"
$P --only c4 > gpurun_out/r02_c4_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_icar -s 2 -c 1 -o gpurun_out/r02_icar $P --only c4 > gpurun_out/r02_c4_ncu.log 2>&1
$P --only c3 > gpurun_out/r02_c3_plain.log 2>&1 && \
  $NCU -k "regex:site_pass_probit|gemm_tn|solve_logit_rsr" -s 40 -c 4 -o gpurun_out/r02_probit $P --only c3 > gpurun_out/r02_c3_ncu.log 2>&1
$P --only q256 > gpurun_out/r02_q256_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_rsr_tiled -s 2 -c 1 -o gpurun_out/r02_tiled $P --only q256 > gpurun_out/r02_q256_ncu.log 2>&1
$P --only c2 > gpurun_out/r02_c2_plain.log 2>&1 && \
  $NCU -k regex:gemm_tn -s 4 -c 2 -o gpurun_out/r02_dmma $P --only c2 > gpurun_out/r02_c2_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/r02_*_plain.log
