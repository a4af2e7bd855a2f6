#!/bin/bash
# ncu --set full of the ICAR solve at BASELINE config 4, raw + source pages as CSV
cd "$GRAFT_REPO_ROOT"
P="python benchmarks/configs.py --steps 6 --warmup 4"
$P --only c4 > gpurun_out/r02q_c4_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "regex:solve_logit_icar_warp" -s 6 -c 1 -o gpurun_out/r02q_icar $P --only c4 > gpurun_out/r02q_c4_ncu.log 2>&1
ncu -i gpurun_out/r02q_icar.ncu-rep --page raw --csv > gpurun_out/r02q_icar_raw.csv 2>/dev/null
ncu -i gpurun_out/r02q_icar.ncu-rep --page source --csv > gpurun_out/r02q_icar_source.csv 2>/dev/null
rm -f gpurun_out/r02q_icar.ncu-rep
tail -1 gpurun_out/r02q_c4_plain.log | cut -c1-400
