#!/bin/bash
# ncu --set full of the probit site pass at 4096 chains, raw + source pages as CSV
cd "$GRAFT_REPO_ROOT"
P="python benchmarks/configs.py --steps 6 --warmup 4"
$P --only c3 > gpurun_out/r02p_c3_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "regex:site_pass_probit" -s 14 -c 1 -o gpurun_out/r02p_probit $P --only c3 > gpurun_out/r02p_c3_ncu.log 2>&1
ncu -i gpurun_out/r02p_probit.ncu-rep --page raw --csv > gpurun_out/r02p_probit_raw.csv 2>/dev/null
ncu -i gpurun_out/r02p_probit.ncu-rep --page source --csv > gpurun_out/r02p_probit_source.csv 2>/dev/null
rm -f gpurun_out/r02p_probit.ncu-rep
tail -2 gpurun_out/r02p_c3_plain.log | cut -c1-400
