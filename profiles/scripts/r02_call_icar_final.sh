#!/bin/bash
# after the ICAR solve rework: all GPU tests, config 4 (+ chain-count series), ncu --set full of the solve
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; tail -2 gpurun_out/r02r_pytest.log
python benchmarks/configs.py --only c4 --c4-chains 148,296,592,1024,2048,4096 --steps 20 --warmup 10 > gpurun_out/r02r_c4_series.jsonl 2> gpurun_out/r02r_c4.err
P="python benchmarks/configs.py --steps 6 --warmup 4"
ncu --set full --clock-control none --import-source on -k "regex:solve_logit_icar_warp" -s 6 -c 1 -o gpurun_out/r02r_icar $P --only c4 > gpurun_out/r02r_c4_ncu.log 2>&1
ncu -i gpurun_out/r02r_icar.ncu-rep --page raw --csv > gpurun_out/r02r_icar_raw.csv 2>/dev/null
rm -f gpurun_out/r02r_icar.ncu-rep
cut -c1-90,225-420 gpurun_out/r02r_c4_series.jsonl
