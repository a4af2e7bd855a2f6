#!/bin/bash
# round 2, call 7: ICAR warp kernel v3 + streamed Z planes: targeted tests first, then the suite, c4 timing and ncu
mkdir -p gpurun_out
python benchmarks/icar_debug_probe.py > gpurun_out/r02g_probe.log 2>&1; tail -6 gpurun_out/r02g_probe.log
timeout 900 python -m pytest tests/test_gpu_logit_icar.py -q --tb=short > gpurun_out/r02g_pytest_icar.log 2>&1; tail -5 gpurun_out/r02g_pytest_icar.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_logit_rsr.py -q --tb=short -k "tcgen05_gram or streamed or large_rank or default_threshold" > gpurun_out/r02g_pytest_stream.log 2>&1; tail -15 gpurun_out/r02g_pytest_stream.log
python benchmarks/configs.py --only c1,c4 --steps 20 --warmup 10 > gpurun_out/r02g_configs.jsonl 2> gpurun_out/r02g_configs.err; tail -c 300 gpurun_out/r02g_configs.err; cat gpurun_out/r02g_configs.jsonl | cut -c1-400
NCU="ncu --set full --clock-control none --import-source on"
exp() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
P="python benchmarks/configs.py --steps 3 --warmup 2"
$P --only c4 > gpurun_out/r02g_c4_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_icar -s 2 -c 1 -o gpurun_out/r02g_icar $P --only c4 > gpurun_out/r02g_c4_ncu.log 2>&1; exp r02g_icar
du -sh gpurun_out
