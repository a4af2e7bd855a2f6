#!/bin/bash
# round 2, call 2: GPU test-suite with the new site passes, a bench line, then ncu --set full captures of the kernels
# that had no committed summary in round 1 (each after a plain run exited 0); reports are exported to CSV on the box.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
python bench.py --steps 30 --warmup 20 --cpu-iters 20 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 600 gpurun_out/r02_bench_a.err
python benchmarks/configs.py --only c2,c3,c4 --steps 20 --warmup 10 > gpurun_out/r02_configs_a.jsonl 2> gpurun_out/r02_configs_a.err
P="python benchmarks/configs.py --steps 3 --warmup 2"
NCU="ncu --set full --clock-control none --import-source on"
exp() {  # export raw + source pages, drop the report (gpurun_out is capped at 64 MiB)
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
$P --only c4 > gpurun_out/r02_c4_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_icar -s 2 -c 1 -o gpurun_out/r02_icar $P --only c4 > gpurun_out/r02_c4_ncu.log 2>&1; exp r02_icar
$P --only c3 > gpurun_out/r02_c3_plain.log 2>&1 && \
  $NCU -k "regex:site_pass_probit|gemm_tn|solve_logit_rsr" -s 30 -c 3 -o gpurun_out/r02_probit $P --only c3 > gpurun_out/r02_c3_ncu.log 2>&1; exp r02_probit
$P --only q256 > gpurun_out/r02_q256_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_rsr_tiled -s 2 -c 1 -o gpurun_out/r02_tiled $P --only q256 > gpurun_out/r02_q256_ncu.log 2>&1; exp r02_tiled
$P --only c2 > gpurun_out/r02_c2_plain.log 2>&1 && \
  $NCU -k regex:gemm_tn -s 4 -c 2 -o gpurun_out/r02_dmma $P --only c2 > gpurun_out/r02_c2_ncu.log 2>&1; exp r02_dmma
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_plain.log 2>&1 && \
  $NCU -k regex:site_pass_logit -s 3 -c 1 -o gpurun_out/r02_site python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_site_ncu.log 2>&1; exp r02_site
ls -la gpurun_out/ | head -40
du -sh gpurun_out
