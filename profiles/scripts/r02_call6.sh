#!/bin/bash
# round 2, call 6: ICAR warp kernel v2 (tests + probe first), full GPU suite, bench, configs, ncu of the ICAR kernel and both site passes
mkdir -p gpurun_out
python benchmarks/icar_debug_probe.py > gpurun_out/r02f_probe.log 2>&1; tail -12 gpurun_out/r02f_probe.log
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -25 gpurun_out/r02f_pytest.log
python bench.py --steps 30 --warmup 20 --cpu-iters 60 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 600 gpurun_out/r02f_bench.err
python benchmarks/configs.py --only c1,c2,c3,c4 --steps 20 --warmup 10 > gpurun_out/r02f_configs.jsonl 2> gpurun_out/r02f_configs.err; tail -c 500 gpurun_out/r02f_configs.err
NCU="ncu --set full --clock-control none --import-source on"
exp() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
P="python benchmarks/configs.py --steps 3 --warmup 2"
$P --only c4 > gpurun_out/r02f_c4_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_icar -s 2 -c 1 -o gpurun_out/r02f_icar $P --only c4 > gpurun_out/r02f_c4_ncu.log 2>&1; exp r02f_icar
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02f_bench_plain.log 2>&1 && \
  $NCU -k regex:site_pass_logit -s 3 -c 1 -o gpurun_out/r02f_site python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02f_site_ncu.log 2>&1; exp r02f_site
$P --only c3 > gpurun_out/r02f_c3_plain.log 2>&1 && \
  $NCU -k "regex:site_pass_probit" -s 10 -c 1 -o gpurun_out/r02f_probit $P --only c3 > gpurun_out/r02f_c3_ncu.log 2>&1; exp r02f_probit
du -sh gpurun_out
