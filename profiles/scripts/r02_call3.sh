#!/bin/bash
# round 2, call 3: GPU test-suite + bench line + configs with site pass v2 (smem-staged tile CSR, FP32 fused psi), the fast FP32
# probit draws, the restructured sample() (device summaries, multi-engine, packed start upload), then ncu of both site passes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -15 gpurun_out/r02c_pytest.log
python bench.py --steps 30 --warmup 20 --cpu-iters 40 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 1500 gpurun_out/r02c_bench.err
python benchmarks/configs.py --only c2,c3,c4 --steps 20 --warmup 10 > gpurun_out/r02c_configs.jsonl 2> gpurun_out/r02c_configs.err; tail -c 500 gpurun_out/r02c_configs.err
NCU="ncu --set full --clock-control none --import-source on"
exp() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02c_bench_plain.log 2>&1 && \
  $NCU -k regex:site_pass_logit -s 3 -c 1 -o gpurun_out/r02c_site python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02c_site_ncu.log 2>&1; exp r02c_site
P="python benchmarks/configs.py --steps 3 --warmup 2"
$P --only c3 > gpurun_out/r02c_c3_plain.log 2>&1 && \
  $NCU -k "regex:site_pass_probit" -s 10 -c 1 -o gpurun_out/r02c_probit $P --only c3 > gpurun_out/r02c_c3_ncu.log 2>&1; exp r02c_probit
du -sh gpurun_out
