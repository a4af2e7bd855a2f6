#!/bin/bash
# round 2, final evidence: full GPU suite, bench (default arguments) + reference arm, all configs, the default call at
# n = 10^4 with streamed Z planes, launch list of the bench command, ncu --set full of the step's kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log
tail -4 gpurun_out/r02z_pytest.log
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; tail -c 300 gpurun_out/r02z_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02z_bench_reference_arm.json 2> gpurun_out/r02z_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_k20.json 2> gpurun_out/r02z_bench_k20.err
python benchmarks/configs.py --only c1,c2,c3,c4,c5,q --steps 20 --warmup 10 > gpurun_out/r02z_configs.jsonl 2> gpurun_out/r02z_configs.err; tail -c 300 gpurun_out/r02z_configs.err
python benchmarks/default_threshold_10k.py 1024 > gpurun_out/r02z_default_threshold_1024.json 2> gpurun_out/r02z_dt1024.err; tail -c 300 gpurun_out/r02z_dt1024.err
python benchmarks/default_threshold_10k.py 4096 > gpurun_out/r02z_default_threshold_4096.json 2> gpurun_out/r02z_dt4096.err; tail -c 300 gpurun_out/r02z_dt4096.err
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r02z_bench_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv $B > gpurun_out/r02z_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
exp() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
$B > gpurun_out/r02z_bench_plain2.log 2>&1 && \
  $NCU -k "regex:site_pass_logit|gram_tc_kernel|solve_logit_rsr_kernel|reduce_partials|alpha_tau" -s 15 -c 5 -o gpurun_out/r02z_step $B > gpurun_out/r02z_step_ncu.log 2>&1; exp r02z_step
P="python benchmarks/configs.py --steps 3 --warmup 2"
$P --only c3 > gpurun_out/r02z_c3_plain.log 2>&1 && \
  $NCU -k "regex:site_pass_probit" -s 10 -c 1 -o gpurun_out/r02z_probit $P --only c3 > gpurun_out/r02z_c3_ncu.log 2>&1; exp r02z_probit
du -sh gpurun_out; ls gpurun_out | head -40
