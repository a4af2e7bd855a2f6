#!/bin/bash
# round 2, call 4: warp-per-chain ICAR solve (tests first), full GPU suite, bench, configs, ncu of the new ICAR kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_logit_icar.py -x -q > gpurun_out/r02d_pytest_icar.log 2>&1; echo "icar pytest rc=$?" >> gpurun_out/r02d_pytest_icar.log
tail -25 gpurun_out/r02d_pytest_icar.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_logit_icar.py > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
python bench.py --steps 30 --warmup 20 --cpu-iters 40 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 800 gpurun_out/r02d_bench.err
python benchmarks/configs.py --only c2,c3,c4 --steps 20 --warmup 10 > gpurun_out/r02d_configs.jsonl 2> gpurun_out/r02d_configs.err; tail -c 500 gpurun_out/r02d_configs.err
OCCU_ICAR_KERNEL=cta python benchmarks/configs.py --only c4 --steps 10 --warmup 5 > gpurun_out/r02d_configs_c4_cta.jsonl 2>&1
NCU="ncu --set full --clock-control none --import-source on"
exp() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
P="python benchmarks/configs.py --steps 3 --warmup 2"
$P --only c4 > gpurun_out/r02d_c4_plain.log 2>&1 && \
  $NCU -k regex:solve_logit_icar -s 2 -c 1 -o gpurun_out/r02d_icar $P --only c4 > gpurun_out/r02d_c4_ncu.log 2>&1; exp r02d_icar
du -sh gpurun_out
