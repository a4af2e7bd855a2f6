#!/bin/bash
mkdir -p gpurun_out
python benchmarks/icar_debug_probe.py > gpurun_out/r02e_probe.log 2>&1; tail -30 gpurun_out/r02e_probe.log
timeout 900 python -m pytest tests/test_gpu_logit_icar.py -q --tb=line > gpurun_out/r02e_pytest_icar.log 2>&1; tail -15 gpurun_out/r02e_pytest_icar.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python benchmarks/icar_debug_probe.py > gpurun_out/r02e_racecheck.log 2>&1; tail -40 gpurun_out/r02e_racecheck.log
python -c "
import json
b=json.load(open('gpurun_out/r02d_bench.json')) if False else None
"
