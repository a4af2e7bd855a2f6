#!/bin/bash
# round 2, call 9: full GPU suite with the final kernels, configs, then the 2-GPU checks are a separate call
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -12 gpurun_out/r02h_pytest.log
python benchmarks/configs.py --only c1,c2,c3,c4 --steps 20 --warmup 10 > gpurun_out/r02h_configs.jsonl 2> gpurun_out/r02h_configs.err; tail -c 300 gpurun_out/r02h_configs.err
python -c "
import json
for l in open('gpurun_out/r02h_configs.jsonl'):
    d=json.loads(l); print(d['config'], round(d['chain_iterations_per_s']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, d['failed_chains'])
"
