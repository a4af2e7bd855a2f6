#!/bin/bash
# round 2, call 10 (2 GPUs): one-process multi-GPU sample(), torchrun bench weak + strong
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_logit_icar.py -q --tb=short > gpurun_out/r02i_pytest_icar.log 2>&1; tail -3 gpurun_out/r02i_pytest_icar.log
python benchmarks/configs.py --only c4 --steps 20 --warmup 10 2>/dev/null | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_logit_rsr.py -q --tb=short -k "summary_from_device" > gpurun_out/r02i_pytest_2gpu.log 2>&1; tail -5 gpurun_out/r02i_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 20 > gpurun_out/r02i_bench_n2_weak.json 2> gpurun_out/r02i_bench_n2_weak.err; tail -c 400 gpurun_out/r02i_bench_n2_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 20 --scaling strong > gpurun_out/r02i_bench_n2_strong.json 2> gpurun_out/r02i_bench_n2_strong.err; tail -c 400 gpurun_out/r02i_bench_n2_strong.err
python - <<'PY'
import json
for f in ('weak','strong'):
    try:
        b=json.load(open(f'gpurun_out/r02i_bench_n2_{f}.json'))
        print(f, b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['timed_calls_s'], b['config']['chains_per_gpu'])
    except Exception as e: print(f, 'failed', e)
PY
