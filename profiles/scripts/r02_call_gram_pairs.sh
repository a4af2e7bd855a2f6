#!/bin/bash
# Gram slice-pair variants: error against long-double NumPy and time per launch
cd "$GRAFT_REPO_ROOT"
out=gpurun_out/gram_pairs.jsonl
: > $out
for lib in default sb5_l6_bn64 sb5_l6_bn80 sb5_l7_bn64 sb6_l6_bn64; do
  if [ $lib = default ]; then unset OCCU_B200_LIB; else export OCCU_B200_LIB=$PWD/occuspytial_b200/csrc/variants/liboccu_b200_$lib.so; fi
  timeout 300 python benchmarks/gram_error_probe.py >> $out 2>> gpurun_out/gram_pairs.err
  timeout 300 python bench.py --steps 30 --warmup 10 --no-e2e --no-cpu-baseline 2>> gpurun_out/gram_pairs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$lib','value':d['value'],'ms_per_step':d['ms_per_step'],'kernels':d['kernels_ms_per_step']}))" >> $out
done
cat $out
