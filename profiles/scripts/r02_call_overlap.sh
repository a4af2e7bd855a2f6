#!/bin/bash
# overlap probe: E engines on separate streams of one device, Gram kernel with fewer smem stages (room for site-pass CTAs)
# and on a high-priority stream; then the bench at the driver's arguments
cd "$GRAFT_REPO_ROOT"
out=gpurun_out/overlap_probe.jsonl
: > $out
for lib in default s3 s2; do
  for prio in 0 1; do
    if [ $lib = default ]; then unset OCCU_B200_LIB; else export OCCU_B200_LIB=$PWD/occuspytial_b200/csrc/variants/liboccu_b200_$lib.so; fi
    timeout 300 python benchmarks/overlap_probe.py 1 2 4 >> $out 2>> gpurun_out/overlap_probe.err
  done
done
unset OCCU_B200_LIB
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err
cat $out
