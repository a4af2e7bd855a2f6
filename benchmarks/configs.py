#!/usr/bin/env python
"""Throughput of the five BASELINE.json configs on ONE B200 (device-resident, CUDA events), one JSON line each.

    python benchmarks/configs.py [--only c2,c3,c4,c5] [--steps 30] [--warmup 20]

c1 (PGGibbs n=150, 2 chains x 2000 iterations) is the reference's CPU-runnable case: it is a parity case
(tests/test_gpu_logit_icar.py), timed here only for completeness.  Per-kernel times come from the engine's
profiling mode (CUDA events per kernel class).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(name, sampler, chains, steps, warmup, note):
    import torch
    seed = 1
    t0 = time.perf_counter()
    eng = sampler._make_engine(chains, seed)
    eng.draw_start(seed, *sampler._DEFAULT_START)
    setup_s = time.perf_counter() - t0
    fails = eng.sweep(warmup)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        e0.record(eng.stream)
        fails += eng.sweep(steps)
        e1.record(eng.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    eng.set_profiling(True)
    eng.sweep(min(steps, 10))
    kms = {k: v[0] / max(v[1], 1) for k, v in eng.kernel_ms().items()}
    eng.close()
    torch.cuda.empty_cache()
    line = {'config': name, 'note': note, 'chains': chains, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms / steps,
            'chain_iterations_per_s': chains * steps / (ms * 1e-3), 'kernels_ms_per_step': kms, 'failed_chains': int(fails),
            'engine_setup_s': setup_s}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='c1,c2,c3,c4,c5,q')
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--c5-chains', default='512,2048,8192,32768')
    ap.add_argument('--c4-chains', default='1024', help='chain counts of config 4 (1024 = BASELINE; others show how the solve scales)')
    args = ap.parse_args()
    only = set(args.only.split(','))
    from occuspytial_b200 import LogitRSRGibbs, ProbitRSRGibbs, LogitICARGibbs
    from occuspytial_b200.utils import make_data

    if 'c1' in only:
        Q, W, X, y, *_ = make_data(n=150, random_state=0)
        run('c1', LogitICARGibbs(Q, W, X, y, random_state=10), 2, 200, 50, 'PGGibbs make_data(n=150), 2 chains')
    if 'c2' in only:
        Q, W, X, y, *_ = make_data(n=1000, min_v=2, max_v=8, lattice=(25, 40), random_state=0)
        s = LogitRSRGibbs(Q, W, X, y, random_state=10)          # default threshold r = 0.5
        run('c2', s, 64, args.steps, args.warmup, f'LogitRSR n=1000 (25x40), default r=0.5 -> q={s.fixed.q}, 64 chains')
    if 'c3' in only:
        Q, W, X, y, *_ = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
        s = ProbitRSRGibbs(Q, W, X, y, random_state=10, q=64)
        run('c3', s, 512, args.steps, args.warmup, 'ProbitRSR n=10000, q=64, 512 chains (one GPU share of 4096 over 8)')
        run('c3x8', s, 4096, args.steps, args.warmup, 'ProbitRSR n=10000, q=64, 4096 chains on one GPU')
    if 'c4' in only:
        Q, W, X, y, *_ = make_data(n=2500, min_v=2, max_v=8, lattice=(50, 50), random_state=0)
        s = LogitICARGibbs(Q, W, X, y, random_state=10)
        for chains in (int(c) for c in args.c4_chains.split(',')):
            run('c4' if chains == 1024 else f'c4_{chains}', s, chains, args.steps, args.warmup,
                f'PGGibbs full ICAR n=2500 (50x50), half bandwidth {s._band()[1]}, {chains} chains')
    if 'c5' in only:
        t0 = time.perf_counter()
        Q, W, X, y, *_ = make_data(n=50000, min_v=2, max_v=8, lattice=(200, 250), random_state=0)
        s = LogitRSRGibbs(Q, W, X, y, random_state=10, q=64)
        ctor = time.perf_counter() - t0
        for chains in (int(c) for c in args.c5_chains.split(',')):
            run(f'c5_{chains}', s, chains, max(5, args.steps // 3), 5,
                f'LogitRSR n=50000 (200x250), q=64, {chains} chains; data + constructor {ctor:.1f} s')
    if only & {'q', 'q256', 'q512'}:
        # second series of SURVEY 8(d): larger RSR ranks (tiled large-q solve, tcgen05 Gram over ~34k / ~51k columns)
        from occuspytial_b200.gibbs import _setup
        Q, W, X, y, *_ = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
        for q in (256, 512):
            if not only & {'q', f'q{q}'}:
                continue
            K = _setup.rsr_basis(Q, X, q=q)
            s = LogitRSRGibbs(Q, W, X, y, random_state=10, K=K)
            run(f'q{q}', s, 1024, max(5, args.steps // 3), 5, f'LogitRSR n=10000, q={q}, 1024 chains')
    if only & {'q', 'r05'}:
        Q, W, X, y, *_ = make_data(n=2500, min_v=2, max_v=8, lattice=(50, 50), random_state=0)
        s = LogitRSRGibbs(Q, W, X, y, random_state=10)          # default threshold r = 0.5
        run('r05_2500', s, 1024, max(5, args.steps // 3), 5,
            f'LogitRSR n=2500 (50x50), default r=0.5 -> q={s.fixed.q}, 1024 chains')


if __name__ == '__main__':
    main()
