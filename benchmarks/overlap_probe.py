"""Do the ALU-bound site pass and the tensor-bound Gram overlap when the chains are split over several engines on
separate streams of ONE device?  Aggregate ms per iteration of E engines x (4096 / E) chains (headline workload).

  python benchmarks/overlap_probe.py [engines ...]      env: OCCU_B200_LIB (library variant), OCCU_GRAM_PRIO
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
import bench   # noqa: E402


def main():
    counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
    sys.argv = sys.argv[:1]
    from occuspytial_b200 import LogitRSRGibbs
    wl = bench.workload(bench.parse())
    sampler = LogitRSRGibbs(wl['Q'], wl['W'], wl['X'], wl['y'], random_state=10, K=wl['K'])
    total, warm, iters, chunk = 4096, 20, 60, 5
    for E in counts:
        per = total // E
        engines = []
        for k in range(E):
            eng = sampler._make_engine(per, 12345, chain_offset=k * per, device='cuda:0', slot=k)
            eng.draw_start(12345, *sampler._DEFAULT_START)
            engines.append(eng)

        def run(n):
            for _ in range(n // chunk):
                for eng in engines:
                    eng.sweep_enqueue(chunk)
            return sum(eng.finish() for eng in engines)
        run(warm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fails = run(iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({'lib': os.path.basename(os.environ.get('OCCU_B200_LIB', 'default')),
                          'gram_prio': os.environ.get('OCCU_GRAM_PRIO', '0'), 'engines': E, 'chains_per_engine': per,
                          'ms_per_iteration': 1e3 * dt / iters, 'chain_its_per_s': total * iters / dt, 'fails': fails}),
              flush=True)
        for eng in engines:
            eng.close()


if __name__ == '__main__':
    main()
