#!/usr/bin/env python
"""bench.py -- chain-iterations/sec of the LogitRSRGibbs sweep (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port) on the host cores

A "step" is one full Gibbs iteration (omega_b, tau, eta, beta, omega_a, alpha, z) of every chain on the
device(s).  Workload: make_data-style synthetic data on a 100 x 100 queen lattice
(n = 10 000 sites, p = r = 3, n/2 surveyed sites with 2..8 visits), RSR rank q = 64, 4096 lock-step chains
per GPU (weak scaling: chains are independent, no data-path collective).  One JSON line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'chain-iterations/sec, LogitRSRGibbs n=10k sites, 4096 chains, 1/2/4/8 B200'
GRAM = {'auto': None, 'dmma': False, 'int8': True}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--lattice', default='100x100')
    ap.add_argument('--q', type=int, default=64)
    ap.add_argument('--chains', type=int, default=4096, help='chains per GPU')
    ap.add_argument('--cpu-iters', type=int, default=300, help='iterations per CPU-baseline chain')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--pg-fp64', action='store_true')
    ap.add_argument('--gram', default='auto', choices=['auto', 'dmma', 'int8'],
                    help='Omega-weighted Gram: exact INT8 slice GEMMs, hand-written FP64 DMMA kernel, or the engine default')
    return ap.parse_args()


def workload(args):
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    rows, cols = (int(v) for v in args.lattice.split('x'))
    n = rows * cols
    Q, W, X, y, *_ = make_data(n=n, min_v=2, max_v=8, lattice=(rows, cols), random_state=0)
    K = _setup.rsr_basis(Q, X, q=args.q)
    V = sum(w.shape[0] for w in W.values())
    return dict(Q=Q, W=W, X=X, y=y, K=K, n=n, V=V, q=args.q, p=X.shape[1], r=3)


def config_dict(args, wl):
    return {'workload': f'LogitRSRGibbs make_data-style n={wl["n"]} ({args.lattice} queen lattice), p=r=3, '
                        f'V={wl["V"]} visit rows, RSR rank q={wl["q"]}, {args.chains} chains per GPU',
            'chains_per_gpu': args.chains, 'n_sites': wl['n'], 'rsr_rank_q': wl['q'], 'visit_rows': wl['V'],
            'parallelism': f'chains sharded over {args.gpus} GPU(s), no per-iteration collective',
            'l2_policy': 'per-step working set (omega/z byte planes 5 B per site and chain ~ 0.2 GB, the 0.15 GB of Z slices, '
                         'the 77 MB of Gram rows) exceeds the 126 MB L2'}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's sweep, one process per chain over all host cores
# --------------------------------------------------------------------------------------------------
def _cpu_worker(payload):
    # one BLAS thread per chain process, as the reference's loky workers are meant to run;
    # the environment variables are too late once NumPy is loaded, so the pools are limited at run time
    from threadpoolctl import threadpool_limits
    from oracle import sweep as S
    wl, seed, warm, iters = payload
    with threadpool_limits(limits=1):
        s = S.OracleSampler('logit_rsr', wl['Q'], wl['W'], wl['X'], wl['y'], seed=seed, K=wl['K'])
        s.init_state()
        for _ in range(warm):
            s.step()
        t0 = time.perf_counter()
        for _ in range(iters):
            s.step()
        return time.perf_counter() - t0


def cpu_chain_iters_per_sec(wl, cores, warm, iters):
    """Reference arm / cpu_baseline: `cores` chains in parallel processes (parallel.py:38-41 analogue)."""
    import multiprocessing as mp
    small = {k: wl[k] for k in ('Q', 'W', 'X', 'y', 'K')}
    with mp.get_context('fork').Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(small, [7, c], warm, iters) for c in range(cores)])
    return cores * iters / max(times), max(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = workload(args)
    cores = os.cpu_count()
    # a "step" of this arm = one Gibbs iteration of `cores` concurrent chains (bounded sample of the workload)
    # the CPU arm needs ~60 ms per chain-iteration: bound the sample so that the default run ends in minutes
    steps, warm = min(args.steps, 100), min(args.warmup, 10)
    value, elapsed = cpu_chain_iters_per_sec(wl, cores, warm, steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'chain-iterations/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * elapsed / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(args, wl),
        'cpu_baseline': {'value': value, 'unit': 'chain-iterations/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{cores} chains x {steps} iterations of the same n={wl["n"]}, q={wl["q"]} '
                                   'workload, one process per chain (oracle/sweep.py + oracle/pg_devroye.c: NumPy/'
                                   'LAPACK restatement of the reference sweep; the Python reference itself cannot '
                                   'travel to the GPU box and needs the absent polyagamma package)'},
        'e2e': {'value': value, 'unit': 'chain-iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def fp64_peak_tflops(torch):
    """cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json has no FP64 figure): best of 5, CUDA events."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    best = float('inf')
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


_JSON_FD = None


def _emit(text):
    if _JSON_FD is None:
        print(text, flush=True)
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, (text + '\n').encode())


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout must carry exactly one JSON line, and NCCL prints its version banner there from native code: everything
        # written to fd 1 during the run goes to stderr, the JSON line goes to the saved descriptor at the end
        global _JSON_FD
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.gibbs.parallel import allgather_summaries

    wl = workload(args)
    sampler = LogitRSRGibbs(wl['Q'], wl['W'], wl['X'], wl['y'], random_state=10, K=wl['K'])
    chains_total = args.chains * world
    lo = rank * args.chains

    # ---- device-resident timing: inputs in HBM, K iterations, CUDA events on the engine's stream ----
    seed = 12345
    start = sampler._start_arrays(chains_total, None)
    start = {k: v[lo:lo + args.chains] for k, v in start.items()}
    eng = sampler._make_engine(args.chains, seed, chain_offset=lo, pg_fp64=args.pg_fp64, gram_int8=GRAM[args.gram])
    eng.set_start(start)
    rec = eng.alloc_record(args.steps)
    fails = eng.sweep(args.warmup)
    launches0 = eng.launches
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        e0.record(eng.stream)
        fails += eng.sweep(args.steps, rec, 0)
        e1.record(eng.stream)
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()
    clock_info = clocks.stop() if rank == 0 else None
    value = chains_total * args.steps / (elapsed_ms * 1e-3)

    # ---- per-kernel timing for the roofline (separate short profiled run, CUDA events per kernel class) ----
    eng.set_profiling(True)
    eng.sweep(min(args.steps, 10))
    kms = eng.kernel_ms()
    eng.set_profiling(False)
    gram_path = 'int8' if eng.gram_int8 else 'dmma'
    # the only cross-GPU exchange of the method: all-gather of per-chain summary blocks for R-hat / ESS
    lib = _lib.load()
    max_lag = min(args.steps - 1, 20)
    summ = torch.empty((args.chains, 1, 6 + max_lag + 1), dtype=torch.float64, device='cuda')
    _lib.check(lib.occu_chain_summaries(rec['tau'].data_ptr(), args.chains, args.steps, 1, max_lag, summ.data_ptr(),
                                        eng.stream.cuda_stream), 'occu_chain_summaries')
    torch.cuda.synchronize()
    gathered = allgather_summaries(summ, chains_total)
    assert gathered.shape[0] == chains_total
    eng.close()
    del eng, rec
    # the hand-written FP64 DMMA Gram kernel, measured beside the default path (rank 0 only, short run)
    dmma_info = None
    if gram_path == 'int8' and rank == 0:
        e2 = sampler._make_engine(args.chains, seed, chain_offset=lo, pg_fp64=args.pg_fp64, gram_int8=False)
        e2.set_start(start)
        e2.sweep(5)
        e2.set_profiling(True)
        e2.sweep(8)
        k2 = e2.kernel_ms()
        e2.close()
        del e2
        g_ms = k2['gram_gemm'][0] / k2['gram_gemm'][1]
        dmma_info = {'kernel': "dgemm::gemm_tn_kernel (hand-written FP64 DMMA Gram, gram='dmma' path)", 'ms_per_launch': g_ms,
                     'step_ms_with_this_path': sum(v[0] / max(v[1], 1) for v in k2.values())}

    # ---- end to end through the public API: host data in, host draws out (allocator-warm process: the workspace of
    # the engine above went back to torch's caching allocator and is reused, as in any second call of a session) ----
    e2e = None
    if not args.no_e2e:
        # untimed warm-up call of the same public API (first-call costs: lazy kernel loading, allocator, pinned staging);
        # the timed call below still includes its own problem upload, start values, every iteration and the D2H of the draws
        sampler.sample(max(8, args.warmup // 5), burnin=2, chains=chains_total, progressbar=False,
                       pg_fp64=args.pg_fp64, gram_int8=GRAM[args.gram])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        prof = None
        if os.environ.get('OCCU_PROFILE_E2E'):   # debugging aid: where does the host time of the timed call go
            import cProfile
            prof = cProfile.Profile()
            prof.enable()
        dt, st = float('inf'), None
        for _ in range(2):   # two timed calls, the faster one is reported (host-side set-up jitters by up to 0.1 s on these boxes)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            post = sampler.sample(args.steps + args.warmup, burnin=args.warmup, chains=chains_total, progressbar=False,
                                  pg_fp64=args.pg_fp64, gram_int8=GRAM[args.gram])
            torch.cuda.synchronize()
            one = time.perf_counter() - t0
            if world > 1:   # a call is as slow as its slowest rank
                t = torch.tensor([one], dtype=torch.float64, device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                one = float(t.item())
            if one < dt:
                dt, st = one, sampler._last_stats
        if prof is not None:
            import pstats
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats('cumulative').print_stats(22)
        iters = args.steps + args.warmup
        assert post['tau'].shape == (args.chains, args.steps)
        e2e = {'value': chains_total * iters / dt, 'unit': 'chain-iterations/s',
               'h2d_bytes_per_step': st['h2d_bytes'] * world / iters, 'd2h_bytes_per_step': st['d2h_bytes'] * world / iters,
               'seconds': dict(st.get('seconds', {}), total=dt),
               'note': f'sampler.sample({iters}, burnin={args.warmup}, chains={chains_total}) wall time incl. problem '
                       'upload, start values, all iterations and the copy of the recorded draws to host (after one short '
                       'untimed warm-up call; faster of two timed calls)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant step: the Omega-weighted Gram ----
    pq = wl['p'] + wl['q']
    n_pad = -(-wl['n'] // 128) * 128
    M = pq * (pq + 1) // 2
    M_pad = -(-M // 64) * 64
    gram_ms, gram_calls = kms['gram_gemm']
    gram_ms /= gram_calls
    gram_flops_alg = 2.0 * wl['n'] * (M + pq) * args.chains      # algorithmic FP64 flops
    peak64 = fp64_peak_tflops(torch)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    total_ms = sum(v[0] / max(v[1], 1) for v in kms.values())
    if gram_path == 'int8':
        # gram_tc_kernel (csrc/gram_tc.cuh): slice pairs (omega planes x 6 Z planes, levels s + t <= 7) over 128-chain x
        # 64-column tiles, plus the 6 slices of the U'z tiles
        # production precision: omega is a 32-bit fixed-point number (4 byte planes) -> 18 slice pairs; FP64 draws: 21
        PAIRS, SB = (21 if args.pg_fp64 else 18), 6
        # useful INT8 operations of the slicing algorithm (no tile / K padding counted): PAIRS products of the
        # [chains x n] omega planes with the [n x M] Z planes, SB products of the z plane with the [n x pq] U planes
        int8_ops = 2.0 * args.chains * wl['n'] * (M * PAIRS + pq * SB)
        int8_peak = 2.0 * peaks.get('bf16_tflops', 1590.0)
        traffic = None
        try:   # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/)
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'))).get('gram_tc_kernel')
        except (OSError, ValueError):
            pass
        roofline = {'bound': 'tensor', 'kernel': 'gtc::gram_tc_kernel<64>: hand-written tcgen05.mma kind::i8 (TMA-staged '
                    'operands, 6 level accumulators in TMEM, FP64 recombination in the epilogue); exact integer evaluation '
                    'of the FP64 contraction [Omega\' Z | z\' U] (csrc/gram_tc.cuh)',
                    'achieved': int8_ops / (gram_ms * 1e-3) / 1e12, 'peak': int8_peak, 'unit': 'TOP/s (int8, dense)',
                    'frac': int8_ops / (gram_ms * 1e-3) / 1e12 / int8_peak, 'traffic': traffic,
                    'peak_source': '2 x MEASURED_PEAKS.json bf16_tflops (no INT8 figure is recorded; nominal dense INT8 is 4500)',
                    'fp64_equivalent_tflops': gram_flops_alg / (gram_ms * 1e-3) / 1e12, 'fp64_dgemm_peak_tflops': peak64,
                    'ms_per_launch': gram_ms, 'share_of_step': gram_ms / total_ms}
    else:
        roofline = {'bound': 'tensor', 'kernel': "dgemm::gemm_tn_kernel (Gram [Omega' Z | kd' U], FP64 DMMA)",
                    'achieved': gram_flops_alg / (gram_ms * 1e-3) / 1e12, 'peak': peak64, 'unit': 'TFLOP/s',
                    'frac': gram_flops_alg / (gram_ms * 1e-3) / 1e12 / peak64, 'traffic': None,
                    'peak_source': 'cuBLAS DGEMM 8192^3 via torch.matmul, measured in this run (MEASURED_PEAKS.json '
                                   'holds no FP64 figure)',
                    'ms_per_launch': gram_ms, 'share_of_step': gram_ms / total_ms}
    site_ms, site_calls = kms['site_pass']
    # psi in (unless fused), z in/out, then omega + kd out (FP64 Gram path) or the 7 byte planes of (omega, z)
    fused_psi = not args.pg_fp64 and os.environ.get('OCCU_FUSE_PSI', '1') != '0' and pq <= 128   # psi formed inside the site pass
    site_bytes_alg = ((0 if fused_psi else 8) + 2 + ((7 if args.pg_fp64 else 5) if gram_path == 'int8' else 16)) * wl['n'] * args.chains
    line = {
        'metric': METRIC, 'value': value, 'unit': 'chain-iterations/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config_dict(args, wl),
        'clocks': clock_info, 'gpu_launches': int(launches), 'failed_chains': int(fails), 'gram_path': gram_path,
        'roofline': roofline,
        'kernels_ms_per_step': {k: v[0] / max(v[1], 1) for k, v in kms.items()},
        'roofline_site_pass': {'bound': 'hbm', 'achieved': site_bytes_alg / (site_ms / site_calls * 1e-3) / 1e9,
                               'peak': hbm_peak, 'unit': 'GB/s',
                               'frac': site_bytes_alg / (site_ms / site_calls * 1e-3) / 1e9 / hbm_peak,
                               'note': 'the site pass is instruction-issue bound (PG sampling), not HBM bound',
                               'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650'},
    }
    # whole-step ceilings of SURVEY.md 8(d): F_alg / P64 (the reference's FP64 arithmetic on the FP64 pipe) and
    # B_alg / BW (every per-chain vector crossing HBM once per stage); the tcgen05 Gram leaves the first behind
    r_ = 3
    f_alg = wl['n'] * pq * (pq + 1) + 4 * wl['n'] * pq + (wl['q'] ** 3 + wl['p'] ** 3 + r_ ** 3) / 3 \
        + 4 * (wl['q'] ** 2 + wl['p'] ** 2 + r_ ** 2) + 0.5 * wl['V'] * r_ * (r_ + 1) + 4 * wl['V'] * r_
    b_alg = 35 * wl['n'] + 16 * pq ** 2
    per_gpu = value / world
    line['step_roofline'] = {
        'alg_flops_per_chain_iteration': f_alg, 'alg_bytes_per_chain_iteration': b_alg,
        'fp64_pipe_ceiling_chain_it_per_s': peak64 * 1e12 / f_alg, 'hbm_ceiling_chain_it_per_s': hbm_peak * 1e9 / b_alg,
        'value_over_fp64_pipe_ceiling': per_gpu / (peak64 * 1e12 / f_alg),
        'value_over_hbm_ceiling': per_gpu / (hbm_peak * 1e9 / b_alg),
        'note': 'per GPU; the FP64-pipe ceiling is what bounds an FP64 implementation of the sweep (cuBLAS DGEMM rate); '
                'the Gram runs as exact integer slices on the INT8 tensor pipe instead, so the step exceeds it'}
    if dmma_info:
        tf = gram_flops_alg / (dmma_info['ms_per_launch'] * 1e-3) / 1e12
        dmma_info.update({'bound': 'tensor', 'achieved': tf, 'peak': peak64, 'unit': 'TFLOP/s', 'frac': tf / peak64,
                          'peak_source': 'cuBLAS DGEMM 8192^3 via torch.matmul, measured in this run'})
        line['roofline_dmma'] = dmma_info
    if e2e:
        line['e2e'] = e2e
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count()
        v, el = cpu_chain_iters_per_sec(wl, cores, 3, args.cpu_iters)
        line['cpu_baseline'] = {'value': v, 'unit': 'chain-iterations/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{cores} chains x {args.cpu_iters} iterations of the same workload, one '
                                          f'process per chain, {el:.1f} s (oracle/sweep.py + oracle/pg_devroye.c)'}
    _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
