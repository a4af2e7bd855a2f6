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
    ap.add_argument('--cpu-iters', type=int, default=150, help='iterations per CPU-baseline chain')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help="weak: --chains per GPU; strong: --chains in total, sharded over the GPUs (BASELINE config 3's wording)")
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
    return dict(Q=Q, W=W, X=X, y=y, K=K, n=n, V=V, q=args.q, p=X.shape[1], r=next(iter(W.values())).shape[1])


def config_dict(args, wl):
    return {'workload': f'LogitRSRGibbs make_data-style n={wl["n"]} ({args.lattice} queen lattice), p=r=3, '
                        f'V={wl["V"]} visit rows, RSR rank q={wl["q"]}, {args.chains} chains ' + ('per GPU' if args.scaling == 'weak' else 'in total'),
            'chains_per_gpu': args.chains if args.scaling == 'weak' else -(-args.chains // args.gpus), 'n_sites': wl['n'], 'rsr_rank_q': wl['q'], 'visit_rows': wl['V'],
            'parallelism': f'chains sharded over {args.gpus} GPU(s), no per-iteration collective; one all-gather of per-chain '
                           'summary blocks (R-hat / ESS) at the end of sample()',
            'l2_policy': 'per-step working set (omega/z byte planes 5 B per site and chain ~ 0.2 GB, the 0.15 GB of Z slices, '
                         'the 77 MB of Gram rows) exceeds the 126 MB L2'}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's sweep, one process per chain over all host cores
# --------------------------------------------------------------------------------------------------
_CPU_SAMPLER = None
ITERS_PER_STEP = 10   # reference arm: one "step" = 10 Gibbs iterations of one chain per host core


def _cpu_init(small):
    """Pool initialiser: build the chain's sampler and warm it up OUTSIDE the timed region."""
    global _CPU_SAMPLER
    from threadpoolctl import threadpool_limits
    from oracle import sweep as S
    threadpool_limits(limits=1)   # one BLAS thread per chain process, as the reference's loky workers run
    _CPU_SAMPLER = S.OracleSampler('logit_rsr', small['Q'], small['W'], small['X'], small['y'],
                                   seed=[7, os.getpid()], K=small['K'])
    _CPU_SAMPLER.init_state()
    for _ in range(3):
        _CPU_SAMPLER.step()


def _cpu_worker(iters):
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        for _ in range(iters):
            _CPU_SAMPLER.step()
        return time.perf_counter() - t0


def cpu_chain_iters_per_sec(wl, cores, warm, iters):
    """Reference arm / cpu_baseline: `cores` chains in parallel processes (parallel.py:38-41 analogue).  The pool is
    started, every sampler built and `warm` iterations run before the clock starts; the timed region is `iters`
    iterations of every chain, wall clock around the whole map (the slowest chain counts)."""
    import multiprocessing as mp
    small = {k: wl[k] for k in ('Q', 'W', 'X', 'y', 'K')}
    with mp.get_context('fork').Pool(cores, initializer=_cpu_init, initargs=(small,)) as pool:
        pool.map(_cpu_worker, [max(1, warm)] * cores, chunksize=1)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [iters] * cores, chunksize=1)
        elapsed = time.perf_counter() - t0
    return cores * iters / elapsed, elapsed


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = workload(args)
    cores = os.cpu_count()
    # a "step" of this arm = ITERS_PER_STEP Gibbs iterations of `cores` concurrent chains: a bounded sample of the
    # workload (one chain-iteration costs ~15-60 ms on a host core) sized so that K = 20 steps already give a stable number
    steps, warm = min(args.steps, 60), min(args.warmup, 5)
    value, elapsed = cpu_chain_iters_per_sec(wl, cores, warm * ITERS_PER_STEP, steps * ITERS_PER_STEP)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'chain-iterations/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * elapsed / steps,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(args, wl),
        'cpu_baseline': {'value': value, 'unit': 'chain-iterations/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{cores} chains x {steps * ITERS_PER_STEP} iterations ({steps} steps of {ITERS_PER_STEP}) '
                                   f'of the same n={wl["n"]}, q={wl["q"]} workload, one process per chain, pool and '
                                   'samplers warmed before the clock (oracle/sweep.py + oracle/pg_devroye.c: NumPy/'
                                   'LAPACK restatement of the reference sweep; the Python reference itself cannot '
                                   'travel to the GPU box and needs the absent polyagamma package; measured in the build '
                                   'container, the port runs 1.37x the unmodified reference per core: '
                                   'profiles/r02_reference_vs_port_cpu.json)'},
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


def int8_peak_tops(torch):
    """Dense INT8 tensor-core rate of this GPU, measured in the run: torch._int_mm (cuBLASLt IMMA) 8192^3, best of 5."""
    n = 8192
    try:
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device='cuda')
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device='cuda')
        torch._int_mm(a, b)
        best = float('inf')
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:   # the library call is only a yardstick; fall back to the bf16-derived proxy
        return None


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
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.gibbs.parallel import shard_bounds

    wl = workload(args)
    sampler = LogitRSRGibbs(wl['Q'], wl['W'], wl['X'], wl['y'], random_state=10, K=wl['K'])
    chains_total = args.chains * world if args.scaling == 'weak' else args.chains
    lo, hi = shard_bounds(chains_total, world, rank)
    local = hi - lo
    dev = f'cuda:{local_rank}'

    # ---- device-resident timing: inputs in HBM, K iterations, CUDA events on the engine's stream ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()          # before the warm-up, so that short timed regions still collect samples
    seed = 12345
    eng = sampler._make_engine(local, seed, chain_offset=lo, device=dev, pg_fp64=args.pg_fp64, gram_int8=GRAM[args.gram])
    eng.draw_start(seed, *sampler._DEFAULT_START)
    rec = eng.alloc_record(args.steps)
    fails = eng.sweep(args.warmup)
    launches0 = eng.launches
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

    # ---- per-kernel timing for the rooflines (separate short profiled run, CUDA events per kernel class) ----
    eng.set_profiling(True)
    eng.sweep(min(args.steps, 10))
    kms = eng.kernel_ms()
    eng.set_profiling(False)
    gram_path = 'int8' if eng.gram_int8 else 'dmma'
    eng.close()
    del eng, rec
    # the same sweep with FP64 latent draws (validation build), and the hand-written FP64 DMMA Gram: measured beside the
    # production path (rank 0 only, short runs)
    fp64_info, dmma_info = None, None
    if rank == 0 and not args.pg_fp64:
        e2 = sampler._make_engine(local, seed, chain_offset=lo, device=dev, pg_fp64=True, gram_int8=GRAM[args.gram])
        e2.draw_start(seed, *sampler._DEFAULT_START)
        e2.sweep(3)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(e2.stream):
            a0.record(e2.stream)
            e2.sweep(6)
            a1.record(e2.stream)
        torch.cuda.synchronize()
        fp64_info = {'value': local * 6 / (a0.elapsed_time(a1) * 1e-3), 'unit': 'chain-iterations/s (this GPU)',
                     'ms_per_step': a0.elapsed_time(a1) / 6, 'steps': 6,
                     'note': 'pg_fp64=True: PG draws and occupancy probabilities in FP64 (validation build), 48-bit omega planes'}
        e2.close()
        del e2
    if gram_path == 'int8' and rank == 0:
        e2 = sampler._make_engine(local, seed, chain_offset=lo, device=dev, pg_fp64=args.pg_fp64, gram_int8=False)
        e2.draw_start(seed, *sampler._DEFAULT_START)
        e2.sweep(5)
        e2.set_profiling(True)
        e2.sweep(8)
        k2 = e2.kernel_ms()
        e2.close()
        del e2
        g_ms = k2['gram_gemm'][0] / k2['gram_gemm'][1]
        dmma_info = {'kernel': "dgemm::gemm_tn_kernel (hand-written FP64 DMMA Gram, gram='dmma' path)", 'ms_per_launch': g_ms,
                     'step_ms_with_this_path': sum(v[0] / max(v[1], 1) for v in k2.values())}

    # ---- end to end through the public API: host data in, host draws out.  Timed calls of sampler.sample() after one short
    # untimed call of the same API (first-call costs: lazy kernel loading, allocator); every timed call includes its own
    # problem upload, start values, every iteration, the device summary blocks and the D2H of the draws ----
    e2e = None
    if not args.no_e2e:
        kw = dict(chains=chains_total, progressbar=False, pg_fp64=args.pg_fp64, gram_int8=GRAM[args.gram])
        if world == 1:
            kw['device'] = dev      # one GPU, as --gpus says (a plain call would fan out over every visible GPU)
        sampler.sample(args.steps + args.warmup, burnin=args.warmup, **kw)   # same shape as the timed calls: sizes the staging buffers
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        calls = []
        for _ in range(3):      # three timed calls, every one reported; the headline is their MEDIAN
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            post = sampler.sample(args.steps + args.warmup, burnin=args.warmup, **kw)
            torch.cuda.synchronize()
            one = time.perf_counter() - t0
            if world > 1:   # a call is as slow as its slowest rank
                t = torch.tensor([one], dtype=torch.float64, device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                one = float(t.item())
            calls.append((one, sampler._last_stats))
        dt, st = sorted(calls, key=lambda c_: c_[0])[1]
        iters = args.steps + args.warmup
        assert post['tau'].shape == (local, args.steps)
        summ = sampler._last_summaries
        assert summ is not None and summ['tau'].shape[0] == chains_total   # all-gathered over the ranks (NCCL)
        h2d = torch.tensor([st['h2d_bytes'], st['d2h_bytes']], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(h2d)
        e2e = {'value': chains_total * iters / dt, 'unit': 'chain-iterations/s',
               'h2d_bytes_per_step': float(h2d[0]) / iters, 'd2h_bytes_per_step': float(h2d[1]) / iters,
               'seconds': dict(st.get('seconds', {}), total=dt), 'timed_calls_s': [c_[0] for c_ in calls],
               'r_hat_tau_from_device_summaries': float(post.summary.loc['tau', 'r_hat']),
               'note': f'median of three timed sampler.sample({iters}, burnin={args.warmup}, chains={chains_total}) calls (all three '
                       'listed in timed_calls_s), wall clock incl. problem upload, start values, all iterations, per-chain '
                       'summary blocks (all-gathered over the ranks, reduced over the chains on the device) and the copy of the recorded draws to host; one '
                       'untimed call of the same API and size before them'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines, per kernel; `roofline` is the kernel with the largest share of the step ----
    pq = wl['p'] + wl['q']
    M = pq * (pq + 1) // 2
    ms = {k: v[0] / max(v[1], 1) for k, v in kms.items()}
    total_ms = sum(ms.values())
    gram_ms = ms['gram_gemm']
    gram_flops_alg = 2.0 * wl['n'] * (M + pq) * local           # algorithmic FP64 flops
    peak64 = fp64_peak_tflops(torch)
    peaks, traffic, ncu_facts = {}, {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    try:   # DRAM bytes per launch and issue statistics from the committed `ncu --set full` captures (profiles/)
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        ncu_facts = traffic.get('_site_pass_ncu', {})
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    hbm_src = 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650'
    scale_traffic = local / 4096.0 * wl['n'] / 10000.0   # the captures were taken on the default workload

    # site pass: z, omega_a, omega_b.  Algorithmic bytes per site and chain: z out (1) + omega/z byte planes out
    # (5 in production precision, 7 with FP64 draws; FP64 Gram path: omega + kd, 16) + psi in (8, unless formed in the pass)
    # + the alpha partial sums (NV doubles per 32-site tile)
    r_ = wl['r']
    nv = r_ * (r_ + 1) // 2 + r_
    fused_psi = not args.pg_fp64 and os.environ.get('OCCU_FUSE_PSI', '1') != '0' and pq <= 128
    site_b = 1 + ((7 if args.pg_fp64 else 5) if gram_path == 'int8' else 16) + (0 if fused_psi else 8) + 8.0 * nv / 32
    site_bytes = site_b * wl['n'] * local
    site_ms = ms['site_pass']
    t_site = traffic.get('site_pass_logit_kernel')
    roof_site = {'bound': 'hbm', 'kernel': 'site_pass_logit_kernel: z Bernoulli update, PG(1, .) draws for omega_a / omega_b with '
                 'pooled refill, alpha partial sums, psi = [X K] theta on the FP64 tensor pipe, byte planes of (omega, z) out '
                 '(csrc/site_pass.cuh)',
                 'achieved': site_bytes / (site_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                 'frac': site_bytes / (site_ms * 1e-3) / 1e9 / hbm_peak,
                 'traffic': None if t_site is None else t_site * scale_traffic, 'peak_source': hbm_src,
                 'alg_bytes_per_site_chain': site_b, 'ms_per_launch': site_ms, 'share_of_step': site_ms / total_ms,
                 'limiter': 'instruction issue, not HBM: ~2.4 PG draws per site and chain at >= 100 instructions each '
                            '(Philox block, proposal, alternating-series test) against 8 bytes of traffic',
                 'ncu': ncu_facts}
    # Gram
    if gram_path == 'int8':
        # gram_tc_kernel (csrc/gram_tc.cuh): slice pairs (omega planes x 6 Z planes, levels s + t <= 7) over 128-chain x
        # 64-column tiles, plus the 6 slices of the U'z tiles; production precision: omega is a 32-bit fixed-point number
        # (4 byte planes) -> 18 slice pairs; FP64 draws: 21.  Useful INT8 operations (no tile / K padding counted)
        PAIRS, SB = (21 if args.pg_fp64 else 18), 6
        int8_ops = 2.0 * local * wl['n'] * (M * PAIRS + pq * SB)
        int8_meas = int8_peak_tops(torch)
        int8_peak = int8_meas if int8_meas else 2.0 * peaks.get('bf16_tflops', 1590.0)
        t_gram = traffic.get('gram_tc_kernel')
        roof_gram = {'bound': 'tensor', 'kernel': 'gtc::gram_tc_kernel<64>: hand-written tcgen05.mma kind::i8 (TMA-staged '
                     'operands, 6 level accumulators in TMEM, FP64 recombination in the epilogue); exact integer evaluation '
                     'of the FP64 contraction [Omega\' Z | z\' U] (csrc/gram_tc.cuh)',
                     'achieved': int8_ops / (gram_ms * 1e-3) / 1e12, 'peak': int8_peak, 'unit': 'TOP/s (int8, dense)',
                     'frac': int8_ops / (gram_ms * 1e-3) / 1e12 / int8_peak,
                     'traffic': None if t_gram is None else t_gram * scale_traffic,
                     'peak_source': ('torch._int_mm 8192^3 (cuBLASLt INT8 GEMM) measured in this run, best of 5' if int8_meas else
                                     '2 x MEASURED_PEAKS.json bf16_tflops (torch._int_mm unavailable)') + '; nominal dense INT8 is 4500',
                     'fp64_equivalent_tflops': gram_flops_alg / (gram_ms * 1e-3) / 1e12, 'fp64_dgemm_peak_tflops': peak64,
                     'ms_per_launch': gram_ms, 'share_of_step': gram_ms / total_ms}
    else:
        roof_gram = {'bound': 'tensor', 'kernel': "dgemm::gemm_tn_kernel (Gram [Omega' Z | kd' U], FP64 DMMA)",
                     'achieved': gram_flops_alg / (gram_ms * 1e-3) / 1e12, 'peak': peak64, 'unit': 'TFLOP/s',
                     'frac': gram_flops_alg / (gram_ms * 1e-3) / 1e12 / peak64, 'traffic': None,
                     'peak_source': 'cuBLAS DGEMM 8192^3 via torch.matmul, measured in this run (MEASURED_PEAKS.json '
                                    'holds no FP64 figure)',
                     'ms_per_launch': gram_ms, 'share_of_step': gram_ms / total_ms}
    # per-chain solve: reads its packed Gram row and H, writes the state and the recorded draw
    solve_ms = ms['solve']
    solve_bytes = 8.0 * (M + 2 * pq + wl['p'] + r_ + 1) * local
    t_solve = traffic.get('solve_logit_rsr_kernel')
    roof_solve = {'bound': 'hbm', 'kernel': 'solve_logit_rsr_kernel: tau, Cholesky of K\'OmegaK + tau K\'QK in shared memory, eta, beta, '
                  'alpha, record (csrc/solve_rsr.cuh)',
                  'achieved': solve_bytes / (solve_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                  'frac': solve_bytes / (solve_ms * 1e-3) / 1e9 / hbm_peak,
                  'traffic': None if t_solve is None else t_solve * local / 4096.0, 'peak_source': hbm_src,
                  'ms_per_launch': solve_ms, 'share_of_step': solve_ms / total_ms,
                  'limiter': 'barrier / dependent-column latency of the q x q factorisation, 6 chains per SM'}
    roofs = {'site_pass': roof_site, 'gram_gemm': roof_gram, 'solve': roof_solve}
    dominant = max(roofs, key=lambda k: ms[k])
    line = {
        'metric': METRIC, 'value': value, 'unit': 'chain-iterations/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None,
        'dtype': 'f64' if args.pg_fp64 else 'f64 algebra (Gram, Cholesky, solves, records) / f32 latent draws (PG, z probability)',
        'data': 'synthetic', 'config': config_dict(args, wl),
        'clocks': clock_info, 'gpu_launches': int(launches), 'failed_chains': int(fails), 'gram_path': gram_path,
        'roofline': dict(roofs[dominant], dominant_kernel_class=dominant),
        'kernels_ms_per_step': ms,
    }
    for k, r in roofs.items():
        if k != dominant:
            line['roofline_' + {'site_pass': 'site_pass', 'gram_gemm': 'gram', 'solve': 'solve'}[k]] = r
    if fp64_info:
        line['value_pg_fp64'] = fp64_info
    # whole-step ceilings of SURVEY.md 8(d): F_alg / P64 (the reference's FP64 arithmetic on the FP64 pipe) and
    # B_alg / BW (every per-chain vector crossing HBM once per stage); the tcgen05 Gram leaves the first behind
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
                                          f'process per chain, pool and samplers warmed before the clock, {el:.1f} s '
                                          '(oracle/sweep.py + oracle/pg_devroye.c)'}
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
