#!/usr/bin/env python
"""Headline benchmark: LTE subframes/s, 20 MHz 64-QAM SIMO 1x4 MRC over ITU Rayleigh.

    python bench.py --gpus N --steps K --warmup W            # CUDA engine (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host cores)

A step is `--batches` passes of the whole link chain (bits -> QAM/grid/IFFT -> 4 fading links -> AWGN ->
FFT -> CRS LS estimate -> MRC -> hard demap -> error count), each over `--trials` subframes at each of the
16 SNR points 0..30 dB, on every GPU (weak scaling): 32 x 4096 = 131 072 subframes and ~60 ms of GPU work per
step by default, so the default 20 timed steps keep the device busy for more than a second.  All pipeline logic
lives in the package (`lte_b200`): this file only calls its public API and measures.  One JSON line is printed
by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'ofdm-lte_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = 'LTE subframes/sec 20MHz 64QAM SIMO-4 MRC Rayleigh'
UNIT = 'subframes/s'
S_SUBFRAME = 14
R_ANT = 4
SNR_POINTS = [float(s) for s in range(0, 31, 2)]
PROFILE, VELOCITY, FC_GHZ = 'Pedestrian_A', 3.0, 2.0
TRAFFIC_JSON = os.path.join(ROOT, 'profiles', 'r02_traffic.json')


def workload_name(trials):
    return (f'20MHz(N=2048,cp=144) 64-QAM SIMO 1x{R_ANT} MRC, ITU {PROFILE} {VELOCITY:g} km/h @ {FC_GHZ:g} GHz '
            f'(time-domain TDL, 16-tone Jakes), CRS LS estimate per 14 symbols, 16 SNR points 0..30 dB x '
            f'{trials} trials')


def bench_config(args, world):
    """The same dict in both arms: the driver compares them key by key."""
    B = args.trials * len(SNR_POINTS)
    return {'workload': workload_name(args.trials), 'subframes_per_batch_per_gpu': B, 'batches_per_step': args.batches,
            'subframes_per_step_per_gpu': B * args.batches,
            'l2_policy': 'inputs and intermediates of one batch (>= 3.5 GB) exceed the 126 MB L2',
            'parallelism': f'trial sharding x{world}, one int64[16] allreduce'}


# ------------------------------------------------------------------------------------ algorithmic bytes / flops
def stage_bytes(N=2048, cp=144, Nd=999, Np=200, Nc=1200, b=6, R=R_ANT, S=S_SUBFRAME, NT=4):
    """Compulsory bytes per subframe of each stage kernel (complex64 = 8 B, one index byte per
    QAM symbol in HBM)."""
    L = N + cp
    R2, ndp, npp = (R + 1) & ~1, (Nd + 1) // 2, (Np + 1) // 2
    coef = S * NT * 6 * R2 * 4
    return {
        # spectral pipeline (sweep default at low Doppler)
        'tx_spectral': S * Nd + S * Nc * 8 + S * cp * 8,
        'channel_spectral': S * Nc * 8 + S * cp * 8 + S * Nd + coef + R * NT * 16 * 4 + R * S * 2 * ndp * 8 + R * 2 * npp * 8,
        'mrc_demap_count_compact': R * S * Nd * 8 + R * 2 * npp * 8 + S * Nd + 8,     # LS estimate formed inside
        # fused time-domain pipeline
        'tx_map_ifft': S * Nd + S * L * 8,
        'channel_rx_fft': S * L * 8 + R * S * Nc * 8,
        'crs_ls_interp': R * Np * 8 + R * Nc * 8,
        'mrc_demap_count': R * S * Nd * 8 + R * Nd * 8 + S * Nd + 8,
        # staged pipeline
        'channel_tdl': S * L * 8 + R * S * L * 8,
        'rx_fft': R * S * L * 8 + R * S * Nc * 8,
    }


def spectral_flops(npairs=600, cp=144, R=R_ANT, S=S_SUBFRAME, NT=4, dmax=13):
    """fp32 flops per subframe of channel_spectral_kernel (csrc/spectral.cu): packed FFMA2 = 4 flops, FMUL2 = 2.
    Per bin pair and symbol: Horner 4 (dmax - 1) FFMA2 (the first step is a copy); per tap with a delay
    V/W = 6 FFMA2 + 4 FMUL2; combine 8 FFMA2 per (tap, antenna) (first tap: 6 + 2 FMUL2); power 2 FFMA2 per antenna.  CP sample pairs:
    6 FFMA2 per (tap, antenna) + 2 for the power."""
    per_pair = 4 * (4 * max(dmax - 1, 0)) + (NT - 1) * (6 * 4 + 4 * 2) + R * ((NT - 1) * 8 * 4 + 6 * 4 + 2 * 2) + R * 2 * 4
    per_cp = (cp // 2) * (NT * R * 6 * 4 + R * 2 * 4)
    return S * (npairs * per_pair + per_cp)


STAGED = ('tx_map_ifft', 'channel_tdl', 'rx_fft', 'crs_ls_interp', 'mrc_demap_count')


def load_traffic(kernel):
    """dram__bytes_read + write per launch of `kernel` from the ncu capture summarised in profiles/r02_traffic.json
    (tools/ncu_traffic.py writes it).  A kernel the capture does not hold is an error, not a silent constant."""
    with open(TRAFFIC_JSON) as fh:
        t = json.load(fh)
    if kernel not in t['kernels']:
        raise KeyError(f"{TRAFFIC_JSON} has no ncu capture of '{kernel}' (has: {sorted(t['kernels'])}); "
                       f"re-run tools/ncu_traffic.py on a fresh `ncu --set full` report of tools/stage_bench.py")
    k = t['kernels'][kernel]
    return {'bytes': k['dram_bytes'], 'subframes_per_launch': t['subframes_per_launch'], 'source': t['source']}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '25', '-i', str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], 0.0, set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                mx = max(mx, float(f[2]))
                if t0 <= ts <= t1:
                    sm.append(float(f[1]))
                    pw.append(float(f[3]))
                    for nme, v in zip(names, f[5:9]):
                        if v.lower().startswith('active'):
                            reasons.add(nme)
            except ValueError:
                continue
        if not sm:
            for ts, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(',')]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm), 'power_w_max': max(pw) if pw else None}


# ------------------------------------------------------------------------------------ CPU arms
def _cpu_one(args):
    seed, snr_db = args
    from oracle import lte_oracle as O
    num = O.Numerology(20.0, 15.0, '64-QAM')
    rs = np.random.RandomState(seed)
    bits = rs.randint(0, 2, 999 * 6 * S_SUBFRAME)
    taps = len(O.ITU[PROFILE][0])
    n = S_SUBFRAME * num.L
    phases = 2 * np.pi * rs.rand(R_ANT, taps, 16)
    z = rs.standard_normal((R_ANT, 2, n))
    r = O.simulate_simo(bits, snr_db, num, R_ANT, 'rayleigh_mp', PROFILE, FC_GHZ, VELOCITY, phases=phases, z=z)
    return r['errors']


def cpu_subframes_per_s(n_subframes, procs):
    """Times the oracle (NumPy fp64 port of the reference path) on `n_subframes` subframes."""
    jobs = [(1000 + i, SNR_POINTS[i % len(SNR_POINTS)]) for i in range(n_subframes)]
    if procs <= 1:
        _cpu_one(jobs[0])
        t0 = time.perf_counter()
        for j in jobs:
            _cpu_one(j)
        return n_subframes / (time.perf_counter() - t0)
    import multiprocessing as mp
    with mp.get_context('fork').Pool(procs) as pool:
        pool.map(_cpu_one, jobs[:procs])            # warm-up: imports, first FFT plans
        t0 = time.perf_counter()
        pool.map(_cpu_one, jobs, chunksize=1)
        return n_subframes / (time.perf_counter() - t0)


_REF_TIMER = r'''
import contextlib, io, json, os, sys, time
sys.path.insert(0, sys.argv[1])
import numpy as np
with contextlib.redirect_stdout(io.StringIO()):
    from config import LTEConfig
    from core.ofdm_core import OFDMSimulator
    sim = OFDMSimulator(LTEConfig(20.0, 15.0, '64-QAM'), channel_type='rayleigh_mp', itu_profile=sys.argv[2],
                        frequency_ghz=2.0, velocity_kmh=float(sys.argv[3]))
    bits = np.random.RandomState(0).randint(0, 2, 999 * 6 * 14)
    n = int(sys.argv[4])
    sim.simulate_simo(bits, snr_db=15.0, num_rx=4, parallel=False)
    ts = []
    for i in range(n):
        t0 = time.perf_counter()
        sim.simulate_simo(bits, snr_db=[0.0, 15.0, 30.0][i % 3], num_rx=4, parallel=False)
        ts.append(time.perf_counter() - t0)
print(json.dumps({'s_per_subframe_median': float(np.median(ts)), 'n': n}))
'''


def reference_subframes_per_s(n_subframes):
    """The UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build) timed on one host core through
    its own `OFDMSimulator.simulate_simo(num_rx=4, parallel=False)` (core/ofdm_core.py:1536), in a child process so
    that its `config` / `core` modules do not meet this repository's.  None when the copy did not travel."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.exists(os.path.join(ref, 'core', 'ofdm_core.py')):
        return None
    try:
        out = subprocess.run([sys.executable, '-c', _REF_TIMER, ref, PROFILE, str(VELOCITY), str(n_subframes)],
                             capture_output=True, text=True, timeout=300, env={**os.environ, 'OMP_NUM_THREADS': '1',
                                                                               'OPENBLAS_NUM_THREADS': '1'})
        r = json.loads(out.stdout.strip().splitlines()[-1])
        return {'value': 1.0 / r['s_per_subframe_median'], 'unit': UNIT, 'cores': 1, 'kind': 'reference',
                'sample': f"median of {r['n']} calls of the unmodified reference's OFDMSimulator.simulate_simo(num_rx=4, "
                          f"parallel=False), one 14-symbol subframe each, SNR 0/15/30 dB, baseline/_ref, 1 process"}
    except Exception as e:                                      # the arm is optional; say why it is missing
        return {'value': None, 'unit': UNIT, 'cores': 1, 'kind': 'reference', 'sample': f'failed: {e!r}'[:200]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    vals = []
    for i in range(args.warmup + args.steps):
        v = cpu_subframes_per_s(per_step, cores)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = (f'{per_step} subframes per step (one 14-symbol subframe per task, SNR cycling 0..30 dB), {cores} processes, '
              f'oracle/lte_oracle.py (NumPy fp64 port of the reference path)')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * per_step / value,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': bench_config(args, world),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'cpu_baseline_reference': reference_subframes_per_s(3),
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world):
    import torch
    import torch.distributed as dist
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    from lte_b200 import _native as nat
    from lte_b200.host_stream import bind_to_gpu_numa

    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa(local)                   # before the pinned staging buffers are allocated
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    cfg = LTEConfig(20.0, 15.0, '64-QAM', 'normal')
    eng = LinkEngine.from_config(cfg, device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, PROFILE, FC_GHZ, VELOCITY)
    n_snr = len(SNR_POINTS)
    B = args.trials * n_snr                     # subframes per batch per GPU
    S, R = S_SUBFRAME, R_ANT
    fused = args.pipeline != 'staged'
    spectral = args.pipeline == 'spectral'
    # two workspaces: LinkEngine.simo_ber_batches keeps two batches in flight on two streams
    inflight = args.inflight or eng.batches_in_flight(chan, B, R, S, fused=fused, spectral=spectral)
    wss = [eng.workspace(B, S, R, fading=True, fused=fused, lazy=spectral) for _ in range(inflight)]
    ws = wss[0]
    snr_lin = torch.tensor([10 ** (s / 10) for s in SNR_POINTS], dtype=torch.float32, device=dev)
    snr_rows = snr_lin.repeat(args.trials).repeat_interleave(R).contiguous()       # [B*R], SNR fastest over b
    nbits = S * eng.Nd * eng.bps
    seed = 2026
    nd = 3 if fused else 1

    # inputs resident in HBM before the timed region: one batch of transmitted symbol indices
    idx = eng.random_indices(B, S, seed, stream_id0=rank * B)
    totals = torch.zeros(n_snr, dtype=torch.int64, device=dev)

    def step(i):
        # global stream ids: independent of the GPU count.  Stream slot b always runs at the same SNR point, so the
        # per-slot counts simply accumulate in the workspaces (the MRC kernel counts with atomics); they are reduced
        # per SNR point once, after the run
        sids = [(((i * world) + rank) * args.batches + j) * B for j in range(args.batches)]
        eng.simo_ber_batches(wss, chan, snr_rows, seed, sids, idx=idx, nbits=nbits, fused=fused, spectral=spectral,
                             noise_domain=nd)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    pipeline_used = 'spectral' if ws.get('spectral') else ('fused' if fused and 'faded' not in ws else 'staged')
    for w in wss:
        w['errors'].zero_()                     # per-slot counters of the timed run start here
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    totals.copy_(sum(w['errors'] for w in wss).view(args.trials, n_snr).sum(0))     # per SNR point, inside the timed region
    if world > 1:
        dist.all_reduce(totals)                 # the only collective: int64[16] error counters
    e1.record()
    sync()
    t1 = time.time()
    launches = eng.launches - l0                # native kernels only (torch fills / NCCL not counted)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    per_step = B * args.batches
    value = world * per_step * args.steps / (ms_total * 1e-3)
    bits_total = world * per_step * args.steps * nbits / n_snr

    # ---- end to end through the package's host-buffer API: packed bits in pinned memory -> errors on host
    nbytes = (nbits + 7) // 8
    host_bits = torch.from_numpy(np.random.RandomState(7 + rank).randint(0, 256, (B, nbytes), dtype=np.uint8)).pin_memory()
    pipe = eng.stream_host_batches(chan, R, snr_rows, B, S, nbits=nbits, seed=seed + 1, noise_domain=nd, fused=fused,
                                   spectral=spectral)
    e2e_steps = max(2, min(args.steps, 10))

    def host_batches(n, i0):
        for k in range(n):
            yield host_bits, ((i0 + k) * world + rank) * B, k
    host_total = 0
    for err, _ in pipe.run(host_batches(2, 0)):                   # warm-up
        host_total += int(err.sum())
    sync()
    e0.record()
    for err, _ in pipe.run(host_batches(e2e_steps * args.batches, 2)):
        host_total += int(err[0])                                  # the step's result is read on the host
    e1.record()
    sync()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * per_step * e2e_steps / (float(ms2.item()) * 1e-3)

    # ---- per-stage device time (CUDA events on the launching stream), rooflines, CPU legs: rank 0 -----
    stages = roofline = cpu = cpu_ref = extra = None
    if rank == 0:
        stages = time_stages(eng, ws, chan, snr_rows, idx, nbits, seed, B, S, R, nat, torch, dev, pipeline_used)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
        sb = stage_bytes()
        for k in stages:
            stages[k]['algo_bytes_per_launch'] = sb[k] * B
            stages[k]['gbs'] = sb[k] * B / (stages[k]['ms'] * 1e-3) / 1e9
            stages[k]['frac'] = stages[k]['gbs'] / peak
        dom = max(stages, key=lambda k: stages[k]['ms'])
        tr = load_traffic(dom)
        fp32 = measure_fp32_peak(eng, nat, torch, dev)
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': stages[dom]['gbs'], 'peak': peak, 'unit': 'GB/s',
                    'frac': stages[dom]['frac'], 'traffic': tr['bytes'] * B / tr['subframes_per_launch'],
                    'traffic_source': tr['source'], 'peak_source': peak_src,
                    'pipeline_bytes_frac': value / world * sum(sb[k] for k in stages) / 1e9 / peak,
                    'pipeline_unfused_equivalent_frac': value / world * sum(sb[k] for k in STAGED) / 1e9 / peak}
        if dom == 'channel_spectral':
            fl = spectral_flops() * B
            roofline['compute'] = {'bound': 'fp32', 'peak_fp32_tflops': fp32, 'achieved': fl / (stages[dom]['ms'] * 1e-3) / 1e12,
                                   'unit': 'TFLOP/s', 'peak_source': 'measured on this box (lte_fp32_peak_launch: independent '
                                   'fma.rn.f32x2 chains, CUDA events)', 'flops_per_launch': fl}
            roofline['compute']['frac'] = roofline['compute']['achieved'] / fp32
            roofline['limiter'] = ('packed-fp32 FMA path (register-file limited: 2.3-2.9 cycles per FFMA2 for this kernel\'s '
                                   'operand patterns against 2.2 in the peak measurement, profiles/r02_ffma2_forms.txt), not HBM: '
                                   'see profiles/r02_spectral_ncu_summary.md and DESIGN.md 4.3a')
        if world == 1:
            n_cpu = args.cpu_subframes
            v = cpu_subframes_per_s(n_cpu, 1)
            cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                   'sample': f'{n_cpu} subframes of the same workload through oracle/lte_oracle.py (NumPy fp64), 1 process'}
            cpu_ref = reference_subframes_per_s(6)
            if not args.no_extra:
                extra = extra_configs(torch, dev, peak, world)

    if rank == 0:
        cfgd = bench_config(args, world)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': cfgd, 'pipeline': pipeline_used, 'batches_in_flight': inflight,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': pipe.h2d_bytes_per_batch * args.batches,
                        'd2h_bytes_per_step': pipe.d2h_bytes_per_batch * args.batches,
                        'api': 'LinkEngine.stream_host_batches (lte_b200/host_stream.py): pinned host bits -> H2D -> '
                               'lte_bits_to_indices -> chain -> int64 error counts D2H, every batch',
                        'numa_cpus_bound': None if numa is None else len(numa)},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
                'cpu_baseline_reference': cpu_ref, 'stages': stages, 'extra': extra,
                'ber': (totals.double() / max(bits_total, 1)).tolist()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_fp32_peak(eng, nat, torch, dev):
    import ctypes as C
    sink = torch.empty(148 * 8 * 256 * 2, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    nat.lib.lte_fp32_peak_launch(C.c_void_p(sink.data_ptr()), 2000, st)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(3):
        a.record()
        fl = nat.lib.lte_fp32_peak_launch(C.c_void_p(sink.data_ptr()), 20000, st)
        b.record()
        torch.cuda.synchronize(dev)
        best = max(best, fl / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


def time_stages(eng, ws, chan, snr_rows, idx, nbits, seed, B, S, R, nat, torch, dev, pipeline, reps=5):
    """Average device time of every stage kernel of the pipeline the bench ran, over `reps` launches after one
    warm-up (channel_* include their Jakes coefficient kernel)."""
    per = R * chan.num_taps * nat.LTE_JAKES_TONES
    ph = eng.random_phases(B, per, seed, 0, out=ws['phases'].view(-1)[:B * per].view(B, per))
    awgn = eng.awgn_desc(ws['power'], snr_rows, seed, 0, combine=True)
    if pipeline == 'spectral':
        calls = {
            'tx_spectral': lambda: eng.tx_spectral(S, idx, out_G=ws['G'], out_tail=ws['tail']),
            'channel_spectral': lambda: eng.channel_spectral(idx, ws['G'], ws['tail'], chan, B, R, S, ph, out=ws['Yd'],
                                                             power=ws['power'], compact=True, out_pilots=ws['Yp']),
            'mrc_demap_count_compact': lambda: eng.mrc_demap_count_compact(ws['Yd'], None, idx, B, R, S, nbits=nbits,
                                                                           errors=ws['errors'], awgn=awgn, Yp=ws['Yp']),
        }
    elif pipeline == 'fused':
        calls = {
            'tx_map_ifft': lambda: eng.modulate(S, idx=idx, want_stats=False, out=ws['tx']),
            'channel_rx_fft': lambda: eng.channel_rx_fft(ws['tx'], chan, B, R, S, ph, nat.WINDOW_USEFUL, out=ws['Y'],
                                                         power=ws['power']),
            'crs_ls_interp': lambda: eng.estimate(ws['Y'], B * R, S, nat.WINDOW_USEFUL, out=ws['H'], awgn=awgn),
            'mrc_demap_count': lambda: eng.mrc_demap_count(ws['Y'], ws['H'], idx, B, R, S, nbits=nbits,
                                                           errors=ws['errors'], awgn=awgn),
        }
    else:
        if 'faded' not in ws:
            ws['faded'] = torch.empty((B, R, S * eng.L), dtype=torch.complex64, device=dev)
        calls = {
            'tx_map_ifft': lambda: eng.modulate(S, idx=idx, want_stats=False, out=ws['tx']),
            'channel_tdl': lambda: eng.channel(ws['tx'], chan, B, R, phases=ph, out=ws['faded'], power=ws['power']),
            'rx_fft': lambda: eng.rx_fft(ws['faded'], B * R, S, nat.WINDOW_USEFUL, power=ws['power'], snr_lin=snr_rows,
                                         seed=seed, out=ws['Y'], noise_domain=1),
            'crs_ls_interp': lambda: eng.estimate(ws['Y'], B * R, S, nat.WINDOW_USEFUL, out=ws['H']),
            'mrc_demap_count': lambda: eng.mrc_demap_count(ws['Y'], ws['H'], idx, B, R, S, nbits=nbits,
                                                           errors=ws['errors']),
        }
    return _time_calls(calls, torch, dev, reps)


def _time_calls(calls, torch, dev, reps):
    out = {}
    for name, fn in calls.items():
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        out[name] = {'ms': a.elapsed_time(b) / reps}
    return out


# ------------------------------------------------------------------------------------ the other BASELINE configs
def extra_configs(torch, dev, hbm_peak, world):
    """Device-timed throughput of the other BASELINE.json configs through the package's batched sweeps, each a few
    batches after one warm-up (inputs generated on the device; no roofline claim beyond bytes of the windowed grid
    per second).  Config 3 proper is the headline chain over Vehicular_A at 30 km/h."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for, tables

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps * 1e-3

    out = {}
    snr16 = torch.tensor([10 ** (s / 10) for s in SNR_POINTS], dtype=torch.float32, device=dev)
    # config 3 proper: SIMO 1x4 MRC 20 MHz 64-QAM over Vehicular_A, 30 km/h (fused time-domain kernel, K = 2)
    cfg = LTEConfig(20.0, 15.0, '64-QAM')
    eng = LinkEngine.from_config(cfg, device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Vehicular_A', FC_GHZ, 30.0)
    B = 4096
    ws = eng.workspace(B, 14, 4, fading=True, fused=True, lazy=True)
    rows = snr16.repeat(B // 16).repeat_interleave(4).contiguous()
    t = timed(lambda: eng.simo_ber(ws, chan, rows, 1, fused=True, noise_domain=3))
    out['config3_vehicular_a_30kmh'] = {'subframes_per_s': B / t, 'ms_per_batch': t * 1e3, 'batch': B,
                                        'pipeline': 'spectral' if ws.get('spectral') else 'fused',
                                        'workload': '20MHz 64-QAM SIMO 1x4 MRC, Vehicular_A 30 km/h, 16 SNR points'}
    del ws
    # config 1: SISO 5 MHz QPSK AWGN, 10 SNR points
    cfg = LTEConfig(5.0, 15.0, 'QPSK')
    eng = LinkEngine.from_config(cfg, device=dev)
    chan = chan_for('awgn', cfg.fs)
    B = 16000
    rows = torch.tensor([10 ** (s / 10) for s in range(0, 20, 2)], dtype=torch.float32, device=dev).repeat(B // 10).contiguous()
    t = timed(lambda: eng.siso_ber(chan, rows, 14, 1))
    out['config1_siso_5mhz_qpsk_awgn'] = {'subframes_per_s': B / t, 'ms_per_batch': t * 1e3, 'batch': B,
                                          'workload': 'SISO 5MHz QPSK AWGN, 10 SNR points, ZF'}
    # config 2: SISO SC-FDM 10 MHz 16-QAM Pedestrian_A, BER + PAPR histogram in one pass
    cfg = LTEConfig(10.0, 15.0, '16-QAM')
    eng = LinkEngine.from_config(cfg, device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', FC_GHZ, 3.0)
    B = 8192
    rows = snr16.repeat(B // 16).contiguous()
    hist = torch.zeros(200, dtype=torch.int64, device=dev)
    t = timed(lambda: eng.siso_ber(chan, rows, 14, 1, sc_fdm=True, papr_hist=hist))
    out['config2_scfdm_10mhz_16qam_peda'] = {'subframes_per_s': B / t, 'ms_per_batch': t * 1e3, 'batch': B,
                                             'workload': 'SISO SC-FDM 10MHz 16-QAM Pedestrian_A 3 km/h, BER + per-symbol PAPR histogram'}
    # config 4: 2x2 SFBC 20 MHz 16-QAM Pedestrian_A; time to 1e6 subframes on the run's GPUs (weak scaling assumed)
    cfg = LTEConfig(20.0, 15.0, '16-QAM')
    eng0 = LinkEngine.from_config(cfg, device=dev)
    eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(2, eng0.Np), device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', FC_GHZ, 3.0)
    B = 2048
    rows = snr16.repeat(B // 16).repeat_interleave(2).contiguous()
    t = timed(lambda: eng.sfbc_ber(chan, rows, 14, 2, 1))
    out['config4_sfbc_2x2_20mhz_16qam'] = {'subframes_per_s': B / t, 'ms_per_batch': t * 1e3, 'batch': B,
                                           'seconds_for_1e6_subframes_at_this_n_gpus': 1e6 / (B / t * world),
                                           'workload': '2x2 SFBC-Alamouti 20MHz 16-QAM Pedestrian_A 3 km/h'}
    # config 5: 4x4 spatial multiplexing 20 MHz 64-QAM MMSE, rank 4 codebook precoder, one SNR point per launch
    cfg = LTEConfig(20.0, 15.0, '64-QAM')
    eng0 = LinkEngine.from_config(cfg, device=dev)
    eng = LinkEngine.from_config(cfg, pilot_sets=tables.mimo_pilot_sets(4, eng0.Np), device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', FC_GHZ, 3.0, gain_conversions=3)
    from core.codebook_lte import LTECodebook
    W = LTECodebook(4, transmission_mode='TM4', rank=4).get_precoder(0)
    B = 512
    t = timed(lambda: eng.sm_ber(chan, W, 20.0, B, 14, 4, 'MMSE', 1))
    out['config5_sm_4x4_20mhz_64qam_mmse'] = {'subframes_per_s': B / t, 'ms_per_batch': t * 1e3, 'batch': B,
                                              'workload': '4x4 spatial multiplexing 20MHz 64-QAM MMSE rank 4, 14 symbols per stream'}
    return out


_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line: everything else written to fd 1 (NCCL prints its version
    # banner there from native code) is sent to stderr; emit() writes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--trials', type=int, default=256, help='trials per SNR point per GPU per batch')
    ap.add_argument('--inflight', type=int, default=0,
                    help='batches in flight side by side (workspaces / streams); 0 = the engine picks (2 through the spectral link)')
    ap.add_argument('--batches', type=int, default=32, help='chain passes (batches of 16 x trials subframes) per step')
    ap.add_argument('--pipeline', default='spectral', choices=['spectral', 'fused', 'staged'],
                    help='spectral: spectral link (one forward transform per OFDM symbol, compact grid, lazy AWGN); '
                         'fused: time-domain channel + RX FFT in one kernel; staged: one kernel per reference stage')
    ap.add_argument('--cpu-subframes', type=int, default=24, help='bounded sample for cpu_baseline')
    ap.add_argument('--no-extra', action='store_true', help='skip the other BASELINE configs')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_gpu(args, rank, world)


if __name__ == '__main__':
    main()
