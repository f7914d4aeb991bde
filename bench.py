#!/usr/bin/env python
"""Headline benchmark: LTE subframes/s, 20 MHz 64-QAM SIMO 1x4 MRC over ITU Rayleigh.

    python bench.py --gpus N --steps K --warmup W            # CUDA engine (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host cores)

A step is one pass of the whole link chain (bits -> QAM/grid/IFFT/CP -> 4 fading links ->
AWGN -> FFT -> CRS LS estimate -> MRC -> hard demap -> error count) over a batch of
`--trials` subframes at each of the 16 SNR points 0..30 dB, on every GPU (weak scaling).
One JSON line is printed by rank 0.
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


def workload_name(trials):
    return (f'20MHz(N=2048,cp=144) 64-QAM SIMO 1x{R_ANT} MRC, ITU {PROFILE} {VELOCITY:g} km/h @ {FC_GHZ:g} GHz '
            f'(time-domain TDL, 16-tone Jakes), CRS LS estimate per 14 symbols, 16 SNR points 0..30 dB x '
            f'{trials} trials')


# ------------------------------------------------------------------------------------ algorithmic bytes
def stage_bytes(N=2048, cp=144, Nd=999, Np=200, Nc=1200, b=6, R=R_ANT, S=S_SUBFRAME):
    """Compulsory bytes per subframe of each stage kernel (complex64 = 8 B)."""
    L = N + cp
    return {
        'tx_map_ifft': S * Nd * b / 8 + S * L * 8,
        'channel_tdl': S * L * 8 + R * S * L * 8,
        'rx_fft': R * S * L * 8 + R * S * (Nd + Np) * 8,
        'crs_ls_interp': R * Np * 8 + R * Nd * 8,
        'mrc_demap_count': R * S * Nd * 8 + R * Nd * 8 + S * Nd * b / 8 + 8,
        # fused pipeline: inputs of the first + outputs of the last fused stage
        'channel_rx_fft': S * L * 8 + R * S * (Nd + Np) * 8,
    }


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures under profiles/
# (round 1, 4096 subframes per launch; summaries: profiles/r01_fused_ncu_summary.md for the fused pipeline,
# r01_staged_ncu_summary.md for channel_tdl / rx_fft)
NCU_TRAFFIC = {
    'tx_map_ifft': 1.0118e9, 'channel_rx_fft': 3.1757e9, 'crs_ls_interp_awgn': 0.2719e9,
    'mrc_demap_count_awgn': 2.4214e9, 'channel_tdl': 4.99e9, 'rx_fft': 5.93e9,
}

STAGED = ('tx_map_ifft', 'channel_tdl', 'rx_fft', 'crs_ls_interp', 'mrc_demap_count')


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
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                mx = max(mx, float(f[2]))
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
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
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------ CPU arm (oracle port)
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


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is
    Python and does not travel to the GPU box) on all host cores."""
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
    sample = f'{per_step} subframes per step (one 14-symbol subframe per task, SNR cycling 0..30 dB), {cores} processes'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * per_step / value,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(args.trials), 'host': 'oracle/lte_oracle.py NumPy port'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
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

    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    cfg = LTEConfig(20.0, 15.0, '64-QAM', 'normal')
    eng = LinkEngine.from_config(cfg, device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, PROFILE, FC_GHZ, VELOCITY)
    n_snr = len(SNR_POINTS)
    B = args.trials * n_snr                     # subframes per step per GPU
    S, R = S_SUBFRAME, R_ANT
    fused = args.pipeline == 'fused'
    ws = eng.workspace(B, S, R, fading=True, fused=fused)
    snr_lin = torch.tensor([10 ** (s / 10) for s in SNR_POINTS], dtype=torch.float32, device=dev)
    snr_rows = snr_lin.repeat(args.trials).repeat_interleave(R).contiguous()       # [B*R], SNR fastest over b
    nbits = S * eng.Nd * eng.bps
    seed = 2026

    # inputs resident in HBM before the timed region: one batch of transmitted symbol indices
    idx = eng.random_indices(B, S, seed, stream_id0=rank * B)
    totals = torch.zeros(n_snr, dtype=torch.int64, device=dev)

    def step(i):
        sid0 = ((i * world) + rank) * B          # global stream ids: independent of the GPU count
        err = eng.simo_ber(ws, chan, snr_rows, seed, stream_id0=sid0, idx=idx, nbits=nbits, fused=fused,
                           noise_domain=3 if fused else 1)
        totals.add_(err.view(args.trials, n_snr).sum(0))

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    totals.zero_()
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
    value = world * B * args.steps / (ms_total * 1e-3)
    bits_total = world * B * args.steps * nbits / n_snr

    # ---- end to end through the host-buffer API: packed bits in pinned memory -> errors on host
    nbytes = (nbits + 7) // 8
    host_bits = torch.from_numpy(np.random.RandomState(7 + rank).randint(0, 256, (B, nbytes), dtype=np.uint8)).pin_memory()
    host_err = torch.empty(B, dtype=torch.int64).pin_memory()
    # The host-buffer API double-buffers its device input: the H2D copy of step i+1 runs on a copy
    # stream while step i computes; both are inside the timed region.
    dev_bits = [torch.empty((B, nbytes), dtype=torch.uint8, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(dev)

    def e2e_upload(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])
            dev_bits[k].copy_(host_bits, non_blocking=True)
            ev_copied[k].record(copy_stream)

    def e2e_step(i, last):
        k = i % 2
        if not last:
            e2e_upload(i + 1)
        main.wait_event(ev_copied[k])
        ix = eng.bits_to_indices(dev_bits[k], nbits, S, packed=True)
        sid0 = ((i * world) + rank) * B
        err = eng.simo_ber(ws, chan, snr_rows, seed + 1, stream_id0=sid0, idx=ix, nbits=nbits, fused=fused,
                           noise_domain=3 if fused else 1)
        host_err.copy_(err, non_blocking=True)
        ev_free[k].record(main)

    for k in range(2):
        ev_free[k].record(main)
    e2e_upload(0)
    for i in range(2):
        e2e_step(i, last=(i == 1))
    sync()
    e2e_steps = max(2, min(args.steps, 50))
    e0.record()
    e2e_upload(2)                               # every upload of a timed step is inside the timed region
    for i in range(e2e_steps):
        e2e_step(2 + i, last=(i == e2e_steps - 1))
    e1.record()
    sync()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(ms2.item()) * 1e-3)

    # ---- per-stage device time (CUDA events on the launching stream), rank 0 ---------------
    stages, roofline, cpu = None, None, None
    if rank == 0:
        stages = time_stages(eng, ws, chan, snr_rows, idx, nbits, seed, B, S, R, nat, torch, dev, fused=fused)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
        sb = stage_bytes()
        for k in stages:
            kb = sb[k.replace('_awgn', '')]
            stages[k]['algo_bytes_per_launch'] = kb * B
            stages[k]['gbs'] = kb * B / (stages[k]['ms'] * 1e-3) / 1e9
            stages[k]['frac'] = stages[k]['gbs'] / peak
        dom = max(stages, key=lambda k: stages[k]['ms'])
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': stages[dom]['gbs'], 'peak': peak, 'unit': 'GB/s',
                    'frac': stages[dom]['frac'], 'traffic': NCU_TRAFFIC.get(dom), 'peak_source': peak_src,
                    'pipeline_bytes_frac': value / world * sum(sb[k.replace('_awgn', '')] for k in stages) / 1e9 / peak,
                    'pipeline_unfused_equivalent_frac': value / world * sum(sb[k] for k in STAGED) / 1e9 / peak}
        if dom == 'channel_rx_fft':
            # the fused kernel trades HBM traffic for arithmetic: ncu (profiles/r01_fused_ncu_summary.md) has it at
            # 62 % fp32-FMA-pipe busy / 57 % issue-active with DRAM at 28 %, i.e. bound by the FMA pipe, not by HBM
            roofline['limiter'] = 'fp32 FMA pipe (ncu: pipe_fma_cycles_active 62 %, dram 28 % of peak); see DESIGN.md 4.3'
        if world == 1:
            n_cpu = args.cpu_subframes
            v = cpu_subframes_per_s(n_cpu, 1)
            cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                   'sample': f'{n_cpu} subframes of the same workload through oracle/lte_oracle.py (NumPy fp64), 1 process'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': workload_name(args.trials), 'subframes_per_step_per_gpu': B,
                           'pipeline': args.pipeline,
                           'l2_policy': 'inputs and intermediates (%.1f GB per step) exceed the 126 MB L2' %
                                        (B * (0.88e6 if fused else 1.8e6) / 1e9),
                           'parallelism': f'trial sharding x{world}, one int64[16] allreduce'},
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * nbytes,
                        'd2h_bytes_per_step': B * 8},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
                'stages': stages, 'ber': (totals.double() / max(bits_total, 1)).tolist()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def time_stages(eng, ws, chan, snr_rows, idx, nbits, seed, B, S, R, nat, torch, dev, reps=5, only=None,
                fused=False):
    """Average device time of every stage kernel over `reps` launches (after one warm-up).
    fused: the stages of the fused pipeline (channel_rx_fft includes its Jakes coefficient kernel)."""
    per = R * chan.num_taps * nat.LTE_JAKES_TONES
    ph = eng.random_phases(B, per, seed, 0, out=ws['phases'].view(-1)[:B * per].view(B, per))
    if fused:
        awgn = eng.awgn_desc(ws['power'], snr_rows, seed, 0, combine=True)
        calls = {
            'tx_map_ifft': lambda: eng.modulate(S, idx=idx, want_stats=False, out=ws['tx']),
            'channel_rx_fft': lambda: eng.channel_rx_fft(ws['tx'], chan, B, R, S, ph, nat.WINDOW_USEFUL, out=ws['Y'],
                                                         power=ws['power']),
            'crs_ls_interp_awgn': lambda: eng.estimate(ws['Y'], B * R, S, nat.WINDOW_USEFUL, out=ws['H'], awgn=awgn),
            'mrc_demap_count_awgn': lambda: eng.mrc_demap_count(ws['Y'], ws['H'], idx, B, R, S, nbits=nbits,
                                                                errors=ws['errors'], awgn=awgn),
        }
    else:
        if 'faded' not in ws:
            ws['faded'] = torch.empty((B, R, S * eng.L), dtype=torch.complex64, device=dev)
        calls = _staged_calls(eng, ws, chan, snr_rows, idx, nbits, seed, B, S, R, nat, ph)
    return _time_calls(calls, torch, dev, reps, only)


def _staged_calls(eng, ws, chan, snr_rows, idx, nbits, seed, B, S, R, nat, ph):
    return {
        'tx_map_ifft': lambda: eng.modulate(S, idx=idx, want_stats=False, out=ws['tx']),
        'channel_tdl': lambda: eng.channel(ws['tx'], chan, B, R, phases=ph, out=ws['faded'], power=ws['power']),
        'rx_fft': lambda: eng.rx_fft(ws['faded'], B * R, S, nat.WINDOW_USEFUL, power=ws['power'], snr_lin=snr_rows,
                                     seed=seed, out=ws['Y'], noise_domain=1),
        'crs_ls_interp': lambda: eng.estimate(ws['Y'], B * R, S, nat.WINDOW_USEFUL, out=ws['H']),
        'mrc_demap_count': lambda: eng.mrc_demap_count(ws['Y'], ws['H'], idx, B, R, S, nbits=nbits,
                                                       errors=ws['errors']),
    }


def _time_calls(calls, torch, dev, reps, only):
    out = {}
    for name, fn in calls.items():
        if only and name not in only:
            continue
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
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--trials', type=int, default=256, help='trials per SNR point per GPU per step')
    ap.add_argument('--pipeline', default='fused', choices=['fused', 'staged'],
                    help='fused: channel + RX FFT in one kernel, AWGN added lazily by the consumers; '
                         'staged: one kernel per reference stage')
    ap.add_argument('--cpu-subframes', type=int, default=24, help='bounded sample for cpu_baseline')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_gpu(args, rank, world)


if __name__ == '__main__':
    main()
