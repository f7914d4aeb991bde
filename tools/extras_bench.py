#!/usr/bin/env python
"""Device-timed throughput of the widened rows of SURVEY 8(f): beamforming link, payload sweep adapter and
the coded chain (CUDA events around the engine calls, warm-up first).  Prints one JSON object; the
oracle timings (`--cpu`) give the CPU figure beside each, on a bounded sample.
usage: python tools/extras_bench.py [--cpu] [--reps 5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from config import LTEConfig  # noqa: E402
from lte_b200 import LinkEngine, chan_for  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    out = {}
    dev = torch.device('cuda', 0)

    # ---- beamforming: 10 MHz 64-QAM 4x2, 14 symbols per stream, 16 SNR points x 512 trials
    cfg = LTEConfig(10.0, 15.0, '64-QAM')
    eng = LinkEngine.from_config(cfg, device=dev)
    from core.codebook_lte import LTECodebook
    cb = LTECodebook(4, 'TM6').codebook
    B, S, R, T = 8192, 14, 2, 4
    idx = eng.random_indices(B, S, 1, 0)
    h = eng.random_channel(B, R, T, 1, 0)
    nstd = torch.full((B,), 0.1, dtype=torch.float32, device=dev)
    W, heff, pmi, gain = eng.bf_weights(h, cb, mode='CODEBOOK')
    ms_w = timed(lambda: eng.bf_weights(h, cb, mode='CODEBOOK'), a.reps)
    ms_l = timed(lambda: eng.bf_link(idx, h, W, heff, nstd, S, seed=1), a.reps)
    sym = B * S * eng.Nd
    out['beamforming'] = {'workload': f'10MHz 64-QAM {T}x{R} codebook, {B} streams x {S} symbols',
                          'bf_weights_ms': ms_w, 'bf_link_ms': ms_l, 'data_symbols_per_s': sym / (ms_l * 1e-3),
                          'stream_symbols_per_s': B * S / (ms_l * 1e-3),
                          # algorithmic bytes: one index byte in, one 8-byte counter per CTA out
                          'bf_link_algo_gbs': sym / (ms_l * 1e-3) / 1e9}

    # ---- coded chain: 5 MHz QPSK AWGN, TB 1000 bits (one K=1024 block), 2048 streams
    cfg2 = LTEConfig(5.0, 15.0, 'QPSK')
    e2 = LinkEngine.from_config(cfg2, device=dev)
    Bc, A = 2048, 1000
    bits = (e2.random_indices(Bc, -(-A // e2.Nd), 2, 0)[:, :A] & 1).contiguous()
    plan = e2.coding_plan(A)
    awgn = chan_for('awgn', cfg2.fs, 'Pedestrian_A', 2.0, 0.0)
    snr = torch.full((Bc,), 4.0, dtype=torch.float32, device=dev)
    ms_enc = timed(lambda: e2.tb_encode(bits, plan), a.reps)
    llr = torch.randn((Bc, plan.sumE), dtype=torch.float32, device=dev) * 3
    ms_dec = timed(lambda: e2.tb_decode(llr, plan, bits_tx=bits, want_bits=False), a.reps)
    ms_all = timed(lambda: e2.siso_coded_ber(bits, awgn, snr, 2, 0), a.reps)
    out['coded'] = {'workload': f'5MHz QPSK AWGN, TB {A} bits (K={plan.Kmax}, C={plan.C}), {Bc} transport blocks',
                    'tb_encode_ms': ms_enc, 'tb_decode_ms': ms_dec, 'chain_ms': ms_all,
                    'transport_blocks_per_s': Bc / (ms_all * 1e-3), 'info_bits_per_s': Bc * A / (ms_all * 1e-3),
                    'decoder_block_iterations_per_s': Bc * plan.C * 8.5 / (ms_dec * 1e-3)}
    # a long block
    A2 = 6120
    p2 = e2.coding_plan(A2)
    B2 = 1024
    llr2 = torch.randn((B2, p2.sumE), dtype=torch.float32, device=dev) * 3
    ms_dec2 = timed(lambda: e2.tb_decode(llr2, p2, want_bits=False), a.reps)
    out['coded']['tb_decode_K6144_ms'] = ms_dec2
    out['coded']['tb_decode_K6144_info_bits_per_s'] = B2 * A2 / (ms_dec2 * 1e-3)

    # ---- payload sweep adapter: 1.25 MHz, 3 modulations x (1,2,4,8) RX x 8 SNR x 32 iterations over 20 kbit
    from lte_b200.sweep import payload_sweep
    cfg3 = LTEConfig(1.25, 15.0, 'QPSK')
    pay = np.random.RandomState(0).randint(0, 2, 20000)
    snrs = list(np.arange(0, 16, 2.0))
    payload_sweep(cfg3, pay, snrs, 4, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    payload_sweep(cfg3, pay, snrs, 32, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    calls = 3 * 4 * len(snrs) * 32
    out['payload_sweep'] = {'workload': '1.25MHz, 3 modulations x RX{1,2,4,8} x 8 SNR x 32 iterations, 20 kbit payload',
                            'seconds': dt, 'simulate_calls_equivalent': calls, 'calls_per_s': calls / dt}

    # ---- BASELINE config 4: 2x2 SFBC 20 MHz 16-QAM over Pedestrian_A, 16 SNR points x 64 trials of one subframe
    from lte_b200 import tables
    from lte_b200.sweep import sfbc_sweep, sm_sweep
    cfg4 = LTEConfig(20.0, 15.0, '16-QAM')
    e4 = LinkEngine.from_config(cfg4, pilot_sets=tables.mimo_pilot_sets(2, 200), device=dev)
    ch4 = chan_for('rayleigh_mp', cfg4.fs, 'Pedestrian_A', 2.0, 3.0)
    B4, R4 = 1024, 2
    rows4 = torch.tensor([10 ** (s / 10) for s in range(0, 32, 2)], dtype=torch.float32, device=dev) \
        .repeat(B4 // 16).repeat_interleave(R4).contiguous()
    ms4 = timed(lambda: e4.sfbc_ber(ch4, rows4, 14, R4, 1, 0), a.reps)
    out['sfbc_cfg4'] = {'workload': f'20MHz 16-QAM 2x2 SFBC Pedestrian_A, {B4} subframes per pass', 'pass_ms': ms4,
                        'subframes_per_s': B4 / (ms4 * 1e-3)}
    # ---- BASELINE config 5: 4x4 SM 20 MHz 64-QAM MMSE rank 4, 256 subframes at one SNR point per launch
    cfg5 = LTEConfig(20.0, 15.0, '64-QAM')
    e5 = LinkEngine.from_config(cfg5, pilot_sets=tables.mimo_pilot_sets(4, 200), device=dev)
    ch5 = chan_for('rayleigh_mp', cfg5.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
    W5 = LTECodebook(4, transmission_mode='TM4', rank=4).get_precoder(0)
    B5 = 256
    ms5 = timed(lambda: e5.sm_ber(ch5, W5, 20.0, B5, 14, 4, 'MMSE', 1, 0), a.reps)
    out['sm_cfg5'] = {'workload': f'20MHz 64-QAM 4x4 SM rank 4 MMSE Pedestrian_A, {B5} subframes per pass',
                      'pass_ms': ms5, 'subframes_per_s': B5 / (ms5 * 1e-3)}

    # ---- shared-channel sweep (common random numbers along the SNR axis) on the headline geometry
    from lte_b200.sweep import simo_sweep_shared_channel
    cfgh = LTEConfig(20.0, 15.0, '64-QAM')
    eh = LinkEngine.from_config(cfgh, device=dev)
    chh = chan_for('rayleigh_mp', cfgh.fs, 'Pedestrian_A', 2.0, 3.0)
    snrs16 = [float(x) for x in range(0, 32, 2)]
    simo_sweep_shared_channel(eh, chh, snrs16, 4096, 4, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        simo_sweep_shared_channel(eh, chh, snrs16, 4096, 4, seed=1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out['shared_channel_sweep'] = {'workload': '20MHz 64-QAM SIMO-4 MRC Pedestrian_A, 4096 trials x 16 SNR points, one '
                                               'channel realisation per trial shared by its SNR points',
                                   'seconds_per_sweep': dt, 'link_evaluations_per_s': 4096 * 16 / dt,
                                   'channel_realisations_per_s': 4096 / dt}

    if a.cpu:
        from oracle import lte_oracle as O
        T_ = np.load(os.path.join(ROOT, 'tests', 'golden', 'coding_tables.npz'))
        qpp = {int(k): (int(f1), int(f2)) for k, f1, f2 in T_['qpp']}
        num = O.Numerology(5.0, 15.0, 'QPSK')
        t0 = time.perf_counter()
        n = 3
        for i in range(n):
            O.simulate_siso_coded(np.random.RandomState(i).randint(0, 2, A), 4.0, num, qpp)
        out['coded']['cpu_oracle_transport_blocks_per_s'] = n / (time.perf_counter() - t0)
        num2 = O.Numerology(10.0, 15.0, '64-QAM')
        t0 = time.perf_counter()
        bb = np.random.RandomState(0).randint(0, 2, 14 * 499 * 6)
        for i in range(n):
            O.simulate_beamforming(bb, 15.0, num2, 4, 2, 'static', global_seed=i)
        out['beamforming']['cpu_oracle_stream_symbols_per_s'] = n * 14 / (time.perf_counter() - t0)
        num4 = O.Numerology(20.0, 15.0, '16-QAM')
        b4 = np.random.RandomState(0).randint(0, 2, 14 * 998 * 4)
        t0 = time.perf_counter()
        O.simulate_sfbc(b4, 10.0, num4, 2, 'rayleigh_mp', 'Pedestrian_A', 2.0, 3.0)
        out['sfbc_cfg4']['cpu_oracle_subframes_per_s'] = 1 / (time.perf_counter() - t0)
        num5 = O.Numerology(20.0, 15.0, '64-QAM')
        b5 = np.random.RandomState(0).randint(0, 2, 999 * 6)
        t0 = time.perf_counter()
        O.simulate_sm(b5, num5, 4, 4, 4, 'MMSE', 20.0, 'rayleigh_mp', 'Pedestrian_A', 3.0, 2.0, global_seed=1)
        out['sm_cfg5']['cpu_oracle_subframes_per_s'] = 1 / 14 / (time.perf_counter() - t0)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
