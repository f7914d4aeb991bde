#!/usr/bin/env python
"""Time (or profile under ncu) the stage kernels of one pipeline of the headline workload in isolation.
usage: python tools/stage_bench.py [--pipeline spectral|fused|staged] [--trials 256] [--reps 5] [--profile NAME]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'ofdm-lte_b200')]
import torch  # noqa: E402

import bench  # noqa: E402
from config import LTEConfig  # noqa: E402
from lte_b200 import LinkEngine, chan_for  # noqa: E402
from lte_b200 import _native as nat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pipeline', default='spectral', choices=['spectral', 'fused', 'staged'])
    ap.add_argument('--trials', type=int, default=256)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--profile', default=bench.PROFILE)
    ap.add_argument('--velocity', type=float, default=bench.VELOCITY)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    cfg = LTEConfig(20.0, 15.0, '64-QAM')
    eng = LinkEngine.from_config(cfg, device=dev)
    chan = chan_for('rayleigh_mp', cfg.fs, a.profile, bench.FC_GHZ, a.velocity)
    B, S, R = a.trials * 16, 14, 4
    fused, spectral = a.pipeline != 'staged', a.pipeline == 'spectral'
    ws = eng.workspace(B, S, R, fading=True, fused=fused, lazy=spectral)
    snr = torch.tensor([10 ** (s / 10) for s in bench.SNR_POINTS], dtype=torch.float32, device=dev)
    snr_rows = snr.repeat(a.trials).repeat_interleave(R).contiguous()
    idx = eng.random_indices(B, S, 1, 0)
    nbits = S * eng.Nd * eng.bps
    eng.simo_ber(ws, chan, snr_rows, 1, 0, idx=idx, nbits=nbits, fused=fused, spectral=spectral,
                 noise_domain=3 if fused else 1)      # populate every buffer
    used = 'spectral' if ws.get('spectral') else ('fused' if fused and 'faded' not in ws else 'staged')
    res = bench.time_stages(eng, ws, chan, snr_rows, idx, nbits, 1, B, S, R, nat, torch, dev, used, reps=a.reps)
    sb = bench.stage_bytes()
    for k, v in res.items():
        v['frac'] = sb[k] * B / (v['ms'] * 1e-3) / 1e9 / 6467.7
    print(json.dumps({'pipeline': used, 'subframes_per_launch': B, 'stages': res}))


if __name__ == '__main__':
    main()
