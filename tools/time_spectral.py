"""Device time of the spectral link's two kernels next to the kernels they replace (headline geometry)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'ofdm-lte_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from config import LTEConfig  # noqa: E402
from lte_b200 import LinkEngine, chan_for  # noqa: E402
from lte_b200 import _native as nat  # noqa: E402


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    prof = sys.argv[2] if len(sys.argv) > 2 else 'Pedestrian_A'
    cfg = LTEConfig(20.0, 15.0, '64-QAM', 'normal')
    eng = LinkEngine.from_config(cfg)
    chan = chan_for('rayleigh_mp', cfg.fs, prof, 2.0, 3.0)
    S, R = 14, 4
    idx = eng.random_indices(B, S, 1)
    per = R * chan.num_taps * nat.LTE_JAKES_TONES
    ph = eng.random_phases(B, per, 1)
    k0, nk = eng.window(nat.WINDOW_USEFUL)
    G = torch.empty((B * S, nk), dtype=torch.complex64, device='cuda')
    tail = torch.empty((B * S, eng.cp), dtype=torch.complex64, device='cuda')
    Y = torch.empty((B * R, S, nk), dtype=torch.complex64, device='cuda')
    tx = torch.empty((B, S * eng.L), dtype=torch.complex64, device='cuda')
    power = torch.zeros((B, R), dtype=torch.float64, device='cuda')
    wsb = eng.spectral_workspace_bytes(chan, B, R, S)
    ws = torch.empty(wsb // 4, dtype=torch.float32, device='cuda')
    out = {'B': B, 'profile': prof}
    out['tx_spectral_ms'] = timeit(lambda: eng.tx_spectral(S, idx, out_G=G, out_tail=tail))
    out['channel_spectral_ms'] = timeit(lambda: eng.channel_spectral(idx, G, tail, chan, B, R, S, ph, out=Y, power=power, workspace=ws))
    nslot = 1
    Yd = torch.empty((B * R, S, 2 * eng.ndp), dtype=torch.complex64, device='cuda')
    Yp = torch.empty((B * R, nslot, 2 * eng.npp), dtype=torch.complex64, device='cuda')
    Hp = torch.empty((B * R, nslot, eng.Np), dtype=torch.complex64, device='cuda')
    errors = torch.zeros(B, dtype=torch.int64, device='cuda')
    out['channel_spectral_compact_ms'] = timeit(lambda: eng.channel_spectral(idx, G, tail, chan, B, R, S, ph, out=Yd, power=power,
                                                                             workspace=ws, compact=True, out_pilots=Yp))
    snr = torch.full((B * R,), 100.0, dtype=torch.float32, device='cuda')
    awgn = eng.awgn_desc(power, snr, 3, 0, combine=True)
    out['crs_ls_compact_ms'] = timeit(lambda: eng.estimate_compact(Yp, B * R, S, out=Hp, awgn=awgn))
    out['mrc_compact_ms'] = timeit(lambda: eng.mrc_demap_count_compact(Yd, Hp, idx, B, R, S, errors=errors, awgn=awgn))
    out['crs_mrc_compact_ms'] = timeit(lambda: eng.mrc_demap_count_compact(Yd, None, idx, B, R, S, errors=errors, awgn=awgn, Yp=Yp))
    wsl = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    out['simo_ber_spectral_ms'] = timeit(lambda: eng.simo_ber(wsl, chan, snr, 5, idx=idx, fused=True, noise_domain=3))
    out['tx_map_ifft_ms'] = timeit(lambda: eng.modulate(S, idx=idx, want_stats=False, out=tx))
    out['channel_rx_fft_ms'] = timeit(lambda: eng.channel_rx_fft(tx, chan, B, R, S, ph, nat.WINDOW_USEFUL, out=Y, power=power))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
