"""Statistical anchors of the sweep engine's noise placements (they use Philox draws, so there is nothing to
replay): (1) the pipeline the bench times -- fast path + one combined noise draw per MRC output ('combined',
noise_domain 3) -- against per-antenna noise on the kept bins (noise_domain 1) at the headline geometry, within
3 sigma of the binomial confidence interval; (2) the GPU BER curve against an oracle Monte Carlo (the fp64 NumPy
restatement of the reference with its own NumPy draws), again within 3 sigma of the combined binomial error."""
import numpy as np
import pytest
import torch

from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu


def test_combined_noise_on_the_fast_path_matches_per_antenna_noise_at_the_headline_geometry():
    """20 MHz, 64-QAM, 1x4, Pedestrian_A 3 km/h -- the bench's geometry, 256 streams per SNR point and arm.  The
    per-stream BER is heavy-tailed (one fading realisation per stream), so the standard error comes from the
    per-stream sample spread, not from a binomial over bits."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    cfg = LTEConfig(20.0, 15.0, '64-QAM')
    eng = LinkEngine.from_config(cfg)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    snr = [0.0, 6.0, 12.0, 18.0]
    n, S, R = 256, 14, 4
    B = n * len(snr)
    rows = torch.tensor([10 ** (s / 10) for s in snr], dtype=torch.float32, device='cuda').repeat(n).repeat_interleave(R).contiguous()
    nbits = S * eng.Nd * eng.bps
    wa = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
    wb = eng.workspace(B, S, R, fading=True)
    fast = eng.simo_ber(wa, chan, rows, 11, stream_id0=0, fused=True, noise_domain=3).view(n, len(snr)).double().cpu() / nbits
    assert wa['spectral']                                   # the pipeline the bench times
    ref = eng.simo_ber(wb, chan, rows, 12, stream_id0=0, fused=False, noise_domain=1).view(n, len(snr)).double().cpu() / nbits
    for i in range(len(snr)):
        pa, pb = float(fast[:, i].mean()), float(ref[:, i].mean())
        se = (float(fast[:, i].var(unbiased=True)) / n + float(ref[:, i].var(unbiased=True)) / n) ** 0.5
        assert abs(pa - pb) <= 3 * se + 1e-7, (snr[i], pa, pb, se)
    assert fast[:, 0].mean() > 10 * fast[:, -1].mean() > 0


def test_gpu_ber_curve_matches_an_oracle_monte_carlo():
    """48 oracle subframes per SNR point (5 MHz, 16-QAM, 1x2 MRC, Pedestrian_A 3 km/h) with NumPy draws against
    1536 GPU subframes per point with Philox draws through the default sweep."""
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    from lte_b200 import sweep
    cfg = LTEConfig(5.0, 15.0, '16-QAM')
    eng = LinkEngine.from_config(cfg)
    chan = chan_for('rayleigh_mp', cfg.fs, 'Pedestrian_A', 2.0, 3.0)
    snr = [2.0, 10.0, 18.0]
    g = sweep.simo_sweep(eng, chan, snr, 1536, 2, seed=5, batch_trials=512)
    num = O.Numerology(5.0, 15.0, '16-QAM')
    nbits = 14 * eng.Nd * eng.bps
    rs = np.random.RandomState(99)
    n_or = 48
    taps = len(O.ITU['Pedestrian_A'][0])
    n = 14 * num.L
    for i, s in enumerate(snr):
        e = 0
        per = []
        for _ in range(n_or):
            bits = rs.randint(0, 2, nbits)
            r = O.simulate_simo(bits, s, num, 2, 'rayleigh_mp', 'Pedestrian_A', 2.0, 3.0,
                                phases=2 * np.pi * rs.rand(2, taps, 16), z=rs.standard_normal((2, 2, n)))
            e += r['errors']
            per.append(r['errors'] / nbits)
        p_or, p_gpu = e / (n_or * nbits), float(g['ber'][i])
        # subframe-level standard error of the oracle mean (fading makes the per-subframe BER heavy-tailed) plus the
        # much smaller one of the GPU mean
        se = np.std(per, ddof=1) / np.sqrt(n_or) * np.sqrt(1 + n_or / 1536.0)
        assert abs(p_or - p_gpu) <= 3 * se + 1e-6, (s, p_or, p_gpu, se)
    assert g['ber'][0] > g['ber'][-1]
