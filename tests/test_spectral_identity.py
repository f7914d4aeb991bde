"""The spectral form of the low-Doppler fading link (tests/spectral_ref.py, the fp64 restatement of
csrc/spectral.cu) against the oracle's sample-by-sample tapped delay line + CP strip + FFT
(core/rayleighchannel.py:44-58, core/lte_receiver.py:444-491): occupied bins to 1e-6, stream power to 1e-6."""
import numpy as np
import pytest

import spectral_ref as SR
from oracle import lte_oracle as O


@pytest.mark.parametrize('bw,mod,prof,R,S', [(20.0, '64-QAM', 'Pedestrian_A', 4, 3), (5.0, 'QPSK', 'Pedestrian_A', 2, 15),
                                             (1.25, '16-QAM', 'Pedestrian_B', 3, 4), (10.0, '16-QAM', 'Vehicular_A', 1, 2)])
def test_spectral_form_matches_time_domain(bw, mod, prof, R, S):
    num = O.Numerology(bw, 15.0, mod)
    rs = np.random.RandomState(3)
    data_idx, _ = O.grid_indices(num.N, num.Nc)
    bits = rs.randint(0, 2, len(data_idx) * num.bits_per_symbol * S)
    sig, _ = O.modulate_stream(bits, num)
    d, g = O.itu_taps(prof, num.fs)
    if max(d) > num.cp_length:
        pytest.skip('delay spread beyond the cyclic prefix')
    fD = O.doppler_hz(2.0, 3.0)
    ph = 2 * np.pi * rs.rand(R, len(d), 16)
    gl = (num.N - num.Nc) // 2
    kept = np.arange(gl, gl + num.Nc)
    Ys, Ps = SR.spectral_rx(sig, num, fD, list(d), g, ph, kept)
    Yt, Pt = SR.time_domain_rx(sig, num, fD, list(d), g, ph)
    assert np.abs(Ys[:, :, kept] - Yt[:, :, kept]).max() / np.abs(Yt[:, :, kept]).max() < 1e-6
    assert np.abs(Ps / Pt - 1).max() < 1e-6
