"""M-point Bluestein DFT kernel (SC-FDM precoder) against numpy.fft for every LTE data-subcarrier
count (none is 2-3-5 smooth: 499 is prime, 999 = 27*37) and the SC-FDM golden case of config 2."""
import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M', [1, 2, 3, 62, 125, 249, 499, 749, 999, 1000, 1024])
def test_dft_m_matches_numpy(M):
    import torch
    from lte_b200 import LinkEngine
    eng = LinkEngine(128, 76, 9, 2, 1.92e6)
    rs = np.random.RandomState(M)
    x = (rs.standard_normal((5, M)) + 1j * rs.standard_normal((5, M))).astype(np.complex64)
    xt = torch.from_numpy(x).cuda()
    fwd = eng.dft_m(xt, M).cpu().numpy()
    want = np.fft.fft(x.astype(np.complex128), axis=1) / np.sqrt(M)
    assert rel_err(fwd, want) < 2e-6
    inv = eng.dft_m(torch.from_numpy(want.astype(np.complex64)).cuda(), M, inverse=True).cpu().numpy()
    assert rel_err(inv, x) < 2e-6
    # unitary: energy preserved
    assert abs(np.linalg.norm(fwd) / np.linalg.norm(x) - 1) < 1e-5


def test_dft_m_rejects_large_sizes():
    import torch
    from lte_b200 import LinkEngine
    from lte_b200._native import LteError
    eng = LinkEngine(128, 76, 9, 2, 1.92e6)
    x = torch.zeros(1500, dtype=torch.complex64, device='cuda')
    with pytest.raises(LteError):
        eng.dft_m(x, 1500)
