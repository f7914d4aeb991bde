"""Parity of the CUDA path (through the C ABI) against the oracle on identical injected
bits, tap phases and noise.  Integer stages bit-exact; floating-point stages within
1e-5 relative (fp32 pipeline vs fp64 oracle), the tolerance BASELINE.json states."""
import numpy as np
import pytest

from cases import BIG_RX_STRIDE, SIMO_CASES, SISO_CASES
from helpers import (golden_bits, golden_bits_rx, load_golden, numerology, oracle_simo, oracle_siso,
                     reference_draws, rel_err)
from oracle import lte_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def explain_mismatch(sym_oracle, modulation, bps, bits_a, bits_b):
    """Every differing bit must belong to a symbol whose oracle value lies within 1e-5 of a
    slicer boundary (hard parts)."""
    from gpu_chain import boundary_distance
    diff = np.flatnonzero(bits_a != bits_b)
    if len(diff) == 0:
        return True
    d = boundary_distance(sym_oracle, modulation)
    return bool(np.all(d[np.unique(diff // bps)] < 1e-5))


def assert_zf_close(got, want, o, num, case):
    """ZF output Y/(H+1e-6): 1e-5 relative wherever the division is well conditioned.  Bins whose
    interpolated |H| falls in a deep fade amplify the fp32 rounding of H by 1/|H| (an fp64 chain has
    the same conditioning, 1e9 times further down), so they are held to the l2 bound scaled by that
    amplification instead."""
    if o['H'] is None or not case.get('equalize', True):
        assert rel_err(got, want) < TOL
        return
    data_idx, _ = O.grid_indices(num.N, num.Nc)
    h = np.abs(o['H'][:, data_idx] + 1e-6).reshape(-1)
    good = h >= 0.05 * np.median(h)
    assert good.mean() > 0.9
    assert rel_err(got[good], want[good]) < TOL
    # the bins that needed the conditioning-scaled bound are few: a Rayleigh |H| falls below 5 % of its median
    # with probability 1 - exp(-ln2 * 0.05^2) = 0.17 %; interpolated estimates spread that over neighbours
    elem0 = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    assert np.mean(elem0 > TOL) < 0.03
    elem = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    assert np.median(elem) < TOL / 10
    assert np.all(elem * (h / np.median(h)) < 10 * TOL)


@pytest.mark.parametrize('case', SISO_CASES, ids=lambda c: c['name'])
def test_siso_chain_matches_oracle(case):
    from gpu_chain import run_chain
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    mode = case.get('mode', 'lte')
    n_per = num.Nc if mode == 'simple' else len(O.grid_indices(num.N, num.Nc)[0])
    S = -(-len(bits) // (n_per * num.bits_per_symbol))
    for snr in case['snrs']:
        if 'global_seed' in case:
            d = O.ReferenceDraws(0, global_seed=case['global_seed'])
            phases, z = np.zeros((1, 0, 16)), np.stack(d.unit_normals(S * num.L))[None]
        else:
            phases, z = reference_draws(case, S * num.L, 1)
        o = oracle_siso(case, bits, snr, phases=phases[0], z=(z[0][0], z[0][1]))
        r = run_chain(num, bits, snr, 1, case['ch'], case['prof'], case['v'], phases, z, mode=mode,
                      equalize=case.get('equalize', True), combine='zf', sc_fdm=case.get('sc_fdm', False))
        assert np.array_equal(r['qam'], o['symbols_tx'].reshape(-1).astype(np.complex64))   # bit-exact map
        assert rel_err(r['signal_tx'], o['signal_tx']) < TOL
        assert rel_err(r['signal_rx'][0], o['signal_rx']) < TOL
        assert rel_err(r['Y'][0], o['Y']) < TOL
        if case.get('sc_fdm'):
            # equaliser output before the IDFT against the oracle's, with the conditioning-aware bound; then the
            # Nd-point IDFT stage by itself against numpy on the very same input (core/lte_receiver.py:319-333)
            data_idx, _ = O.grid_indices(num.N, num.Nc)
            assert_zf_close(r['symbols_zf'], O.zf_equalize(o['Y'], o['H'])[:, data_idx].reshape(-1), o, num, case)
            zf = r['symbols_zf'].astype(np.complex128).reshape(S, -1)
            assert rel_err(r['symbols'], O.dft_precode(zf, inverse=True).reshape(-1)) < 2e-6
            assert np.median(np.abs(r['symbols'] - o['symbols_rx']) / np.maximum(np.abs(o['symbols_rx']), 1e-30)) < TOL
        else:
            assert_zf_close(r['symbols'], o['symbols_rx'], o, num, case)
        papr_db = 10 * np.log10(r['stats'][0, 0] / (r['stats'][0, 1] / (S * num.L)))
        assert abs(papr_db - o['papr_db']) < 1e-4
        # demap of the oracle's own symbols (cast to fp32) is bit-exact by construction
        assert r['errors'] == o['errors'] or explain_mismatch(o['symbols_rx'], num.modulation,
                                                              num.bits_per_symbol, r['bits_rx'], o['bits_rx'])
        # and against the reference's golden output
        assert o['errors'] == int(g[f'errors_{snr}'])
        gb = golden_bits_rx(g, snr)
        assert np.array_equal(r['bits_rx'], gb) or explain_mismatch(o['symbols_rx'], num.modulation,
                                                                   num.bits_per_symbol, r['bits_rx'], gb)


@pytest.mark.parametrize('case', SIMO_CASES, ids=lambda c: c['name'])
def test_simo_chain_matches_oracle(case):
    from gpu_chain import run_chain
    g = load_golden(case['name'])
    bits = golden_bits(g)
    num = numerology(case)
    R = case['R']
    S = -(-len(bits) // (len(O.grid_indices(num.N, num.Nc)[0]) * num.bits_per_symbol))
    phases, z = reference_draws(case, S * num.L, R)
    for snr in case['snrs']:
        o = oracle_simo(case, bits, snr, phases=phases, z=z)
        r = run_chain(num, bits, snr, R, case['ch'], case['prof'], case['v'], phases, z)
        assert rel_err(r['signal_tx'], o['signal_tx']) < TOL
        if r['signal_faded'] is not None:
            assert rel_err(r['signal_faded'], o['signal_faded']) < TOL
        assert rel_err(r['power'], np.sum(np.abs(o['signal_faded']) ** 2, axis=1)) < TOL
        assert rel_err(r['signal_rx'], o['signal_rx']) < TOL
        assert rel_err(r['Y'], o['Y']) < TOL
        assert rel_err(r['H'], o['H'][:, ::14, :]) < TOL
        assert rel_err(r['symbols'], o['symbols_combined']) < TOL
        assert r['errors'] == o['errors'] or explain_mismatch(o['symbols_combined'], num.modulation,
                                                              num.bits_per_symbol, r['bits_rx'], o['bits_rx'])
        gb = golden_bits_rx(g, snr)
        assert np.array_equal(r['bits_rx'], gb) or explain_mismatch(o['symbols_combined'], num.modulation,
                                                                   num.bits_per_symbol, r['bits_rx'], gb)
        if snr == case['full_snr']:
            big = case.get('big')
            rx = r['signal_rx'][:, ::BIG_RX_STRIDE] if big else r['signal_rx']
            assert rel_err(rx, g['signal_rx']) < TOL
            assert rel_err(r['symbols'], g['symbols_combined']) < TOL
            assert rel_err(r['H'], g['H']) < TOL


@pytest.mark.parametrize('mod', ['QPSK', '16-QAM', '64-QAM'])
def test_demap_is_bit_exact_on_reference_vectors(mod):
    """Slicer incl. exact ties, constellation points and the zero symbol (tables.npz)."""
    import torch
    from lte_b200 import LinkEngine
    g = load_golden('tables')
    bps = O.BITS_PER_SYMBOL[mod]
    eng = LinkEngine(128, 76, 9, bps, 1.92e6)
    y = g[f'demapsyms_{mod}'].astype(np.complex64)
    want = O.qam_demap(y.astype(np.complex128), mod)      # oracle on the fp32-representable inputs
    sym = torch.from_numpy(y[None, :]).cuda()
    _, idx_rx = eng.demap_count(sym, want_idx=True)
    got = eng.indices_to_bits(idx_rx, len(y) * bps).cpu().numpy().reshape(-1)
    assert np.array_equal(got, want)
    # the reference's own output differs from `want` only where fp32 rounding moved a tie
    ref = g[f'demapbits_{mod}']
    moved = y.astype(np.complex128) != g[f'demapsyms_{mod}']
    bad = np.flatnonzero(got != ref) // bps
    assert np.all(moved[bad])
    # map: bits -> constellation points, bit-exact in fp32
    bits = g[f'mapbits_{mod}']
    nsym = -(-len(bits) // bps)
    eng2 = LinkEngine(128, 76, 9, bps, 1.92e6, mode='simple')
    S = -(-nsym // 76)
    idx = eng2.bits_to_indices(torch.from_numpy(bits[None, :]).cuda(), len(bits), S)
    _, qam, _ = eng2.modulate(S, idx=idx, want_qam=True)
    assert np.array_equal(qam.cpu().numpy().reshape(-1)[:nsym], g[f'mapsyms_{mod}'].astype(np.complex64))


def test_fused_time_noise_equals_awgn_then_fft():
    """lte_rx_fft with fused Philox noise (time domain) is bit-identical to lte_awgn_add + lte_rx_fft."""
    import torch
    from lte_b200 import LinkEngine
    from lte_b200 import _native as nat
    eng = LinkEngine(512, 300, 36, 4, 7.68e6)
    rows, S = 6, 3
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(rows, S * eng.L, 2, device='cuda', generator=g)
    x = torch.view_as_complex(x).contiguous()
    power = (x.abs() ** 2).sum(dim=1).double()
    snr = torch.tensor([1.0, 3.0, 10.0, 30.0, 100.0, 1000.0], dtype=torch.float32, device='cuda')
    y = eng.awgn(x, 1, power, snr, rows, seed=99, row_id0=40)
    Y1 = eng.rx_fft(y, rows, S, nat.WINDOW_USEFUL)
    Y2 = eng.rx_fft(x, rows, S, nat.WINDOW_USEFUL, power=power, snr_lin=snr, seed=99, row_id0=40, noise_domain=0)
    assert torch.equal(Y1, Y2)
    # a different seed or row id gives different noise
    Y3 = eng.rx_fft(x, rows, S, nat.WINDOW_USEFUL, power=power, snr_lin=snr, seed=100, row_id0=40, noise_domain=0)
    assert not torch.equal(Y1, Y3)


def test_frequency_domain_noise_statistics():
    """noise_domain=1 draws CN(0, 2 sigma^2) directly on the kept bins: same variance as the
    time-domain AWGN pushed through the unitary FFT, white, Gaussian tails."""
    import torch
    from lte_b200 import LinkEngine
    from lte_b200 import _native as nat
    eng = LinkEngine(2048, 1200, 144, 6, 30.72e6)
    rows, S = 64, 14
    x = torch.zeros(rows, S * eng.L, dtype=torch.complex64, device='cuda')
    n = S * eng.L
    power = torch.full((rows,), float(n), dtype=torch.float64, device='cuda')       # mean power 1
    snr = torch.full((rows,), 0.5, dtype=torch.float32, device='cuda')               # sigma^2 = 1 per component
    for dom in (0, 1):
        Y = eng.rx_fft(x, rows, S, nat.WINDOW_USEFUL, power=power, snr_lin=snr, seed=5, row_id0=0, noise_domain=dom)
        v = torch.view_as_real(Y).double().reshape(-1, 2)
        m = v.shape[0]
        assert abs(v.mean().item()) < 5 / m ** 0.5
        assert abs(v.var(dim=0).mean().item() - 1.0) < 5 * (2 / m) ** 0.5
        assert abs((v[:, 0] * v[:, 1]).mean().item()) < 5 / m ** 0.5
        kurt = (v ** 4).mean().item()
        assert abs(kurt - 3.0) < 0.05
        assert abs((v[:-1, 0] * v[1:, 0]).mean().item()) < 5 / m ** 0.5            # adjacent bins uncorrelated
        assert v.abs().max().item() > 4.5                                            # tails present (1e6+ draws)


@pytest.mark.parametrize('mod', ['QPSK', '16-QAM', '64-QAM'])
@pytest.mark.parametrize('nbits_off', [0, 1, 5, 13, 47])
def test_packed_bits_to_indices_matches_the_byte_per_bit_path(mod, nbits_off):
    """np.packbits rows (the e2e path of bench.py): the four-symbols-per-thread kernel and the scalar fallback give
    the indices of QAMModulator.bits_to_symbols, including rows that end inside a symbol / a group of four."""
    import torch
    from config import LTEConfig
    from lte_b200 import LinkEngine
    eng = LinkEngine.from_config(LTEConfig(1.25, 15.0, mod))
    b = eng.bps
    for S in (4, 3):                                   # S * Nd = 248 (x4 kernel) and 186 (not a multiple of 4: fallback)
        nsym = S * eng.Nd
        nbits = nsym * b - nbits_off
        rs = np.random.RandomState(nbits)
        bits = rs.randint(0, 2, (3, nbits)).astype(np.uint8)
        packed = np.stack([np.packbits(r) for r in bits])
        want = np.stack([O.bits_to_indices(np.concatenate([r, np.zeros(nsym * b - nbits, dtype=np.uint8)]), b) for r in bits])
        got_p = eng.bits_to_indices(torch.from_numpy(packed).cuda(), nbits, S, packed=True).cpu().numpy()
        got_u = eng.bits_to_indices(torch.from_numpy(bits).cuda(), nbits, S).cpu().numpy()
        assert np.array_equal(got_u, want) and np.array_equal(got_p, want)


@pytest.mark.parametrize('bw,mod,S,B', [(20.0, '64-QAM', 14, 5), (10.0, '16-QAM', 3, 4), (5.0, 'QPSK', 15, 3), (2.5, '64-QAM', 1, 7)])
def test_packed_bits_sixteen_symbols_per_thread(bw, mod, S, B):
    """The e2e path's unpack kernel (16 symbols per thread, one 128-bit store) on the headline row length (13 986
    symbols: rows neither start nor end on a 16-symbol group) and on ragged bit budgets."""
    import torch
    from config import LTEConfig
    from lte_b200 import LinkEngine
    eng = LinkEngine.from_config(LTEConfig(bw, 15.0, mod))
    b = eng.bps
    nsym = S * eng.Nd
    for off in (0, 3, 2 * b + 1):
        nbits = nsym * b - off
        rs = np.random.RandomState(nsym + off)
        bits = rs.randint(0, 2, (B, nbits)).astype(np.uint8)
        packed = np.stack([np.packbits(r) for r in bits])
        pad = np.zeros((B, nsym * b - nbits), dtype=np.uint8)
        full = np.concatenate([bits, pad], axis=1).reshape(B, nsym, b)
        want = (full * (1 << np.arange(b - 1, -1, -1))).sum(axis=2).astype(np.uint8)
        got = eng.bits_to_indices(torch.from_numpy(packed).cuda(), nbits, S, packed=True).cpu().numpy()
        assert np.array_equal(got, want), (bw, mod, off)
