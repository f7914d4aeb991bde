"""Out-of-bounds guard for the kernels added late in round 1 (compute-sanitizer is closed on this pool): every
buffer the engine allocates gets sentinel-filled guard zones on both sides; after the beamforming, coded-chain and
sweep passes the guards must be untouched."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PAD = 512


class Guard:
    def __init__(self):
        self.blocks = []

    def empty(self, eng):
        def _empty(shape, dtype):
            n = int(np.prod(shape)) if len(shape) else 1
            flat = torch.empty(n + 2 * PAD, dtype=dtype, device=eng.device)
            flat.view(torch.uint8).fill_(0xA5)
            self.blocks.append(flat)
            return flat[PAD:PAD + n].view(shape)
        return _empty

    def check(self):
        assert self.blocks
        for flat in self.blocks:
            raw = flat.view(torch.uint8)
            item = flat.element_size()
            assert bool((raw[:PAD * item] == 0xA5).all()) and bool((raw[-PAD * item:] == 0xA5).all()), \
                f'guard zone of a {flat.dtype} buffer of {flat.numel() - 2 * PAD} elements was overwritten'


def _engine(bw, mod, **kw):
    from config import LTEConfig
    from lte_b200 import LinkEngine
    eng = LinkEngine.from_config(LTEConfig(bw, 15.0, mod), **kw)
    g = Guard()
    eng._empty = g.empty(eng)
    return eng, g


@pytest.mark.parametrize('A,B', [(1, 3), (488, 5), (6121, 3), (19987, 2)])
def test_coded_chain_stays_inside_its_buffers(A, B):
    from lte_b200 import chan_for
    eng, g = _engine(1.25, '16-QAM')
    bits = torch.from_numpy(np.random.RandomState(A).randint(0, 2, (B, A)).astype(np.uint8)).cuda()
    chan = chan_for('rayleigh_mp', eng.fs, 'Vehicular_A', 2.0, 30.0)
    snr = torch.full((B,), 30.0, dtype=torch.float32, device='cuda')
    err, crc = eng.siso_coded_ber(bits, chan, snr, seed=1, stream_id0=3)
    torch.cuda.synchronize()
    g.check()
    plan = eng.coding_plan(A)
    coded = eng.tb_encode(bits, plan)
    got, ok, e = eng.tb_decode((3.0 * (1.0 - 2.0 * coded.float())).contiguous(), plan, bits_tx=bits)
    torch.cuda.synchronize()
    g.check()
    assert torch.equal(got, bits) and bool(ok.all()) and int(e.sum()) == 0


@pytest.mark.parametrize('T,R', [(2, 1), (4, 2), (8, 8)])
def test_beamforming_stays_inside_its_buffers(T, R):
    from oracle import lte_oracle as O
    eng, g = _engine(2.5, '64-QAM')
    B, S = 7, 3
    idx = eng.random_indices(B, S, 1, 0)
    h = eng.random_channel(B, R, T, 1, 0)
    for mode in ('MRT', 'CODEBOOK'):
        W, heff, pmi, gain = eng.bf_weights(h, O.codebook(T, 1), mode=mode)
        nstd = torch.full((B,), 0.05, dtype=torch.float32, device='cuda')
        err, sym = eng.bf_link(idx, h, W, heff, nstd, S, seed=2, want_symbols=True)
        z = torch.randn((B, S, 2, R, eng.Nd), dtype=torch.float32, device='cuda')
        err2, _ = eng.bf_link(idx, h, W, heff, nstd, S, z=z)
    torch.cuda.synchronize()
    g.check()


def test_sweep_passes_stay_inside_their_buffers():
    from lte_b200 import chan_for, tables
    eng, g = _engine(1.25, '16-QAM')
    chan = chan_for('rayleigh_mp', eng.fs, 'Pedestrian_B', 2.0, 3.0)
    rows = torch.full((6,), 10.0, dtype=torch.float32, device='cuda')
    eng.siso_ber(chan, rows, 15, seed=1)
    ws = eng.workspace(3, 15, 2, fading=True, fused=True)
    eng.simo_ber(ws, chan, rows, 1, fused=True, noise_domain=3)
    e2, g2 = _engine(1.25, '16-QAM', pilot_sets=tables.mimo_pilot_sets(2, eng.Np))
    e2.sfbc_ber(chan, rows, 15, 2, seed=1)
    e4, g4 = _engine(1.25, 'QPSK', pilot_sets=tables.mimo_pilot_sets(4, eng.Np))
    from core.codebook_lte import LTECodebook
    ch3 = chan_for('rayleigh_mp', eng.fs, 'Pedestrian_A', 2.0, 3.0, gain_conversions=3)
    e4.sm_ber(ch3, LTECodebook(4, 'TM4', 3).get_precoder(1), 12.0, 3, 2, 4, 'SIC', seed=1)
    torch.cuda.synchronize()
    for gg in (g, g2, g4):
        gg.check()


@pytest.mark.parametrize('bw,mod,prof,B,S,R', [(2.5, '16-QAM', 'Pedestrian_A', 3, 15, 3), (5.0, '64-QAM', 'Pedestrian_B', 5, 14, 4),
                                              (20.0, '64-QAM', 'Pedestrian_A', 2, 29, 1), (10.0, 'QPSK', 'Pedestrian_A', 4, 1, 2)])
def test_spectral_link_stays_inside_its_buffers(bw, mod, prof, B, S, R):
    """The spectral kernels (bulk-copy ring, 128-bit pair stores into the padded compact rows, per-warp power atomics)
    on odd stream / symbol / antenna counts: every buffer of the workspace -- G, tails, compact grid, pilot side
    buffer, pilot estimates, errors -- keeps its guard zones, and so does the caller-owned coefficient workspace."""
    from lte_b200 import chan_for
    eng, g = _engine(bw, mod)
    scratch = {}

    def guarded_scratch(nbytes):                       # same contract as LinkEngine._scratch, with guard zones
        n = max(int(nbytes), 256)
        if 'buf' not in scratch or scratch['n'] < n:
            flat = torch.empty(n + 2 * PAD, dtype=torch.uint8, device=eng.device)
            flat.fill_(0xA5)
            g.blocks.append(flat)
            scratch.update(buf=flat[PAD:PAD + n], n=n)
        return scratch['buf']
    eng._scratch = guarded_scratch
    chan = chan_for('rayleigh_mp', eng.fs, prof, 2.0, 3.0)
    rows = torch.full((B * R,), 31.6, dtype=torch.float32, device='cuda')
    for nd in (2, 3):
        ws = eng.workspace(B, S, R, fading=True, fused=True, lazy=True)
        err = eng.simo_ber(ws, chan, rows, 1, stream_id0=7, fused=True, noise_domain=nd)
        assert ws.get('spectral') is True and err.shape == (B,)
    # the windowed (non-compact) output of the same kernel
    idx = eng.random_indices(B, S, 1, 0)
    G, tail = eng.tx_spectral(S, idx)
    ph = eng.random_phases(B, R * chan.num_taps * 16, 1, 0)
    Y, pw = eng.channel_spectral(idx, G, tail, chan, B, R, S, ph)
    torch.cuda.synchronize()
    g.check()
    assert bool(torch.isfinite(torch.view_as_real(Y)).all()) and bool((pw > 0).all())
