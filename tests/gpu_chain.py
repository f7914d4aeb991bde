"""Drives the CUDA stages (through lte_b200.LinkEngine -> C ABI) on the same inputs
as the oracle, stage by stage, for the parity tests."""
import numpy as np
import torch

from lte_b200 import LinkEngine, chan_for
from lte_b200 import _native as nat
from oracle import lte_oracle as O


def to_dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to('cuda', dtype=dtype)


def draws_to_device(phases, z):
    """phases [links][taps][16] radians -> u in [0,1) float32; z [links][2][n] -> complex64 [links][n]."""
    u = to_dev(phases / (2 * np.pi), torch.float32) if phases.size else None
    zc = to_dev(z[:, 0, :] + 1j * z[:, 1, :], torch.complex64)
    return u, zc


def run_chain(num, bits, snr_db, R, channel_type, itu_profile, velocity_kmh, phases, z, mode='lte',
              equalize=True, combine='mrc', window=nat.WINDOW_FULL, sc_fdm=False):
    """Returns a dict of numpy arrays mirroring oracle.simulate_siso / simulate_simo."""
    eng = LinkEngine(num.N, num.Nc, num.cp_length, num.bits_per_symbol, num.fs, mode=mode)
    nbits = len(bits)
    S = eng.symbols_for_bits(nbits)
    bits_d = to_dev(np.asarray(bits, dtype=np.uint8)[None, :], torch.uint8)
    idx = eng.bits_to_indices(bits_d, nbits, S)
    if sc_fdm:      # SC-FDM: Nd-point unitary DFT of every OFDM symbol's data before the grid (core/dft_precoding.py:67-93)
        qam = eng.qam_map(idx)
        pre = eng.dft_m(qam.view(S, eng.Nd), eng.Nd).view(1, -1)
        tx, _, stats = eng.modulate(S, symbols=pre)
    else:
        tx, qam, stats = eng.modulate(S, idx=idx, want_qam=True)
    chan = chan_for(channel_type, num.fs, itu_profile, 2.0, velocity_kmh)
    u, zc = draws_to_device(phases, z)
    faded, power = eng.channel(tx, chan, 1, R, phases=u)
    snr_lin = torch.full((R,), 10 ** (snr_db / 10), dtype=torch.float32, device='cuda')
    if faded is None:
        rx = eng.awgn(tx, R, power.view(-1), snr_lin, R, z=zc)
    else:
        rx = eng.awgn(faded.view(R, -1), 1, power.view(-1), snr_lin, R, z=zc)
    Y = eng.rx_fft(rx, R, S, window)
    out = dict(idx=idx, S=S, eng=eng)
    if mode == 'simple':
        sym = Y[:, :, :num.Nc].reshape(1, -1) if window == nat.WINDOW_FULL else Y.reshape(1, -1)
        H = None
    else:
        H = eng.estimate(Y, R, S, window)
        if combine == 'mrc':
            sym = eng.mrc(Y, H, 1, R, S, window)
        else:
            sym = eng.zf(Y, H if equalize else None, 1, S, window)
            out['symbols_zf'] = sym.cpu().numpy().reshape(-1)
            if sc_fdm:  # ... and its inverse after the equaliser (core/lte_receiver.py:319-333)
                sym = eng.dft_m(sym.view(S, eng.Nd), eng.Nd, inverse=True).view(1, -1)
    errors, idx_rx = eng.demap_count(sym, idx_tx=idx, nbits=nbits, want_idx=True)
    bits_rx = eng.indices_to_bits(idx_rx, nbits)
    torch.cuda.synchronize()
    n = S * num.L
    out.update(signal_tx=tx.cpu().numpy().reshape(-1), qam=qam.cpu().numpy().reshape(-1),
               stats=stats.cpu().numpy(), power=power.cpu().numpy().reshape(-1),
               signal_faded=None if faded is None else faded.cpu().numpy().reshape(R, n),
               signal_rx=rx.cpu().numpy().reshape(R, n), Y=Y.cpu().numpy(),
               H=None if H is None else H.cpu().numpy(), symbols=sym.cpu().numpy().reshape(-1),
               errors=int(errors.item()), bits_rx=bits_rx.cpu().numpy().reshape(-1).astype(np.int64))
    return out


def boundary_distance(symbols, modulation):
    """Distance of every symbol axis value to the nearest slicer boundary (oracle units)."""
    lv = np.unique(O.constellation(modulation).real)
    mids = (lv[:-1] + lv[1:]) / 2
    y = np.concatenate([symbols.real, symbols.imag])
    return np.min(np.abs(y[:, None] - mids[None, :]), axis=1).reshape(2, -1).min(axis=0)
