"""Batched link-chain engine: torch tensors in HBM, stages launched through the C ABI.

Shapes (B streams, S OFDM symbols per stream, L = N + cp, R/T antennas):
    idx      uint8      [B, S*Nd]      natural-binary constellation indices
    tx       complex64  [B*T, S*L]
    faded    complex64  [B, R, S*L]
    Y        complex64  [rows, S, nk]  window of the FFT grid (full or occupied bins)
    H        complex64  [rows, ceil(S/14), nk]
    errors   int64      [B]
"""
import ctypes as C
import functools

import numpy as np
import torch

from . import _native as nat
from . import tables


def chan_for(channel_type, fs, itu_profile='Pedestrian_A', frequency_ghz=2.0, velocity_kmh=0.0,
             gain_conversions=2, faithful_gains=True):
    return tables.channel_desc(channel_type, fs, itu_profile, frequency_ghz, velocity_kmh,
                               gain_conversions, faithful_gains)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class LinkEngine:
    """One numerology + modulation + pilot layout on one GPU."""

    def __init__(self, N, Nc, cp, bits_per_symbol, fs, mode='lte', pilot_sets=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("lte_b200.LinkEngine needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.N, self.Nc, self.cp, self.L = int(N), int(Nc), int(cp), int(N) + int(cp)
        self.bps, self.fs = int(bits_per_symbol), float(fs)
        self.simple = mode == 'simple'
        if self.simple:
            self.data_idx, self.pilot_idx = np.arange(self.Nc), np.zeros(0, dtype=np.int64)
        else:
            self.data_idx, self.pilot_idx = tables.grid_indices(self.N, self.Nc)
        self.Nd, self.Np = len(self.data_idx), len(self.pilot_idx)
        if pilot_sets is None:
            pilot_sets = tables.pilot_values(0, self.Np)[None, :]
        pilot_sets = np.ascontiguousarray(np.asarray(pilot_sets, dtype=np.complex64))
        self.num_pilot_sets = pilot_sets.shape[0]
        desc = nat.PlanDesc(self.N, self.Nc, self.cp, self.bps, 1 if self.simple else 0,
                            self.num_pilot_sets, self.fs)
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(nat.lib.lte_plan_create(C.byref(desc), pilot_sets.ctypes.data_as(C.c_void_p),
                                              C.byref(self._plan)), 'lte_plan_create')
        assert nat.lib.lte_plan_num_data(self._plan) == self.Nd
        assert nat.lib.lte_plan_num_pilots(self._plan) == self.Np
        ndp, npp = C.c_int32(), C.c_int32()
        nat.check(nat.lib.lte_plan_compact_shape(self._plan, C.byref(ndp), C.byref(npp)), 'lte_plan_compact_shape')
        self.ndp, self.npp = ndp.value, npp.value     # data / pilot bin pairs of the compact sweep layout
        self.gl = (self.N - self.Nc) // 2
        self.launches = 0          # native kernels launched through this engine (bench accounting)

    @classmethod
    def from_config(cls, config, mode='lte', pilot_sets=None, device=None):
        return cls(config.N, config.Nc, config.cp_length, config.bits_per_symbol, config.fs,
                   mode=mode, pilot_sets=pilot_sets, device=device)

    def __del__(self):
        try:
            if getattr(self, '_plan', None):
                nat.lib.lte_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _scratch(self, nbytes):
        """Caller-owned workspace for the C ABI (the library never allocates): one cached buffer per CUDA
        stream of this engine, grown on demand, 256-byte aligned by the torch allocator."""
        key = torch.cuda.current_stream(self.device).cuda_stream
        pool = self.__dict__.setdefault('_scratch_pool', {})
        buf = pool.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = pool[key] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        return buf

    def window(self, window):
        return (0, self.N) if window == nat.WINDOW_FULL else ((0, self.Nc) if self.simple else (self.gl, self.Nc))

    def symbols_for_bits(self, nbits):
        """OFDM symbols needed for nbits (reference core/modulator.py:264-275)."""
        per = self.Nd * self.bps
        return int(-(-int(nbits) // per))

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ bits
    def bits_to_indices(self, bits, nbits, S, packed=False):
        """bits: uint8 [B, nbits] of 0/1 (or np.packbits rows if packed) -> idx [B, S*Nd]."""
        B = bits.shape[0]
        idx = self._empty((B, S * self.Nd), torch.uint8)
        nat.check(nat.lib.lte_bits_to_indices(self._plan, _ptr(bits), -int(nbits) if packed else int(nbits),
                                              _ptr(idx), S * self.Nd, B, self._stream()), 'lte_bits_to_indices')
        self.launches += 1
        return idx

    def indices_to_bits(self, idx, nbits):
        B, nsym = idx.shape
        bits = self._empty((B, int(nbits)), torch.uint8)
        nat.check(nat.lib.lte_indices_to_bits(self._plan, _ptr(idx), nsym, _ptr(bits), int(nbits), B,
                                              self._stream()), 'lte_indices_to_bits')
        self.launches += 1
        return bits

    def random_indices(self, B, S, seed, stream_id0=0, out=None, nsym=None):
        """Philox symbol indices [B, nsym] (default nsym = S * Nd), keyed (seed, stream id, position): a shorter
        row is a prefix of a longer one."""
        nsym = S * self.Nd if nsym is None else int(nsym)
        idx = out if out is not None else self._empty((B, nsym), torch.uint8)
        nat.check(nat.lib.lte_random_indices(self._plan, _ptr(idx), nsym, B, int(seed), int(stream_id0),
                                             self._stream()), 'lte_random_indices')
        self.launches += 1
        return idx

    def random_phases(self, B, per_stream, seed, stream_id0=0, out=None):
        ph = out if out is not None else self._empty((B, per_stream), torch.float32)
        nat.check(nat.lib.lte_random_phases(_ptr(ph), per_stream, B, int(seed), int(stream_id0), self._stream()),
                  'lte_random_phases')
        self.launches += 1
        return ph

    # ------------------------------------------------------------------ stage 1+2 TX
    def modulate(self, S, idx=None, symbols=None, T=1, want_qam=False, want_stats=True, out=None, stats=None):
        """-> (tx [B*T, S*L], qam [B, S*Nd] or None, stats [B*T, 2] float64 or None)."""
        if idx is not None:
            B = idx.shape[0]
        else:
            B = symbols.shape[0] // T
        tx = out if out is not None else self._empty((B * T, S * self.L), torch.complex64)
        qam = self._empty((B, S * self.Nd), torch.complex64) if (want_qam and idx is not None) else None
        if want_stats:
            if stats is None:
                stats = torch.zeros((B * T, 2), dtype=torch.float64, device=self.device)
            else:
                stats.zero_()
        else:
            stats = None
        nat.check(nat.lib.lte_tx_map_ifft(self._plan, _ptr(idx) if symbols is None else None, _ptr(symbols), T,
                                          _ptr(tx), _ptr(qam), _ptr(stats), B, S, self._stream()),
                  'lte_tx_map_ifft')
        self.launches += 1
        return tx, qam, stats

    def qam_map(self, idx):
        """Symbol indices (uint8, any shape) -> constellation points (complex64, same shape)."""
        out = self._empty(tuple(idx.shape), torch.complex64)
        nat.check(nat.lib.lte_qam_map(self._plan, _ptr(idx), _ptr(out), idx.numel(), self._stream()), 'lte_qam_map')
        self.launches += 1
        return out

    # ------------------------------------------------------------------ PAPR / CCDF
    def modulate_papr(self, S, idx=None, symbols=None, T=1, write_tx=False, hist=None, hist_lo=0.0,
                      hist_step=0.1, want_db=True, want_peak_mean=False):
        """TX stage with the per-symbol PAPR epilogue (useful part, no CP).
        -> (papr_db [B*T, S] float32 or None, peak_mean [B*T, S, 2] or None, tx [B*T, S*L] or None).
        `hist` (int64 [bins], zeroed by the caller) accumulates the papr_db histogram in place."""
        B = idx.shape[0] if idx is not None else symbols.shape[0] // T
        tx = self._empty((B * T, S * self.L), torch.complex64) if write_tx else None
        db = self._empty((B * T, S), torch.float32) if want_db else None
        pm = self._empty((B * T, S, 2), torch.float32) if want_peak_mean else None
        nat.check(nat.lib.lte_tx_papr(self._plan, _ptr(idx) if symbols is None else None, _ptr(symbols), T,
                                      _ptr(tx), None, _ptr(db), _ptr(pm), _ptr(hist), float(hist_lo),
                                      float(hist_step), 0 if hist is None else hist.numel(), B, S,
                                      self._stream()), 'lte_tx_papr')
        self.launches += 1
        return db, pm, tx

    def papr_symbols(self, x, include_cp=False, hist=None, hist_lo=0.0, hist_step=0.1, want_peak_mean=False):
        """Per-OFDM-symbol PAPR of time-domain streams x [rows, S*L] -> (papr_db [rows, S], peak_mean or None)."""
        x = x.reshape(-1, x.shape[-1])
        rows, S = x.shape[0], x.shape[1] // self.L
        db = self._empty((rows, S), torch.float32)
        pm = self._empty((rows, S, 2), torch.float32) if want_peak_mean else None
        if S == 0:
            return db, pm
        if x.shape[1] != S * self.L:
            x = x[:, :S * self.L].contiguous()
        nat.check(nat.lib.lte_papr_symbols(self._plan, _ptr(x), 1 if include_cp else 0, _ptr(db), _ptr(pm),
                                           _ptr(hist), float(hist_lo), float(hist_step),
                                           0 if hist is None else hist.numel(), rows, S, self._stream()),
                  'lte_papr_symbols')
        self.launches += 1
        return db, pm

    def histogram(self, x, lo, step, bins, hist=None):
        """Counts of float32 x in `bins` bins of width `step` from `lo` (outliers in the edge bins)."""
        x = x.reshape(-1).to(torch.float32).contiguous()
        if hist is None:
            hist = torch.zeros(bins, dtype=torch.int64, device=self.device)
        nat.check(nat.lib.lte_histogram(_ptr(x), x.numel(), float(lo), float(step), int(bins), _ptr(hist),
                                        self._stream()), 'lte_histogram')
        self.launches += 1
        return hist

    # ------------------------------------------------------------------ SC-FDM M-point DFT
    def dft_m(self, x, M, inverse=False, out=None):
        """Unitary M-point DFT (or IDFT) of every length-M row of x (any leading shape)."""
        rows = x.numel() // M
        self._need_dft(M)
        y = out if out is not None else torch.empty_like(x)
        nat.check(nat.lib.lte_dft_m(self._plan, _ptr(x), _ptr(y), int(M), 1 if inverse else 0, rows,
                                    self._stream()), 'lte_dft_m')
        self.launches += 1
        return y

    def _need_dft(self, M):
        have = self.__dict__.setdefault('_dft_sizes', set())
        if int(M) not in have:                 # chirp tables are plan state, built once per M (synchronous)
            nat.check(nat.lib.lte_plan_add_dft(self._plan, int(M)), 'lte_plan_add_dft')
            have.add(int(M))

    def dft_qam(self, idx, M):
        """QAM map + unitary M-point DFT of every length-M row of symbol indices (SC-FDM precoder input), one kernel."""
        self._need_dft(M)
        rows = idx.numel() // M
        out = self._empty(tuple(idx.shape), torch.complex64)
        nat.check(nat.lib.lte_dft_qam(self._plan, _ptr(idx), _ptr(out), int(M), rows, self._stream()), 'lte_dft_qam')
        self.launches += 1
        return out

    # ------------------------------------------------------------------ stage 3 channel
    def channel(self, tx, chan, B, R, T=1, phases=None, out=None, power=None):
        """-> (faded [B, R, n] or None for the AWGN channel type, power [B, R] float64)."""
        n = tx.shape[-1]
        if power is None:
            power = torch.zeros((B, R), dtype=torch.float64, device=self.device)
        else:
            power.zero_()
        faded = None
        if chan.num_taps > 0:
            faded = out if out is not None else self._empty((B, R, n), torch.complex64)
            if phases is None:
                raise ValueError("phases are required for a fading channel")
        need = nat.lib.lte_channel_tdl_workspace_bytes(self._plan, C.byref(chan), B, R, T, n)
        if need < 0:
            nat.check(int(need), 'lte_channel_tdl_workspace_bytes')
        work = self._scratch(need) if need > 0 else None
        nat.check(nat.lib.lte_channel_tdl(self._plan, C.byref(chan), _ptr(tx), _ptr(phases), _ptr(faded),
                                          _ptr(power), _ptr(work), B, R, T, n, self._stream()), 'lte_channel_tdl')
        self.launches += 2 if chan.num_taps > 0 else 1     # Jakes coefficient kernel + TDL kernel
        return faded, power

    def channel_rx_fft(self, tx, chan, B, R, S, phases, window=nat.WINDOW_FULL, out=None, power=None, T=1):
        """Fused fading channel + CP strip + FFT: -> (Y [B*R, S, nk] noise-free, power [B, R]),
        or None when the configuration needs the staged `channel` + `rx_fft` pair.
        T > 1: tx [B, T, S*L], phases [B, R*T*taps*16]; the R x T links are summed per receive antenna."""
        k0, nk = self.window(window)
        if T > 1:
            need = nat.lib.lte_channel_rx_fft_mimo_workspace_bytes(self._plan, C.byref(chan), B, R, T, S)
        else:
            need = nat.lib.lte_channel_rx_fft_workspace_bytes(self._plan, C.byref(chan), B, R, S)
        if need == nat.LTE_ERR_UNSUPPORTED:
            return None
        if need < 0:
            nat.check(int(need), 'lte_channel_rx_fft_workspace_bytes')
        work = self._scratch(need)
        if power is None:
            power = torch.zeros((B, R), dtype=torch.float64, device=self.device)
        else:
            power.zero_()
        Y = out if out is not None else self._empty((B * R, S, nk), torch.complex64)
        if T > 1:
            rc = nat.lib.lte_channel_rx_fft_mimo(self._plan, C.byref(chan), _ptr(tx), _ptr(phases), _ptr(Y), _ptr(power),
                                                 _ptr(work), window, B, R, T, S, self._stream())
            if rc == nat.LTE_ERR_UNSUPPORTED:                # shared-memory footprint: known only at launch
                return None
        else:
            rc = nat.lib.lte_channel_rx_fft(self._plan, C.byref(chan), _ptr(tx), _ptr(phases), _ptr(Y), _ptr(power),
                                            _ptr(work), window, B, R, S, self._stream())
        nat.check(rc, 'lte_channel_rx_fft')
        self.launches += 2                                   # Jakes coefficient kernel + fused kernel
        return Y, power

    # ------------------------------------------------------------------ spectral fading link
    def tx_spectral(self, S, idx, out_G=None, out_tail=None):
        """TX side of the spectral link (csrc/spectral.cu): idx [B, S*Nd] ->
        (G [B*S, Nc] ramp-weighted spectrum on the occupied window, tail [B*S, cp] symbol tails)."""
        B = idx.shape[0]
        k0, nk = self.window(nat.WINDOW_USEFUL)
        G = out_G if out_G is not None else self._empty((B * S, nk), torch.complex64)
        tail = out_tail if out_tail is not None else self._empty((B * S, self.cp), torch.complex64)
        nat.check(nat.lib.lte_tx_spectral(self._plan, _ptr(idx), _ptr(G), _ptr(tail), B, S, self._stream()),
                  'lte_tx_spectral')
        self.launches += 1
        return G, tail

    def spectral_workspace_bytes(self, chan, B, R, S):
        """Bytes of caller-owned scratch lte_channel_spectral needs, or None when the configuration is
        outside the spectral link's validity range (Doppler, delay spread, identity link)."""
        n = nat.lib.lte_channel_spectral_workspace_bytes(self._plan, C.byref(chan), B, R, S)
        if n == nat.LTE_ERR_UNSUPPORTED:
            return None
        if n < 0:
            nat.check(int(n), 'lte_channel_spectral_workspace_bytes')
        return int(n)

    def channel_spectral(self, idx, G, tail, chan, B, R, S, phases, out=None, power=None, workspace=None,
                         compact=False, out_pilots=None):
        """Channel side of the spectral link: -> (Y [B*R, S, Nc] noise-free on the occupied window,
        power [B, R]), or None when unsupported (use modulate + channel_rx_fft).
        compact: -> ((Yd [B*R, S, 2 ndp], Yp [B*R, slots, 2 npp]), power): data bins in data-symbol order and
        slot-head pilot bins, rows padded to an even length (see include/lte_b200.h)."""
        need = self.spectral_workspace_bytes(chan, B, R, S)
        if need is None:
            return None
        if workspace is None or workspace.numel() * workspace.element_size() < need:
            workspace = self._scratch(need)
        k0, nk = self.window(nat.WINDOW_USEFUL)
        if power is None:
            power = torch.zeros((B, R), dtype=torch.float64, device=self.device)
        else:
            power.zero_()
        Yp = None
        if compact:
            nslot = -(-S // nat.LTE_SLOT_SYMBOLS)
            Y = out if out is not None else self._empty((B * R, S, 2 * self.ndp), torch.complex64)
            Yp = out_pilots if out_pilots is not None else self._empty((B * R, nslot, 2 * self.npp), torch.complex64)
        else:
            Y = out if out is not None else self._empty((B * R, S, nk), torch.complex64)
        nat.check(nat.lib.lte_channel_spectral(self._plan, C.byref(chan), _ptr(idx), _ptr(G), _ptr(tail),
                                               _ptr(phases), _ptr(Y), _ptr(Yp), _ptr(power), _ptr(workspace), B, R, S,
                                               self._stream()), 'lte_channel_spectral')
        self.launches += 2                                   # Jakes coefficient kernel + channel kernel
        return ((Y, Yp) if compact else Y), power

    def awgn(self, x, x_div, power, snr_lin, rows, z=None, seed=0, row_id0=0, out=None):
        n = x.shape[-1]
        y = out if out is not None else self._empty((rows, n), torch.complex64)
        nat.check(nat.lib.lte_awgn_add(self._plan, _ptr(x), x_div, _ptr(power), _ptr(snr_lin), _ptr(z), int(seed),
                                       int(row_id0), _ptr(y), rows, n, self._stream()), 'lte_awgn_add')
        self.launches += 1
        return y

    # ------------------------------------------------------------------ stage 2 RX
    def rx_fft(self, rx, rows, S, window=nat.WINDOW_FULL, rx_div=1, power=None, snr_lin=None, z=None, seed=0,
               row_id0=0, out=None, noise_domain=0):
        k0, nk = self.window(window)
        Y = out if out is not None else self._empty((rows, S, nk), torch.complex64)
        nat.check(nat.lib.lte_rx_fft(self._plan, _ptr(rx), rx_div, _ptr(power), _ptr(snr_lin), _ptr(z),
                                     int(noise_domain), int(seed), int(row_id0), _ptr(Y), window, rows, S,
                                     self._stream()), 'lte_rx_fft')
        self.launches += 1
        return Y

    # ------------------------------------------------------------------ stage 4
    @staticmethod
    def awgn_desc(power, snr_lin, seed, row_id0=0, combine=False):
        """Lazy frequency-domain AWGN: the consumers of a noise-free Y add the noise
        lte_rx_fft(noise_domain=1) would have added (see include/lte_b200.h).  combine: the MRC
        kernel draws one equivalent sample per combiner output instead of one per antenna."""
        return nat.AwgnDesc(power.data_ptr(), snr_lin.data_ptr(), int(seed), int(row_id0), 1 if combine else 0)

    def estimate(self, Y, rows, S, window=nat.WINDOW_FULL, pilot_set=0, out=None, awgn=None):
        k0, nk = self.window(window)
        nslot = -(-S // nat.LTE_SLOT_SYMBOLS)
        H = out if out is not None else self._empty((rows, nslot, nk), torch.complex64)
        if awgn is None:
            nat.check(nat.lib.lte_crs_ls_interp(self._plan, _ptr(Y), _ptr(H), window, pilot_set, rows, S,
                                                self._stream()), 'lte_crs_ls_interp')
        else:
            nat.check(nat.lib.lte_crs_ls_interp_awgn(self._plan, _ptr(Y), _ptr(H), window, pilot_set, rows, S,
                                                     C.byref(awgn), self._stream()), 'lte_crs_ls_interp_awgn')
        self.launches += 1
        return H

    def estimate_compact(self, Yp, rows, S, out=None, awgn=None):
        """LS estimate at the pilot positions of the compact layout: Yp [rows, slots, 2 npp] ->
        Hp [rows, slots, Np] (awgn: lazy frequency-domain AWGN, the draws of `estimate(..., awgn=...)`)."""
        Hp = out if out is not None else self._empty((rows, Yp.shape[1], self.Np), torch.complex64)
        nat.check(nat.lib.lte_crs_ls_compact(self._plan, _ptr(Yp), _ptr(Hp), rows, S,
                                             C.byref(awgn) if awgn is not None else None, self._stream()),
                  'lte_crs_ls_compact')
        self.launches += 1
        return Hp

    # ------------------------------------------------------------------ stage 5
    def zf(self, Y, H, B, S, window=nat.WINDOW_FULL, out=None, awgn=None):
        """Zero forcing at the data bins; awgn: Y is noise free and the AWGN joins as the bins are read."""
        o = out if out is not None else self._empty((B, S * self.Nd), torch.complex64)
        if awgn is not None:
            nat.check(nat.lib.lte_equalize_zf_awgn(self._plan, _ptr(Y), _ptr(H), _ptr(o), window, B, S, C.byref(awgn),
                                                   self._stream()), 'lte_equalize_zf_awgn')
        else:
            nat.check(nat.lib.lte_equalize_zf(self._plan, _ptr(Y), _ptr(H), _ptr(o), window, B, S, self._stream()),
                      'lte_equalize_zf')
        self.launches += 1
        return o

    def mrc(self, Y, H, B, R, S, window=nat.WINDOW_FULL, out=None):
        o = out if out is not None else self._empty((B, S * self.Nd), torch.complex64)
        nat.check(nat.lib.lte_equalize_mrc(self._plan, _ptr(Y), _ptr(H), _ptr(o), window, B, R, S, self._stream()),
                  'lte_equalize_mrc')
        self.launches += 1
        return o

    # ------------------------------------------------------------------ SFBC Alamouti
    def sfbc_encode(self, S, idx=None, symbols=None, want_qam=False):
        """idx / symbols [B, S*2*(Nd//2)] -> (per-antenna data symbols [B*2, S*Nd], qam or None)."""
        src = idx if idx is not None else symbols
        B = src.shape[0]
        out = self._empty((B * 2, S * self.Nd), torch.complex64)
        qam = self._empty((B, S * 2 * (self.Nd // 2)), torch.complex64) if (want_qam and idx is not None) else None
        nat.check(nat.lib.lte_sfbc_encode(self._plan, _ptr(idx) if symbols is None else None, _ptr(symbols),
                                          _ptr(out), _ptr(qam), B, S, self._stream()), 'lte_sfbc_encode')
        self.launches += 1
        return out, qam

    def tx_sfbc(self, S, idx, out=None):
        """Alamouti encode + per-TX resource grid + IFFT + CP in one kernel: idx [B, S*2*(Nd//2)] -> tx [B, 2, S*L]
        (= sfbc_encode + modulate(symbols=..., T=2) without the encoded symbols in memory)."""
        B = idx.shape[0]
        tx = out if out is not None else self._empty((B, 2, S * self.L), torch.complex64)
        nat.check(nat.lib.lte_tx_sfbc_ifft(self._plan, _ptr(idx), _ptr(tx), B, S, self._stream()), 'lte_tx_sfbc_ifft')
        self.launches += 1
        return tx

    def sfbc_decode(self, Y, H0, H1, B, R, S, window=nat.WINDOW_FULL):
        out = self._empty((B, S * 2 * (self.Nd // 2)), torch.complex64)
        nat.check(nat.lib.lte_sfbc_decode(self._plan, _ptr(Y), _ptr(H0), _ptr(H1), _ptr(out), window, B, R, S,
                                          self._stream()), 'lte_sfbc_decode')
        self.launches += 1
        return out

    # ------------------------------------------------------------------ spatial multiplexing
    DETECTORS = {'MMSE': 0, 'IRC': 0, 'ZF': 1, 'SIC': 2, 'MRC': 3}
    SPECTRAL_MAX_DELAY = 32          # longest tap delay in samples for which simo_ber picks the spectral link by itself

    def sfbc_decode_count(self, Y, H0, H1, idx_tx, B, R, S, window=nat.WINDOW_FULL, nbits=None, errors=None, awgn=None):
        """Alamouti decode + slicer + bit-error count in one kernel: int64 [B] errors against idx_tx [B, S*2*(Nd//2)]."""
        nd2 = 2 * (self.Nd // 2)
        nbits = S * nd2 * self.bps if nbits is None else int(nbits)
        if errors is None:
            errors = torch.zeros(B, dtype=torch.int64, device=self.device)
        else:
            errors.zero_()
        nat.check(nat.lib.lte_sfbc_decode_count(self._plan, _ptr(Y), _ptr(H0), _ptr(H1), _ptr(idx_tx), _ptr(errors), nbits,
                                                window, B, R, S, C.byref(awgn) if awgn is not None else None,
                                                self._stream()), 'lte_sfbc_decode_count')
        self.launches += 1
        return errors

    @staticmethod
    def _w_host(W):
        W = np.ascontiguousarray(np.asarray(W, dtype=np.complex64))
        return W, W.ctypes.data_as(C.c_void_p), W.shape[0], W.shape[1]

    def sm_precode(self, S, W, idx=None, symbols=None, want_qam=False):
        """idx / symbols [B, S*Nd], W [T, L] -> (per-antenna data symbols [B*T, S*Nd], qam or None)."""
        Wc, wp, T, L = self._w_host(W)
        src = idx if idx is not None else symbols
        B = src.shape[0]
        out = self._empty((B * T, S * self.Nd), torch.complex64)
        qam = self._empty((B, S * self.Nd), torch.complex64) if (want_qam and idx is not None) else None
        nat.check(nat.lib.lte_sm_precode(self._plan, _ptr(idx) if symbols is None else None, _ptr(symbols), wp, T, L,
                                         _ptr(out), _ptr(qam), B, S, self._stream()), 'lte_sm_precode')
        self.launches += 1
        return out, qam

    def flat_mimo(self, tx, h, B, R, T):
        """tx [B*T, n], h complex64 [B, R, T] -> (out [B*R, n], power [B, R])."""
        n = tx.shape[-1]
        out = self._empty((B * R, n), torch.complex64)
        power = torch.zeros((B, R), dtype=torch.float64, device=self.device)
        nat.check(nat.lib.lte_flat_mimo(self._plan, _ptr(tx), _ptr(h), _ptr(out), _ptr(power), B, R, T, n,
                                        self._stream()), 'lte_flat_mimo')
        self.launches += 1
        return out, power

    def estimate_pilots(self, Y, rows, S, window=nat.WINDOW_FULL, awgn=None):
        """LS estimates at the pilots of every pilot set on every OFDM symbol: Y [rows, S, nk] -> Hp [sets, rows*S, Np]."""
        Hp = self._empty((self.num_pilot_sets, rows * S, self.Np), torch.complex64)
        nat.check(nat.lib.lte_crs_ls_pilots(self._plan, _ptr(Y), _ptr(Hp), window, rows, S,
                                            C.byref(awgn) if awgn is not None else None, self._stream()), 'lte_crs_ls_pilots')
        self.launches += 1
        return Hp

    def mimo_detect(self, Y, H, W, sigma2, detector, B, R, S, window=nat.WINDOW_FULL, awgn=None, Hpilot=None):
        """Y [B*R, S, nk], H [T, B*R, S, nk] -> detected symbols [B, S*Nd] (demapped layer order).
        H=None: the detector forms the per-symbol CRS estimates itself from Y's pilot bins (same values, no H tensor).
        sigma2: one float for every stream, or a float64 device tensor [B] (one noise variance per stream)."""
        Wc, wp, T, L = self._w_host(W)
        per_stream = None
        if torch.is_tensor(sigma2):
            per_stream = sigma2.to(device=self.device, dtype=torch.float64).contiguous()
            if per_stream.numel() != B:
                raise ValueError("per-stream sigma2 must hold one value per stream")
            sigma2 = 0.0
        det = self.DETECTORS.get(str(detector).upper())
        if det is None:
            raise ValueError(f"Detector '{detector}' no soportado")
        out = torch.zeros((B, S * self.Nd), dtype=torch.complex64, device=self.device)
        nat.check(nat.lib.lte_mimo_detect(self._plan, _ptr(Y), _ptr(H) if H is not None else None,
                                          _ptr(Hpilot) if Hpilot is not None else None, wp, T, L, float(sigma2),
                                          _ptr(per_stream) if per_stream is not None else None, det, _ptr(out),
                                          window, B, R, S, C.byref(awgn) if awgn is not None else None,
                                          self._stream()), 'lte_mimo_detect')
        self.launches += 1
        return out

    def siso_ber(self, chan, snr_lin_rows, S, seed, stream_id0=0, idx=None, nbits=None, sc_fdm=False, noise_domain=1,
                 papr_hist=None, papr_lo=0.0, papr_step=0.1, fused=True):
        """One pass of the SISO chain with the zero-forcing equaliser Y / (H + 1e-6)
        (reference simulate_siso, core/ofdm_core.py:660-737) over B independent streams: what the GUIs
        run for num_rx = 1.  noise_domain 1: AWGN on the kept bins in the RX epilogue; 0: per time sample,
        the very draws of the per-call API with rng='philox' (stream b = call with stream id stream_id0 + b).
        sc_fdm: SC-FDM uplink (reference enable_sc_fdm=True): the Nd data symbols of every OFDM symbol go
        through the unitary Nd-point DFT before the grid (core/dft_precoding.py:67-93) and through its
        inverse after the equaliser (core/lte_receiver.py:319-333).
        papr_hist (int64 [bins], accumulated): histogram of the per-OFDM-symbol PAPR in dB of the useful
        samples (core/ofdm_system.py:173-229), taken in the TX kernel's epilogue -- BER and PAPR in one pass.
        fused (fading channels, noise_domain 1): fading + CP strip + FFT in one kernel and lazy AWGN in the estimator
        and the equaliser instead of the staged TDL / FFT pair; the SC-FDM precoder takes the symbol indices directly."""
        B = snr_lin_rows.shape[0]
        if idx is None:
            idx = self.random_indices(B, S, seed, stream_id0)
        symbols = None
        if sc_fdm:
            symbols = self.dft_qam(idx.view(B * S, self.Nd), self.Nd).view(B, S * self.Nd)   # QAM map inside the DFT's load
        if papr_hist is not None:
            _, _, tx = self.modulate_papr(S, idx=None if sc_fdm else idx, symbols=symbols, write_tx=True,
                                          hist=papr_hist, hist_lo=papr_lo, hist_step=papr_step, want_db=False)
        else:
            tx, _, _ = self.modulate(S, idx=None if sc_fdm else idx, symbols=symbols, want_stats=False)
        if chan.num_taps > 0:
            ph = self.random_phases(B, chan.num_taps * nat.LTE_JAKES_TONES, seed, stream_id0)
            if noise_domain == 1 and fused:
                # fading + CP strip + FFT in one kernel, noise-free grid; the AWGN (the draws of the RX epilogue, from
                # the power this kernel measures) joins in the estimator and in the equaliser
                got = self.channel_rx_fft(tx, chan, B, 1, S, ph, nat.WINDOW_USEFUL)
                if got is not None:
                    Y, power = got
                    awgn = self.awgn_desc(power, snr_lin_rows, seed, stream_id0)
                    H = self.estimate(Y, B, S, nat.WINDOW_USEFUL, awgn=awgn)
                    return self._siso_tail(self.zf(Y, H, B, S, nat.WINDOW_USEFUL, awgn=awgn), idx, B, S, nbits, sc_fdm)
            rx, power = self.channel(tx, chan, B, 1, phases=ph)
        else:
            _, power = self.channel(tx, chan, B, 1)
            rx = tx
        Y = self.rx_fft(rx, B, S, nat.WINDOW_USEFUL, power=power, snr_lin=snr_lin_rows, seed=seed,
                        row_id0=stream_id0, noise_domain=noise_domain)
        H = self.estimate(Y, B, S, nat.WINDOW_USEFUL)
        return self._siso_tail(self.zf(Y, H, B, S, nat.WINDOW_USEFUL), idx, B, S, nbits, sc_fdm)

    def _siso_tail(self, data, idx, B, S, nbits, sc_fdm):
        if sc_fdm:
            data = self.dft_m(data.view(B * S, self.Nd), self.Nd, inverse=True).view(B, S * self.Nd)
        errors, _ = self.demap_count(data, idx_tx=idx, nbits=nbits)
        return errors

    def sfbc_ber(self, chan, snr_lin_rows, S, R, seed, stream_id0=0, idx=None, fused=True):
        """One pass of the 2-TX Alamouti SFBC chain (reference simulate_miso / simulate_mimo,
        core/ofdm_core.py:1850-2258) over B independent streams: SFBC encode, per-TX interleaved CRS,
        R x 2 independently faded links summed per RX antenna, AWGN of power (P_rx / 2) / snr, CRS
        estimates of both TX antennas per 14-symbol slot, Alamouti decode averaged over RX, slicer, count.
        The engine must have been built with tables.mimo_pilot_sets(2, Np); snr_lin_rows: float32 [B*R]
        linear SNR per (stream, antenna).  Returns int64 [B] bit errors over S * 2 * (Nd // 2) symbols.
        fused (default): channel + FFT in one kernel and lazy AWGN, staged kernels where that does not apply."""
        if self.num_pilot_sets != 2:
            raise ValueError("sfbc_ber needs an engine with the two SFBC pilot sets")
        B = snr_lin_rows.shape[0] // R
        nd2 = 2 * (self.Nd // 2)
        if idx is None:
            idx = self.random_indices(B, S, seed, stream_id0, nsym=S * nd2)      # = the first S * nd2 of S * Nd draws
        if chan.num_taps == 0:
            raise ValueError("sfbc_ber models the fading links of the reference's MIMO channel (rayleigh_mp)")
        tx = self.tx_sfbc(S, idx)
        ph = self.random_phases(B, R * 2 * chan.num_taps * nat.LTE_JAKES_TONES, seed, stream_id0)
        snr2 = (snr_lin_rows * 2.0).contiguous()
        if fused:
            # fading + CP strip + FFT in one kernel, noise-free grid; the AWGN joins in the estimator and the decoder
            # (the draws of the staged path's RX epilogue, from the stream power this kernel measures)
            got = self.channel_rx_fft(tx, chan, B, R, S, ph, nat.WINDOW_USEFUL, T=2)
            if got is not None:
                Y, power = got
                awgn = self.awgn_desc(power, snr2, seed, stream_id0 * R)
                H0 = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, pilot_set=0, awgn=awgn)
                H1 = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, pilot_set=1, awgn=awgn)
                return self.sfbc_decode_count(Y, H0, H1, idx, B, R, S, nat.WINDOW_USEFUL, nbits=S * nd2 * self.bps, awgn=awgn)
        rx, power = self.channel(tx, chan, B, R, T=2, phases=ph)
        Y = self.rx_fft(rx.view(B * R, -1), B * R, S, nat.WINDOW_USEFUL, power=power,
                        snr_lin=snr2, seed=seed, row_id0=stream_id0 * R, noise_domain=1)
        H0 = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, pilot_set=0)
        H1 = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, pilot_set=1)
        return self.sfbc_decode_count(Y, H0, H1, idx, B, R, S, nat.WINDOW_USEFUL, nbits=S * nd2 * self.bps)

    def sm_ber(self, chan, W, snr_db, B, S, R, detector, seed, stream_id0=0, idx=None, estimate_tensor=False,
               fused=True):
        """One pass of the TM4-like spatial-multiplexing chain (reference simulate_spatial_multiplexing,
        core/ofdm_core.py:2489-2815) over B independent streams that share the precoder W [T, L].  `snr_db` is
        one float for every stream or a sequence / tensor of B values (stream b runs at snr_db[b]; the detectors
        then take sigma^2 = 10^(-snr/10) per stream, so a sweep's SNR points share one pass): layer mapping +
        precoding, per-TX interleaved CRS, R x T independently faded links summed per RX antenna
        (`chan` built with gain_conversions=3 as that path does), AWGN from each antenna's measured power,
        CRS estimate of every TX antenna on EVERY OFDM symbol (formed inside the detector from the pilot bins
        unless `estimate_tensor`), MMSE / ZF / SIC / MRC detection on
        H_eff = H W, slicer, count.  The engine must carry tables.mimo_pilot_sets(T, Np)."""
        W = np.asarray(W, dtype=complex)
        T = W.shape[0]
        if self.num_pilot_sets != T:
            raise ValueError("sm_ber needs an engine built with the pilot sets of its T TX antennas")
        if idx is None:
            idx = self.random_indices(B, S, seed, stream_id0)
        data, _ = self.sm_precode(S, W, idx=idx)
        tx, _, _ = self.modulate(S, symbols=data, T=T, want_stats=False)
        ph = self.random_phases(B, R * T * chan.num_taps * nat.LTE_JAKES_TONES, seed, stream_id0)
        got = self.channel_rx_fft(tx, chan, B, R, S, ph, nat.WINDOW_USEFUL, T=T) if (fused and not estimate_tensor) else None
        if got is None:
            rx, power = self.channel(tx, chan, B, R, T=T, phases=ph)
        if torch.is_tensor(snr_db) or np.ndim(snr_db) > 0:
            snr_b = torch.as_tensor(snr_db, dtype=torch.float64, device=self.device).reshape(-1)
            if snr_b.numel() != B:
                raise ValueError("per-stream snr_db must hold one value per stream")
            snr_rows = torch.pow(10.0, snr_b / 10).to(torch.float32).repeat_interleave(R).contiguous()
            sigma2 = torch.pow(10.0, -snr_b / 10)
        else:
            snr_rows = torch.full((B * R,), float(10 ** (snr_db / 10)), dtype=torch.float32, device=self.device)
            sigma2 = 10 ** (-snr_db / 10)
        # only the occupied bins travel: FFT window, per-TX estimates (written straight into one [T, ...] tensor) and detector
        k0, nk = self.window(nat.WINDOW_USEFUL)
        if got is not None:
            # fused channel + FFT: noise-free grid, the AWGN joins inside the detector (data and pilot bins)
            Y, power = got
            awgn = self.awgn_desc(power, snr_rows, seed, stream_id0 * R)
            Hp = self.estimate_pilots(Y, B * R, S, nat.WINDOW_USEFUL, awgn=awgn)     # each pilot's noise sample once
            sym = self.mimo_detect(Y, None, W, sigma2, detector, B, R, S, nat.WINDOW_USEFUL, awgn=awgn, Hpilot=Hp)
            errors, _ = self.demap_count(sym, idx_tx=idx, nbits=S * self.Nd * self.bps)
            return errors
        Y = self.rx_fft(rx.view(B * R, -1), B * R, S, nat.WINDOW_USEFUL, power=power, snr_lin=snr_rows, seed=seed,
                        row_id0=stream_id0 * R, noise_domain=1)
        H = None                        # the detector estimates from Y's pilot bins itself (bit-identical values)
        if estimate_tensor:             # the explicit per-TX estimate passes, kept for the API path and the tests
            H = self._empty((T, B * R, S, nk), torch.complex64)
            for t in range(T):
                self.estimate(Y.view(B * R * S, 1, nk), B * R * S, 1, nat.WINDOW_USEFUL, pilot_set=t,
                              out=H[t].view(B * R * S, 1, nk))
        sym = self.mimo_detect(Y, H, W, sigma2, detector, B, R, S, nat.WINDOW_USEFUL)
        errors, _ = self.demap_count(sym, idx_tx=idx, nbits=S * self.Nd * self.bps)
        return errors

    # ------------------------------------------------------------------ coded chain
    def coding_plan(self, A):
        """Cached layout tables of a transport block of A bits (lte_b200/coding.py)."""
        from .coding import CodingPlan
        plans = self.__dict__.setdefault('_coding_plans', {})
        if A not in plans:
            plans[A] = CodingPlan(A, self.device)
        return plans[A]

    def tb_encode(self, bits, plan):
        """bits uint8 [B, A] of 0/1 -> rate-matched coded bits uint8 [B, sumE]
        (CRC-24A, segmentation + CRC-24B, turbo encoder, rate matching)."""
        B = bits.shape[0]
        crc = self._empty((B, 24), torch.uint8)
        cb = self._empty((B, plan.sumK), torch.uint8)
        enc = self._empty((B, plan.sumE), torch.uint8)
        coded = self._empty((B, plan.sumE), torch.uint8)
        nat.check(nat.lib.lte_tb_encode(_ptr(bits), plan.A, _ptr(plan.blk), plan.C, plan.sumK, plan.sumE,
                                        _ptr(plan.rm_table), _ptr(plan.pi_tab), _ptr(crc), _ptr(cb), _ptr(enc),
                                        _ptr(coded), B, self._stream()), 'lte_tb_encode')
        self.launches += 4
        return coded

    def symbol_interleave(self, idx, rows):
        """idx uint8 [B, nsym] -> QAM symbols through the rows x Nd block interleaver, complex64 [B, rows*Nd]."""
        B, nsym = idx.shape
        out = self._empty((B, rows * self.Nd), torch.complex64)
        nat.check(nat.lib.lte_symbol_interleave(self._plan, _ptr(idx), nsym, rows, _ptr(out), B, self._stream()),
                  'lte_symbol_interleave')
        self.launches += 1
        return out

    def soft_demap(self, data, H, sigma2, fading, nsym, rows, window=nat.WINDOW_FULL):
        """data [B, rows*Nd] equalised symbols (received order), H [B, nslot, nk], sigma2 float32 [B] ->
        de-interleaved LLRs float32 [B, nsym*bps]."""
        B = data.shape[0]
        llr = self._empty((B, nsym * self.bps), torch.float32)
        nat.check(nat.lib.lte_soft_demap(self._plan, _ptr(data), _ptr(H), window, _ptr(sigma2), 1 if fading else 0,
                                         nsym, rows, _ptr(llr), B, self._stream()), 'lte_soft_demap')
        self.launches += 1
        return llr

    def tb_decode(self, llr, plan, iterations=8, bits_tx=None, want_bits=True, logmap=False):
        """llr float32 [B, sumE] -> (bits_rx uint8 [B, A] or None, crc_ok int32 [B], errors int64 [B] or None)."""
        B = llr.shape[0]
        dem = self._empty((B, plan.sumE), torch.float32)
        work = self._empty((-(-B * plan.C // 4) * 4, plan.work_floats), torch.float32)     # whole warps of 4 blocks
        cbdec = self._empty((B, plan.sumK), torch.uint8)
        bits_rx = self._empty((B, plan.A), torch.uint8) if want_bits else None
        crc_ok = self._empty((B,), torch.int32)
        errors = torch.zeros(B, dtype=torch.int64, device=self.device) if bits_tx is not None else None
        nat.check(nat.lib.lte_tb_decode(_ptr(llr), _ptr(plan.blk), plan.C, plan.sumK, plan.sumE, plan.Kmax,
                                        _ptr(plan.dm_table), _ptr(plan.pi_tab), int(iterations), 1 if logmap else 0, _ptr(dem),
                                        _ptr(work),
                                        _ptr(cbdec), plan.A, _ptr(bits_tx), _ptr(bits_rx), _ptr(crc_ok), _ptr(errors),
                                        B, self._stream()), 'lte_tb_decode')
        self.launches += 3
        return bits_rx, crc_ok, errors

    def coded_tx(self, bits, plan):
        """bits [B, A] -> (tx [B, rows*L], rows, nsym, stats [B, 2] peak / power sums): coding chain, QAM map,
        block interleaver, resource grid + IFFT + CP."""
        coded = self.tb_encode(bits, plan)
        nsym = -(-plan.sumE // self.bps)
        idx = self._coded_indices(coded, plan, nsym)
        rows = -(-nsym // self.Nd)
        sym = self.symbol_interleave(idx, rows)
        tx, _, stats = self.modulate(rows, symbols=sym, want_stats=True)
        return tx, rows, nsym, stats

    def _coded_indices(self, coded, plan, nsym):
        B = coded.shape[0]
        idx = self._empty((B, nsym), torch.uint8)
        nat.check(nat.lib.lte_bits_to_indices(self._plan, _ptr(coded), plan.sumE, _ptr(idx), nsym, B, self._stream()),
                  'lte_bits_to_indices')
        self.launches += 1
        return idx

    def coded_rx(self, rx, plan, rows, nsym, sigma2, fading, bits_tx=None, iterations=8, want_bits=True):
        """rx [B, rows*L] noisy received streams -> dict(bits_rx, crc_ok, errors, symbols, H, llr)."""
        B = rx.shape[0]
        Y = self.rx_fft(rx, B, rows, nat.WINDOW_USEFUL)
        H = self.estimate(Y, B, rows, nat.WINDOW_USEFUL)
        data = self.zf(Y, H, B, rows, nat.WINDOW_USEFUL)
        llr = self.soft_demap(data, H, sigma2, fading, nsym, rows, nat.WINDOW_USEFUL)
        bits_rx, crc_ok, errors = self.tb_decode(llr[:, :plan.sumE].contiguous(), plan, iterations, bits_tx, want_bits)
        return dict(bits_rx=bits_rx, crc_ok=crc_ok, errors=errors, data=data, H=H, llr=llr)

    def siso_coded_ber(self, bits, chan, snr_db_rows, seed, stream_id0=0, iterations=8):
        """One pass of the coded SISO chain over B streams with Philox channel / noise draws (time-domain
        AWGN from the measured stream power, as the reference).  bits uint8 [B, A]; snr_db_rows: float32 [B] in dB.
        Returns (errors int64 [B], crc_ok int32 [B])."""
        B, A = bits.shape
        plan = self.coding_plan(A)
        tx, rows, nsym, _ = self.coded_tx(bits, plan)
        snr_lin = torch.pow(10.0, snr_db_rows / 10.0).contiguous()
        if chan.num_taps > 0:
            ph = self.random_phases(B, chan.num_taps * nat.LTE_JAKES_TONES, seed, stream_id0)
            faded, power = self.channel(tx, chan, B, 1, phases=ph)
            rx = self.awgn(faded.view(B, -1), 1, power, snr_lin, B, seed=seed, row_id0=stream_id0)
        else:
            _, power = self.channel(tx, chan, B, 1)
            rx = self.awgn(tx, 1, power, snr_lin, B, seed=seed, row_id0=stream_id0)
        sigma2 = (1.0 / snr_lin).contiguous()
        r = self.coded_rx(rx, plan, rows, nsym, sigma2, chan.num_taps > 0, bits_tx=bits, iterations=iterations,
                          want_bits=False)
        return r['errors'], r['crc_ok']

    # ------------------------------------------------------------------ beamforming
    def random_channel(self, B, R, T, seed, stream_id0=0):
        """Flat channel matrices h [B, R, T] ~ CN(0, 1), Philox keyed (seed, stream_id0 + b)."""
        h = self._empty((B, R, T), torch.complex64)
        nat.check(nat.lib.lte_random_channel(_ptr(h), B, R, T, int(seed), int(stream_id0), self._stream()),
                  'lte_random_channel')
        self.launches += 1
        return h

    def bf_weights(self, h, codebook, mode='MRT'):
        """h complex64 [B, R, T]; codebook: list of [T, 1] rank-1 precoders (host) ->
        (W [B, T], heff [B, R], pmi int32 [B], gain_db float32 [B])."""
        B, R, T = h.shape
        cb = np.ascontiguousarray(np.stack([np.asarray(w).reshape(-1) for w in codebook]).astype(np.complex64)) \
            if codebook is not None and len(codebook) else np.zeros((0, T), np.complex64)
        W = self._empty((B, T), torch.complex64)
        heff = self._empty((B, R), torch.complex64)
        pmi = self._empty((B,), torch.int32)
        gain = self._empty((B,), torch.float32)
        m = {'MRT': nat.BF_MRT, 'CODEBOOK': nat.BF_CODEBOOK}[str(mode).upper()]
        nat.check(nat.lib.lte_bf_weights(_ptr(h), cb.ctypes.data_as(C.c_void_p), cb.shape[0], m, _ptr(W), _ptr(heff),
                                         _ptr(pmi), _ptr(gain), B, R, T, self._stream()), 'lte_bf_weights')
        self.launches += 1
        return W, heff, pmi, gain

    def rank_codebook(self, num_tx, max_rank):
        """The TM4 codebooks of ranks 1..max_rank as the device table lte_rank_feedback reads:
        (complex64 [max_rank, ncb_stride, T, 4] with W [T, rank] in the first `rank` columns, int32 sizes per rank)."""
        from core.codebook_lte import LTECodebook
        books = [LTECodebook(num_tx, transmission_mode='TM4', rank=r) for r in range(1, max_rank + 1)]
        sizes = np.array([b.codebook_size for b in books], dtype=np.int32)
        tab = np.zeros((max_rank, int(sizes.max()), num_tx, 4), dtype=np.complex64)
        for r, b in enumerate(books):
            for i in range(b.codebook_size):
                tab[r, i, :, :r + 1] = np.asarray(b.get_precoder(i)).reshape(num_tx, r + 1)
        return torch.from_numpy(tab).to(self.device), sizes

    def rank_feedback(self, H, snr_db, codebook, sizes, rank_threshold=0.15, max_rank=None):
        """RankAdaptation.get_feedback (eigenvalue RI, capacity PMI) for H complex64 [n, R, T] and snr_db float64 [n]
        (device) against `codebook, sizes` of rank_codebook() -> (ri int32 [n], pmi int32 [n])."""
        n, R, T = H.shape
        max_rank = int(max_rank) if max_rank is not None else min(R, T, 4)
        ri = self._empty((n,), torch.int32)
        pmi = self._empty((n,), torch.int32)
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        assert codebook.shape[0] >= max_rank and codebook.shape[2] == T and codebook.shape[3] == 4
        assert snr_db.dtype == torch.float64 and snr_db.numel() == n and H.dtype == torch.complex64
        nat.check(nat.lib.lte_rank_feedback(_ptr(H.contiguous()), _ptr(snr_db), _ptr(codebook), codebook.shape[1],
                                            sizes.ctypes.data_as(C.c_void_p), float(rank_threshold), max_rank, _ptr(ri),
                                            _ptr(pmi), n, R, T, self._stream()), 'lte_rank_feedback')
        self.launches += 1
        return ri, pmi

    def bf_link(self, idx, h, W, heff, noise_std, S, nbits=None, z=None, seed=0, row_id0=0, want_symbols=False,
                errors=None, count=True):
        """idx uint8 [B, S*Nd]; h [B, R, T]; W [B, T]; heff [B, R]; noise_std float32 [B]; z optional float32
        [B, S, 2, R, Nd] replayed normals -> (errors int64 [B] or None, equalised symbols [B, S*Nd] or None)."""
        B, R, T = h.shape
        out = self._empty((B, S * self.Nd), torch.complex64) if want_symbols else None
        if count:
            if errors is None:
                errors = torch.zeros(B, dtype=torch.int64, device=self.device)
            else:
                errors.zero_()
        else:
            errors = None
        nb = int(nbits) if nbits is not None else S * self.Nd * self.bps
        nat.check(nat.lib.lte_bf_link(self._plan, _ptr(idx), _ptr(h), _ptr(W), _ptr(heff), _ptr(noise_std), _ptr(z),
                                      int(seed), int(row_id0), _ptr(out), _ptr(errors), nb, B, R, T, S,
                                      self._stream()), 'lte_bf_link')
        self.launches += 1
        return errors, out

    # ------------------------------------------------------------------ stage 6
    def demap_count(self, syms, idx_tx=None, nbits=None, want_idx=False, errors=None):
        """-> (errors int64 [B] or None, idx_rx uint8 [B, nsym] or None)."""
        B, nsym = syms.shape
        idx_rx = self._empty((B, nsym), torch.uint8) if want_idx else None
        if idx_tx is not None:
            if errors is None:
                errors = torch.zeros(B, dtype=torch.int64, device=self.device)
            else:
                errors.zero_()
        nb = int(nbits) if nbits is not None else nsym * self.bps
        nat.check(nat.lib.lte_demap_count(self._plan, _ptr(syms), _ptr(idx_tx), _ptr(idx_rx), _ptr(errors), nsym, nb,
                                          B, self._stream()), 'lte_demap_count')
        self.launches += 1
        return errors, idx_rx

    def mrc_demap_count(self, Y, H, idx_tx, B, R, S, nbits=None, window=nat.WINDOW_USEFUL, errors=None, awgn=None,
                        accumulate=False):
        """accumulate: add to `errors` (the kernels count with atomics) instead of starting from zero."""
        if errors is None:
            errors = torch.zeros(B, dtype=torch.int64, device=self.device)
        elif not accumulate:
            errors.zero_()
        nb = int(nbits) if nbits is not None else S * self.Nd * self.bps
        if awgn is None:
            nat.check(nat.lib.lte_mrc_demap_count(self._plan, _ptr(Y), _ptr(H), _ptr(idx_tx), _ptr(errors), window,
                                                  nb, B, R, S, self._stream()), 'lte_mrc_demap_count')
        else:
            nat.check(nat.lib.lte_mrc_demap_count_awgn(self._plan, _ptr(Y), _ptr(H), _ptr(idx_tx), _ptr(errors),
                                                       window, nb, B, R, S, C.byref(awgn), self._stream()),
                      'lte_mrc_demap_count_awgn')
        self.launches += 1
        return errors

    def mrc_demap_count_compact(self, Yd, Hp, idx_tx, B, R, S, nbits=None, errors=None, awgn=None, accumulate=False,
                                Yp=None):
        """MRC + slicer + bit-error count on the compact layout (Yd [B*R, S, 2 ndp], Hp [B*R, slots, Np]): every
        thread interpolates its own bin between its two pilots; counts are bit-identical to
        estimate + mrc_demap_count on the windowed layout.  accumulate: add to `errors` instead of zeroing it.
        With Hp = None and Yp [B*R, slots, 2 npp] the LS estimate of estimate_compact is formed inside the same
        launch (lte_crs_mrc_demap_count_compact): same counts, no estimate tensor."""
        if errors is None:
            errors = torch.zeros(B, dtype=torch.int64, device=self.device)
        elif not accumulate:
            errors.zero_()
        nb = int(nbits) if nbits is not None else S * self.Nd * self.bps
        aw = C.byref(awgn) if awgn is not None else None
        if Hp is None:
            nat.check(nat.lib.lte_crs_mrc_demap_count_compact(self._plan, _ptr(Yd), _ptr(Yp), _ptr(idx_tx), _ptr(errors),
                                                              nb, B, R, S, aw, self._stream()),
                      'lte_crs_mrc_demap_count_compact')
        else:
            nat.check(nat.lib.lte_mrc_demap_count_compact(self._plan, _ptr(Yd), _ptr(Hp), _ptr(idx_tx), _ptr(errors), nb, B,
                                                          R, S, aw, self._stream()), 'lte_mrc_demap_count_compact')
        self.launches += 1
        return errors

    def side_streams(self, n):
        """n CUDA streams of this engine's device for batches in flight side by side (created once)."""
        have = self.__dict__.setdefault('_side', [])
        while len(have) < n:
            have.append(torch.cuda.Stream(device=self.device))
        return have[:n]

    def batches_in_flight(self, chan, B, R, S, fused=True, spectral=None):
        """How many batches simo_ber_batches should keep in flight: two where the spectral link runs (measured
        2.28-2.32 -> 2.39 M subframes/s on one box), one through the time-domain kernels, where a second batch in
        flight gains nothing (fused pipeline on one box: 1.75 / 1.79 M with one, 1.75 / 1.75 M with two -- it runs at
        the board's power cap either way)."""
        if spectral is None:
            spectral = fused and chan.num_taps > 0 and max(chan.delay[:chan.num_taps]) <= self.SPECTRAL_MAX_DELAY
        ok = bool(spectral) and fused and chan.num_taps > 0 and self.Np > 0 and self.num_pilot_sets == 1 \
            and self.spectral_workspace_bytes(chan, B, R, S) is not None
        return 2 if ok else 1

    def simo_ber_batches(self, workspaces, chan, snr_lin_rows, seed, stream_ids, idx=None, nbits=None, noise_domain=3,
                         fused=True, spectral=None):
        """Several batches of one shape through simo_ber, round-robin over len(workspaces) workspaces and as many
        side streams: the kernels of a batch stay in order on their stream, while the front of the next batch (TX
        spectra) fills the SMs the tail of the previous one (MRC, the persistent channel kernel's last CTAs) leaves
        idle -- measured 1.739 -> 1.687 ms per 4096-subframe batch with two workspaces, no further gain with three
        or four.  stream_ids: stream_id0 of every batch.  Per-slot error counts ACCUMULATE in each workspace's
        'errors' (the caller zeroes them before and sums them over the workspaces after).  Forks from and joins the
        current stream, so events recorded on it bracket all the work."""
        main = torch.cuda.current_stream(self.device)
        streams = self.side_streams(len(workspaces))
        for st in streams:
            st.wait_stream(main)
        for i, sid in enumerate(stream_ids):
            k = i % len(workspaces)
            with torch.cuda.stream(streams[k]):
                self.simo_ber(workspaces[k], chan, snr_lin_rows, seed, stream_id0=int(sid), idx=idx, nbits=nbits,
                              fused=fused, spectral=spectral, noise_domain=noise_domain, accumulate=True)
        for st in streams:
            main.wait_stream(st)

    # ------------------------------------------------------------------ host-buffer front end
    def stream_host_batches(self, chan, num_rx, snr_lin_rows, B, S, nbits=None, seed=0, noise_domain=3, fused=True,
                            depth=3, spectral=None):
        """Pipeline that takes payload batches from HOST memory (np.packbits rows, ideally pinned) and returns the
        per-stream bit-error counts in pinned host memory, overlapping the H2D copy of the next batch with the
        kernels of the current one (lte_b200/host_stream.py)."""
        from .host_stream import HostBatchPipeline
        return HostBatchPipeline(self, chan, num_rx, snr_lin_rows, B, S, nbits=nbits, seed=seed,
                                 noise_domain=noise_domain, fused=fused, depth=depth, spectral=spectral)

    # ------------------------------------------------------------------ batched SIMO chain
    def workspace(self, B, S, R, fading, fused=False, lazy=False):
        """Pre-allocated HBM buffers for `simo_ber` so a sweep re-uses them every step
        (fused: no buffer for the faded streams; lazy: the time-domain stream and the windowed grid are only
        allocated if a step actually takes the fused / staged path -- the spectral link needs neither)."""
        k0, nk = self.window(nat.WINDOW_USEFUL)
        nslot = -(-S // nat.LTE_SLOT_SYMBOLS)
        n = S * self.L
        ws = dict(B=B, S=S, R=R,
                  idx=self._empty((B, S * self.Nd), torch.uint8),
                  stats=torch.zeros((B, 2), dtype=torch.float64, device=self.device),
                  power=torch.zeros((B, R), dtype=torch.float64, device=self.device),
                  errors=torch.zeros(B, dtype=torch.int64, device=self.device))
        if not lazy:
            ws.update(tx=self._empty((B, n), torch.complex64), Y=self._empty((B * R, S, nk), torch.complex64),
                      H=self._empty((B * R, nslot, nk), torch.complex64))
        if fading:
            if not fused:
                ws['faded'] = self._empty((B, R, n), torch.complex64)
            ws['phases'] = self._empty((B, R * nat.LTE_MAX_TAPS * nat.LTE_JAKES_TONES), torch.float32)
        return ws

    def _spectral_buffers(self, ws, chan):
        """Lazily adds the spectral link's buffers to a workspace; False when the link is unsupported."""
        if 'spectral' not in ws:
            B, S, R = ws['B'], ws['S'], ws['R']
            ok = self.Np > 0 and self.num_pilot_sets == 1 and self.spectral_workspace_bytes(chan, B, R, S) is not None
            ws['spectral'] = ok
            if ok:
                k0, nk = self.window(nat.WINDOW_USEFUL)
                nslot = -(-S // nat.LTE_SLOT_SYMBOLS)
                ws['G'] = self._empty((B * S, nk), torch.complex64)
                ws['tail'] = self._empty((B * S, self.cp), torch.complex64)
                ws['Yd'] = self._empty((B * R, S, 2 * self.ndp), torch.complex64)
                ws['Yp'] = self._empty((B * R, nslot, 2 * self.npp), torch.complex64)
                ws['Hp'] = self._empty((B * R, nslot, self.Np), torch.complex64)
        return ws['spectral']

    def simo_ber(self, ws, chan, snr_lin_rows, seed, stream_id0=0, idx=None, nbits=None, noise_domain=1,
                 fused=False, spectral=None, accumulate=False):
        """One pass of the SIMO-MRC link chain over B independent streams.

        ws: workspace(); snr_lin_rows: float32 [B*R] linear SNR per (stream, antenna);
        idx: transmitted symbol indices [B, S*Nd] or None to draw them with Philox
        keyed (seed, stream_id0 + b).  Returns the int64 [B] bit-error counts (in ws).
        noise_domain: 0 time-domain AWGN, 1 on the kept bins in the RX epilogue, 2 the same draws
        added lazily by the CRS / MRC kernels, 3 like 2 but one equivalent draw per MRC output
        (statistically identical, 1/R of the generator work).  fused: fading channel + RX FFT in
        one kernel (the faded streams are never written; needs noise_domain 2 or 3 -- 1 is promoted
        to 2; falls back to the staged kernels when lte_channel_rx_fft reports the configuration
        unsupported).  spectral (default: chosen automatically when `fused`): at low Doppler and short delay spreads
        (<= SPECTRAL_MAX_DELAY samples; True forces it wherever it is supported) use the spectral link
        (lte_tx_spectral + lte_channel_spectral, compact grid) instead of the fused time-domain kernel;
        silently falls back to `fused` outside its validity range.
        accumulate: the counts are ADDED to ws['errors'] (a sweep whose batches keep the same SNR per stream slot
        then needs no per-batch zeroing / reduction: the caller zeroes ws['errors'] once and reads it at the end).
        """
        B, S, R = ws['B'], ws['S'], ws['R']
        if idx is None:
            idx = self.random_indices(B, S, seed, stream_id0, out=ws['idx'])
        if spectral is None:
            # automatic choice: the spectral link pays while the Horner sweep over the delay spread is short; measured
            # at 20 MHz / 1x4 per 4096 subframes: Pedestrian_A (13 samples) 1.85 ms against 2.25 ms through the fused
            # time-domain kernel, Vehicular_A (77) 3.09 against 2.48 ms, Pedestrian_B (114) 3.66 against 2.48 ms
            spectral = fused and chan.num_taps > 0 and max(chan.delay[:chan.num_taps]) <= self.SPECTRAL_MAX_DELAY
        if spectral and chan.num_taps > 0 and self._spectral_buffers(ws, chan):
            # low Doppler: the spectral link (csrc/spectral.cu) -- no time-domain stream, one forward transform
            # per OFDM symbol, compact grid; the lazy AWGN draws are those of the other paths
            per = R * chan.num_taps * nat.LTE_JAKES_TONES
            ph = self.random_phases(B, per, seed, stream_id0, out=ws['phases'].view(-1)[:B * per].view(B, per))
            G, tail = self.tx_spectral(S, idx, out_G=ws['G'], out_tail=ws['tail'])
            (Yd, Yp), power = self.channel_spectral(idx, G, tail, chan, B, R, S, ph, out=ws['Yd'], power=ws['power'],
                                                    compact=True, out_pilots=ws['Yp'])
            awgn = self.awgn_desc(power, snr_lin_rows, seed, stream_id0 * R, combine=(noise_domain == 3))
            # LS estimate, interpolation, MRC, slicer and count in one launch
            return self.mrc_demap_count_compact(Yd, None, idx, B, R, S, nbits=nbits, errors=ws['errors'], awgn=awgn,
                                                accumulate=accumulate, Yp=Yp)
        if 'tx' not in ws:
            ws['tx'] = self._empty((B, S * self.L), torch.complex64)
            k0, nk = self.window(nat.WINDOW_USEFUL)
            ws['Y'] = self._empty((B * R, S, nk), torch.complex64)
            ws['H'] = self._empty((B * R, -(-S // nat.LTE_SLOT_SYMBOLS), nk), torch.complex64)
        tx, _, _ = self.modulate(S, idx=idx, want_stats=False, out=ws['tx'])
        if fused and chan.num_taps > 0:
            per = R * chan.num_taps * nat.LTE_JAKES_TONES
            ph = self.random_phases(B, per, seed, stream_id0, out=ws['phases'].view(-1)[:B * per].view(B, per))
            got = self.channel_rx_fft(tx, chan, B, R, S, ph, nat.WINDOW_USEFUL, out=ws['Y'], power=ws['power'])
            if got is not None:
                Y, power = got
                awgn = self.awgn_desc(power, snr_lin_rows, seed, stream_id0 * R, combine=(noise_domain == 3))
                H = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, out=ws['H'], awgn=awgn)
                return self.mrc_demap_count(Y, H, idx, B, R, S, nbits=nbits, errors=ws['errors'], awgn=awgn, accumulate=accumulate)
            noise_domain = 3 if noise_domain == 3 else 2
            if 'faded' not in ws:
                ws['faded'] = self._empty((B, R, S * self.L), torch.complex64)
        if chan.num_taps > 0:
            per = R * chan.num_taps * nat.LTE_JAKES_TONES
            ph = self.random_phases(B, per, seed, stream_id0, out=ws['phases'].view(-1)[:B * per].view(B, per))
            faded, power = self.channel(tx, chan, B, R, phases=ph, out=ws['faded'], power=ws['power'])
            rx, div = faded, 1
        else:
            _, power = self.channel(tx, chan, B, R, power=ws['power'])
            rx, div = tx, R
        if noise_domain >= 2:       # noise-free grid, AWGN added lazily by the consumers (2: same draws as 1)
            Y = self.rx_fft(rx, B * R, S, nat.WINDOW_USEFUL, rx_div=div, out=ws['Y'])
            awgn = self.awgn_desc(power, snr_lin_rows, seed, stream_id0 * R, combine=(noise_domain == 3))
            H = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, out=ws['H'], awgn=awgn)
            return self.mrc_demap_count(Y, H, idx, B, R, S, nbits=nbits, errors=ws['errors'], awgn=awgn, accumulate=accumulate)
        Y = self.rx_fft(rx, B * R, S, nat.WINDOW_USEFUL, rx_div=div, power=power, snr_lin=snr_lin_rows, seed=seed,
                        row_id0=stream_id0 * R, out=ws['Y'], noise_domain=noise_domain)
        H = self.estimate(Y, B * R, S, nat.WINDOW_USEFUL, out=ws['H'])
        return self.mrc_demap_count(Y, H, idx, B, R, S, nbits=nbits, errors=ws['errors'], accumulate=accumulate)


def _on_own_device(fn):
    """Every native call runs with the engine's device current: streams, workspaces and the SM-count queries
    of the launchers belong to it (an engine on cuda:1 may be driven from a process whose current device is 0)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


for _name, _fn in list(vars(LinkEngine).items()):
    if not _name.startswith('_') and not isinstance(_fn, (staticmethod, classmethod, property)) and callable(_fn) \
            and _name not in ('window', 'symbols_for_bits'):
        setattr(LinkEngine, _name, _on_own_device(_fn))
