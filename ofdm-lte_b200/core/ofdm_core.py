"""Orchestration layer with the reference's class names, signatures and result-dict keys
(reference core/ofdm_core.py): OFDMTransmitter / OFDMReceiver / OFDMChannel / OFDMSimulator.

The bodies are sequences of CUDA stage launches on tensors that stay in HBM; NumPy arrays
appear only at the API boundary.  Engine-only options are keyword-only and default to the
reference's behaviour:

    rng='numpy'    draw channel phases / noise from NumPy's legacy global RNG in the reference's
                   order, including its re-seeding side effects (bit-compatible replay);
    rng='philox'   counter-based draws on the GPU (independent trials), keyed by `seed`.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from config import LTEConfig
from lte_b200 import _native as nat
from lte_b200 import tables

from . import _backend as be
from .channel import ChannelSimulator
from .demodulator import OFDMDemodulator
from .modulator import OFDMModulator


class OFDMTransmitter:
    """reference core/ofdm_core.py:42-155."""

    def __init__(self, config: LTEConfig, mode: str = 'lte', enable_sc_fdm: bool = False):
        self.config = config
        self.mode = mode
        self.enable_sc_fdm = enable_sc_fdm
        self.modulator = OFDMModulator(config, mode=mode, enable_sc_fdm=enable_sc_fdm)
        self.last_signal_tx = None
        self.last_symbols_tx = None
        self.last_mapping_infos = None

    def modulate(self, bits) -> Tuple[np.ndarray, List[np.ndarray], Dict]:
        if not isinstance(bits, (np.ndarray, torch.Tensor)):
            bits = np.array(bits, dtype=int)
        if (bits.numel() if isinstance(bits, torch.Tensor) else bits.size) == 0:
            raise ValueError("Bits array cannot be empty")
        signal_tx, symbols_tx, mapping_infos = self.modulator.modulate_stream(bits)
        self.last_signal_tx, self.last_symbols_tx, self.last_mapping_infos = signal_tx, symbols_tx, mapping_infos
        return signal_tx, symbols_tx, mapping_infos

    @staticmethod
    def _papr_from_stats(stats, n):
        peak, total = float(stats[0]), float(stats[1])
        avg = total / n
        if avg > 0:
            lin = peak / avg
            return {'papr_db': 10 * np.log10(lin), 'papr_linear': lin, 'peak_power': peak, 'avg_power': avg}
        return {'papr_db': 0.0, 'papr_linear': 1.0, 'peak_power': peak, 'avg_power': avg}

    def calculate_papr(self, signal) -> Dict:
        """max |x|^2 / mean |x|^2 over the whole stream (reference :114-147)."""
        x = be.as_complex_tensor(signal).reshape(-1)
        p = x.real ** 2 + x.imag ** 2
        return self._papr_from_stats((p.max().item(), p.double().sum().item()), x.numel())

    def get_config(self) -> LTEConfig:
        return self.config

    def __repr__(self) -> str:
        return f"OFDMTransmitter({self.config.modulation}, {'SC-FDM' if self.enable_sc_fdm else 'OFDM'})"


class OFDMReceiver:
    """reference core/ofdm_core.py:158-276."""

    def __init__(self, config: LTEConfig, mode: str = 'lte', enable_equalization: bool = True,
                 enable_sc_fdm: bool = False):
        self.config = config
        self.mode = mode
        self.enable_equalization = enable_equalization
        self.enable_sc_fdm = enable_sc_fdm
        self.demodulator = OFDMDemodulator(config, mode=mode, enable_equalization=enable_equalization,
                                           enable_sc_fdm=enable_sc_fdm)
        self.last_symbols_rx = None
        self.last_bits_rx = None
        self.channel_estimate = None

    def demodulate(self, signal_rx) -> Tuple[np.ndarray, np.ndarray]:
        if (signal_rx.numel() if isinstance(signal_rx, torch.Tensor) else np.size(signal_rx)) == 0:
            raise ValueError("Received signal cannot be empty")
        symbols, bits_rx = self.demodulator.demodulate_stream(signal_rx)
        self.last_symbols_rx, self.last_bits_rx = symbols, bits_rx
        return symbols, bits_rx

    def estimate_channel(self) -> Dict:
        return {'estimated': False, 'method': 'none'}

    def calculate_ber(self, bits_tx, bits_rx) -> float:
        n = min(len(bits_tx), len(bits_rx))
        if n == 0:
            return 0.0
        return float(np.sum(np.asarray(bits_tx[:n]) != np.asarray(bits_rx[:n])) / n)

    def get_config(self) -> LTEConfig:
        return self.config

    def __repr__(self) -> str:
        return f"OFDMReceiver({self.config.modulation}, {'SC-FDM' if self.enable_sc_fdm else 'OFDM'})"


class OFDMChannel:
    """reference core/ofdm_core.py:279-557."""

    def __init__(self, channel_type: str = 'awgn', snr_db: float = 10.0, fs: float = 15.36e6,
                 itu_profile: str = 'Pedestrian_A', frequency_ghz: float = 2.0, velocity_kmh: float = 0):
        self.channel_type = channel_type
        self.snr_db = snr_db
        self.fs = fs
        self.profile = itu_profile
        self.frequency_ghz = frequency_ghz
        self.velocity_kmh = velocity_kmh
        if channel_type == 'rayleigh_mp':
            self.channel = ChannelSimulator(channel_type='rayleigh_mp', snr_db=snr_db, fs=fs, itu_profile=itu_profile,
                                            frequency_ghz=frequency_ghz, velocity_kmh=velocity_kmh)
        else:
            self.channel = ChannelSimulator('awgn', snr_db=snr_db)

    def set_snr(self, snr_db: float) -> None:
        self.snr_db = snr_db
        self.channel.set_snr(snr_db)

    def transmit(self, signal_tx) -> np.ndarray:
        return self.channel.transmit(signal_tx)

    def _chan_desc(self):
        return tables.channel_desc(self.channel_type, self.fs, self.profile, self.frequency_ghz, self.velocity_kmh)

    def _transmit_simo_device(self, eng, tx_t, num_rx, draws):
        """R independent links of one TX stream (reference :361-412).
        -> (rx [R, n], faded [R, n] or None)."""
        n = tx_t.shape[-1]
        chan = self._chan_desc()
        snr = torch.full((num_rx,), float(10 ** (self.snr_db / 10)), dtype=torch.float32, device=tx_t.device)
        if draws.kind == 'numpy':
            # reference order per antenna: taps x rand(16), normal(n), normal(n)
            us, zs = [], []
            for _ in range(num_rx):
                us.append(draws.phases(chan.num_taps))
                zs.append(draws.unit_normals(n))
            u = torch.stack(us).reshape(1, -1) if chan.num_taps else None
            z = torch.stack(zs)
            faded, power = eng.channel(tx_t, chan, 1, num_rx, phases=u)
            src, div = (faded.reshape(num_rx, n), 1) if faded is not None else (tx_t, num_rx)
            rx = eng.awgn(src, div, power.reshape(-1), snr, num_rx, z=z)
        else:
            sid = draws.next_stream()
            u = eng.random_phases(1, num_rx * chan.num_taps * nat.LTE_JAKES_TONES, draws.seed, sid) \
                if chan.num_taps else None
            faded, power = eng.channel(tx_t, chan, 1, num_rx, phases=u)
            src, div = (faded.reshape(num_rx, n), 1) if faded is not None else (tx_t, num_rx)
            rx = eng.awgn(src, div, power.reshape(-1), snr, num_rx, seed=draws.seed, row_id0=sid * num_rx)
        return rx, (faded.reshape(num_rx, n) if faded is not None else None)

    def transmit_simo(self, signal_tx, num_rx: int = 2) -> List[np.ndarray]:
        if num_rx < 1:
            raise ValueError("num_rx must be >= 1")
        eng = be.engine_for(_FsOnly(self.fs))
        tx_t = be.as_complex_tensor(signal_tx).reshape(1, -1)
        rx, _ = self._transmit_simo_device(eng, tx_t, num_rx, be.NumpyDraws())
        out = be.to_numpy(rx)
        return [out[r] for r in range(num_rx)]


    def _transmit_mimo_device(self, eng, tx_t, num_rx, draws):
        """T x num_rx links (reference :434-543): every (rx, tx) link is faded independently and, in
        the reference, carries its own AWGN at 100 dB; the links are summed per RX antenna and one
        AWGN with noise power (P_rx / T) / snr is added.  tx_t: [T, n] -> (rx [R, n], H [R, T])."""
        T, n = tx_t.shape
        chan = self._chan_desc()
        dev = tx_t.device
        snr_lin = float(10 ** (self.snr_db / 10))
        chan_mat = torch.zeros((num_rx, T), dtype=torch.complex128, device=dev)
        if chan.num_taps == 0:
            # fixed links h = exp(j pi tx / 2) (reference :479-486)
            h = torch.tensor([np.exp(1j * (t * np.pi / 2)) if t else 1.0 + 0j for t in range(T)],
                             dtype=torch.complex64, device=dev)
            acc = (tx_t * h[:, None]).sum(dim=0, keepdim=True).repeat(num_rx, 1).contiguous()
            chan_mat[:] = h.to(torch.complex128)[None, :]
            link_z = None
        else:
            if draws.kind == 'numpy':          # per rx, per tx: taps x rand(16), normal(n), normal(n)
                us, zl, zr = [], [], []
                for _ in range(num_rx):
                    ur, zlr = [], []
                    for _ in range(T):
                        ur.append(draws.phases(chan.num_taps))
                        zlr.append(draws.unit_normals(n))
                    us.append(torch.stack(ur))
                    zl.append(torch.stack(zlr))
                    zr.append(draws.unit_normals(n))
                u = torch.stack(us)             # [R, T, taps, 16]
                link_z = torch.stack(zl)        # [R, T, n]
                rx_z = torch.stack(zr)
            else:
                sid = draws.next_stream()
                u = eng.random_phases(1, num_rx * T * chan.num_taps * nat.LTE_JAKES_TONES, draws.seed, sid) \
                    .reshape(num_rx, T, chan.num_taps, nat.LTE_JAKES_TONES)
                link_z = None
            acc = torch.zeros((num_rx, n), dtype=torch.complex64, device=dev)
            hi_snr = torch.full((num_rx,), 1e10, dtype=torch.float32, device=dev)
            tx_pow = (tx_t.real.double() ** 2 + tx_t.imag.double() ** 2).mean(dim=1)
            for t in range(T):
                faded, power = eng.channel(tx_t[t:t + 1], chan, 1, num_rx, phases=u[:, t].reshape(1, -1))
                link = faded.reshape(num_rx, n)
                if link_z is not None:          # the reference's per-link noise at 100 dB (:495-505)
                    link = eng.awgn(link, 1, power.reshape(-1), hi_snr, num_rx, z=link_z[:, t].contiguous())
                rx_pow = (link.real.double() ** 2 + link.imag.double() ** 2).mean(dim=1)
                corr = (link.to(torch.complex128) * tx_t[t].conj().to(torch.complex128)[None, :]).mean(dim=1)
                chan_mat[:, t] = torch.sqrt(rx_pow / tx_pow[t]) * torch.exp(1j * torch.angle(corr))
                acc += link
        _, p_acc = eng.channel(acc, tables.channel_desc('awgn', self.fs), num_rx, 1)
        snr = torch.full((num_rx,), snr_lin * T, dtype=torch.float32, device=dev)   # (P / T) / snr
        if draws.kind == 'numpy':
            if chan.num_taps == 0:
                rx_z = torch.stack([draws.unit_normals(n) for _ in range(num_rx)])
            rx = eng.awgn(acc, 1, p_acc.reshape(-1), snr, num_rx, z=rx_z)
        else:
            sid2 = draws.next_stream()
            rx = eng.awgn(acc, 1, p_acc.reshape(-1), snr, num_rx, seed=draws.seed, row_id0=sid2 * num_rx)
        return rx, chan_mat

    def transmit_mimo(self, signals_tx, num_rx: int = 1):
        if len(signals_tx) == 0:
            raise ValueError("No transmitted signals provided")
        n = len(signals_tx[0])
        for t, sig in enumerate(signals_tx):
            if len(sig) != n:
                raise ValueError(f"TX signal {t} length mismatch")
        eng = be.engine_for(_FsOnly(self.fs))
        tx_t = torch.stack([be.as_complex_tensor(s).reshape(-1) for s in signals_tx])
        rx, H = self._transmit_mimo_device(eng, tx_t, num_rx, be.NumpyDraws())
        out = be.to_numpy(rx)
        return [out[r] for r in range(num_rx)], be.to_numpy(H)

    def get_config(self) -> Dict:
        return {'type': self.channel_type, 'snr_db': self.snr_db, 'fs': self.fs, 'profile': self.profile,
                'frequency_ghz': self.frequency_ghz, 'velocity_kmh': self.velocity_kmh}

    def __repr__(self) -> str:
        return f"OFDMChannel({self.channel_type}, SNR={self.snr_db}dB, {self.profile})"


class _FsOnly:
    N, Nc, cp_length, bits_per_symbol = 128, 76, 9, 2

    def __init__(self, fs):
        self.fs = fs


class OFDMSimulator:
    """SISO / SIMO orchestrator (reference core/ofdm_core.py:560-737, 1340-1679, 1795-1846)."""

    def __init__(self, config: Optional[LTEConfig] = None, channel_type: str = 'awgn', mode: str = 'lte',
                 enable_sc_fdm: bool = False, enable_equalization: bool = True, num_channels: int = 1,
                 itu_profile: str = 'Pedestrian_A', frequency_ghz: float = 2.0, velocity_kmh: float = 0.0,
                 *, rng: str = 'numpy', seed: int = 0):
        if config is None:
            config = LTEConfig()
        self.config = config
        self.channel_type = channel_type
        self.mode = mode
        self.enable_sc_fdm = enable_sc_fdm
        self.enable_equalization = enable_equalization
        self.itu_profile = itu_profile
        self.frequency_ghz = frequency_ghz
        self.velocity_kmh = velocity_kmh
        self._draws = be.make_draws(rng, seed)
        self.tx = OFDMTransmitter(config, mode=mode, enable_sc_fdm=enable_sc_fdm)
        self.rx = OFDMReceiver(config, mode=mode, enable_equalization=enable_equalization,
                               enable_sc_fdm=enable_sc_fdm)
        if self.rx.demodulator.lte_receiver is not None:
            self.rx.demodulator.lte_receiver.faithful_rng = (rng == 'numpy')
        fs = getattr(config, 'fs', 15.36e6)
        self.channels = []
        for _ in range(num_channels):
            if channel_type == 'rayleigh_mp':
                ch = OFDMChannel(channel_type='rayleigh_mp', snr_db=10.0, fs=fs, itu_profile=itu_profile,
                                 frequency_ghz=frequency_ghz, velocity_kmh=velocity_kmh)
            else:        # unknown channel types silently mean AWGN (reference :644-654)
                ch = OFDMChannel('awgn', snr_db=10.0, fs=fs)
            ch.channel._set_draws(self._draws)
            self.channels.append(ch)
        self.last_results = None

    # ------------------------------------------------------------------ helpers
    def _engine(self):
        return self.tx.modulator._engine()

    @staticmethod
    def _check_bits(bits):
        if not isinstance(bits, (np.ndarray, torch.Tensor)):
            bits = np.array(bits, dtype=int)
        if (bits.numel() if isinstance(bits, torch.Tensor) else bits.size) == 0:
            raise ValueError("Bits array cannot be empty")
        return bits

    def _pilot_side_effect(self):
        if self._draws.kind == 'numpy' and self.mode != 'simple':
            be.reference_pilot_side_effect(0, self._engine().Np)

    def _finish(self, eng, data, idx_tx, bits, nbits):
        """slicer + error count + received bits (reference :712-718)."""
        nsym = idx_tx.shape[1]
        errors, idx_rx = eng.demap_count(data[:, :nsym].contiguous(), idx_tx=idx_tx, nbits=nbits, want_idx=True)
        bits_rx = be.to_numpy(eng.indices_to_bits(idx_rx, nbits).reshape(-1), np.int64)
        return int(errors.item()), bits_rx

    # ------------------------------------------------------------------ SISO
    def simulate_siso(self, bits, snr_db: float = 10.0) -> Dict:
        bits = self._check_bits(bits)
        eng = self._engine()
        b_t = be.as_bits_tensor(bits)
        nbits = b_t.shape[1]
        tx, qam, idx, stats, S = self.tx.modulator._modulate_stream_device(b_t)
        self._pilot_side_effect()
        papr = self.tx._papr_from_stats(be.to_numpy(stats)[0], tx.shape[1])
        self.channels[0].set_snr(snr_db)
        sid = self._draws.next_stream() if self._draws.kind == 'philox' else 0
        rx = self.channels[0].channel._transmit_device(tx, row_id0=sid)
        data = self.rx.demodulator._demodulate_stream_device(rx)
        bit_errors, bits_rx = self._finish(eng, data, idx, bits, nbits)
        qam_np = be.to_numpy(qam)
        results = {
            'transmitted_bits': int(nbits), 'received_bits': int(nbits), 'bits_received_array': bits_rx,
            'bit_errors': bit_errors, 'errors': bit_errors, 'ber': float(bit_errors / nbits), 'snr_db': float(snr_db),
            'papr_db': float(papr['papr_db']), 'papr_linear': float(papr['papr_linear']),
            'signal_tx': be.to_numpy(tx.reshape(-1)), 'signal_rx': be.to_numpy(rx.reshape(-1)),
            'symbols_tx': [qam_np[s] for s in range(S)], 'symbols_rx': be.to_numpy(data.reshape(-1)),
        }
        self.last_results = results
        return results

    # ------------------------------------------------------------------ SIMO
    def _demodulate_with_channel_est(self, signal_rx) -> Tuple[np.ndarray, List[np.ndarray]]:
        """One antenna: FFT + CRS estimate, no equalisation (reference :1340-1403)."""
        eng = be.engine_for(self.config)
        lte = self.rx.demodulator.lte_receiver
        rx = be.as_complex_tensor(signal_rx).reshape(1, -1)
        Y, S = lte._fft_device(rx)
        H = eng.estimate(Y, 1, S, nat.WINDOW_FULL)
        if self._draws.kind == 'numpy':
            be.reference_pilot_side_effect(0, eng.Np)
        data = eng.zf(Y, None, 1, S, nat.WINDOW_FULL)
        Hn = be.to_numpy(H.reshape(-1, eng.N))
        return be.to_numpy(data.reshape(-1)), [Hn[s // nat.LTE_SLOT_SYMBOLS] for s in range(S)]

    def _combine_symbols_mrc(self, symbols_rx_list, h_estimates_per_antenna, regularization: float = 1e-10):
        """sum_i conj(H_i) Y_i / (sum_i |H_i|^2 + reg) on data symbols (reference :1405-1534)."""
        eng = be.engine_for(self.config)
        R = len(symbols_rx_list)
        n = max(len(s) for s in symbols_rx_list) if R else 0
        if n == 0:
            raise ValueError("No symbols received from any antenna")
        dev = be.device()
        data_idx = torch.from_numpy(eng.data_idx).to(dev)
        Y = torch.zeros((R, n), dtype=torch.complex64, device=dev)
        for r, s in enumerate(symbols_rx_list):
            m = min(len(s), n)
            Y[r, :m] = be.as_complex_tensor(s).reshape(-1)[:m]
        S = -(-n // eng.Nd)
        H = torch.ones((R, S, eng.Nd), dtype=torch.complex64, device=dev)
        for r, hl in enumerate(h_estimates_per_antenna):
            for s in range(S):
                if len(hl):
                    h = be.as_complex_tensor(hl[min(s, len(hl) - 1)]).reshape(-1)
                    H[r, s] = h[data_idx]
        Hf = H.reshape(R, -1)[:, :n]
        num = (Hf.conj() * Y).sum(dim=0)
        den = (Hf.real ** 2 + Hf.imag ** 2).sum(dim=0) + regularization
        return be.to_numpy(num / den)

    def simulate_simo(self, bits, snr_db: float = 10.0, num_rx: int = 2, combining: str = 'mrc',
                      parallel: bool = True) -> Dict:
        """1 x num_rx with MRC (reference :1536-1679).  `parallel` is accepted for compatibility:
        the antennas are a batch dimension of every kernel launch."""
        bits = self._check_bits(bits)
        eng = be.engine_for(self.config)
        b_t = be.as_bits_tensor(bits)
        nbits = b_t.shape[1]
        tx, qam, idx, stats, S = self.tx.modulator._modulate_stream_device(b_t)
        self._pilot_side_effect()
        papr = self.tx._papr_from_stats(be.to_numpy(stats)[0], tx.shape[1])
        self.channels[0].set_snr(snr_db)
        rx, _ = self.channels[0]._transmit_simo_device(eng, tx, num_rx, self._draws)
        Y, S_rx = self.rx.demodulator.lte_receiver._fft_device(rx)
        H = eng.estimate(Y, num_rx, S_rx, nat.WINDOW_FULL)
        self._pilot_side_effect()
        comb = eng.mrc(Y, H, 1, num_rx, S_rx, nat.WINDOW_FULL)
        bit_errors, bits_rx = self._finish(eng, comb, idx, bits, nbits)
        per_ant = eng.zf(Y, None, num_rx, S_rx, nat.WINDOW_FULL)
        Hn = be.to_numpy(H)                       # [R, nslot, N]
        rx_np, qam_np, per_np = be.to_numpy(rx), be.to_numpy(qam), be.to_numpy(per_ant)
        results = {
            'transmitted_bits': int(nbits), 'received_bits': int(nbits), 'bits_received_array': bits_rx,
            'bit_errors': bit_errors, 'errors': bit_errors, 'ber': float(bit_errors / nbits), 'snr_db': float(snr_db),
            'papr_db': float(papr['papr_db']), 'papr_linear': float(papr['papr_linear']),
            'signal_tx': be.to_numpy(tx.reshape(-1)), 'signal_rx_list': [rx_np[r] for r in range(num_rx)],
            'symbols_tx': [qam_np[s] for s in range(S)], 'symbols_rx_combined': be.to_numpy(comb.reshape(-1)),
            'symbols_rx_list': [per_np[r] for r in range(num_rx)],
            'channel_estimates_per_antenna': [[Hn[r, s // nat.LTE_SLOT_SYMBOLS] for s in range(S_rx)]
                                              for r in range(num_rx)],
            'num_rx': num_rx, 'combining_method': combining, 'diversity_level': num_rx,
            'parallel_processing': parallel,
        }
        self.last_results = results
        return results

    # ------------------------------------------------------------------ SFBC transmit diversity
    def _simulate_sfbc(self, bits, snr_db, num_rx, mode_name):
        """2 TX Alamouti SFBC, num_rx RX, averaging combiner (reference :1850-2258)."""
        bits = self._check_bits(bits)
        eng0 = be.engine_for(self.config)
        sets = tables.mimo_pilot_sets(2, eng0.Np)
        eng = be.engine_for(self.config, pilot_sets=sets)
        nd2 = 2 * (eng.Nd // 2)
        b_t = be.as_bits_tensor(bits)
        nbits = b_t.shape[1]
        S = int(-(-nbits // (nd2 * eng.bps)))
        # symbol indices laid out [S, nd2] (bits_per_ofdm = nd2 * b, reference :1888-1892)
        idx = torch.zeros((1, S * nd2), dtype=torch.uint8, device=b_t.device)
        nat.check(nat.lib.lte_bits_to_indices(eng._plan, b_t.data_ptr(), nbits, idx.data_ptr(), S * nd2, 1,
                                              eng._stream()), 'lte_bits_to_indices')
        data, qam = eng.sfbc_encode(S, idx=idx, want_qam=True)
        tx, _, _ = eng.modulate(S, symbols=data, T=2, want_stats=False)          # [2, S*L]
        if self._draws.kind == 'numpy':                 # TX ends on seed(1); choice(len(pilot_idx[1::2]))
            be.reference_pilot_side_effect(1, len(np.arange(eng.Np)[1::2]))
        # per-OFDM-symbol PAPR averaged in dB (reference :1946-1953, :2016-2017)
        p = (tx.real ** 2 + tx.imag ** 2).reshape(2, S, eng.L)
        papr_sym = 10 * torch.log10(p.max(dim=2).values / p.mean(dim=2))
        papr0, papr1 = float(papr_sym[0].mean()), float(papr_sym[1].mean())
        self.channels[0].set_snr(snr_db)
        rx, chan_mat = self.channels[0]._transmit_mimo_device(eng, tx, num_rx, self._draws)
        S_rx = rx.shape[1] // eng.L
        Y = eng.rx_fft(rx[:, :S_rx * eng.L].contiguous(), num_rx, S_rx, nat.WINDOW_FULL)
        H0 = eng.estimate(Y, num_rx, S_rx, nat.WINDOW_FULL, pilot_set=0)
        H1 = eng.estimate(Y, num_rx, S_rx, nat.WINDOW_FULL, pilot_set=1)
        if self._draws.kind == 'numpy':
            be.reference_pilot_side_effect(1, len(np.arange(eng.Np)[1::2]))
        dec = eng.sfbc_decode(Y, H0, H1, 1, num_rx, S_rx, nat.WINDOW_FULL)
        bit_errors, bits_rx = self._finish(eng, dec, idx[:, :S_rx * nd2].contiguous() if S_rx < S else idx, bits,
                                           min(nbits, S_rx * nd2 * eng.bps))
        if len(bits_rx) < nbits:
            bits_rx = np.pad(bits_rx, (0, nbits - len(bits_rx)), 'constant')
        papr_db = float(np.mean([papr0, papr1]))
        results = {
            'transmitted_bits': int(nbits), 'received_bits': int(nbits), 'bits_received_array': bits_rx,
            'bit_errors': bit_errors, 'errors': bit_errors, 'ber': float(bit_errors / nbits), 'snr_db': float(snr_db),
            'num_tx': 2, 'num_rx': num_rx, 'mode': mode_name, 'diversity_order': 2 * num_rx,
            'channel_matrix': be.to_numpy(chan_mat), 'papr_db_tx0': papr0, 'papr_db_tx1': papr1,
            'papr_db': papr_db, 'papr_linear': 10 ** (papr_db / 10),
            'symbols_rx': be.to_numpy(dec.reshape(-1)),
        }
        self.last_results = results
        return results

    def simulate_miso(self, bits, snr_db: float = 10.0) -> Dict:
        return self._simulate_sfbc(bits, snr_db, 1, 'MISO-SFBC')

    def simulate_mimo(self, bits, snr_db: float = 10.0, num_rx: int = 2) -> Dict:
        return self._simulate_sfbc(bits, snr_db, num_rx, 'MIMO-SFBC')

    # ------------------------------------------------------------------ coded SISO
    def calculate_noise_var_zf(self, H_estimate, snr_db: float) -> float:
        """Effective noise variance after ZF from the harmonic mean of |H|^2 (reference :739-789)."""
        H = np.atleast_1d(np.asarray(H_estimate))
        nv = 1.0 / (10 ** (snr_db / 10))
        if H.size == 0:
            return nv
        hp = np.maximum(np.abs(H) ** 2, 1e-12)
        return float(nv / hp[0]) if len(hp) == 1 else float(nv / (len(hp) / np.sum(1.0 / hp)))

    def _calculate_llrs_qpsk(self, symbols, noise_var):
        """reference :791-815."""
        from .modulator import symbols_to_llrs
        return symbols_to_llrs(symbols, noise_var, 2)

    def _calculate_llrs_16qam(self, symbols, noise_var):
        """reference :817-869 (max-log over the natural-binary raster constellation, clipped to +-10)."""
        from .modulator import symbols_to_llrs
        return symbols_to_llrs(symbols, noise_var, 4)

    def _calculate_llrs_64qam(self, symbols, noise_var):
        """reference :871-923."""
        from .modulator import symbols_to_llrs
        return symbols_to_llrs(symbols, noise_var, 6)

    def simulate_siso_coded(self, bits, snr_db: float = 10.0) -> Dict:
        """SISO link with CRC-24A, code-block segmentation, the rate-1/3 turbo code, rate matching, a
        symbol block interleaver, max-log LLRs and 8 max-log BCJR iterations (reference :925-1338).
        TX coding, interleaving, soft demapping and decoding are the kernels of csrc/coding.cu; the OFDM
        transmitter, channel, FFT, CRS estimate and ZF equaliser are the ones simulate_siso uses."""
        bits = self._check_bits(bits)
        eng = be.engine_for(self.config)
        b_t = be.as_bits_tensor(bits)
        nbits = b_t.shape[1]
        plan = eng.coding_plan(nbits)
        tx, rows, nsym, stats = eng.coded_tx(b_t, plan)
        self._pilot_side_effect()
        papr = self.tx._papr_from_stats(be.to_numpy(stats)[0], tx.shape[1])
        self.channels[0].set_snr(snr_db)
        sid = self._draws.next_stream() if self._draws.kind == 'philox' else 0
        rx = self.channels[0].channel._transmit_device(tx, row_id0=sid)
        fading = self.channels[0].channel_type != 'awgn'
        sigma2 = torch.full((1,), 1.0 / (10 ** (snr_db / 10)), dtype=torch.float32, device=tx.device)
        r = eng.coded_rx(rx.reshape(1, -1), plan, rows, nsym, sigma2, fading, bits_tx=b_t)
        self._pilot_side_effect()
        bits_rx = be.to_numpy(r['bits_rx'].reshape(-1), np.int64)
        bit_errors = int(r['errors'].item())
        # de-interleaved views of what the demapper consumed (reference result keys)
        Nd = eng.Nd
        k0, nk = eng.window(nat.WINDOW_USEFUL)
        data = r['data'].reshape(Nd, rows).T.reshape(-1)[:nsym]
        didx = torch.as_tensor(np.asarray(eng.data_idx) - k0, device=tx.device, dtype=torch.long)
        Hd = r['H'].reshape(-1, nk)[:, didx].repeat_interleave(nat.LTE_SLOT_SYMBOLS, dim=0)[:rows]
        Hd = Hd.reshape(-1).reshape(Nd, rows).T.reshape(-1)[:nsym]
        s2 = 1.0 / (10 ** (snr_db / 10))
        if fading:
            nv = torch.clamp(s2 / torch.clamp(Hd.abs() ** 2, 1e-6, 1e6), min=s2 / 4.0)
            nv_mean = float(nv.mean())
        else:
            nv_mean = s2
        results = {
            'transmitted_bits': int(nbits), 'received_bits': int(nbits), 'bits_received_array': bits_rx,
            'bit_errors': bit_errors, 'ber': float(bit_errors / nbits), 'crc_pass': bool(r['crc_ok'].item()),
            'snr_db': float(snr_db), 'papr_db': float(papr['papr_db']), 'papr_linear': float(papr['papr_linear']),
            'coded_bits_length': int(plan.sumE), 'signal_tx': be.to_numpy(tx.reshape(-1)),
            'signal_rx': be.to_numpy(rx.reshape(-1)), 'symbols_rx': be.to_numpy(data), 'H_estimate': be.to_numpy(Hd),
            'channel_snr_db': 0.0, 'noise_var_mean': nv_mean, 'llrs': be.to_numpy(r['llr'].reshape(-1)),
        }
        self.last_results = results
        return results

    # ------------------------------------------------------------------ beamforming
    def simulate_beamforming(self, bits, snr_db: float = 10.0, num_tx: int = 2, num_rx: int = 1,
                             codebook_type: str = 'TM6', velocity_kmh: float = 3.0,
                             update_mode: str = 'adaptive') -> Dict:
        """Codebook / MRT beamforming over a flat num_rx x num_tx channel (reference :2260-2477):
        per OFDM symbol x = W s, y = H x + n, then MRC with H_eff = H W, slicer and BER.  The whole
        per-symbol loop is one `lte_bf_link` launch; PMI, W, H_eff and the array gain come from
        `lte_bf_weights`.  With rng='numpy' H and the noise replay the caller's global RNG exactly
        as the reference consumes it (randn(R,T) x2, then randn(R,Nd) x2 per OFDM symbol)."""
        from .beamforming_precoder import AdaptiveBeamforming
        from .csi_feedback import CSIFeedback
        bits = self._check_bits(bits)
        eng = be.engine_for(self.config)
        csi = CSIFeedback(num_tx, num_rx, codebook_type=codebook_type)
        b_t = be.as_bits_tensor(bits)
        nbits = b_t.shape[1]
        S = eng.symbols_for_bits(nbits)
        idx = eng.bits_to_indices(b_t, nbits, S)
        Nd = eng.Nd
        if self._draws.kind == 'numpy':
            H = (np.random.randn(num_rx, num_tx) + 1j * np.random.randn(num_rx, num_tx)) / np.sqrt(2)
            h = be.as_complex_tensor(H.astype(np.complex64)[None])
            z = np.empty((S, 2, num_rx, Nd), dtype=np.float32)
            for s in range(S):
                z[s, 0] = np.random.randn(num_rx, Nd)
                z[s, 1] = np.random.randn(num_rx, Nd)
            z_d, seed, sid = torch.from_numpy(z).to(h.device), 0, 0
        else:
            sid = self._draws.next_stream()
            seed = self._draws.seed
            h = eng.random_channel(1, num_rx, num_tx, seed, sid)
            H = be.to_numpy(h[0]).astype(complex)
            z_d = None
        adaptive = update_mode == 'adaptive'
        W, heff, pmi, gain = eng.bf_weights(h, csi.codebook.codebook, mode='MRT' if adaptive else 'CODEBOOK')
        noise_std = torch.full((1,), float(np.sqrt(10 ** (-snr_db / 10) / 2)), dtype=torch.float32, device=h.device)
        errors, sym = eng.bf_link(idx, h, W, heff, noise_std, S, nbits=nbits, z=z_d, seed=seed, row_id0=sid * num_rx,
                                  want_symbols=True)
        _, idx_rx = eng.demap_count(sym, want_idx=True)
        bits_rx = be.to_numpy(eng.indices_to_bits(idx_rx, nbits).reshape(-1), np.int64)
        bit_errors = int(errors.item())
        pmi0 = int(pmi.item())
        # the channel is static over the call, so every symbol reports the same PMI (reference :2366-2369)
        pmi_history = [pmi0] * S
        csi.pmi_history.extend(pmi_history)
        csi.total_feedbacks += S
        # update_mode='static': the reference's precoder object never gets a W, so its gain reads 0.0 (:190-191)
        avg_gain = float(gain.item()) if adaptive else 0.0
        results = {
            'transmitted_bits': int(nbits), 'received_bits': int(nbits), 'bits_received_array': bits_rx,
            'bit_errors': bit_errors, 'errors': bit_errors, 'ber': float(bit_errors / nbits), 'snr_db': float(snr_db),
            'num_tx': num_tx, 'num_rx': num_rx, 'mode': 'Beamforming', 'codebook_type': codebook_type,
            'beamforming_gain_db': avg_gain, 'channel_matrix': H, 'pmi_history': pmi_history,
            'unique_pmis': len(set(pmi_history)), 'velocity_kmh': velocity_kmh,
            'precoder': be.to_numpy(W[0]).astype(complex).reshape(-1, 1), 'symbols_rx': be.to_numpy(sym.reshape(-1)),
        }
        if adaptive:
            results['update_period'] = int(AdaptiveBeamforming(num_tx, velocity_kmh, 2.0).update_period)
        self.last_results = results
        return results

    # ------------------------------------------------------------------ sweeps
    def run_ber_sweep(self, num_bits: int, snr_range, num_trials: int = 1,
                      progress_callback: Optional[callable] = None) -> Dict:
        """reference :1795-1846 (one random bit vector, SISO).  With rng='philox' on the LTE chain (OFDM or
        SC-FDM) the whole SNR x trials grid is one batch of streams through LinkEngine.siso_ber (same bits, independent
        channel / noise draws per point and trial); with rng='numpy' every point replays the reference's draws
        call by call."""
        bits = np.random.randint(0, 2, num_bits)
        snr_values = np.atleast_1d(snr_range)
        if self._draws.kind == 'philox' and self.mode == 'lte' and self.enable_equalization and num_bits > 0:
            eng = self._engine()
            b_t = be.as_bits_tensor(bits)
            S = eng.symbols_for_bits(num_bits)
            idx1 = eng.bits_to_indices(b_t, num_bits, S)
            if self.enable_sc_fdm:
                pre = eng.dft_m(eng.qam_map(idx1).view(S, eng.Nd), eng.Nd).view(1, -1)
                _, _, stats = eng.modulate(S, symbols=pre, want_stats=True)
            else:
                _, _, stats = eng.modulate(S, idx=idx1, want_stats=True)
            papr_db = float(self.tx._papr_from_stats(be.to_numpy(stats)[0], S * eng.L)['papr_db'])
            n_snr = len(snr_values)
            B = n_snr * num_trials
            rows = torch.tensor([10 ** (float(s) / 10) for s in snr_values], dtype=torch.float32,
                                device=idx1.device).repeat(num_trials).contiguous()
            err = eng.siso_ber(self.channels[0]._chan_desc(), rows, S, self._draws.seed,
                               stream_id0=self._draws.next_stream() * B, idx=idx1.expand(B, -1).contiguous(),
                               nbits=num_bits, sc_fdm=self.enable_sc_fdm)
            ber = be.to_numpy(err.view(num_trials, n_snr).double().mean(dim=0)) / num_bits
            if progress_callback:
                progress_callback(100, f"{B} streams in one batch")
            return {'snr_db': snr_values, 'ber_mean': ber, 'ber_values': ber.copy(),
                    'papr_values': np.full(n_snr, papr_db)}
        ber_values, papr_values = [], []
        total, cur = len(snr_values) * num_trials, 0
        for snr in snr_values:
            ber_t, papr_t = [], []
            for trial in range(num_trials):
                r = self.simulate_siso(bits, snr_db=snr)
                ber_t.append(r['ber'])
                papr_t.append(r['papr_db'])
                cur += 1
                if progress_callback:
                    progress_callback(int(cur / total * 100), f"SNR: {snr:.1f} dB - Trial {trial+1}/{num_trials}")
            ber_values.append(np.mean(ber_t))
            papr_values.append(np.mean(papr_t))
        return {'snr_db': snr_values, 'ber_mean': np.array(ber_values), 'ber_values': np.array(ber_values),
                'papr_values': np.array(papr_values)}


    def run_full_sweep(self, bits, snr_range, n_iterations: int = 1, modulations=('QPSK', '16-QAM', '64-QAM'),
                       num_rx_values=(1, 2, 4, 8), progress_callback: Optional[callable] = None, seed: int = 0) -> Dict:
        """The GUIs' modulations x num_rx x SNR x iterations sweep over one payload
        (reference SIMO/gui/main_window.py:128-273) as batched GPU launches; same result dict as the
        worker thread emits.  See lte_b200.sweep.payload_sweep."""
        from lte_b200.sweep import payload_sweep
        bits = self._check_bits(bits)
        return payload_sweep(self.config, bits, snr_range, n_iterations, modulations, num_rx_values,
                             channel_type=self.channel_type, itu_profile=self.itu_profile,
                             frequency_ghz=self.frequency_ghz, velocity_kmh=self.velocity_kmh, seed=seed,
                             progress_callback=progress_callback)


def simulate_spatial_multiplexing(bits, num_tx=4, num_rx=2, rank='adaptive', detector_type='MMSE',
                                  modulation='64-QAM', snr_db=15, config=None, channel_type='awgn',
                                  itu_profile='Pedestrian_A', velocity_kmh=3, frequency_ghz=2.0,
                                  enable_csi_feedback=True, coherence_time_symbols=None, enable_parallel=False,
                                  codebook_type='TM4', *, rng='numpy', seed=0):
    """TM4-like spatial multiplexing (reference core/ofdm_core.py:2489-2815): layer mapping, codebook
    precoding on the first ceil(Nd / rank) data bins, per-TX interleaved CRS, R x T channel,
    per-OFDM-symbol CRS estimation and MMSE / ZF / SIC / MRC detection on H_eff = H W."""
    from .codebook_lte import LTECodebook
    from .rank_adaptation import RankAdaptation
    if config is None:
        config = LTEConfig(modulation=modulation)
    bits = OFDMSimulator._check_bits(bits) if len(bits) else bits
    draws = be.make_draws(rng, seed)
    nbits = len(bits)
    eng0 = be.engine_for(config, bits_per_symbol={'QPSK': 2, '16-QAM': 4, '64-QAM': 6}[modulation])
    sets = tables.mimo_pilot_sets(num_tx, eng0.Np)
    eng = be.engine_for(config, pilot_sets=sets, bits_per_symbol=eng0.bps)
    S = int(-(-nbits // (eng.Nd * eng.bps)))
    # H_initial comes from the caller's global RNG state and is unrelated to the channel (:2574)
    if draws.kind == 'numpy':
        H_initial = (np.random.randn(num_rx, num_tx) + 1j * np.random.randn(num_rx, num_tx)) / np.sqrt(2 * num_tx)
    else:
        rs = np.random.RandomState(seed)
        H_initial = (rs.randn(num_rx, num_tx) + 1j * rs.randn(num_rx, num_tx)) / np.sqrt(2 * num_tx)
    if rank == 'adaptive' and enable_csi_feedback:
        fb = RankAdaptation(num_tx, num_rx, snr_db=snr_db).get_feedback(H_initial)
        rank_used, pmi_used, W = fb['ri'], fb['pmi'], fb['W']
    else:
        rank_used = int(rank) if rank != 'adaptive' else min(num_tx, num_rx)
        pmi_used = 0
        W = LTECodebook(num_tx, transmission_mode='TM4', rank=rank_used).get_precoder(0)
    W = np.asarray(W, dtype=complex)
    if num_rx < rank_used:
        raise ValueError(f"num_rx ({num_rx}) debe ser >= num_layers ({rank_used})")
    b_t = be.as_bits_tensor(bits)
    idx = eng.bits_to_indices(b_t, nbits, S)
    data, _ = eng.sm_precode(S, W, idx=idx)
    tx, _, _ = eng.modulate(S, symbols=data, T=num_tx, want_stats=False)          # [T, S*L]
    own_last = len(np.arange(eng.Np)[(num_tx - 1) % min(num_tx, 4)::min(num_tx, 4)])
    if draws.kind == 'numpy':
        be.reference_pilot_side_effect((num_tx - 1) % 4, own_last)
    from .channel import ChannelSimulator
    fs = config.fs if hasattr(config, 'fs') else 15.36e6
    channel_sim = ChannelSimulator(channel_type=channel_type, snr_db=snr_db, fs=fs, itu_profile=itu_profile,
                                   frequency_ghz=frequency_ghz, velocity_kmh=velocity_kmh, verbose=False)
    rx, H_channel = channel_sim._transmit_sm_device(eng, tx, num_rx, draws)
    Y = eng.rx_fft(rx, num_rx, S, nat.WINDOW_FULL)
    # CRS estimate on every OFDM symbol (:2743-2752): rows = R*S single-symbol streams per TX pilot set
    H = torch.stack([eng.estimate(Y.reshape(num_rx * S, 1, eng.N), num_rx * S, 1, nat.WINDOW_FULL, pilot_set=t)
                     .reshape(num_rx, S, eng.N) for t in range(num_tx)])            # [T, R, S, N]
    if draws.kind == 'numpy':
        be.reference_pilot_side_effect((num_tx - 1) % 4, own_last)
    sym = eng.mimo_detect(Y, H, W, 10 ** (-snr_db / 10), detector_type, 1, num_rx, S, nat.WINDOW_FULL)
    errors, idx_rx = eng.demap_count(sym, idx_tx=idx, nbits=nbits, want_idx=True)
    bits_rx = be.to_numpy(eng.indices_to_bits(idx_rx, nbits).reshape(-1), np.int64)
    bit_errors = int(errors.item())
    return {
        'transmitted_bits': int(nbits), 'received_bits': int(nbits), 'bits_received_array': bits_rx,
        'bit_errors': bit_errors, 'errors': bit_errors, 'ber': float(bit_errors / nbits) if nbits else 0.0,
        'snr_db': float(snr_db), 'num_tx': num_tx, 'num_rx': num_rx, 'rank': rank_used,
        'detector_type': detector_type, 'mode': 'Spatial Multiplexing TM4', 'codebook_type': codebook_type,
        'channel_matrix': H_channel, 'precoder_matrix': W, 'pmi_used': pmi_used, 'velocity_kmh': velocity_kmh,
        'modulation': modulation, 'symbols_rx': be.to_numpy(sym.reshape(-1)),
    }
