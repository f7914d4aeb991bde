#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (Darioxavierl/OFDM-LTE) from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference draws its channel phases and noise from NumPy's legacy global RNG,
which every ResourceMapper.map_symbols() call re-seeds with cell_id
(core/resource_mapper.py:148).  The draws are therefore a deterministic function
of the configuration; the tests regenerate them with
oracle.lte_oracle.ReferenceDraws instead of storing megabytes of noise.

Large complex arrays are stored as complex64 (6e-8 relative rounding, well inside
the 1e-5 parity budget); small ones as complex128.
"""
import contextlib
import io
import os
import sys

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from config import LTEConfig  # noqa: E402  (reference)
from core.ofdm_core import OFDMSimulator  # noqa: E402  (reference)
from core.resource_mapper import LTEResourceGrid, PilotPattern  # noqa: E402
from core.modulator import QAMModulator  # noqa: E402

sys.path.insert(0, HERE)
from cases import SISO_CASES, SIMO_CASES, SFBC_CASES, SM_CASES, PAPR_CASES, BF_CASES, CODED_CASES, BIG_RX_STRIDE  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def make_bits(seed, n):
    return np.random.RandomState(seed).randint(0, 2, n).astype(np.uint8)


def nd_of(cfg):
    return len(LTEResourceGrid(cfg.N, cfg.Nc).get_data_indices())


def tables():
    out = {}
    for bw in (1.25, 2.5, 5.0, 10.0, 15.0, 20.0):
        for cp in ('normal', 'extended'):
            c = LTEConfig(bw, 15.0, 'QPSK', cp)
            out[f'num_{bw}_{cp}'] = np.array([c.N, c.Nc, c.cp_length, c.fs], dtype=np.float64)
        g = LTEResourceGrid(c.N, c.Nc)
        out[f'data_idx_{bw}'] = g.get_data_indices().astype(np.int32)
        out[f'pilot_idx_{bw}'] = g.get_pilot_indices().astype(np.int32)
    c = LTEConfig(3.0, 15.0, 'QPSK')          # non-profile bandwidth, config.py:108-111
    out['num_3.0_normal'] = np.array([c.N, c.Nc, c.cp_length, c.fs], dtype=np.float64)
    c = LTEConfig(5.0, 7.5, 'QPSK', 'extended')
    out['num_5.0_7.5_extended'] = np.array([c.N, c.Nc, c.cp_length, c.fs], dtype=np.float64)
    for cell in range(4):
        out[f'pilots_cell{cell}'] = PilotPattern(cell).generate_pilots(200)
    for mod in ('QPSK', '16-QAM', '64-QAM'):
        q = QAMModulator(mod)
        out[f'const_{mod}'] = q.get_constellation()
        b = int(np.log2(len(q.constellation)))
        bits = make_bits(11, b * 257 - 1)                     # exercises the zero padding
        out[f'mapbits_{mod}'] = bits
        out[f'mapsyms_{mod}'] = q.bits_to_symbols(bits)
        rs = np.random.RandomState(5)
        y = (rs.standard_normal(600) + 1j * rs.standard_normal(600)) * 0.8
        # exact ties and exact constellation points, plus the all-zero symbol
        lv = np.unique(q.constellation.real)
        mids = (lv[:-1] + lv[1:]) / 2 if len(lv) > 1 else np.array([0.0])
        ties = np.array([a + 1j * bb for a in np.concatenate([mids, [0.0]])
                         for bb in np.concatenate([lv, mids, [0.0]])])
        y = np.concatenate([y, ties, q.constellation, [0j]])
        out[f'demapsyms_{mod}'] = y
        out[f'demapbits_{mod}'] = q.symbols_to_bits(y).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, 'tables.npz'), **out)


def n_bits_of(case, cfg):
    n_per = cfg.Nc if case.get('mode') == 'simple' else nd_of(cfg)
    return n_per * cfg.bits_per_symbol * case['nsym'] - case.get('drop_bits', 0)


def siso_case(case):
    cfg = LTEConfig(case['bw'], 15.0, case['mod'], case.get('cp_type', 'normal'))
    sim = quiet(OFDMSimulator, cfg, channel_type=case['ch'], itu_profile=case['prof'], frequency_ghz=2.0,
                velocity_kmh=case['v'], mode=case.get('mode', 'lte'),
                enable_sc_fdm=case.get('sc_fdm', False), enable_equalization=case.get('equalize', True))
    bits = make_bits(case['seed'], n_bits_of(case, cfg))
    ctype = np.complex64 if case.get('big') else np.complex128
    out = dict(bits=np.packbits(bits), nbits=len(bits))
    for snr in case['snrs']:
        if 'global_seed' in case:      # mode='simple' does not re-seed the global RNG itself
            np.random.seed(case['global_seed'])
        r = quiet(sim.simulate_siso, bits, snr_db=snr)
        out[f'errors_{snr}'] = r['errors']
        out[f'bits_rx_{snr}'] = np.packbits(r['bits_received_array'].astype(np.uint8))
        out[f'papr_db_{snr}'] = r['papr_db']
        if snr == case['full_snr']:
            out['signal_tx'] = r['signal_tx'].astype(ctype)
            out['signal_rx'] = r['signal_rx'].astype(ctype)
            out['symbols_rx'] = r['symbols_rx'].astype(ctype)
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print(case['name'], {s: int(out[f'errors_{s}']) for s in case['snrs']})


def simo_case(case):
    cfg = LTEConfig(case['bw'], 15.0, case['mod'], 'normal')
    sim = quiet(OFDMSimulator, cfg, channel_type=case['ch'], itu_profile=case['prof'], frequency_ghz=2.0,
                velocity_kmh=case['v'])
    bits = make_bits(case['seed'], n_bits_of(case, cfg))
    big = case.get('big')
    ctype = np.complex64 if big else np.complex128
    out = dict(bits=np.packbits(bits), nbits=len(bits))
    for snr in case['snrs']:
        r = quiet(sim.simulate_simo, bits, snr_db=snr, num_rx=case['R'], parallel=False)
        out[f'errors_{snr}'] = r['errors']
        out[f'bits_rx_{snr}'] = np.packbits(r['bits_received_array'].astype(np.uint8))
        if snr == case['full_snr']:
            out['signal_tx'] = r['signal_tx'].astype(ctype)
            rx = np.stack(r['signal_rx_list'])
            out['signal_rx'] = (rx[:, ::BIG_RX_STRIDE] if big else rx).astype(ctype)
            out['symbols_combined'] = r['symbols_rx_combined'].astype(ctype)
            H = np.array(r['channel_estimates_per_antenna'])               # [R][S][N]
            out['H'] = H[:, ::14, :].astype(ctype)                         # one per 14-symbol slot
            out['papr_db'] = r['papr_db']
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print(case['name'], {s: int(out[f'errors_{s}']) for s in case['snrs']})


def install_sfbc_shim():
    """simulate_miso / simulate_mimo raise at HEAD: estimate_channel_periodic unpacks three values
    from estimate_channel_from_grid, which returns two (core/mimo_channel_estimator_periodic.py:219-222
    vs :185).  The runtime patch below is the intended behaviour -- estimate on the
    first symbol of every 14-symbol slot, hold for the slot -- and touches nothing else."""
    from core.mimo_channel_estimator_periodic import MIMOChannelEstimatorPeriodic

    def estimate_channel_periodic(self, grids):
        H0s, H1s = [], []
        for s0 in range(0, len(grids), self.slot_size):
            H, _ = self.estimate_channel_from_grid(grids[s0])
            for _ in range(min(self.slot_size, len(grids) - s0)):
                H0s.append(H[0, 0, :])
                H1s.append(H[0, 1, :])
        return H0s, H1s, 0.0

    MIMOChannelEstimatorPeriodic.estimate_channel_periodic = estimate_channel_periodic


def sfbc_case(case):
    cfg = LTEConfig(case['bw'], 15.0, case['mod'], 'normal')
    sim = quiet(OFDMSimulator, cfg, channel_type=case['ch'], itu_profile=case['prof'], frequency_ghz=2.0,
                velocity_kmh=case['v'])
    nd2 = nd_of(cfg) - nd_of(cfg) % 2
    bits = make_bits(case['seed'], nd2 * cfg.bits_per_symbol * case['nsym'] - case.get('drop_bits', 0))
    out = dict(bits=np.packbits(bits), nbits=len(bits))
    for snr in case['snrs']:
        if case['R'] == 1:
            r = quiet(sim.simulate_miso, bits, snr_db=snr)
        else:
            r = quiet(sim.simulate_mimo, bits, snr_db=snr, num_rx=case['R'])
        out[f'errors_{snr}'] = r['errors']
        out[f'bits_rx_{snr}'] = np.packbits(r['bits_received_array'].astype(np.uint8))
        out[f'channel_matrix_{snr}'] = r['channel_matrix']
        out[f'papr_{snr}'] = np.array([r['papr_db_tx0'], r['papr_db_tx1'], r['papr_db'], r['papr_linear']])
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print(case['name'], {s: int(out[f'errors_{s}']) for s in case['snrs']})


def sm_case(case):
    from core.ofdm_core import simulate_spatial_multiplexing
    cfg = LTEConfig(case['bw'], 15.0, case['mod'], 'normal')
    bits = make_bits(case['gseed'], nd_of(cfg) * cfg.bits_per_symbol * case['nsym'] - case.get('drop_bits', 0))
    out = dict(bits=np.packbits(bits), nbits=len(bits))
    for snr in case['snrs']:
        np.random.seed(case['gseed'])          # H_initial is drawn from the caller's global RNG state
        r = quiet(simulate_spatial_multiplexing, bits, num_tx=case['T'], num_rx=case['R'], rank=case['rank'],
                  detector_type=case['det'], modulation=case['mod'], snr_db=snr, config=cfg,
                  channel_type=case['ch'], itu_profile=case['prof'], velocity_kmh=case['v'], frequency_ghz=2.0)
        out[f'errors_{snr}'] = r['errors']
        out[f'bits_rx_{snr}'] = np.packbits(r['bits_received_array'].astype(np.uint8))
        out[f'rank_pmi_{snr}'] = np.array([r['rank'], r['pmi_used']])
        out[f'W_{snr}'] = np.asarray(r['precoder_matrix'], dtype=complex)
        out[f'channel_matrix_{snr}'] = r['channel_matrix']
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print(case['name'], {s: int(out[f'errors_{s}']) for s in case['snrs']})


def papr_case(case):
    """Per-symbol PAPR of the reference's OFDMSystem for OFDM and SC-FDM on the same bits."""
    from core.ofdm_system import OFDMSystem
    cfg = LTEConfig(case['bw'], 15.0, case['mod'])
    n = case['nsym'] * nd_of(cfg) * cfg.bits_per_symbol - case.get('drop_bits', 0)
    bits = make_bits(case['seed'], n)
    out = {'bits': bits}
    for tag, flag in (('ofdm', False), ('scfdm', True)):
        sysm = quiet(OFDMSystem, LTEConfig(case['bw'], 15.0, case['mod']), 'awgn', mode='lte', enable_sc_fdm=flag)
        sig = quiet(sysm.modulator.modulate_stream, bits)[0]
        a = sysm.calculate_papr_without_cp(sig)
        b = sysm.calculate_papr_per_symbol(sig)
        c = sysm.calculate_papr(sig)
        out[f'{tag}_no_cp_db'] = np.asarray(a['papr_per_symbol'], dtype=np.float64)
        out[f'{tag}_cp_db'] = np.asarray(b['papr_per_symbol'], dtype=np.float64)
        out[f'{tag}_cp_peak'] = np.asarray(b['power_peak_per_symbol'], dtype=np.float64)
        out[f'{tag}_cp_avg'] = np.asarray(b['power_avg_per_symbol'], dtype=np.float64)
        out[f'{tag}_stream'] = np.array([c['papr_db'], c['papr_linear'], c['peak_power'], c['avg_power']])
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print('wrote', case['name'])


def bf_case(case):
    """OFDMSimulator.simulate_beamforming with the caller's global RNG seeded explicitly."""
    cfg = LTEConfig(case['bw'], 15.0, case['mod'], 'normal')
    sim = quiet(OFDMSimulator, cfg, channel_type='awgn')
    bits = make_bits(case['gseed'], nd_of(cfg) * cfg.bits_per_symbol * case['nsym'] - case.get('drop_bits', 0))
    out = dict(bits=np.packbits(bits), nbits=len(bits))
    for snr in case['snrs']:
        np.random.seed(case['gseed'])
        r = quiet(sim.simulate_beamforming, bits, snr_db=snr, num_tx=case['T'], num_rx=case['R'],
                  codebook_type=case['cb'], velocity_kmh=case['v'], update_mode=case['upd'])
        out[f'errors_{snr}'] = r['errors']
        out[f'bits_rx_{snr}'] = np.packbits(r['bits_received_array'].astype(np.uint8))
        out[f'channel_matrix_{snr}'] = r['channel_matrix']
        out[f'pmi_history_{snr}'] = np.array(r['pmi_history'])
        out[f'gain_unique_{snr}'] = np.array([r['beamforming_gain_db'], r['unique_pmis']])
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print(case['name'], {s: int(out[f'errors_{s}']) for s in case['snrs']})


def coding_tables():
    """Known-answer vectors of the channel-coding building blocks (core/channel_coding/*.py) and of the
    soft demappers (core/ofdm_core.py:791-923)."""
    from core.channel_coding.crc import calculate_crc24a, calculate_crc24b
    from core.channel_coding.rate_matching import rate_dematching_turbo, rate_match_turbo
    from core.channel_coding.segmentation import segment_code_blocks
    from core.channel_coding.turbo_decoder import turbo_decode
    from core.channel_coding.turbo_encoder import QPP_INTERLEAVER_PARAMS, turbo_encode
    out = {'qpp': np.array([[k, f1, f2] for k, (f1, f2) in sorted(QPP_INTERLEAVER_PARAMS.items())], dtype=np.int32)}
    rs = np.random.RandomState(7)
    for n in (1, 24, 100, 1001):
        b = rs.randint(0, 2, n).astype(np.uint8)
        out[f'crc_in_{n}'] = b
        out[f'crc24a_{n}'] = calculate_crc24a(b)
        out[f'crc24b_{n}'] = calculate_crc24b(b)
    for B in (40, 41, 100, 6144, 6145, 13000, 20011):
        tb = rs.randint(0, 2, B).astype(np.uint8)
        blocks, meta = quiet(segment_code_blocks, tb)
        out[f'seg_in_{B}'] = tb
        out[f'seg_sizes_{B}'] = np.array(meta['block_sizes'])
        out[f'seg_out_{B}'] = np.concatenate(blocks)
    for K in (40, 104, 512, 6144):
        u = rs.randint(0, 2, K).astype(np.uint8)
        enc = turbo_encode(u)
        out[f'enc_in_{K}'], out[f'enc_out_{K}'] = u, enc
        out[f'rm_out_{K}'] = rate_match_turbo(enc, len(enc), K, rv_idx=0)
        l = rs.randn(3 * K + 12)
        out[f'dm_in_{K}'], out[f'dm_out_{K}'] = l, rate_dematching_turbo(l, K, rv_idx=0)
    for K, sigma in ((40, 0.9), (104, 1.1), (512, 1.25)):
        u = rs.randint(0, 2, K).astype(np.uint8)
        enc = turbo_encode(u)
        l = 2.0 * (1.0 - 2.0 * enc + sigma * rs.randn(len(enc))) / sigma ** 2
        out[f'dec_bits_{K}'], out[f'dec_llr_{K}'] = u, l
        out[f'dec_out_{K}'] = turbo_decode(l, K=K, num_iterations=8)
    # exact log-MAP mode (set_decoder_mode(False)): decisions and one BCJR pass with a-priori input
    from core.channel_coding import turbo_decoder as td
    quiet(td.set_decoder_mode, False)
    for K in (40, 104):
        out[f'dec_out_logmap_{K}'] = td.turbo_decode(out[f'dec_llr_{K}'], K=K, num_iterations=8)
    l = out['dec_llr_40']
    Ls, Lp = np.concatenate([l[0:120:3], l[120:123]]), np.concatenate([l[1:120:3], l[123:126]])
    La = np.concatenate([0.4 * np.cos(np.arange(40)), np.zeros(3)])
    out['bcjr_logmap_in'] = np.stack([Ls, Lp, La])
    out['bcjr_logmap_ext'] = td.LogMAPDecoder().decode(Ls, Lp, La, return_extrinsic=True)[1]
    quiet(td.set_decoder_mode, True)
    out['bcjr_maxlog_ext'] = td.LogMAPDecoder().decode(Ls, Lp, La, return_extrinsic=True)[1]
    sim = quiet(OFDMSimulator, LTEConfig(1.25, 15.0, 'QPSK'))
    y = (rs.randn(300) + 1j * rs.randn(300)) * 0.8
    nv = 0.05 + rs.rand(300)
    out['llr_sym'], out['llr_nv'] = y, nv
    out['llr_QPSK'] = sim._calculate_llrs_qpsk(y, nv)
    out['llr_16-QAM'] = sim._calculate_llrs_16qam(y, nv)
    out['llr_64-QAM'] = sim._calculate_llrs_64qam(y, nv)
    np.savez_compressed(os.path.join(HERE, 'coding_tables.npz'), **out)
    print('wrote coding_tables')


def coded_case(case):
    cfg = LTEConfig(case['bw'], 15.0, case['mod'], 'normal')
    sim = quiet(OFDMSimulator, cfg, channel_type=case['ch'], itu_profile=case['prof'], frequency_ghz=2.0,
                velocity_kmh=case['v'])
    bits = make_bits(case['seed'], case['nbits'])
    out = dict(bits=np.packbits(bits), nbits=len(bits))
    for snr in case['snrs']:
        r = quiet(sim.simulate_siso_coded, bits, snr_db=snr)
        out[f'errors_{snr}'] = r['bit_errors']
        out[f'bits_rx_{snr}'] = np.packbits(r['bits_received_array'].astype(np.uint8))
        out[f'scalars_{snr}'] = np.array([float(r['crc_pass']), r['coded_bits_length'], r['papr_db'],
                                          r['noise_var_mean']])
        out[f'symbols_rx_{snr}'] = r['symbols_rx']
        out[f'H_{snr}'] = r['H_estimate']
    out['signal_tx'] = r['signal_tx']
    np.savez_compressed(os.path.join(HERE, case['name'] + '.npz'), **out)
    print(case['name'], {s: int(out[f'errors_{s}']) for s in case['snrs']})


def main():
    if 'coded' in sys.argv[1:]:       # only the channel-coding fixtures
        coding_tables()
        for case in CODED_CASES:
            coded_case(case)
        return
    if 'bf' in sys.argv[1:]:          # only the beamforming fixtures
        for case in BF_CASES:
            bf_case(case)
        return
    tables()
    coding_tables()
    for case in CODED_CASES:
        coded_case(case)
    for case in BF_CASES:
        bf_case(case)
    for case in PAPR_CASES:
        papr_case(case)
    for case in SM_CASES:
        sm_case(case)
    install_sfbc_shim()
    for case in SFBC_CASES:
        sfbc_case(case)
    for case in SISO_CASES:
        siso_case(case)
    for case in SIMO_CASES:
        simo_case(case)


if __name__ == '__main__':
    main()
