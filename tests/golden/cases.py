"""Case table shared by make_golden.py (which runs the reference) and the tests
(which run the oracle and the CUDA path on the same inputs)."""

SISO_CASES = [
    # BASELINE.json config 1: SISO 5 MHz QPSK AWGN
    dict(name='siso_cfg1_5mhz_qpsk_awgn', bw=5.0, mod='QPSK', ch='awgn', prof='Pedestrian_A', v=0.0,
         nsym=14, snrs=[0.0, 6.0, 12.0], full_snr=6.0, seed=1),
    # config 2: SISO SC-FDM 10 MHz 16-QAM Pedestrian_A 3 km/h
    dict(name='siso_cfg2_10mhz_16qam_scfdm_peda', bw=10.0, mod='16-QAM', ch='rayleigh_mp',
         prof='Pedestrian_A', v=3.0, nsym=14, snrs=[10.0, 20.0], full_snr=20.0, seed=2, sc_fdm=True),
    # variants reachable through the same API (last paragraph)
    dict(name='siso_ext_cp_bad_urban', bw=5.0, mod='16-QAM', ch='rayleigh_mp', prof='Bad_Urban', v=60.0,
         nsym=3, snrs=[12.0], full_snr=12.0, seed=3, drop_bits=7, cp_type='extended'),
    dict(name='siso_simple_mode', bw=2.5, mod='64-QAM', ch='awgn', prof='Pedestrian_A', v=0.0,
         nsym=4, snrs=[18.0], full_snr=18.0, seed=4, drop_bits=3, mode='simple', global_seed=1234),
    dict(name='siso_no_eq_2p5mhz', bw=2.5, mod='QPSK', ch='awgn', prof='Pedestrian_A', v=0.0,
         nsym=5, snrs=[4.0], full_snr=4.0, seed=5, equalize=False),
    dict(name='siso_nonprofile_3mhz_vehb', bw=3.0, mod='64-QAM', ch='rayleigh_mp', prof='Vehicular_B',
         v=120.0, nsym=16, snrs=[25.0], full_snr=25.0, seed=6, drop_bits=11),
    dict(name='siso_15mhz_pedb_static', bw=15.0, mod='16-QAM', ch='rayleigh_mp', prof='Pedestrian_B',
         v=0.0, nsym=2, snrs=[15.0], full_snr=15.0, seed=7, big=True),
]

SIMO_CASES = [
    # small full-tensor case crossing a 14-symbol slot boundary, ragged bit count
    dict(name='simo_small_1p25mhz_veha', bw=1.25, mod='64-QAM', ch='rayleigh_mp', prof='Vehicular_A',
         v=30.0, nsym=15, R=4, snrs=[10.0, 25.0], full_snr=25.0, seed=8, drop_bits=5),
    dict(name='simo_awgn_5mhz_16qam_r2', bw=5.0, mod='16-QAM', ch='awgn', prof='Pedestrian_A', v=0.0,
         nsym=3, R=2, snrs=[8.0], full_snr=8.0, seed=9),
    dict(name='simo_10mhz_qpsk_vehb_r8', bw=10.0, mod='QPSK', ch='rayleigh_mp', prof='Vehicular_B',
         v=120.0, nsym=2, R=8, snrs=[0.0], full_snr=0.0, seed=10, big=True),
    # config 3: SIMO 1x4 MRC 20 MHz 64-QAM Vehicular_A 30 km/h, one subframe
    dict(name='simo_cfg3_20mhz_64qam_veha', bw=20.0, mod='64-QAM', ch='rayleigh_mp', prof='Vehicular_A',
         v=30.0, nsym=14, R=4, snrs=[10.0, 20.0, 30.0], full_snr=20.0, seed=12, big=True),
    # headline variant: Pedestrian_A 3 km/h ("EPA-style")
    dict(name='simo_headline_20mhz_64qam_peda', bw=20.0, mod='64-QAM', ch='rayleigh_mp',
         prof='Pedestrian_A', v=3.0, nsym=14, R=4, snrs=[0.0, 16.0, 30.0], full_snr=16.0, seed=13,
         big=True),
]

SFBC_CASES = [
    # MISO 2x1 over AWGN (fixed h = exp(j pi tx / 2) links, core/ofdm_core.py:479-486)
    dict(name='sfbc_miso_5mhz_qpsk_awgn', bw=5.0, mod='QPSK', ch='awgn', prof='Pedestrian_A', v=0.0,
         nsym=3, R=1, snrs=[4.0], full_snr=4.0, seed=21, drop_bits=1),
    # small 2x2 crossing a slot boundary
    dict(name='sfbc_2x2_1p25mhz_16qam_peda', bw=1.25, mod='16-QAM', ch='rayleigh_mp', prof='Pedestrian_A',
         v=3.0, nsym=16, R=2, snrs=[6.0, 12.0], full_snr=12.0, seed=22, drop_bits=3),
    dict(name='sfbc_2x4_2p5mhz_64qam_veha', bw=2.5, mod='64-QAM', ch='rayleigh_mp', prof='Vehicular_A',
         v=30.0, nsym=15, R=4, snrs=[18.0], full_snr=18.0, seed=23),
    # BASELINE.json config 4: 2x2 SFBC 20 MHz 16-QAM over Rayleigh multipath, one subframe
    dict(name='sfbc_cfg4_20mhz_16qam_peda', bw=20.0, mod='16-QAM', ch='rayleigh_mp', prof='Pedestrian_A',
         v=3.0, nsym=14, R=2, snrs=[10.0, 20.0], full_snr=10.0, seed=24, big=True),
]

SM_CASES = [
    # BASELINE.json config 5: 4x4 spatial multiplexing 20 MHz 64-QAM MMSE (called per OFDM symbol by the GUI)
    dict(name='sm_cfg5_20mhz_64qam_4x4_mmse_r4', bw=20.0, mod='64-QAM', T=4, R=4, rank=4, det='MMSE',
         ch='rayleigh_mp', prof='Pedestrian_A', v=3.0, nsym=1, snrs=[15.0, 25.0], gseed=77, big=True),
    dict(name='sm_1p25mhz_64qam_4x4_mmse_r4', bw=1.25, mod='64-QAM', T=4, R=4, rank=4, det='MMSE',
         ch='rayleigh_mp', prof='Pedestrian_A', v=3.0, nsym=2, snrs=[25.0], gseed=78, drop_bits=2),
    dict(name='sm_1p25mhz_16qam_4x4_zf_r2', bw=1.25, mod='16-QAM', T=4, R=4, rank=2, det='ZF',
         ch='rayleigh_mp', prof='Vehicular_A', v=30.0, nsym=3, snrs=[20.0], gseed=79, drop_bits=2),
    dict(name='sm_2p5mhz_qpsk_2x2_sic_r2_flat', bw=2.5, mod='QPSK', T=2, R=2, rank=2, det='SIC',
         ch='awgn', prof='Pedestrian_A', v=3.0, nsym=2, snrs=[3.0], gseed=80),
    dict(name='sm_1p25mhz_16qam_4x2_mrc_r1_flat', bw=1.25, mod='16-QAM', T=4, R=2, rank=1, det='MRC',
         ch='awgn', prof='Pedestrian_A', v=3.0, nsym=1, snrs=[6.0], gseed=81),
    dict(name='sm_5mhz_64qam_4x4_adaptive_mmse', bw=5.0, mod='64-QAM', T=4, R=4, rank='adaptive', det='MMSE',
         ch='rayleigh_mp', prof='Pedestrian_A', v=3.0, nsym=1, snrs=[7.0, 22.0], gseed=82),
    dict(name='sm_1p25mhz_16qam_4x4_sic_r3', bw=1.25, mod='16-QAM', T=4, R=4, rank=3, det='SIC',
         ch='rayleigh_mp', prof='Pedestrian_A', v=3.0, nsym=2, snrs=[25.0], gseed=83, drop_bits=2),
    dict(name='sm_1p25mhz_qpsk_8x4_mmse_r4_flat', bw=1.25, mod='QPSK', T=8, R=4, rank=4, det='MMSE',
         ch='awgn', prof='Pedestrian_A', v=3.0, nsym=1, snrs=[18.0], gseed=84),
]

BIG_RX_STRIDE = 8   # 'big' SIMO cases store every 8th sample of signal_rx

# per-symbol PAPR (OFDM and SC-FDM on the same bits), core/ofdm_system.py:116-229
PAPR_CASES = [
    dict(name='papr_cfg2_10mhz_16qam', bw=10.0, mod='16-QAM', nsym=14, seed=21),
    dict(name='papr_5mhz_qpsk_ragged', bw=5.0, mod='QPSK', nsym=6, seed=22, drop_bits=5),
    dict(name='papr_20mhz_64qam', bw=20.0, mod='64-QAM', nsym=3, seed=23),
]

# beamforming path, OFDMSimulator.simulate_beamforming (core/ofdm_core.py:2260-2477).
# gseed seeds the caller's global RNG (nothing on this path re-seeds it).
BF_CASES = [
    dict(name='bf_2x1_adaptive_1p25mhz_qpsk', bw=1.25, mod='QPSK', T=2, R=1, cb='TM6', upd='adaptive', v=3.0,
         nsym=3, snrs=[0.0, 10.0], gseed=91, drop_bits=1),
    dict(name='bf_4x2_static_tm6_2p5mhz_16qam', bw=2.5, mod='16-QAM', T=4, R=2, cb='TM6', upd='static', v=3.0,
         nsym=2, snrs=[8.0], gseed=92),
    dict(name='bf_8x1_adaptive_5mhz_64qam', bw=5.0, mod='64-QAM', T=8, R=1, cb='TM6', upd='adaptive', v=30.0,
         nsym=2, snrs=[6.0, 15.0], gseed=93),
    dict(name='bf_4x4_static_tm4_1p25mhz_64qam', bw=1.25, mod='64-QAM', T=4, R=4, cb='TM4', upd='static', v=3.0,
         nsym=4, snrs=[12.0], gseed=94, drop_bits=5),
    # the configuration of results/beamforming/resultados_comparacion.txt: 10 MHz, 64-QAM, 15 dB, 3 km/h
    dict(name='bf_2x2_adaptive_10mhz_64qam', bw=10.0, mod='64-QAM', T=2, R=2, cb='TM6', upd='adaptive', v=3.0,
         nsym=2, snrs=[15.0], gseed=95),
]

# coded chain, OFDMSimulator.simulate_siso_coded (core/ofdm_core.py:925-1338)
CODED_CASES = [
    dict(name='coded_1p25mhz_qpsk_awgn', bw=1.25, mod='QPSK', ch='awgn', prof='Pedestrian_A', v=0.0,
         nbits=200, snrs=[-1.0, 3.0], seed=61),
    dict(name='coded_1p25mhz_16qam_peda', bw=1.25, mod='16-QAM', ch='rayleigh_mp', prof='Pedestrian_A', v=3.0,
         nbits=333, snrs=[8.0, 16.0], seed=62),
    dict(name='coded_2p5mhz_64qam_veha', bw=2.5, mod='64-QAM', ch='rayleigh_mp', prof='Vehicular_A', v=30.0,
         nbits=500, snrs=[16.0, 26.0], seed=63),
    # two code blocks (B = 6224 > 6144: K- / K+ split, CRC-24B per block)
    dict(name='coded_5mhz_qpsk_awgn_2blocks', bw=5.0, mod='QPSK', ch='awgn', prof='Pedestrian_A', v=0.0,
         nbits=6200, snrs=[1.0, 5.0], seed=64),
]
