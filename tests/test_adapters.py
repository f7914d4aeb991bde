"""Payload adapters: ImageProcessor bit I/O against the reference's behaviour
(utils/image_processing.py) on CPU, and the GUIs' full diversity sweep on the GPU."""
import importlib.util
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _image_processor():
    """utils/image_processing.py is host-only: load it without the package's native loader."""
    spec = importlib.util.spec_from_file_location(
        'lte_image_processing', os.path.join(ROOT, 'ofdm-lte_b200', 'utils', 'image_processing.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.ImageProcessor


def test_image_round_trip_and_psnr(tmp_path):
    from PIL import Image
    IP = _image_processor()
    rs = np.random.RandomState(0)
    arr = rs.randint(0, 256, (9, 7, 3)).astype(np.uint8)
    path = str(tmp_path / 'a.png')
    Image.fromarray(arr, 'RGB').save(path)
    bits, meta = IP.image_to_bits(path)
    assert meta == {'height': 9, 'width': 7, 'channels': 3, 'dtype': 'uint8'}
    assert np.array_equal(bits, np.unpackbits(arr.flatten()))            # MSB-first, row-major RGB
    back = IP.bits_to_image(bits, meta)
    assert np.array_equal(np.array(back), arr) and IP.calculate_psnr(Image.open(path), back) == float('inf')
    # truncated streams are zero-padded, long ones cut (reference :66-72)
    short = np.array(IP.bits_to_image(bits[:100], meta)).flatten()
    assert np.array_equal(np.unpackbits(short)[:100], bits[:100]) and not np.unpackbits(short)[100:].any()
    assert np.array_equal(np.array(IP.bits_to_image(np.concatenate([bits, bits]), meta)), arr)
    # grayscale input is converted to RGB
    g = str(tmp_path / 'g.png')
    Image.fromarray(arr[:, :, 0], 'L').save(g)
    gb, gm = IP.image_to_bits(g)
    assert gm['channels'] == 3 and len(gb) == 9 * 7 * 3 * 8
    # PSNR from bits: flip the MSB of the first byte
    b2 = bits.copy()
    b2[0] ^= 1
    mse = 128.0 ** 2 / (len(bits) // 8)
    assert abs(IP.calculate_psnr_bits(bits, b2) - 20 * np.log10(255.0 / np.sqrt(mse))) < 1e-9
    assert IP.calculate_psnr_bits(bits, bits[:-3]) == float('inf')       # compared over the common prefix
    noisy = np.array(back).astype(int)
    noisy[0, 0, 0] ^= 0x10
    want = 20 * np.log10(255.0 / np.sqrt(16.0 ** 2 / arr.size))
    assert abs(IP.calculate_psnr(arr, noisy.astype(np.uint8)) - want) < 1e-9


@pytest.mark.gpu
def test_payload_sweep_matches_direct_engine_calls_and_shards():
    from config import LTEConfig
    from lte_b200 import LinkEngine, chan_for
    from lte_b200.sweep import payload_sweep
    cfg = LTEConfig(1.25, 15.0, 'QPSK')
    bits = np.random.RandomState(4).randint(0, 2, 62 * 4 * 15 - 5)          # ragged, crosses a slot boundary
    snrs = [2.0, 10.0]
    kw = dict(modulations=('16-QAM',), num_rx_values=(1, 2), channel_type='rayleigh_mp', itu_profile='Vehicular_A',
              velocity_kmh=30.0, seed=3)
    one = payload_sweep(cfg, bits, snrs, 6, **kw)
    assert one['mode'] == 'sweep_full' and one['modulations'] == ['16-QAM'] and one['num_rx_values'] == [1, 2]
    assert cfg.modulation == 'QPSK'                                        # the caller's config is untouched
    parts = [payload_sweep(cfg, bits, snrs, 6, rank=r, world=2, max_batch_bytes=1 << 20, **kw) for r in range(2)]
    for key in ('1RX', '2RX'):
        a = one['data']['16-QAM'][key]
        assert np.array_equal(a['snr_values'], snrs) and a['num_rx'] == int(key[0])
        assert np.array_equal(parts[0]['data']['16-QAM'][key]['errors'] + parts[1]['data']['16-QAM'][key]['errors'],
                              a['errors'])
        assert np.array_equal(a['bits'], [6 * len(bits)] * 2)
        assert np.allclose(a['ber_values'], a['errors'] / a['bits'])
    # the same streams through the engine directly
    cfg16 = LTEConfig(1.25, 15.0, '16-QAM')
    eng = LinkEngine.from_config(cfg16)
    chan = chan_for('rayleigh_mp', cfg16.fs, 'Vehicular_A', 2.0, 30.0)
    S = eng.symbols_for_bits(len(bits))
    idx = eng.bits_to_indices(torch.from_numpy(bits.astype(np.uint8)).cuda()[None], len(bits), S).expand(12, -1).contiguous()
    snr_lin = torch.tensor([10 ** (s / 10) for s in snrs], dtype=torch.float32, device='cuda').repeat(6)
    e1 = eng.siso_ber(chan, snr_lin.contiguous(), S, 3, 0, idx=idx, nbits=len(bits))
    assert np.array_equal(e1.view(6, 2).sum(0).cpu().numpy(), one['data']['16-QAM']['1RX']['errors'])
    ws = eng.workspace(12, S, 2, fading=True, fused=True)
    e2 = eng.simo_ber(ws, chan, snr_lin.repeat_interleave(2).contiguous(), 3, 0, idx=idx, nbits=len(bits),
                      noise_domain=2, fused=True)
    assert np.array_equal(e2.view(6, 2).sum(0).cpu().numpy(), one['data']['16-QAM']['2RX']['errors'])


@pytest.mark.gpu
def test_full_sweep_api_tracks_the_per_call_api():
    """Batched sweep vs the loop the GUI runs over simulate_siso / simulate_simo (independent draws:
    statistical agreement), plus the diversity ordering the GUI plots."""
    from config import LTEConfig
    from core.ofdm_core import OFDMSimulator
    cfg = LTEConfig(1.25, 15.0, 'QPSK')
    sim = OFDMSimulator(cfg, channel_type='rayleigh_mp', itu_profile='Pedestrian_A', velocity_kmh=3.0, rng='philox',
                        seed=1)
    bits = np.random.RandomState(2).randint(0, 2, 62 * 2 * 28)
    seen = []
    r = sim.run_full_sweep(bits, [0.0, 6.0, 12.0], n_iterations=150, modulations=('QPSK',), num_rx_values=(1, 2, 4),
                           progress_callback=lambda p, m: seen.append((p, m)))
    d = r['data']['QPSK']
    b1, b2, b4 = d['1RX']['ber_values'], d['2RX']['ber_values'], d['4RX']['ber_values']
    assert np.all(np.diff(b1) < 0) and np.all(b2 < b1) and np.all(b4 < b2)
    assert seen[-1][0] == 95 and len(seen) == 3
    loop1 = np.mean([sim.simulate_siso(bits, 6.0)['ber'] for _ in range(150)])
    loop2 = np.mean([sim.simulate_simo(bits, 6.0, num_rx=2)['ber'] for _ in range(150)])
    assert abs(loop1 - b1[1]) / b1[1] < 0.25 and abs(loop2 - b2[1]) / b2[1] < 0.3
    with pytest.raises(ValueError):
        sim.run_full_sweep([], [0.0])


@pytest.mark.gpu
def test_run_ber_sweep_batched_matches_the_per_call_loop_statistically():
    """OFDMSimulator / OFDMModule.run_ber_sweep with rng='philox': one batch for the whole SNR x trials grid."""
    from config import LTEConfig
    from core.ofdm_core import OFDMSimulator
    from ofdm_module import OFDMModule
    cfg = LTEConfig(1.25, 15.0, 'QPSK')
    sim = OFDMSimulator(cfg, channel_type='rayleigh_mp', itu_profile='Pedestrian_A', velocity_kmh=3.0, rng='philox', seed=4)
    np.random.seed(1)
    seen = []
    r = sim.run_ber_sweep(62 * 2 * 14, [0.0, 8.0, 16.0], num_trials=200, progress_callback=lambda p, m: seen.append(p))
    assert list(r['snr_db']) == [0.0, 8.0, 16.0] and r['ber_mean'].shape == (3,) and seen == [100]
    assert np.all(np.diff(r['ber_mean']) < 0) and np.array_equal(r['ber_mean'], r['ber_values'])
    assert np.allclose(r['papr_values'], r['papr_values'][0]) and 3.0 < r['papr_values'][0] < 14.0
    bits = np.random.RandomState(3).randint(0, 2, 62 * 2 * 14)
    loop = np.mean([sim.simulate_siso(bits, 8.0)['ber'] for _ in range(200)])
    assert abs(loop - r['ber_mean'][1]) / r['ber_mean'][1] < 0.25
    # the replaying mode still goes call by call (and OFDMModule delegates)
    m = OFDMModule(cfg, channel_type='awgn')
    np.random.seed(2)
    r2 = m.run_ber_sweep(300, [2.0, 6.0], num_trials=2)
    assert r2['ber_mean'].shape == (2,) and r2['ber_mean'][0] >= r2['ber_mean'][1]
