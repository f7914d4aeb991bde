"""Host-side sharding / reduction logic of the multi-GPU sweep on CPU with the gloo backend,
world_size 2 (the GPU path itself is covered by tests/test_gpu_sweep.py)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_sweep():
    """sweep.py has no CUDA dependency; import it without triggering the package's native loader."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('lte_sweep', os.path.join(ROOT, 'ofdm-lte_b200', 'lte_b200', 'sweep.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def fake_errors(trial_lo, n, n_snr):
    """Deterministic per-stream 'error count' that depends only on the global stream id."""
    sid = torch.arange(trial_lo * n_snr, (trial_lo + n) * n_snr, dtype=torch.int64)
    return (sid * 2654435761 % 97)


def _worker(rank, world, port, n_trials, n_snr, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sweep = _load_sweep()
    r = sweep.run_sweep(lambda lo, n: fake_errors(lo, n, n_snr), n_snr, n_trials, bits_per_stream=100,
                        batch_trials=7, rank=rank, world=world)
    if rank == 0:
        torch.save(r, out)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    sweep = _load_sweep()
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [sweep.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sweep_equals_single_rank(tmp_path):
    n_trials, n_snr = 53, 5
    sweep = _load_sweep()
    single = sweep.run_sweep(lambda lo, n: fake_errors(lo, n, n_snr), n_snr, n_trials, 100, batch_trials=16)
    out = str(tmp_path / 'r.pt')
    mp.spawn(_worker, args=(2, _free_port(), n_trials, n_snr, out), nprocs=2, join=True)
    multi = torch.load(out)
    assert torch.equal(multi['errors'], single['errors'])
    assert torch.equal(multi['bits'], single['bits'])
    assert torch.equal(single['bits'], torch.full((n_snr,), n_trials * 100, dtype=torch.int64))


class _FakePaprEngine:
    """CPU stand-in with the three engine calls papr_sweep makes; the 'PAPR' of a symbol is a
    deterministic function of its global stream id so that sharding can be checked exactly."""
    device = torch.device('cpu')
    Nd = 4

    def random_indices(self, B, S, seed, stream_id0=0, out=None):
        return (torch.arange(stream_id0, stream_id0 + B, dtype=torch.int64)[:, None] * S +
                torch.arange(S, dtype=torch.int64)[None, :])

    def modulate_papr(self, S, idx=None, symbols=None, T=1, write_tx=False, hist=None, hist_lo=0.0, hist_step=0.1,
                      want_db=True, want_peak_mean=False):
        db = ((idx * 2654435761 % 1201).float() / 100.0)
        b = torch.clamp(torch.floor((db - hist_lo) / hist_step).long(), 0, hist.numel() - 1)
        hist += torch.bincount(b.reshape(-1), minlength=hist.numel())
        return db, None, None


def _papr_worker(rank, world, port, n_streams, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from lte_b200_papr_host import papr_sweep
    r = papr_sweep(_FakePaprEngine(), n_streams, symbols_per_stream=3, batch_streams=5, hist_step=0.25, hist_bins=50,
                   rank=rank, world=world)
    if rank == 0:
        torch.save(r, out)
    dist.destroy_process_group()


def test_two_rank_papr_histogram_equals_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from lte_b200_papr_host import papr_sweep
    single = papr_sweep(_FakePaprEngine(), 41, symbols_per_stream=3, batch_streams=16, hist_step=0.25, hist_bins=50)
    out = str(tmp_path / 'p.pt')
    mp.spawn(_papr_worker, args=(2, _free_port(), 41, out), nprocs=2, join=True)
    multi = torch.load(out, weights_only=False)
    assert (multi['hist'] == single['hist']).all() and multi['count'] == single['count'] == 41 * 3
    assert abs(multi['mean_db'] - single['mean_db']) < 1e-9 and multi['max_db'] == single['max_db']
    assert (multi['ccdf'] == single['ccdf']).all()


class _FakeBfEngine:
    """CPU stand-in with the four engine calls beamforming_sweep makes; every quantity is a deterministic
    function of the global stream id, so the gloo reduction of errors / gain / PMI histogram can be checked
    exactly against a single-rank run."""
    device = torch.device('cpu')
    Nd, bps = 5, 2

    def random_indices(self, B, S, seed, stream_id0=0, out=None):
        return torch.arange(stream_id0, stream_id0 + B, dtype=torch.int64)[:, None].repeat(1, S * self.Nd)

    def random_channel(self, B, R, T, seed, stream_id0=0):
        return torch.arange(stream_id0, stream_id0 + B, dtype=torch.float64)

    def bf_weights(self, h, codebook, mode='MRT'):
        pmi = (h.long() * 7) % len(codebook)
        return h, h, pmi.to(torch.int32), (h % 5).float() * 0.5          # gains are multiples of 0.5: exact sums

    def bf_link(self, idx, h, W, heff, noise_std, S, seed=0, row_id0=0, **kw):
        return (idx[:, 0] * 2654435761 % 11), None


def _bf_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sweep = _load_sweep()
    r = sweep.beamforming_sweep(_FakeBfEngine(), [None] * 4, [0.0, 5.0, 10.0], 37, 4, 2, mode='CODEBOOK',
                                symbols_per_stream=2, batch_trials=6, rank=rank, world=world)
    if rank == 0:
        torch.save(r, out)
    dist.destroy_process_group()


def test_two_rank_beamforming_sweep_equals_single_rank(tmp_path):
    sweep = _load_sweep()
    single = sweep.beamforming_sweep(_FakeBfEngine(), [None] * 4, [0.0, 5.0, 10.0], 37, 4, 2, mode='CODEBOOK',
                                     symbols_per_stream=2, batch_trials=16)
    out = str(tmp_path / 'bf.pt')
    mp.spawn(_bf_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    multi = torch.load(out, weights_only=False)
    assert torch.equal(multi['errors'], single['errors']) and torch.equal(multi['bits'], single['bits'])
    assert torch.equal(multi['pmi_hist'], single['pmi_hist']) and int(single['pmi_hist'].sum()) == 37 * 3
    assert multi['mean_gain_db'] == single['mean_gain_db']


class _FakeCodedEngine:
    device = torch.device('cpu')
    Nd = 8

    def random_indices(self, B, S, seed, stream_id0=0, out=None):
        return torch.arange(stream_id0, stream_id0 + B, dtype=torch.int64)[:, None].repeat(1, S * self.Nd)

    def siso_coded_ber(self, bits, chan, snr_db_rows, seed, stream_id0=0, iterations=8):
        sid = torch.arange(stream_id0, stream_id0 + bits.shape[0], dtype=torch.int64)
        err = sid * 40503 % 13
        return err, (err == 0).to(torch.int32)


def _coded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sweep = _load_sweep()
    r = sweep.coded_sweep(_FakeCodedEngine(), None, [1.0, 4.0], 29, tb_bits=20, batch_trials=4, rank=rank, world=world)
    if rank == 0:
        torch.save(r, out)
    dist.destroy_process_group()


def test_two_rank_coded_sweep_equals_single_rank(tmp_path):
    sweep = _load_sweep()
    single = sweep.coded_sweep(_FakeCodedEngine(), None, [1.0, 4.0], 29, tb_bits=20, batch_trials=29)
    out = str(tmp_path / 'c.pt')
    mp.spawn(_coded_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    multi = torch.load(out, weights_only=False)
    assert torch.equal(multi['errors'], single['errors']) and torch.equal(multi['block_errors'], single['block_errors'])
    assert torch.equal(multi['bler'], single['bler']) and torch.equal(single['bits'], torch.tensor([29 * 20] * 2))


class _FakeScfdmEngine:
    """CPU stand-in for scfdm_sweep: errors and the PAPR bin of a stream's symbols depend only on the global
    stream id, so the gloo reduction of counters + histogram can be checked exactly."""
    device = torch.device('cpu')
    Nd, bps = 6, 4

    def siso_ber(self, chan, snr_lin_rows, S, seed, stream_id0=0, sc_fdm=False, papr_hist=None, papr_lo=0.0,
                 papr_step=0.1, **kw):
        assert sc_fdm
        sid = torch.arange(stream_id0, stream_id0 + snr_lin_rows.shape[0], dtype=torch.int64)
        db = ((sid[:, None] * S + torch.arange(S)[None, :]) * 2654435761 % 701).float() / 100.0
        b = torch.clamp(torch.floor((db - papr_lo) / papr_step).long(), 0, papr_hist.numel() - 1)
        papr_hist += torch.bincount(b.reshape(-1), minlength=papr_hist.numel())
        return sid * 40503 % 17


def _scfdm_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sweep = _load_sweep()
    r = sweep.scfdm_sweep(_FakeScfdmEngine(), None, [0.0, 6.0, 12.0], 31, symbols_per_stream=3, batch_trials=5,
                          papr_bins=40, papr_step=0.2, rank=rank, world=world)
    if rank == 0:
        torch.save(r, out)
    dist.destroy_process_group()


def test_two_rank_scfdm_sweep_equals_single_rank(tmp_path):
    sweep = _load_sweep()
    single = sweep.scfdm_sweep(_FakeScfdmEngine(), None, [0.0, 6.0, 12.0], 31, symbols_per_stream=3, batch_trials=31,
                               papr_bins=40, papr_step=0.2)
    out = str(tmp_path / 's.pt')
    mp.spawn(_scfdm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    multi = torch.load(out, weights_only=False)
    assert torch.equal(multi['errors'], single['errors']) and torch.equal(multi['bits'], single['bits'])
    assert torch.equal(multi['papr_hist'], single['papr_hist']) and int(single['papr_hist'].sum()) == 31 * 3 * 3
    assert torch.equal(multi['papr_ccdf'], single['papr_ccdf'])
