"""Compatibility shim for the reference's thread-per-antenna helper (core/parallel_processing.py:23-223).

In this engine the antennas are a batch dimension of every kernel launch and trials shard over GPUs
(lte_b200/sweep.py), so there is nothing left to parallelise on the host: the methods keep their names and
return shapes and simply map the caller's function over the antenna list in order.  (The reference's pool
gives no speed-up either -- the work holds the GIL -- and its threaded path is racy)"""
import time


class MIMOParallelProcessor:
    def __init__(self, num_antennas, enable_parallel=True, threshold=4, max_workers=None):
        self.num_antennas = num_antennas
        self.threshold = threshold
        self.enable_parallel = enable_parallel and (num_antennas >= threshold)
        self.max_workers = min(num_antennas, 8) if max_workers is None else min(max_workers, num_antennas)

    @staticmethod
    def _map(func, items, *extra):
        return [func(x, *extra) for x in items]

    def parallel_ofdm_modulate(self, modulator_func, freq_symbols_per_ant):
        return self._map(modulator_func, freq_symbols_per_ant)

    def parallel_ofdm_demodulate(self, demodulator_func, rx_signals_per_ant):
        return self._map(demodulator_func, rx_signals_per_ant)

    def parallel_channel_estimation(self, estimator_func, received_per_ant, tx_pilots=None):
        if tx_pilots is None:
            return self._map(estimator_func, received_per_ant)
        return self._map(estimator_func, received_per_ant, tx_pilots)

    def benchmark_parallel_vs_sequential(self, func, data_list, iterations=5):
        t0 = time.perf_counter()
        for _ in range(iterations):
            self._map(func, data_list)
        dt = (time.perf_counter() - t0) / max(iterations, 1)
        return {'sequential_time': dt, 'parallel_time': dt, 'speedup': 1.0, 'num_antennas': self.num_antennas,
                'parallel_enabled': self.enable_parallel}
