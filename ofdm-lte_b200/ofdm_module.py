"""Backward-compatible façade (reference ofdm_module.py:32-207): OFDMModule is a SISO OFDMSimulator under the
older single-object API; every stage runs on the CUDA engine."""
from config import LTEConfig
from core.ofdm_core import OFDMSimulator


def _delegate(path):
    """Read-only property that follows `path` from the wrapped simulator."""
    def get(self):
        obj = self.simulator
        for name in path:
            obj = obj[name] if isinstance(name, int) else getattr(obj, name)
        return obj
    return property(get)


class OFDMModule:
    # the pieces older callers reach into
    channel = _delegate(('channels', 0))
    modulator = _delegate(('tx', 'modulator'))
    demodulator = _delegate(('rx', 'demodulator'))
    tx = _delegate(('tx',))
    rx = _delegate(('rx',))

    def __init__(self, config=None, channel_type='awgn', mode='lte', enable_sc_fdm=False,
                 enable_equalization=True, **engine_options):
        self.config = LTEConfig() if config is None else config
        self.channel_type, self.mode = channel_type, mode
        self.enable_sc_fdm, self.enable_equalization = enable_sc_fdm, enable_equalization
        self.last_results = None
        self.simulator = OFDMSimulator(config=self.config, channel_type=channel_type, mode=mode, num_channels=1,
                                       enable_sc_fdm=enable_sc_fdm, enable_equalization=enable_equalization,
                                       **engine_options)

    def transmit(self, bits, snr_db=10.0):
        self.last_results = self.simulator.simulate_siso(bits, snr_db=snr_db)
        return self.last_results

    def run_ber_sweep(self, num_bits, snr_range, num_trials=1, progress_callback=None):
        return self.simulator.run_ber_sweep(num_bits, snr_range, num_trials=num_trials,
                                            progress_callback=progress_callback)

    def _calculate_papr(self, signal):
        return self.simulator.tx.calculate_papr(signal)

    def get_config(self):
        return self.config

    def __repr__(self):
        waveform = 'SC-FDM' if self.enable_sc_fdm else 'OFDM'
        return f"OFDMModule({self.config.modulation}, {waveform}, {self.channel_type})"
