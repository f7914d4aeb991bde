"""Backward-compatible façade (reference ofdm_module.py:32-207): OFDMModule wraps a SISO
OFDMSimulator whose stages run on the CUDA engine."""
from config import LTEConfig
from core.ofdm_core import OFDMSimulator


class OFDMModule:
    def __init__(self, config=None, channel_type='awgn', mode='lte', enable_sc_fdm=False,
                 enable_equalization=True, **engine_options):
        if config is None:
            config = LTEConfig()
        self.config = config
        self.channel_type = channel_type
        self.mode = mode
        self.enable_sc_fdm = enable_sc_fdm
        self.enable_equalization = enable_equalization
        self.simulator = OFDMSimulator(config=config, channel_type=channel_type, mode=mode,
                                       enable_sc_fdm=enable_sc_fdm, enable_equalization=enable_equalization,
                                       num_channels=1, **engine_options)
        self.last_results = None

    def transmit(self, bits, snr_db=10.0):
        results = self.simulator.simulate_siso(bits, snr_db=snr_db)
        self.last_results = results
        return results

    def _calculate_papr(self, signal):
        return self.simulator.tx.calculate_papr(signal)

    @property
    def channel(self):
        return self.simulator.channels[0]

    @property
    def modulator(self):
        return self.simulator.tx.modulator

    @property
    def demodulator(self):
        return self.simulator.rx.demodulator

    @property
    def tx(self):
        return self.simulator.tx

    @property
    def rx(self):
        return self.simulator.rx

    def run_ber_sweep(self, num_bits, snr_range, num_trials=1, progress_callback=None):
        return self.simulator.run_ber_sweep(num_bits, snr_range, num_trials=num_trials,
                                            progress_callback=progress_callback)

    def get_config(self):
        return self.config

    def __repr__(self):
        return f"OFDMModule({self.config.modulation}, {'SC-FDM' if self.enable_sc_fdm else 'OFDM'}, {self.channel_type})"
