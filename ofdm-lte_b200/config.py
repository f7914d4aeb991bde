"""LTE numerology and ITU-R M.1225 channel tables.

Drop-in for the reference's `config` module (reference config.py:11-215): same names, constructor
signature, derived attributes and error behaviour.  Pure host arithmetic; the values feed
`lte_b200.LinkEngine` plans.
"""
import math


def _table(rows, keys):
    return {name: dict(zip(keys, vals)) for name, *vals in rows}


# bandwidth [MHz] -> occupied subcarriers Nc and FFT size N (reference config.py:11-18)
LTE_PROFILES = _table([(1.25, 76, 128), (2.5, 150, 256), (5.0, 300, 512), (10.0, 600, 1024), (15.0, 900, 2048),
                       (20.0, 1200, 2048)], ('Nc', 'N'))

# cyclic-prefix durations in microseconds (reference config.py:21-25)
CP_VALUES = {'normal': 4.7, 'extended_15khz': 16.6, 'extended_7.5khz': 33.0}

_BITS = {'QPSK': 2, '16-QAM': 4, '64-QAM': 6}
MODULATION_SCHEMES = list(_BITS)
SUBCARRIER_SPACING = [15.0, 7.5]

# ITU-R M.1225 tapped-delay-line profiles: tap delays in microseconds, tap powers in dB (reference config.py:34-60)
ITU_CHANNEL_MODELS = _table([
    ('Pedestrian_A', [0.0, 0.11, 0.19, 0.41], [0.0, -9.7, -19.2, -22.8], 'Pedestrian, low velocity, short distance'),
    ('Pedestrian_B', [0.0, 0.2, 0.8, 1.2, 2.3, 3.7], [0.0, -0.9, -4.9, -8.0, -7.8, -23.9], 'Pedestrian, high velocity'),
    ('Vehicular_A', [0.0, 0.31, 0.71, 1.09, 1.73, 2.51], [0.0, -1.0, -9.0, -10.0, -15.0, -20.0],
     'Vehicular, low velocity, short distance'),
    ('Vehicular_B', [0.0, 0.3, 0.7, 1.09, 1.73, 2.51, 3.7, 4.53], [0.0, -1.0, -9.0, -10.0, -13.0, -16.0, -21.6, -24.0],
     'Vehicular, high velocity, long distance'),
    ('Bad_Urban', [0.0, 0.1, 0.3, 0.5, 0.9, 1.3, 1.9, 2.6], [0.0, -3.0, -5.0, -7.0, -9.0, -11.0, -13.0, -15.0],
     'Urban with severe multipath'),
], ('delays_us', 'power_db', 'description'))

# label of every entry of get_info(), with the attribute (and display scale) it reports
_INFO = (('Bandwidth (MHz)', 'bandwidth', 1), ('Subcarrier Spacing (kHz)', 'delta_f', 1), ('Modulation', 'modulation', None),
         ('CP Type', 'cp_type', None), ('Useful Subcarriers (Nc)', 'Nc', 1), ('FFT Points (N)', 'N', 1),
         ('Sampling Frequency (MHz)', 'fs', 1e-6), ('Sampling Period (ns)', 'Ts', 1e9),
         ('OFDM Symbol Duration (μs)', 'T_symbol', 1e6), ('CP Duration (μs)', 'cp_duration', 1),
         ('CP Length (samples)', 'cp_length', 1), ('Bits per Symbol', 'bits_per_symbol', 1),
         ('Samples per OFDM Symbol', 'samples_per_ofdm_symbol', 1))


class LTEConfig:
    """Primary parameters (bandwidth, spacing, modulation, CP type) -> derived numerology (reference :63-198)."""

    def __init__(self, bandwidth=5.0, delta_f=15.0, modulation='QPSK', cp_type='normal'):
        if modulation not in _BITS:
            raise ValueError(f"Unsupported modulation: {modulation}. Options: {MODULATION_SCHEMES}")
        self.bandwidth, self.delta_f, self.modulation, self.cp_type = bandwidth, delta_f, modulation, cp_type
        self._calculate_parameters()

    def _next_power_of_2(self, x):
        return int(2 ** math.ceil(math.log2(x)))

    def _get_cp_duration(self):
        if self.cp_type != 'extended':
            return CP_VALUES['normal']
        return CP_VALUES['extended_15khz' if self.delta_f == 15.0 else 'extended_7.5khz']

    def _get_bits_per_symbol(self):
        return _BITS.get(self.modulation, 2)

    def _calculate_parameters(self):
        known = LTE_PROFILES.get(self.bandwidth)
        if known:
            self.Nc, self.N = known['Nc'], known['N']
        else:                                   # any other bandwidth: fill it with subcarriers, round the FFT up (:108-111)
            self.Nc = int((self.bandwidth * 1e3) / self.delta_f)
            self.N = self._next_power_of_2(self.Nc)
        self.fs = self.N * self.delta_f * 1e3
        self.Ts = 1 / self.fs
        self.T_symbol = self.N * self.Ts
        self.cp_duration = self._get_cp_duration()
        self.cp_length = int(self.cp_duration * 1e-6 * self.fs)          # truncation, not rounding (:124)
        self.bits_per_symbol = self._get_bits_per_symbol()
        self.samples_per_ofdm_symbol = self.N + self.cp_length

    def get_info(self):
        return {label: getattr(self, attr) if scale is None else getattr(self, attr) * scale
                for label, attr, scale in _INFO}

    def __str__(self):
        return "\n".join(["LTE OFDM Configuration:"] + [f"  {k}: {v}" for k, v in self.get_info().items()])

    def __repr__(self):
        return (f"LTEConfig(bandwidth={self.bandwidth}, delta_f={self.delta_f}, "
                f"modulation='{self.modulation}', cp_type='{self.cp_type}')")

    def copy(self):
        return LTEConfig(self.bandwidth, self.delta_f, self.modulation, self.cp_type)


def _preset(bandwidth, modulation):
    return lambda: LTEConfig(bandwidth=bandwidth, delta_f=15.0, modulation=modulation, cp_type='normal')


create_config_5MHz_QPSK = _preset(5.0, 'QPSK')
create_config_20MHz_16QAM = _preset(20.0, '16-QAM')
create_config_10MHz_64QAM = _preset(10.0, '64-QAM')
