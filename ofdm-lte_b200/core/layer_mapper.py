"""Round-robin symbol <-> layer mapping (reference core/layer_mapper.py:14-160).  Pure index
arithmetic; inside the link chain it is folded into `lte_sm_precode` / `lte_mimo_detect`."""
import numpy as np


class LayerMapper:
    def __init__(self, num_layers):
        if num_layers < 1 or num_layers > 8:
            raise ValueError(f"num_layers debe estar en [1,8], recibido: {num_layers}")
        self.num_layers = num_layers

    def map_to_layers(self, symbols):
        symbols = np.asarray(symbols)
        if self.num_layers == 1:
            return symbols.reshape(1, -1)
        rem = len(symbols) % self.num_layers
        if rem:
            symbols = np.concatenate([symbols, np.zeros(self.num_layers - rem, dtype=symbols.dtype)])
        return symbols.reshape(-1, self.num_layers).T

    def demap_from_layers(self, layers, original_length=None):
        layers = np.asarray(layers)
        symbols = layers.flatten() if self.num_layers == 1 else layers.T.flatten()
        return symbols if original_length is None else symbols[:original_length]

    def get_symbols_per_layer(self, total_symbols):
        return total_symbols if self.num_layers == 1 else int(np.ceil(total_symbols / self.num_layers))

    def get_padded_length(self, total_symbols):
        rem = total_symbols % self.num_layers
        return total_symbols if (self.num_layers == 1 or rem == 0) else total_symbols + self.num_layers - rem


class LayerDemapper:
    def __init__(self, num_layers):
        self.mapper = LayerMapper(num_layers)

    def demap(self, layers, original_length=None):
        return self.mapper.demap_from_layers(layers, original_length)
