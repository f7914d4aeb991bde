"""Reference-compatible `core` package (same module and class names as
Darioxavierl/OFDM-LTE `core/`), backed by the CUDA engine in `lte_b200`."""
