"""B200-native TDVP backend (``backend="cuda"``) behind a PyTDSCF-style Model / Simulator API."""
__version__ = "0.1.0"
