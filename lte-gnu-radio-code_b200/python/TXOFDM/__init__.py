"""TXOFDM -- same public name as GRR/gr-TXOFDM/python/__init__.py:23."""
from .tx_signal_transmitter import tx_signal_transmitter
from .tx_modulator_vec import tx_modulator_vec

__all__ = ["tx_signal_transmitter", "tx_modulator_vec"]     # the second one is additive
