"""TXOFDM -- same public name as GRR/gr-TXOFDM/python/__init__.py:23."""
from .tx_signal_transmitter import tx_signal_transmitter

__all__ = ["tx_signal_transmitter"]
