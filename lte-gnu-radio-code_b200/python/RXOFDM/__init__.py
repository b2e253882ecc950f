"""RXOFDM -- same public name as GRR/gr-RXOFDM/python/__init__.py:23."""
from .synch_and_chan_est import synch_and_chan_est

__all__ = ["synch_and_chan_est"]
