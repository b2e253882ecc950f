"""RXOFDM -- same public name as GRR/gr-RXOFDM/python/__init__.py:23."""
from .synch_and_chan_est import synch_and_chan_est
from .synch_and_chan_est_vec import synch_and_chan_est_vec

__all__ = ["synch_and_chan_est", "synch_and_chan_est_vec"]     # the second one is additive
