"""utsa_ofdm -- same public names as GRR/gr-utsa_ofdm/python/__init__.py:23-24."""
from .SynchAndChanEst import SynchAndChanEst
from .SynchAndChanEstVec import SynchAndChanEstVec
from .TxSignalTransmitter import TxSignalTransmitter

__all__ = ["SynchAndChanEst", "TxSignalTransmitter", "SynchAndChanEstVec"]     # the last one is additive
