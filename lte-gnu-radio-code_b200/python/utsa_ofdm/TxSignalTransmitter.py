"""utsa_ofdm.TxSignalTransmitter -- drop-in for GRR/gr-utsa_ofdm/python/TxSignalTransmitter.py:15-29."""
from ofdm_b200.tx_block import TxReplayBase


class TxSignalTransmitter(TxReplayBase):
    def __init__(self, pickle_directory, pickle_file, **kw):
        TxReplayBase.__init__(self, pickle_directory, pickle_file, **kw)
