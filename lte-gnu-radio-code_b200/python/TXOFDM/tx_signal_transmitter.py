"""TXOFDM.tx_signal_transmitter -- drop-in for GRR/gr-TXOFDM/python/tx_signal_transmitter.py:13-27.

The reference file defines the class as `TxSignalTransmitter` while its package imports
`tx_signal_transmitter`; the name used by ofdm_chain.py:82 and the GRC template is
exported here, with the other kept as an alias."""
from ofdm_b200.tx_block import TxReplayBase


class tx_signal_transmitter(TxReplayBase):
    def __init__(self, case, pickle_directory, pickle_file, **kw):
        self.case = case
        TxReplayBase.__init__(self, pickle_directory, pickle_file, **kw)


TxSignalTransmitter = tx_signal_transmitter
