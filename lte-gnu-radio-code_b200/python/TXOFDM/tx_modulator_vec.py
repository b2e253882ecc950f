"""TXOFDM.tx_modulator_vec -- live modulator with one frame per item (additive; the reference's TX blocks replay a
waveform that GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/SDRScript.py:128-134 modulated offline).  See
ofdm_b200/vec_block.py.  zc_root defaults to 23, the root the repository's transmitter uses (SynchSignal.py:21-30)."""
from ofdm_b200.vec_block import TxVecBlockBase


class tx_modulator_vec(TxVecBlockBase):
    def __init__(self, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, *, modulation=2, zc_root=23, device=None):
        self._setup_vec("tx_modulator_vec", nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, modulation, zc_root,
                        "utsa", device)
