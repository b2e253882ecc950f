"""RXOFDM.synch_and_chan_est_vec -- vector-item form of RXOFDM.synch_and_chan_est (additive; no counterpart in the
reference, whose blocks carry single samples: GRR/gr-RXOFDM/python/synch_and_chan_est.py:121-124).  One frame per item
in, that frame's equalised data symbols per item out; see ofdm_b200/vec_block.py.  Variant constants are the gr-RXOFDM
ones: Zadoff-Chu root 37 keyed on the parity of num_synch_bins (:54-64), gate 0.4 (:170), `snr` as a linear regulariser
(:184,:247); pass zc_root=23 to receive what the repository's transmitter sends."""
from ofdm_b200.vec_block import RxVecBlockBase


class synch_and_chan_est_vec(RxVecBlockBase):
    compat = "rxofdm"

    def __init__(self, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, *, modulation=2, zc_root=None,
                 device=None, soft=False, noise_var=1.0):
        self._setup_vec("synch_and_chan_est_vec", nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, 0.4,
                        modulation, zc_root, device, soft, noise_var)
