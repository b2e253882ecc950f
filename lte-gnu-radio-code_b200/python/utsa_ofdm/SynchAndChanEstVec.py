"""utsa_ofdm.SynchAndChanEstVec -- vector-item form of utsa_ofdm.SynchAndChanEst (additive; no counterpart in the
reference, whose blocks carry single samples: GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:112-115).  One frame per
item in, that frame's equalised data symbols per item out; see ofdm_b200/vec_block.py.  Constructor arguments keep the
reference's names, meaning and order where they still apply (num_ofdm_symb, the diagnostics file arguments and `genie`
have no meaning without cross-call state and are gone); variant constants are the utsa ones: Zadoff-Chu root 23,
gate = scale_factor_gate, `snr` as the MMSE regulariser (:184,:243)."""
from ofdm_b200.vec_block import RxVecBlockBase


class SynchAndChanEstVec(RxVecBlockBase):
    compat = "utsa"

    def __init__(self, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, scale_factor_gate, *,
                 modulation=2, zc_root=None, device=None, soft=False, noise_var=1.0):
        self._setup_vec("SynchAndChanEstVec", nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr,
                        scale_factor_gate, modulation, zc_root, device, soft, noise_var)
