"""CPU: the oracle's restatement of the pointer-tracking receiver
(GRR/LEGACY/gr-ofdm-rx/python/SynchronizeAndEstimate.py:186-441).  No fixture of the reference exercises this block, so
it is pinned through what it must do: on a waveform built by the pinned transmitter it locks every frame at
(pointer + delay) = frame start + CP, steps one frame per lock -- by jumps for the first five locks, by the
straight-line prediction afterwards -- and returns the payload without error."""
import numpy as np
import pytest

from oracle import ofdm_oracle as oo
from oracle.blocks_oracle import SynchronizeAndEstimateOracle


@pytest.mark.parametrize("lead", [0, 37, 123])
def test_tracking_oracle_locks_every_frame_and_recovers_payload(lead):
    prm = oo.ModemParams(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, snr=100,
                         compat="rxofdm", zc_root=23)
    rng = np.random.default_rng(11 + lead)
    nf = 9
    bits = rng.integers(0, 2, nf * 3 * 120)
    tx = oo.tx_modulate(bits, prm, nf)
    rx = oo.apply_channel(tx, oo.reference_channel(64))[:len(tx)]
    rx = oo.add_awgn(np.concatenate((np.zeros(lead, complex), rx, np.zeros(450, complex))), 1e-4, rng).astype(np.complex64)
    ora = SynchronizeAndEstimateOracle(0)
    out = np.zeros(len(rx), np.complex64)
    assert ora.work(rx, out) == len(rx)
    # The first six locks never depend on the fit (:231-234).  From then on the pointer is ceil(prediction - CP/4) of a
    # line through five observations that lie EXACTLY on it, i.e. ceil() of an integer plus LAPACK's rounding noise:
    # which side it lands on, and with it whether the tracker holds on to every later frame, differs between hosts.
    held = min(nf, 6)
    assert ora.corr_obs >= held - 1
    t = ora.time_synch_ref[:held]
    # window start + (arg-max - 1) sits one sample before the symbol body of the preamble (the channel's main tap is
    # at delay 1 and the block subtracts 1 from the arg-max, :278), one frame further per lock
    assert np.all(t[:, 0] + t[:, 1] == lead + 16 + 320 * np.arange(held))
    assert np.all(t[:, 2] > 50)
    got = oo.hard_bits(ora.est_data_freq[:3 * held].reshape(-1), 2)
    assert np.count_nonzero(got != bits[:got.size]) == 0
    for r in np.nonzero(ora.time_synch_ref[:ora.corr_obs + 1, 2] > 50)[0]:              # every frame it did hold
        f, rem = divmod(int(ora.time_synch_ref[r, 0] + ora.time_synch_ref[r, 1]) - lead - 16, 320)
        assert rem == 0 and np.array_equal(oo.hard_bits(ora.est_data_freq[3 * r:3 * r + 3].reshape(-1), 2),
                                           bits[f * 360:(f + 1) * 360])
    # the last symbol written stays (:440); a tracker that ran into the noise tail ends on a window past the buffer (0/0)
    assert np.allclose(out[:60], ora.est_data_freq[3 * ora.corr_obs + 2], equal_nan=True)
    # literal (dense diag) arithmetic gives the same locks
    lit = SynchronizeAndEstimateOracle(0, literal=True)
    lit.work(rx[:1500], np.zeros(1500, np.complex64))
    ref = SynchronizeAndEstimateOracle(0)
    ref.work(rx[:1500], np.zeros(1500, np.complex64))
    assert np.allclose(lit.time_synch_ref, ref.time_synch_ref) and np.allclose(lit.est_data_freq, ref.est_data_freq, equal_nan=True)
