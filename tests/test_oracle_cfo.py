"""CPU checks of the frequency-offset hypothesis bank in the oracle: the bank undoes a carrier offset
that breaks the plain receiver, the literal (diag-matmul) and vectorised forms agree, and the legacy
block's stateful restatement (oracle/blocks_oracle.py, SynchEstAndFOOracle) locks frame after frame."""
import numpy as np

from oracle import ofdm_oracle as oo
from oracle.blocks_oracle import SynchEstAndFOOracle, SYNCH_EST_AND_FO_CASES


def _waveform(prm, n_frames, fs, cfo_hz, seed, lead=0, tail=0):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, n_frames * prm.d * prm.num_data_bins * prm.bps)
    tx = oo.tx_modulate(bits, prm, n_frames)
    rx = oo.apply_channel(tx, oo.reference_channel(prm.nfft))[:len(tx)]
    rx = np.concatenate((np.zeros(lead), rx, np.zeros(tail)))
    rx = rx * np.exp(-2j * np.pi * cfo_hz / fs * np.arange(len(rx)))
    return bits, oo.add_awgn(rx, 1e-5, rng)


def test_bank_matches_reference_formula():
    fo, fs, n = [-75.0, 0.0, 50.0], 960000.0, 64
    c = oo.cfo_bank(fo, fs, n)
    assert c.shape == (3, 64)
    assert np.allclose(c[2], np.exp(1j * 2 * np.pi * 50.0 / fs * np.arange(64)))
    assert np.allclose(c[1], 1.0)


def test_window_local_bank_is_the_reference_form():
    """The reference restarts the phasor at every window (SynchEstAndFO.py:254,:332), so it only serves offsets
    whose rotation over one symbol is negligible -- its own example is range(-75, 75, 25) Hz at 960 kHz (:157)."""
    fs = 960000.0
    prm = oo.ModemParams(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=1, bps=2, snr=60, gate=0.5)
    fo = list(range(-75, 75, 25))
    bits, rx = _waveform(prm, 6, fs, 50.0, 3)
    bank = oo.rx_frames(rx, 6, prm, cfo=oo.cfo_bank(fo, fs, 64))
    assert bank["locked"].all() and np.count_nonzero(oo.hard_bits(bank["eq"].reshape(-1), 2) != bits) == 0
    lit = oo.rx_frames(rx, 2, prm, cfo=oo.cfo_bank(fo, fs, 64), literal=True)
    assert np.allclose(lit["eq"], bank["eq"][:2], atol=1e-9) and np.array_equal(lit["hyp"], bank["hyp"][:2])
    # a large offset defeats it: the windows are de-rotated but the symbols keep their relative phase
    step = 0.3 * fs / 64
    bits, rx = _waveform(prm, 6, fs, step, 4)
    big = oo.rx_frames(rx, 6, prm, cfo=oo.cfo_bank([-step, 0.0, step], fs, 64))
    assert set(big["hyp"]) == {2}
    assert np.count_nonzero(oo.hard_bits(big["eq"].reshape(-1), 2) != bits) > 100


def test_phase_continuous_bank_recovers_payload_that_the_plain_receiver_loses():
    """EXTENSION: the phasor continued from the buffer origin undoes a real carrier offset, also across the M0
    preamble repeats that the window-local form adds up incoherently."""
    fs = 960000.0
    prm = oo.ModemParams(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=2, d=3, bps=2, snr=60, gate=0.5)
    step = 0.3 * fs / 64
    fo = [-step, 0.0, step]
    nus = [f / fs for f in fo]
    bits, rx = _waveform(prm, 8, fs, step, 3)
    plain = oo.rx_frames(rx, 8, prm)
    local = oo.rx_frames(rx, 8, prm, cfo=oo.cfo_bank(fo, fs, 64))
    cont = oo.rx_frames(rx, 8, prm, cfo=oo.cfo_bank(fo, fs, 64), nus=nus)
    assert set(cont["hyp"]) == {2} and cont["locked"].all()
    assert cont["peak"].min() > 1.5 * local["peak"].max() > 1.5 * plain["peak"].max()
    assert np.count_nonzero(oo.hard_bits(cont["eq"].reshape(-1), 2) != bits) == 0
    assert np.count_nonzero(oo.hard_bits(plain["eq"].reshape(-1), 2) != bits) > 500


def test_legacy_block_restatement_locks_every_frame():
    ora = SynchEstAndFOOracle(1, [-3000.0, 0.0, 3000.0])
    assert SYNCH_EST_AND_FO_CASES[1] == (48, 960000, 64, [1, 1], 36)
    prm = ora.prm
    assert (prm.zc_root, prm.stride, prm.cp_len) == (37, 15, 16)
    _, rx = _waveform(prm, 10, ora.fs, 3000.0, 8, lead=9, tail=40)
    out = np.zeros(len(rx), np.complex64)
    ora.work(rx, out)
    assert ora.last_written == 0
    ora.work(rx, out)
    assert ora.last_written == 24 * 36
    locks = ora.time_synch_ref[:, 0]
    assert np.count_nonzero(locks) == 10
    # one lock per [1 synch, 1 data] frame, on the stride-15 search grid just below the true position 9 + 16 + 160 f
    off = (25 + 160 * np.arange(10)) - locks[:10]
    assert np.all((off >= 0) & (off < 15))
