"""CPU checks of the DSSS restatement (oracle/ofdm_oracle.py dsss_*, oracle/blocks_oracle.py
SynchEstFOAndDSSSOracle): code definition, spread/despread round trip, and the legacy block's behaviour
(output written on the first call, state of row 0 kept across calls)."""
import numpy as np

from oracle import ofdm_oracle as oo
from oracle.blocks_oracle import SynchEstFOAndDSSSOracle, SYNCH_EST_FO_AND_DSSS_CASES


def spread_waveform(ora, n_frames, seed, lead=9, tail=40, noise=1e-5):
    """Frames of M0 synch symbols + one data symbol whose ndb bins carry ndb/DSSS QPSK symbols spread by ora.SC,
    through the reference multipath channel."""
    p = ora.prm
    rng = np.random.default_rng(seed)
    ng = p.num_data_bins // ora.DSSS
    syms = oo.bits_to_symbols(rng.integers(0, 2, n_frames * ng * 2), 2).reshape(n_frames, ng)
    grid = oo.tx_freq_grid(np.zeros(n_frames * p.d * p.num_data_bins * p.bps, dtype=int), p, n_frames)
    db = oo.bins_used(p.num_data_bins, p.nfft)
    for f in range(n_frames):
        row = f * p.frame_syms + p.m0
        grid[row, db] = 0
        grid[row, db[:ng * ora.DSSS]] = oo.dsss_spread(syms[f], ora.SC)
    tx = oo.tx_symb_gen(grid, p.cp_len)
    rx = oo.apply_channel(tx, oo.reference_channel(p.nfft))[:len(tx)]
    rx = np.concatenate((np.zeros(lead), rx, np.zeros(tail)))
    return syms, oo.add_awgn(rx, noise, rng).astype(np.complex64)


def test_code_and_round_trip():
    for sf in (1, 6, 7, 12, 24):
        c = oo.dsss_code(sf, 37)
        assert np.allclose(np.abs(c), 1.0) and c[0] == 1.0
        k = np.arange(sf)
        xx = k * k if sf % 2 == 0 else k * (k + 1)
        assert np.allclose(c, np.exp(-1j * np.pi * 37 * xx / sf))
        rng = np.random.default_rng(sf)
        s = rng.standard_normal((5, 4)) + 1j * rng.standard_normal((5, 4))
        assert np.allclose(oo.dsss_despread(oo.dsss_spread(s, c), c), s)
    # bins past the last whole group are ignored
    r = np.arange(10, dtype=complex)[None, :]
    assert oo.dsss_despread(r, oo.dsss_code(4, 37)).shape == (1, 2)


def test_dsss_block_restatement_quirks():
    assert SYNCH_EST_FO_AND_DSSS_CASES[1] == (48, 960000, 64, [3, 1], 36, 6)
    # one hypothesis: with several, the legacy block estimates the channel from the LAST one and de-rotates the data
    # with the winner of the last (noise) trial (SynchEstAndFOOracle), which only decodes if those happen to be right
    ora = SynchEstFOAndDSSSOracle(1, [0.0])
    p = ora.prm
    assert (p.m0, p.d, p.num_data_bins, ora.DSSS) == (3, 1, 36, 6)
    syms, rx = spread_waveform(ora, 9, 4, tail=30)
    out = np.zeros(len(rx), np.complex64)
    ora.work(rx, out)
    assert ora.last_written == 12 * 6                               # written on the very first call (:404-406)
    locks = int(np.count_nonzero(ora.time_synch_ref[:, 0]))
    assert locks == 9
    got = ora.est_data_freq_d[:8]
    assert np.array_equal(oo.hard_bits(got.reshape(-1), 2), oo.hard_bits(syms[:8].reshape(-1), 2))
    # (the row copy of :392 needs a lock whose data window does not fit; loop A's own guard (:274) is stricter than
    # loop B's (:366), so within one buffer it cannot happen -- the statement is restated but unreachable here)
    short = np.zeros(120, np.complex64)
    ora.work(short, out)                                            # nothing fits: state of row 0 survives, output rewritten
    assert ora.last_written == 12 * 6 and np.count_nonzero(ora.time_synch_ref[0]) == 3
