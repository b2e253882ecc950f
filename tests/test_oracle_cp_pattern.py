"""Oracle, CPU only: the LTE normal-CP pattern (extension, our definition -- the reference has one prefix length,
GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:79).  Checks the framing arithmetic and a TX -> channel -> RX round trip."""
import numpy as np

from oracle import ofdm_oracle as oo


def test_lte_subframe_is_30720_samples():
    prm = oo.ModemParams(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200,
                         m0=1, d=13, bps=2)
    assert [prm.sym_cp(k) for k in range(14)] == [160] + [144] * 6 + [160] + [144] * 6
    assert prm.frame_len == 30720                         # TS 36.211 5.6: one subframe at 30.72 Msps
    assert prm.sym_start(7) == 15360                      # the second slot starts half-way


def test_uniform_prefix_is_unchanged():
    a = oo.ModemParams()
    b = oo.ModemParams(cp_first=16, slot_syms=7)          # cp_first == cp_len: no pattern
    assert not a.cp_pattern and not b.cp_pattern and a.frame_len == b.frame_len == 4 * 80
    bits = np.random.default_rng(0).integers(0, 2, 2 * 3 * 60 * 2)
    assert np.array_equal(oo.tx_modulate(bits, a, 2), oo.tx_modulate(bits, b, 2))


def test_round_trip_with_pattern():
    prm = oo.ModemParams(nfft=64, cp_len=16, cp_first=20, slot_syms=2, num_synch_bins=62, num_data_bins=60, m0=1,
                         d=4, bps=2, snr=100, gate=0.5)
    n_frames = 3
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, n_frames * prm.d * 60 * 2)
    tx = oo.tx_modulate(bits, prm, n_frames)
    assert len(tx) == n_frames * prm.frame_len == n_frames * (5 * 80 + 3 * 4)
    # every symbol: the prefix repeats the tail of the body, and the symbol has unit variance
    at = 0
    for k in range(prm.frame_syms):
        cp = prm.sym_cp(k)
        sym = tx[at:at + 64 + cp]
        assert np.allclose(sym[:cp], sym[64:64 + cp])
        assert abs(np.var(sym) - 1.0) < 1e-9
        at += 64 + cp
    rx = oo.add_awgn(oo.apply_channel(tx, oo.reference_channel(64))[:len(tx)], 1e-5, rng)
    res = oo.rx_frames(rx, n_frames, prm)
    assert res["locked"].all() and (res["delay"] == 1).all()
    assert np.array_equal(oo.hard_bits(res["eq"].reshape(-1), 2), bits)
