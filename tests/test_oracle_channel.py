"""CPU: the oracle's batched test-bench channel (GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/MultiAntennaSystem.py:221-260) and
its counter-based noise stream.  Philox4x32-10 is pinned to the Random123 known-answer vectors (kat_vectors of the
reference implementation of Salmon et al., SC'11), so the CUDA generator -- compared sample by sample with this
restatement in tests/test_gpu_bench_parity.py -- is pinned to the published algorithm as well."""
import numpy as np

from oracle import ofdm_oracle as oo


def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = oo.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(v) for v in got) == want
    # vectorised over a batch of counters
    both = oo.philox4x32_10(np.array([kat[0][0], kat[1][0]], dtype=np.uint32), (0, 0))
    assert tuple(int(v) for v in both[0]) == kat[0][2]


def test_philox_awgn_is_counter_addressed_and_gaussian():
    z = oo.philox_awgn(200000, 0.5, 2024)
    assert np.array_equal(oo.philox_awgn(1000, 0.5, 2024, first=5000), z[5000:6000])       # any stretch, any order
    assert abs(z.real.var() - 0.25) < 0.01 and abs(z.imag.var() - 0.25) < 0.01 and abs(z.mean()) < 0.01
    assert abs(np.mean(z[1:] * np.conj(z[:-1]))) < 0.01
    assert not np.allclose(oo.philox_awgn(100, 0.5, 2025), z[:100])
    # the 64-bit sample index continues into the second counter word
    hi = oo.philox_awgn(4, 1.0, 1, first=(1 << 32) - 2)
    assert not np.allclose(hi[2:], oo.philox_awgn(2, 1.0, 1, first=0))


def test_batched_channel_is_the_reference_convolution_per_frame(golden):
    prm = oo.ModemParams()
    n_frames, fl = 6, prm.frame_len
    tx = np.asarray(golden["tx_online"]).reshape(-1)[:n_frames * fl]
    h = oo.REF_TAPS / np.linalg.norm(oo.REF_TAPS)
    y = oo.channel_awgn_frames(tx, n_frames, fl, h)
    for f in range(n_frames):
        assert np.allclose(y[f * fl:(f + 1) * fl], np.convolve(tx[f * fl:(f + 1) * fl], h)[:fl])
    # one frame = the reference's whole-buffer np.convolve (rx_signal_gen) truncated to the buffer
    assert np.allclose(oo.channel_awgn_frames(tx, 1, len(tx), h), oo.apply_channel(tx, h)[:len(tx)])
    per = np.stack([np.roll(h, 0) * (f + 1) for f in range(n_frames)])
    y2 = oo.channel_awgn_frames(tx, n_frames, fl, per)
    assert np.allclose(y2[2 * fl:3 * fl], 3 * y[2 * fl:3 * fl])
    yn = oo.channel_awgn_frames(tx, n_frames, fl, h, 0.1, 7)
    assert np.allclose(yn - y, oo.philox_awgn(n_frames * fl, 0.1, 7))
