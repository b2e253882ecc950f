"""Oracle, CPU only: the time-domain form of the sliding search (no transform per trial) gives the same peaks and delays
as the reference's per-trial formulation (sync_trial, GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:145-164)."""
import numpy as np
import pytest

from oracle import ofdm_oracle as oo


@pytest.mark.parametrize("kw,stride", [
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2), 1),
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=2, d=2, bps=2), 3),
    (dict(nfft=128, cp_len=9, num_synch_bins=120, num_data_bins=100, m0=1, d=2, bps=2), 1),     # eight unused bins
], ids=["ref", "m0_2_stride3", "n128_nsb120"])
def test_time_domain_search_equals_per_trial_transform(kw, stride):
    prm = oo.ModemParams(snr=100, **kw)
    rng = np.random.default_rng(1)
    nf = 3
    bits = rng.integers(0, 2, nf * prm.d * prm.num_data_bins * prm.bps)
    tx = oo.tx_modulate(bits, prm, nf)
    rx = oo.add_awgn(oo.apply_channel(tx, oo.reference_channel(prm.nfft))[:len(tx)], 1e-3, rng).astype(np.complex64)
    first = prm.cp_len
    n_trials = (len(rx) - first - 5) // stride               # the last trials run off the end: zero padding
    pk, dl = oo.sync_search_time_domain(rx, first, stride, n_trials, prm)
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    ref = [oo.sync_trial(rx, first + i * stride, prm, zc, bank)[:2] for i in range(n_trials)]
    rp = np.array([r[1] for r in ref])
    rd = np.array([r[0] for r in ref])
    assert np.max(np.abs(pk - rp)) < 1e-9 * np.max(rp)
    assert np.array_equal(dl, rd)


@pytest.mark.parametrize("m0,continuous", [(1, False), (2, False), (2, True)], ids=["m0_1", "m0_2", "m0_2_continuous"])
def test_time_domain_search_with_hypothesis_bank(m0, continuous):
    """The frequency-offset hypothesis bank (SynchEstAndFO.py:247-277) in the time-domain form: per-hypothesis taps
    h_s[m] u^m, two prefix terms per delay, sliding sums for DC and the band edge -- against sync_trial_cfo per trial."""
    prm = oo.ModemParams(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=m0, d=2, bps=2, snr=100)
    rng = np.random.default_rng(1)
    nf = 2
    bits = rng.integers(0, 2, nf * prm.d * prm.num_data_bins * prm.bps)
    tx = oo.tx_modulate(bits, prm, nf)
    fs = 960000.0
    step = 0.3 * fs / 64
    rx = oo.apply_channel(tx, oo.reference_channel(prm.nfft))[:len(tx)] * np.exp(-2j * np.pi * step / fs * np.arange(len(tx)))
    rx = oo.add_awgn(rx, 1e-3, rng).astype(np.complex64)
    fo = [-step, 0.0, step, 50.0]
    nus = [f / fs for f in fo]
    cfo = oo.cfo_bank(fo, fs, 64)
    first = prm.cp_len
    n_trials = len(rx) - first - 5
    pk, dl, hy = oo.sync_search_time_domain(rx, first, 1, n_trials, prm, nus=nus, continuous=continuous)
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    ref = [oo.sync_trial_cfo(rx, first + i, prm, cfo, zc, bank, nus=nus if continuous else None) for i in range(n_trials)]
    rp = np.array([r[2] for r in ref])
    assert np.max(np.abs(pk - rp)) < 1e-9 * np.max(rp)
    assert np.array_equal(dl, [r[1] for r in ref]) and np.array_equal(hy, [r[0] for r in ref])
