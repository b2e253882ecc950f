"""Parity on what bench.py actually times, and on the test-bench channel that builds its inputs.

* channel_kernel (ofdm_channel_awgn) against the oracle's batched restatement of
  MultiAntennaSystem.rx_signal_gen / additive_noise (GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/MultiAntennaSystem.py:221-260):
  the FIR per frame, and the Philox4x32-10 noise stream sample by sample.
* The device-generated batches of the two benchmarked workloads (bench.make_batch: rx_n64_16qam_1m = BASELINE
  configs[1], rx_lte20_64qam = configs[2] with an independent multipath per frame): the fused receiver's outputs on
  slices of >= 4096 data symbols taken from the start, the middle and the end of the batch against oo.rx_frames on the
  same samples -- delay / locked exact, peak / eq / LLR <= 1e-4 of the RMS, hard bits exact wherever the oracle's own
  LLR is not within fp32 rounding of a decision boundary.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

import bench  # noqa: E402
from oracle import ofdm_oracle as oo  # noqa: E402
from test_gpu_parity import make, oracle_prm, unpack_symbols, rel_err, RTOL  # noqa: E402

REF = dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4)


def _tx_batch(m, n_frames, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    bits = torch.randint(0, 256, (n_frames * m.cfg.d * m.cfg.hard_stride,), dtype=torch.uint8, device="cuda", generator=g)
    return m.tx_modulate(bits, n_frames)


@pytest.mark.parametrize("kw,n_frames", [(REF, 37), (dict(nfft=256, cp_len=20, num_synch_bins=254, num_data_bins=200,
                                                         m0=2, d=2, bps=2), 5)], ids=["N64", "N256"])
def test_channel_fir_matches_oracle_convolution(kw, n_frames):
    """noise_sigma = 0: shared taps and one tap row per frame, every frame filtered on its own."""
    m = make(dict(kw, snr=100))
    fl = m.cfg.frame_len
    tx = _tx_batch(m, n_frames, 3)
    tx_h = tx.cpu().numpy()
    shared = (oo.REF_TAPS / np.linalg.norm(oo.REF_TAPS)).astype(np.complex64)
    got = m.channel_awgn(tx, n_frames, torch.from_numpy(shared).cuda(), 0.0, 0).cpu().numpy()
    ref = oo.channel_awgn_frames(tx_h, n_frames, fl, shared)
    assert rel_err(got, ref) < 1e-5
    # the frame-local convolution equals np.convolve over the whole stream except for the taps' reach into each frame
    whole = oo.apply_channel(tx_h, shared)[:len(tx_h)]
    inner = np.ones(len(tx_h), bool)
    for f in range(1, n_frames):
        inner[f * fl:f * fl + len(shared) - 1] = False
    assert rel_err(got[inner], whole[inner]) < 1e-5
    rng = np.random.default_rng(11)
    per = ((rng.standard_normal((n_frames, 7)) + 1j * rng.standard_normal((n_frames, 7))) / 3).astype(np.complex64)
    got = m.channel_awgn(tx, n_frames, torch.from_numpy(per).cuda(), 0.0, 0).cpu().numpy()
    assert rel_err(got, oo.channel_awgn_frames(tx_h, n_frames, fl, per)) < 1e-5


def test_channel_noise_is_the_philox_stream_of_the_oracle():
    """With noise: sample i carries sigma * BoxMuller(Philox4x32-10(counter i, key seed)) -- the oracle restates the
    generator (pinned to the Random123 known-answer vectors in the CPU suite) and the two streams agree sample by
    sample; different seeds give different streams; the stream does not depend on the launch shape."""
    m = make(dict(REF, snr=100))
    n_frames, fl = 700, m.cfg.frame_len
    tx = _tx_batch(m, n_frames, 4)
    taps = torch.tensor([1.0], dtype=torch.complex64).cuda()
    sigma, seed = 0.37, (5 << 32) + 2024
    got = m.channel_awgn(tx, n_frames, taps, sigma, seed).cpu().numpy().astype(complex) - tx.cpu().numpy()
    ref = oo.philox_awgn(n_frames * fl, sigma, seed)
    # tx is O(1) in fp32, so subtracting it back leaves 6e-8 of rounding on top of the generator's own accuracy
    assert np.max(np.abs(got - ref)) < 2e-6
    again = m.channel_awgn(tx[:50 * fl].contiguous(), 50, taps, sigma, seed).cpu().numpy() - tx[:50 * fl].cpu().numpy()
    assert np.array_equal(again, (got[:50 * fl]).astype(np.complex64)) or np.max(np.abs(again - got[:50 * fl])) < 1e-6
    other = m.channel_awgn(tx, n_frames, taps, sigma, seed + 1).cpu().numpy() - tx.cpu().numpy()
    assert np.abs(np.vdot(other, got)) / (np.linalg.norm(other) * np.linalg.norm(got)) < 0.02


def test_channel_noise_statistics():
    """4.8 M samples: zero mean, per-component variance sigma^2, I/Q uncorrelated, white (lags 1..3), Gaussian kurtosis."""
    m = make(dict(REF, snr=100))
    n_frames = 15000
    zero = torch.zeros(n_frames * m.cfg.frame_len, dtype=torch.complex64, device="cuda")
    z = m.channel_awgn(zero, n_frames, torch.tensor([1.0], dtype=torch.complex64).cuda(), 0.5, 99).cpu().numpy().astype(complex)
    n = z.size
    assert abs(z.mean()) < 5 * 0.5 * np.sqrt(2.0 / n)
    assert abs(z.real.var() - 0.25) < 0.25 * 6 * np.sqrt(2.0 / n) and abs(z.imag.var() - 0.25) < 0.25 * 6 * np.sqrt(2.0 / n)
    assert abs(np.mean(z.real * z.imag)) < 0.25 * 6 / np.sqrt(n)
    for lag in (1, 2, 3):
        assert abs(np.mean(z[lag:] * np.conj(z[:-lag]))) < 0.5 * 6 / np.sqrt(n)
    assert abs(np.mean(z.real ** 4) / z.real.var() ** 2 - 3.0) < 0.02


def _compare_slice(m, prm, rx, got, f0, nf, tag):
    cfg = m.cfg
    lo, hi = f0 * cfg.frame_len, (f0 + nf) * cfg.frame_len
    tail = min(rx.numel() - hi, cfg.frame_len)              # the last frame's delayed windows may look past its end
    rx_h = rx[lo:hi + tail].cpu().numpy()
    ref = oo.rx_frames(rx_h, nf, prm)
    s0, s1 = f0 * cfg.d, (f0 + nf) * cfg.d
    assert np.array_equal(got["delay"][f0:f0 + nf].cpu().numpy(), ref["delay"]), tag
    assert np.array_equal(got["locked"][f0:f0 + nf].cpu().numpy(), ref["locked"]), tag
    assert ref["locked"].all(), tag
    assert rel_err(got["peak"][f0:f0 + nf].cpu().numpy(), ref["peak"]) < RTOL, tag
    # The 1e-4 bar is relative to the RMS of the compared quantity for a well-conditioned bin.  The one-tap equaliser
    # amplifies every rounding error of a bin by 1/|H_k|, so a bin in a fade (the per-frame random multipath of the LTE
    # workload has them) gets the bar scaled by its conditioning kappa_k = rms|H| / |H_k| >= 1 -- the fp32 error bound
    # of the quotient, not a looser test: on the fixed reference channel kappa stays below 3.
    hk = np.abs(ref["h"][:, oo.bins_used(cfg.num_data_bins, cfg.nfft)])              # [frames, ndb]
    kappa = np.maximum(1.0, np.sqrt(np.mean(hk ** 2)) / np.maximum(hk, 1e-30))[:, None, :]
    eq = got["eq"][s0:s1].cpu().numpy().reshape(ref["eq"].shape)
    rms_eq = np.sqrt(np.mean(np.abs(ref["eq"]) ** 2))
    assert np.all(np.abs(eq - ref["eq"]) <= RTOL * rms_eq * kappa), (tag, float(np.max(np.abs(eq - ref["eq"]) / (rms_eq * kappa))))
    well = np.broadcast_to(kappa < 2.0, eq.shape)
    assert well.mean() > 0.5 and rel_err(eq[well], ref["eq"][well]) < RTOL, tag
    rows = ref["eq"].reshape(nf * cfg.d, -1)                 # brute force over the constellation, a few rows at a time
    ref_llr = np.concatenate([oo.maxlog_llr(rows[i:i + 256], cfg.bps, 1.0 / cfg.noise_var) for i in range(0, len(rows), 256)])
    rms = np.sqrt(np.mean(ref_llr ** 2))
    tol_llr = RTOL * rms * np.repeat(np.broadcast_to(kappa, eq.shape).reshape(nf * cfg.d, -1), cfg.bps, axis=1)
    assert np.all(np.abs(got["llr"][s0:s1].cpu().numpy() - ref_llr) <= tol_llr), tag
    per_sym = cfg.num_data_bins * cfg.bps
    ref_bits = oo.hard_bits(ref["eq"].reshape(-1), cfg.bps)
    got_bits = unpack_symbols(got["hard"][s0:s1].cpu().numpy(), per_sym)
    # a decision is only defined in fp32 where the oracle's LLR is not within that bar of zero
    decisive = (np.abs(ref_llr) > tol_llr).reshape(-1)
    assert np.array_equal(got_bits[decisive], ref_bits[decisive]), tag
    assert np.mean(~decisive) < 1e-3, tag
    return int(np.sum(got_bits != ref_bits)), int(np.sum(~decisive)), ref_bits


@pytest.mark.parametrize("wname,sub,edge", [("rx_n64_16qam_1m", 1366, 1366), ("rx_lte20_64qam", 316, 16)])
def test_benchmarked_batch_matches_oracle(wname, sub, edge):
    w = bench.WORKLOADS[wname]
    from ofdm_b200 import Modem
    m = Modem(bench.modem_config(w))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    nf = w["n_frames"]
    bits, tx, taps, rx = bench.make_batch(m, w, nf, 0)
    got = m.rx_frames(rx, nf)
    torch.cuda.synchronize()
    assert sub * cfg.d >= 4096
    flips = ties = 0
    # first, middle (odd phase, >= 4096 data symbols) and last frames of the batch
    for f0, n in ((0, edge), ((nf - sub) // 2 + 1, sub), (nf - edge, edge)):
        a, b, ref_bits = _compare_slice(m, prm, rx, got, f0, n, "%s frames %d.." % (wname, f0))
        flips += a
        ties += b
    assert flips <= ties
    # the batch is what the workload says it is: the oracle's channel + Philox stream reproduces a stretch of it
    fl = cfg.frame_len
    k = 8
    tp = taps[:k].cpu().numpy() if taps.dim() == 2 else taps.cpu().numpy()
    ref_rx = oo.channel_awgn_frames(tx[:k * fl].cpu().numpy(), k, fl, tp, w["noise_sigma"], 2024)
    assert rel_err(rx[:k * fl].cpu().numpy(), ref_rx) < 1e-5
    # BER of the whole batch from the device counters equals the oracle's decisions vs the payload on the last slice
    # to within sampling error
    cnt = m.counters(bits, nf, bits_rx=got["hard"], eq=got["eq"], locked=got["locked"]).cpu().numpy()
    assert cnt[4] == nf and cnt[5] == nf
    per_sym = cfg.num_data_bins * cfg.bps
    tx_bits = unpack_symbols(bits.view(nf * cfg.d, cfg.hard_stride)[(nf - edge) * cfg.d:].cpu().numpy(), per_sym)
    ber_ref = np.mean(ref_bits != tx_bits)
    assert cnt[0] / cnt[1] == pytest.approx(ber_ref, rel=0.3, abs=1e-4)


def test_host_receiver_survives_a_hypothesis_bank_change():
    """set_cfo_bank used to free the host-buffer receiver's per-frame result buffer without forgetting it.
    rx_frames_host -> set_cfo_bank -> rx_frames_host -> clear bank -> rx_frames_host -> close, results identical."""
    m = make(dict(REF, bps=2, snr=100, gate=0.5))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    n_frames = 64
    rng = np.random.default_rng(2)
    bits = rng.integers(0, 2, n_frames * cfg.d * cfg.num_data_bins * cfg.bps)
    tx = oo.tx_modulate(bits, prm, n_frames)
    rx = oo.add_awgn(oo.apply_channel(tx, oo.reference_channel(64))[:len(tx)], 1e-4, rng).astype(np.complex64)
    a = m.rx_frames_host(rx, n_frames)
    m.set_cfo_bank([-100.0, 0.0, 100.0], 1e6)
    filler = [torch.empty(1 << 16, device="cuda") for _ in range(8)]      # anything freed by mistake gets reused here
    b = m.rx_frames_host(rx, n_frames)
    m.set_cfo_bank([], 1e6)
    c = m.rx_frames_host(rx, n_frames)
    del filler
    for k in ("delay", "locked", "hard"):
        assert np.array_equal(a[k], c[k]), k
        assert np.array_equal(a[k], b[k]), k                 # the zero-offset hypothesis wins on a clean capture
    assert np.array_equal(a["delay"], oo.rx_frames(rx, n_frames, prm)["delay"])
    m.close()


@pytest.mark.parametrize("kw", [dict(nfft=4096, cp_len=1024, num_synch_bins=4094, num_data_bins=3840, m0=1, d=1, bps=2),
                                dict(nfft=2048, cp_len=1030, num_synch_bins=2046, num_data_bins=1200, m0=1, d=1, bps=2),
                                dict(nfft=2048, cp_len=512, num_synch_bins=2046, num_data_bins=1200, m0=1, d=1, bps=2)],
                         ids=["N4096_CP1024", "N2048_CP1030", "N2048_CP512"])
def test_stride1_search_with_long_prefixes(kw):
    """the time-domain search tiles 1024 positions per CTA and yields 1024 - CP trials from them; with
    CP >= 1024 that is zero or negative (division by zero on the host).  Long prefixes take the per-trial transforms;
    CP = 512 is the longest prefix the time-domain kernel still serves."""
    m = make(dict(kw, snr=100, compat="utsa", zc_root=23))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    rng = np.random.default_rng(8)
    bits = rng.integers(0, 2, 2 * cfg.d * cfg.num_data_bins * cfg.bps)
    tx = oo.tx_modulate(bits, prm, 2)
    rx = oo.add_awgn(oo.apply_channel(tx, oo.reference_channel(cfg.nfft))[:len(tx)], 1e-4, rng).astype(np.complex64)
    n_trials = 2 * cfg.cp_len + 40
    first = cfg.frame_len                                     # around the second frame's preamble
    pk, dl = (a.cpu().numpy() for a in m.sync_search(torch.from_numpy(rx).cuda(), first, 1, n_trials))
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    for t in (0, 1, cfg.cp_len - 1, cfg.cp_len, n_trials - 1):
        d_i, d_v, _ = oo.sync_trial(rx, first + t, prm, zc, bank)
        assert abs(pk[t] - d_v) < RTOL * prm.mm
        if d_v > 0.5 * prm.mm:
            assert dl[t] == d_i
    assert pk.max() > 0.7 * prm.mm
