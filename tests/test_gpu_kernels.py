"""More GPU parity: the modular kernels behind the stream blocks (sliding search, channel estimate,
symbol-list demod), the host-buffer entry points, counters, edge cases, and size-independent
properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import ofdm_oracle as oo  # noqa: E402
from test_gpu_parity import make, oracle_prm, pack_symbols, unpack_symbols, rel_err, RTOL  # noqa: E402

REF = dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2)


def synth(cfg, prm, n_frames, seed, noise=1e-4, pad=0):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, n_frames * cfg.d * cfg.num_data_bins * cfg.bps)
    tx = oo.tx_modulate(bits, prm, n_frames)
    rx = oo.apply_channel(tx, oo.reference_channel(cfg.nfft))[:len(tx)]
    rx = np.concatenate((rx, np.zeros(pad)))
    return bits, oo.add_awgn(rx, noise, rng).astype(np.complex64)


@pytest.mark.parametrize("compat,stride", [("utsa", 1), ("rxofdm", 15)])
def test_sliding_search_matches_oracle_trials(compat, stride):
    m = make(dict(REF, snr=100, compat=compat, zc_root=23))
    prm = oracle_prm(m.cfg)
    _, rx = synth(m.cfg, prm, 4, 3)
    n_trials = (len(rx) - 80 - 64 - 16 - 1) // stride + 1
    peak, delay = m.sync_search(torch.from_numpy(rx).cuda(), 16, stride, n_trials)
    peak, delay = peak.cpu().numpy(), delay.cpu().numpy()
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    ref = [oo.sync_trial(rx, 16 + t * stride, prm, zc, bank)[:2] for t in range(n_trials)]
    rp = np.array([r[1] for r in ref])
    rd = np.array([r[0] for r in ref])
    assert rel_err(peak, rp) < RTOL
    strong = rp > 0.3 * prm.mm                # argmax is only well defined where there is a peak
    assert strong.sum() > 8
    assert np.array_equal(delay[strong], rd[strong])
    # literal (diag-matmul) oracle agrees with the vectorised one on a few trials
    for t in (0, 5, n_trials - 1):
        a = oo.sync_trial(rx, 16 + t * stride, prm, zc, bank, literal=True)
        assert a[0] == rd[t] and abs(a[1] - rp[t]) < 1e-9


def test_chanest_and_symbol_list_demod_match_oracle():
    m = make(dict(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=240, m0=2, d=3, bps=4, snr=40,
                  noise_var=0.02))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    bits, rx = synth(cfg, prm, 3, 11)
    d_rx = torch.from_numpy(rx).cuda()
    L = cfg.sym_len
    pos = np.array([cfg.cp_len - 2 + f * cfg.frame_len for f in range(3)], dtype=np.int64)
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    dl, hs, tims, esfs = [], [], [], []
    for p in pos:
        d_i, d_v, sd = oo.sync_trial(rx, int(p), prm, zc, bank)
        ce, ce1, dr = oo.chan_est(sd, d_i, prm, zc, bank)
        tim, esf = oo.chan_est_aux(ce, ce1, dr, prm)
        dl.append(d_i), hs.append(ce1), tims.append(tim), esfs.append(esf)
    h, ht, se = m.chanest(d_rx, torch.from_numpy(pos).cuda(), torch.tensor(dl, dtype=torch.int32).cuda(), want_aux=True)
    assert rel_err(h.cpu().numpy(), np.array(hs)) < RTOL
    assert rel_err(ht.cpu().numpy(), np.array(tims)) < RTOL
    assert rel_err(se.cpu().numpy(), np.array(esfs)) < RTOL
    # every data symbol, shuffled, plus one window that runs off the end of the buffer (zero padded)
    starts, delays, rows, refs = [], [], [], []
    for f in range(3):
        for n in range(cfg.d):
            starts.append(int(pos[f] + cfg.m0 * L + n * L)), delays.append(dl[f]), rows.append(f)
    starts.append(len(rx) - 100), delays.append(3), rows.append(1)
    order = np.random.default_rng(0).permutation(len(starts))
    starts, delays, rows = (np.array(a)[order] for a in (starts, delays, rows))
    for s, d, r in zip(starts, delays, rows):
        refs.append(oo.demod_symbol(rx, int(s), int(d), hs[r], prm))
    eq, llr, hard = m.demod(d_rx, torch.from_numpy(starts.astype(np.int64)).cuda(),
                            torch.from_numpy(delays.astype(np.int32)).cuda(),
                            torch.from_numpy(rows.astype(np.int32)).cuda(), h, want_llr=True, want_hard=True)
    refs = np.array(refs)
    assert rel_err(eq.cpu().numpy(), refs) < RTOL
    assert rel_err(llr.cpu().numpy(), oo.maxlog_llr(refs, 4, 1.0 / cfg.noise_var)) < RTOL
    assert np.array_equal(unpack_symbols(hard.cpu().numpy(), 240 * 4), oo.hard_bits(refs, 4).reshape(-1))


@pytest.mark.parametrize("bps", [1, 2, 4, 6])
def test_demapper_closed_forms_equal_brute_force(bps):
    """Feed arbitrary points straight through the demod epilogue (H = 1, no rotation) by building
    frequency-domain symbols; compares LLRs and hard bits with the constellation brute force."""
    m = make(dict(REF, bps=bps, snr=200, noise_var=0.37, equalizer="zf"))
    cfg = m.cfg
    rng = np.random.default_rng(bps)
    n_sym = 40
    pts = (rng.uniform(-1.6, 1.6, (n_sym, 60)) + 1j * rng.uniform(-1.6, 1.6, (n_sym, 60)))
    grid = np.zeros((n_sym, 64), complex)
    grid[:, oo.bins_used(60, 64)] = pts
    time = np.fft.ifft(grid, axis=1).astype(np.complex64).reshape(-1)
    hmat = np.zeros((1, 64), np.complex64)
    hmat[0, oo.bins_used(62, 64)] = 1.0
    eq, llr, hard = m.demod(torch.from_numpy(time).cuda(), torch.arange(n_sym, dtype=torch.int64).cuda() * 64,
                            torch.zeros(n_sym, dtype=torch.int32).cuda(), torch.zeros(n_sym, dtype=torch.int32).cuda(),
                            torch.from_numpy(hmat).cuda(), want_llr=True, want_hard=True)
    eq = eq.cpu().numpy()
    scale = np.sqrt(60 / np.sum(np.abs(pts) ** 2, axis=1, keepdims=True))       # the unit-power step
    assert rel_err(eq, pts * scale) < RTOL
    ref_llr = oo.maxlog_llr(eq.astype(complex), bps, 1.0 / 0.37)
    assert rel_err(llr.cpu().numpy(), ref_llr) < RTOL
    got = unpack_symbols(hard.cpu().numpy(), 60 * bps)
    ref_bits = oo.hard_bits(eq.astype(complex), bps).reshape(-1)
    margin = np.abs(ref_llr).reshape(-1) > 1e-3          # skip points sitting on a decision boundary
    assert margin.mean() > 0.99
    assert np.array_equal(got[margin], ref_bits[margin])


def test_host_entry_points_equal_device_path():
    m = make(dict(REF, bps=4, snr=100, noise_var=0.01))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    n_frames = 257
    bits, rx = synth(cfg, prm, n_frames, 5)
    dev = m.rx_frames(torch.from_numpy(rx).cuda(), n_frames)
    host = m.rx_frames_host(rx, n_frames)
    for k in ("delay", "peak", "locked", "eq", "llr", "hard"):
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k
    packed = pack_symbols(bits, n_frames * cfg.d, 240, cfg.hard_stride)
    t_dev = m.tx_modulate(torch.from_numpy(packed).cuda(), n_frames).cpu().numpy()
    t_host = m.tx_modulate_host(packed.reshape(-1), n_frames)
    assert np.array_equal(t_dev, t_host)


def test_counters_match_oracle_and_accumulate():
    m = make(dict(REF, bps=4, snr=100, noise_var=0.01))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    n_frames = 50
    bits, rx = synth(cfg, prm, n_frames, 8, noise=3e-2)
    got = m.rx_frames(torch.from_numpy(rx).cuda(), n_frames)
    packed = torch.from_numpy(pack_symbols(bits, n_frames * cfg.d, 240, cfg.hard_stride)).cuda()
    cnt = m.counters(packed, n_frames, bits_rx=got["hard"], eq=got["eq"], locked=got["locked"])
    rx_bits = unpack_symbols(got["hard"].cpu().numpy(), 240)
    ref = oo.counters(rx_bits, bits, got["eq"].cpu().numpy().reshape(-1), oo.bits_to_symbols(bits, 4))
    c = cnt.cpu().numpy()
    assert c[0] == ref[0] and c[1] == ref[1] and c[0] > 0
    assert abs(c[2] - ref[2]) < 1e-4 * ref[2] and abs(c[3] - ref[3]) < 1e-4 * ref[3]
    assert c[4] == got["locked"].sum().item() and c[5] == n_frames
    m.counters(packed, n_frames, bits_rx=got["hard"], eq=got["eq"], locked=got["locked"], cnt=cnt)
    assert np.allclose(cnt.cpu().numpy(), 2 * c)


def test_edge_cases():
    from ofdm_b200 import Modem, ModemConfig, OfdmError
    m = make(dict(REF, snr=100))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    # zero frames / zero trials are no-ops
    z = torch.zeros(1000, dtype=torch.complex64).cuda()
    out = m.rx_frames(z, 0)
    assert out["eq"].shape == (0, 60)
    # a single frame; a frame whose tail is cut off (zero padded like np.fft.fft does)
    bits, rx = synth(cfg, prm, 2, 21)
    cut = rx[:len(rx) - 37].copy()
    got = m.rx_frames(torch.from_numpy(cut).cuda(), 2)
    ref = oo.rx_frames(np.concatenate((cut, np.zeros(64, np.complex64))), 2, prm)
    assert rel_err(got["eq"].cpu().numpy().reshape(ref["eq"].shape), ref["eq"]) < RTOL
    # all-zero input: no lock, finite outputs
    got = m.rx_frames(z, 3)
    assert not got["locked"].any().item()
    assert torch.isfinite(torch.view_as_real(got["eq"])).all().item()
    # unsupported numerologies are refused with a message, not a crash
    for bad in (dict(nfft=48), dict(nfft=8192), dict(bps=3), dict(num_data_bins=64), dict(cp_len=64)):
        with pytest.raises(OfdmError):
            Modem(ModemConfig(**dict(REF, **bad)))


def test_full_size_round_trip_properties():
    """BASELINE configs[1] size: 333 334 frames / 1 000 002 data symbols of 16-QAM.  The oracle cannot
    run this in seconds, so check size-independent properties: TX -> clean channel -> RX returns the
    payload bit for bit; a 4096-symbol subsample agrees with the oracle; linearity of the LLRs in
    the noise scale."""
    m = make(dict(REF, bps=4, snr=100, gate=0.5, noise_var=0.01))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    n_frames = 333334
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    bits = torch.randint(0, 256, (n_frames * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda", generator=g)
    tx = m.tx_modulate(bits, n_frames)
    assert abs(float(torch.mean(torch.abs(tx) ** 2)) - 1.0) < 1e-2           # both normalisations -> unit power
    taps = torch.tensor([0.0, 1.0], dtype=torch.complex64).cuda()            # the reference's 'AWGN' channel: 1-sample delay
    rx = m.channel_awgn(tx, n_frames, taps, 0.0, 0)
    got = m.rx_frames(rx, n_frames)
    cnt = m.counters(bits, n_frames, bits_rx=got["hard"], eq=got["eq"], locked=got["locked"]).cpu().numpy()
    assert cnt[5] == n_frames and cnt[4] == n_frames
    assert cnt[1] == n_frames * 3 * 240
    # the reference's per-symbol unit-power step rescales QAM symbols by their sample power, which can
    # push outer points over a boundary even without noise: the oracle shows the same on the subsample
    assert int(got["delay"].min()) == 1 and int(got["delay"].max()) == 1
    sub = 1366                                                                 # frames -> 4098 data symbols
    rx_h = rx[:sub * cfg.frame_len].cpu().numpy()
    ref = oo.rx_frames(rx_h, sub, prm)
    assert rel_err(got["eq"][:sub * 3].cpu().numpy().reshape(ref["eq"].shape), ref["eq"]) < RTOL
    ref_bits = oo.hard_bits(ref["eq"].reshape(-1), 4)
    assert np.array_equal(unpack_symbols(got["hard"][:sub * 3].cpu().numpy(), 240), ref_bits)
    tx_bits = unpack_symbols(bits[:sub * 3 * cfg.hard_stride].cpu().numpy().reshape(sub * 3, -1), 240)
    assert cnt[0] / cnt[1] == pytest.approx(np.mean(ref_bits != tx_bits), rel=0.2, abs=2e-5)
    # QPSK at the same size is constant-modulus: zero errors end to end
    m2 = make(dict(REF, bps=2, snr=100, gate=0.5))
    bits2 = bits[:n_frames * 3 * m2.cfg.hard_stride].contiguous()
    rx2 = m2.channel_awgn(m2.tx_modulate(bits2, n_frames), n_frames, taps, 0.0, 0)
    got2 = m2.rx_frames(rx2, n_frames)
    cnt2 = m2.counters(bits2, n_frames, bits_rx=got2["hard"], eq=got2["eq"], locked=got2["locked"]).cpu().numpy()
    assert cnt2[0] == 0 and cnt2[1] == n_frames * 3 * 120 and cnt2[4] == n_frames
    assert np.sqrt(cnt2[2] / cnt2[3]) < 1e-3
    # LLRs scale linearly with 1/noise_var
    m3 = make(dict(REF, bps=2, snr=100, gate=0.5, noise_var=0.5))
    got3 = m3.rx_frames(rx2, 1000)
    assert torch.allclose(got3["llr"], got2["llr"][:3000] * 2.0, rtol=1e-5, atol=1e-6)


def test_on_device_loopback_counts_match_oracle_counters():
    """TX -> channel -> RX -> counters without leaving the device (ofdm_b200.loopback); the counter vector equals
    the one the oracle computes from the same tensors copied back."""
    from ofdm_b200 import ModemConfig
    from ofdm_b200.loopback import Loopback
    lb = Loopback(ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4, snr=100, gate=0.5,
                              noise_var=0.02))
    n_frames = 50
    r = lb.run(n_frames, seed=5, noise_sigma=0.1)
    s = r["summary"]
    assert s["frames"] == n_frames and s["frames_locked"] == n_frames and s["bits"] == n_frames * 3 * 240
    tx_bits = np.unpackbits(r["bits"].cpu().numpy().reshape(n_frames * 3, 30), axis=1).reshape(-1)
    rx_bits = np.unpackbits(r["out"]["hard"].cpu().numpy(), axis=1).reshape(-1)
    eq = r["out"]["eq"].cpu().numpy().reshape(-1)
    ref = oo.counters(rx_bits, tx_bits, eq, oo.bits_to_symbols(tx_bits, 4))
    assert s["bit_errors"] == ref[0] and 0 < s["ber"] < 0.2
    assert abs(s["evm_rms"] - np.sqrt(ref[2] / ref[3])) < 1e-4
    # noiseless reference channel: error free, and per-frame random taps are accepted
    assert lb.run(n_frames, seed=6)["summary"]["bit_errors"] == 0
    taps = lb.taps_random(n_frames, [0.2, 0.7, 0.1], 3)
    assert lb.run(n_frames, seed=7, taps=taps, noise_sigma=0.01)["summary"]["frames"] == n_frames


@pytest.mark.parametrize("kw,n_frames", [
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2), 40),
    (dict(nfft=256, cp_len=20, num_synch_bins=254, num_data_bins=200, m0=1, d=2, bps=4), 9),
    (dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=1, bps=2), 2),
], ids=["N64", "N256", "N2048"])
def test_time_domain_acquisition_equals_transform_search(kw, n_frames, monkeypatch):
    """Stride-1 search of the utsa configuration: the time-domain kernel (no transform per trial) against the
    per-trial transform kernel on the same capture, incl. windows that run off the end of the buffer, and against
    the oracle's derivation on a stretch of trials."""
    m = make(dict(kw, snr=100, compat="utsa", zc_root=23))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    _, rx = synth(cfg, prm, n_frames, 5, noise=1e-3)
    d_rx = torch.from_numpy(rx).cuda()
    n_trials = len(rx) - cfg.cp_len - 7                      # the last trials read past the end (zero padding)
    monkeypatch.setenv("OFDM_SYNC_TD", "0")
    pk0, dl0 = (a.cpu().numpy() for a in m.sync_search(d_rx, cfg.cp_len, 1, n_trials))
    strong = pk0 > 0.3 * prm.mm
    assert strong.sum() >= n_frames
    n_chk = min(300, n_trials)
    rp, rd = oo.sync_search_time_domain(rx, cfg.cp_len, 1, n_chk, prm)
    s2 = rp > 0.3 * prm.mm
    for mode in ("1", "2", "3"):                             # plain / register-blocked time-domain kernels, overlap-save
        monkeypatch.setenv("OFDM_SYNC_TD", mode)
        pk1, dl1 = (a.cpu().numpy() for a in m.sync_search(d_rx, cfg.cp_len, 1, n_trials))
        assert rel_err(pk1, pk0) < RTOL, mode
        assert np.array_equal(dl1[strong], dl0[strong]), mode
        assert np.mean(dl1 == dl0) > 0.99, mode              # elsewhere only near-ties between noise lags may differ
        assert rel_err(pk1[:n_chk], rp) < RTOL, mode
        assert np.array_equal(dl1[:n_chk][s2], rd[s2]), mode


@pytest.mark.parametrize("kw,bank,continuous", [
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=2, d=2, bps=2), None, False),
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2), "small", False),
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=2, d=2, bps=2), "large", False),
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=2, d=2, bps=2), "large", True),
    (dict(nfft=256, cp_len=24, num_synch_bins=254, num_data_bins=200, m0=2, d=2, bps=4), "large", True),
    (dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=1, bps=2), "small", False),
], ids=["N64_m2", "N64_bank", "N64_m2_bank", "N64_m2_bank_cont", "N256_m2_bank_cont", "N2048_bank"])
def test_general_time_domain_acquisition(kw, bank, continuous, monkeypatch):
    """M0 = 2 and / or the frequency-offset hypothesis bank (SynchEstAndFO.py:247-277) on the time-domain search:
    against the per-trial transform kernel over a whole capture (windows running off the end included) and against
    the oracle's derivation (sync_search_time_domain with nus) on a stretch of trials."""
    m = make(dict(kw, snr=100, compat="utsa", zc_root=23))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    n_frames = 2 if cfg.nfft >= 2048 else 6
    _, rx = synth(cfg, prm, n_frames, 9, noise=1e-3)
    fs = 960000.0
    step = 0.3 * fs / cfg.nfft
    nus = None
    if bank:
        fo = [-75.0, -25.0, 0.0, 50.0] if bank == "small" else [-step, 0.0, step]
        rx = (rx * np.exp(-2j * np.pi * fo[-1] / fs * np.arange(len(rx)))).astype(np.complex64)
        nus = [f / fs for f in fo]
        m.set_cfo_bank(fo, fs, phase_continuous=continuous)
    d_rx = torch.from_numpy(rx).cuda()
    n_trials = len(rx) - cfg.cp_len - 7
    monkeypatch.setenv("OFDM_SYNC_TDG", "0")
    r0 = [a.cpu().numpy() for a in m.sync_search(d_rx, cfg.cp_len, 1, n_trials)]
    monkeypatch.setenv("OFDM_SYNC_TDG", "1")
    r1 = [a.cpu().numpy() for a in m.sync_search(d_rx, cfg.cp_len, 1, n_trials)]
    strong = r0[0] > 0.3 * prm.mm
    assert strong.sum() >= n_frames
    assert rel_err(r1[0], r0[0]) < RTOL
    assert np.array_equal(r1[1][strong], r0[1][strong]) and np.mean(r1[1] == r0[1]) > 0.98
    if bank == "large":                                     # hypotheses a third of a bin apart: one clear winner
        assert np.array_equal(r1[2][strong], r0[2][strong]) and np.mean(r1[2] == r0[2]) > 0.95
        if continuous:                                       # (the window-local phasor adds the two repeats incoherently)
            top = r0[0] > 0.9 * r0[0].max()                   # the trials that land on a frame start
            assert top.sum() >= n_frames - 1 and np.mean(r1[2][top] == 2) > 0.9
    n_chk = min(200 if cfg.nfft <= 256 else 40, n_trials)
    ref = oo.sync_search_time_domain(rx, cfg.cp_len, 1, n_chk, prm, nus=nus, continuous=continuous)
    s2 = ref[0] > 0.3 * prm.mm
    assert rel_err(r1[0][:n_chk], ref[0]) < RTOL
    assert np.array_equal(r1[1][:n_chk][s2], ref[1][s2])
    if bank == "large":
        assert np.array_equal(r1[2][:n_chk][s2], ref[2][s2])
    elif bank:
        # the reference's own example bank (range(-75, 75, 25) Hz at 960 kHz, SynchEstAndFO.py:157) turns a window by
        # hundredths of a radian: the hypotheses' peaks differ in the 5th digit, so the winner is only defined up to
        # rounding -- the chosen hypothesis must be within the tolerance of the best one
        zc, dbank = oo.zc_for(prm), oo.delay_bank(prm)
        cfo = oo.cfo_bank(fo, fs, cfg.nfft)
        for t in np.flatnonzero(s2)[:25]:
            per_h = [r[1] for r in oo.sync_trial_cfo(rx, cfg.cp_len + int(t), prm, cfo, zc, dbank)[3]]
            assert per_h[r1[2][t]] > (1 - RTOL) * max(per_h)
