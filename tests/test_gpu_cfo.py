"""GPU parity of the frequency-offset hypothesis bank (reference arithmetic at
GRR/LEGACY/gr-ofdm-rx/python/SynchEstAndFO.py:196,:247-277,:331-333): framed receiver, sliding search,
per-item channel estimate / demod, and the legacy block's work() against its stateful oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import ofdm_oracle as oo  # noqa: E402
from oracle.blocks_oracle import SynchEstAndFOOracle  # noqa: E402
from test_gpu_parity import make, oracle_prm, pack_symbols, unpack_symbols, rel_err, RTOL  # noqa: E402


def synth_cfo(cfg, prm, n_frames, seed, fs, cfo_hz, noise=1e-4, lead=0, tail=0):
    """TX -> reference multipath -> carrier offset -cfo_hz (so hypothesis +cfo_hz undoes it) -> AWGN."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, n_frames * cfg.d * cfg.num_data_bins * cfg.bps)
    tx = oo.tx_modulate(bits, prm, n_frames)
    rx = oo.apply_channel(tx, oo.reference_channel(cfg.nfft))[:len(tx)]
    rx = np.concatenate((np.zeros(lead), rx, np.zeros(tail)))
    rx = rx * np.exp(-2j * np.pi * cfo_hz / fs * np.arange(len(rx)))
    return bits, oo.add_awgn(rx, noise, rng).astype(np.complex64)


FRAMED = [
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2), 960000.0),
    (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=36, m0=2, d=2, bps=6), 960000.0),
    (dict(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=240, m0=1, d=2, bps=4), 3840000.0),
    (dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=2, bps=4), 30720000.0),
]


@pytest.mark.parametrize("cont", [False, True], ids=["window_local", "phase_continuous"])
@pytest.mark.parametrize("kw,fs", FRAMED, ids=lambda v: "N%d_m%d" % (v["nfft"], v["m0"]) if isinstance(v, dict) else "")
def test_rx_frames_with_cfo_bank_matches_oracle(kw, fs, cont):
    m = make(dict(kw, snr=40, noise_var=0.05, sync_advance=2))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    step = 0.2 * fs / cfg.nfft                         # hypotheses 0.2 cycles per window apart: a clear winner
    fo = [-2 * step, -step, 0.0, step, 2 * step]
    m.set_cfo_bank(fo, fs, phase_continuous=cont)
    cfo = oo.cfo_bank(fo, fs, cfg.nfft)
    nus = [f / fs for f in fo] if cont else None
    n_frames = 7
    bits, rx = synth_cfo(cfg, prm, n_frames, 5, fs, step, noise=2e-4)
    got = m.rx_frames(torch.from_numpy(rx).cuda(), n_frames, want_h=True)
    ref = oo.rx_frames(rx, n_frames, prm, cfo=cfo, nus=nus)
    assert np.array_equal(got["hyp"].cpu().numpy(), ref["hyp"])
    if cont or cfg.m0 == 1:
        assert set(ref["hyp"]) == {3}
    assert np.array_equal(got["delay"].cpu().numpy(), ref["delay"])
    assert np.array_equal(got["locked"].cpu().numpy(), ref["locked"])
    if cont:
        assert ref["locked"].all()
    assert rel_err(got["peak"].cpu().numpy(), ref["peak"]) < RTOL
    assert rel_err(got["h"].cpu().numpy(), ref["h"]) < RTOL
    eq = got["eq"].cpu().numpy().reshape(ref["eq"].shape)
    assert rel_err(eq, ref["eq"]) < RTOL
    per_sym = cfg.num_data_bins * cfg.bps
    hb = unpack_symbols(got["hard"].cpu().numpy(), per_sym)
    ref_hb = oo.hard_bits(ref["eq"].reshape(-1), cfg.bps)
    # decisions that sit closer to a boundary than the fp32 error are not comparable
    llr = oo.maxlog_llr(ref["eq"].reshape(n_frames * cfg.d, -1), cfg.bps, 1.0).reshape(-1)
    safe = np.abs(llr) > 1e-3
    assert np.array_equal(hb[safe], ref_hb[safe])
    if cont and cfg.bps <= 4:                          # the continued phasor really undoes the offset
        assert np.array_equal(hb, bits)
    # removing the bank restores the plain receiver
    m.set_cfo_bank([], fs)
    plain = m.rx_frames(torch.from_numpy(rx).cuda(), n_frames)
    ref0 = oo.rx_frames(rx, n_frames, prm)
    assert plain["hyp"] is None
    assert rel_err(plain["eq"].cpu().numpy().reshape(ref0["eq"].shape), ref0["eq"]) < RTOL


def test_sliding_search_with_cfo_bank_matches_oracle():
    fs = 960000.0
    m = make(dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, snr=100, compat="rxofdm",
                  zc_root=23))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    step = 0.25 * fs / cfg.nfft
    fo = [-step, 0.0, step, 2 * step]
    m.set_cfo_bank(fo, fs)
    cfo = oo.cfo_bank(fo, fs, cfg.nfft)
    _, rx = synth_cfo(cfg, prm, 4, 9, fs, step, lead=5)
    stride = cfg.stride
    n_trials = (len(rx) - 80 - 64 - 16 - 1) // stride + 1
    peak, delay, hyp = (t.cpu().numpy() for t in m.sync_search(torch.from_numpy(rx).cuda(), 16, stride, n_trials))
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    ref = [oo.sync_trial_cfo(rx, 16 + t * stride, prm, cfo, zc, bank) for t in range(n_trials)]
    rp = np.array([r[2] for r in ref])
    assert rel_err(peak, rp) < RTOL
    strong = rp > 0.4 * prm.mm
    assert strong.sum() >= 4
    assert np.array_equal(delay[strong], np.array([r[1] for r in ref])[strong])
    assert np.array_equal(hyp[strong], np.array([r[0] for r in ref])[strong]) and 2 in set(hyp[strong])
    # elsewhere the winner is only defined where the two best hypotheses are not within rounding of each other
    vals = np.array([[h[1] for h in r[3]] for r in ref])
    top2 = np.sort(vals, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-3 * top2[:, 1]
    assert clear.sum() > n_trials // 2
    assert np.array_equal(hyp[clear], np.array([r[0] for r in ref])[clear])


def test_chanest_and_demod_take_a_hypothesis_per_item():
    fs = 3840000.0
    m = make(dict(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=240, m0=2, d=3, bps=4, snr=40, noise_var=0.02))
    cfg = m.cfg
    prm = oracle_prm(cfg)
    step = 0.2 * fs / cfg.nfft
    fo = [-step, 0.0, step]
    m.set_cfo_bank(fo, fs)
    cfo = oo.cfo_bank(fo, fs, cfg.nfft)
    _, rx = synth_cfo(cfg, prm, 3, 11, fs, step)
    d_rx = torch.from_numpy(rx).cuda()
    L = cfg.sym_len
    pos = np.array([cfg.cp_len - 2 + f * cfg.frame_len for f in range(3)], dtype=np.int64)
    hyps = np.array([2, 0, 1], dtype=np.int32)                 # deliberately not all the winner
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    dl, hs, tims, esfs = [], [], [], []
    for p, h in zip(pos, hyps):
        d_i, d_v, sd = oo.sync_trial(rx, int(p), prm, zc, bank, rot=cfo[h])
        ce, ce1, dr = oo.chan_est(sd, d_i, prm, zc, bank)
        tim, esf = oo.chan_est_aux(ce, ce1, dr, prm)
        dl.append(d_i), hs.append(ce1), tims.append(tim), esfs.append(esf)
    h, ht, se = m.chanest(d_rx, torch.from_numpy(pos).cuda(), torch.tensor(dl, dtype=torch.int32).cuda(), want_aux=True,
                          hyp=torch.from_numpy(hyps).cuda())
    assert rel_err(h.cpu().numpy(), np.array(hs)) < RTOL
    assert rel_err(ht.cpu().numpy(), np.array(tims)) < RTOL
    assert rel_err(se.cpu().numpy(), np.array(esfs)) < RTOL
    starts, delays, rows, hy, refs = [], [], [], [], []
    for f in range(3):
        for n in range(cfg.d):
            starts.append(int(pos[f] + cfg.m0 * L + n * L)), delays.append(dl[f]), rows.append(f), hy.append((f + n) % 3)
    starts.append(len(rx) - 100), delays.append(3), rows.append(1), hy.append(2)       # window past the end: zero padded
    for s, d, r, q in zip(starts, delays, rows, hy):
        refs.append(oo.demod_symbol(rx, int(s), int(d), hs[r], prm, rot=cfo[q]))
    eq, _, _ = m.demod(d_rx, torch.tensor(starts, dtype=torch.int64).cuda(), torch.tensor(delays, dtype=torch.int32).cuda(),
                       torch.tensor(rows, dtype=torch.int32).cuda(), h, hyp=torch.tensor(hy, dtype=torch.int32).cuda())
    assert rel_err(eq.cpu().numpy(), np.array(refs)) < RTOL
    # hyp=None leaves the windows alone even though a bank is installed
    eq0, _, _ = m.demod(d_rx, torch.tensor(starts[:3], dtype=torch.int64).cuda(),
                        torch.tensor(delays[:3], dtype=torch.int32).cuda(), torch.tensor(rows[:3], dtype=torch.int32).cuda(), h)
    ref0 = np.array([oo.demod_symbol(rx, int(s), int(d), hs[r], prm) for s, d, r in zip(starts[:3], delays[:3], rows[:3])])
    assert rel_err(eq0.cpu().numpy(), ref0) < RTOL


# M0 > 1 adds the preamble repeats up coherently, which the window-local phasor only allows at (near) zero offset
@pytest.mark.parametrize("case,cfo_hz", [(1, 3000.0), (4, 0.0), (6, 0.0)])
def test_synch_est_and_fo_block_matches_its_oracle(case, cfo_hz):
    import OFDMReceiver
    fo_range = [-3000.0, 0.0, 3000.0, 6000.0]
    blk = OFDMReceiver.SynchEstAndFO(case, fo_range, "/tmp/", "cest", 0)
    ora = SynchEstAndFOOracle(case, fo_range)
    prm = ora.prm
    n_frames = 10
    _, rx = synth_cfo(blk.cfg, prm, n_frames, 21 + case, ora.fs, cfo_hz, noise=1e-4, lead=9, tail=40)
    out_b = np.zeros(len(rx), np.complex64)
    out_o = np.zeros(len(rx), np.complex64)
    for call in range(2):
        assert blk.work([rx], [out_b]) == len(rx)
        assert ora.work(rx, out_o) == len(rx)
        assert np.array_equal(blk.time_synch_ref, ora.time_synch_ref), call
        assert blk.dmax_tmp_ind == ora.dmax_tmp_ind
        assert blk.last_written == ora.last_written
        assert rel_err(blk.est_chan_freq_P, ora.est_chan_freq_P) < RTOL
        assert rel_err(blk.est_data_freq, ora.est_data_freq) < RTOL
        assert rel_err(blk.est_chan_time, ora.est_chan_time) < RTOL
        assert rel_err(blk.est_synch_freq, ora.est_synch_freq) < RTOL
    n = ora.last_written
    assert n > 0 and rel_err(out_b[:n], out_o[:n]) < RTOL
    assert np.count_nonzero(ora.time_synch_ref[:, 0]) >= n_frames - 1


def test_cfo_bank_argument_checks():
    from ofdm_b200 import OfdmError
    m = make(dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2))
    with pytest.raises(OfdmError):
        m.set_cfo_bank(list(range(65)), 1e6)
    mp = make(dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, pilot_spacing=6))
    with pytest.raises(OfdmError):
        mp.set_cfo_bank([0.0, 100.0], 1e6)
