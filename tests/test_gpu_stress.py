"""Race evidence without the sanitizer (compute-sanitizer is closed on this GPU pool): every kernel family, every
transform size and modulation, launched many times on the same input with all SMs loaded for several waves -- the
outputs must be bit-identical from launch to launch.

The kernels share per-group scratch between steps and rely on the exchange barriers to order its reuse (the reduction
slots redf / redu in fft_core.cuh: group_sum / group_max_u64, the row buffers of the transmitter, the staged LLR / eq rows
of the bulk epilogue, named barriers shared by the groups of a CTA).  A missing barrier shows up as a value that depends
on warp timing; with the grid oversubscribed and tens of repeats per configuration such a dependence does not stay
hidden.  `OFDM_B200_LIB=variants/<tag>.so` reruns the same loop on the single-buffer (OFDM_SB_*) and two-deep prefetch
(OFDM_PF_DEPTH=2) builds (tools/build_variant.sh; profiles/README.md records the runs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from test_gpu_parity import make  # noqa: E402

REPS = 30
SIZES = [(64, 16), (128, 32), (256, 64), (512, 36), (1024, 72), (2048, 144), (4096, 288)]


def frames_for(nfft, d):
    # a few waves of the persistent grid: 148 SMs x 4 CTAs x max(1, 128 / (nfft / 8)) groups per CTA
    groups = 148 * 4 * max(1, 1024 // nfft)
    return max(64, int(2.3 * groups))


def same(a, b):
    return a is None or torch.equal(a.view(torch.uint8).reshape(-1), b.view(torch.uint8).reshape(-1))


def build_batch(m, n_frames, seed, sigma=0.05):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    cfg = m.cfg
    bits = torch.randint(0, 256, (n_frames * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda", generator=g)
    tx = m.tx_modulate(bits, n_frames)
    taps = torch.tensor([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398], dtype=torch.complex64, device="cuda")
    rx = m.channel_awgn(tx, n_frames, taps / torch.linalg.norm(taps), sigma, seed)
    return bits, tx, rx


@pytest.mark.parametrize("bps", [1, 2, 4, 6])
@pytest.mark.parametrize("nfft,cp", SIZES, ids=["N%d" % s[0] for s in SIZES])
def test_tx_and_framed_rx_are_bit_identical_across_launches(nfft, cp, bps):
    ndb = int(0.9375 * nfft) // 4 * 4
    m = make(dict(nfft=nfft, cp_len=cp, num_synch_bins=nfft - 2, num_data_bins=ndb, m0=1, d=3, bps=bps, snr=100, gate=0.5,
                  noise_var=0.02))
    nf = frames_for(nfft, 3)
    bits, tx0, rx = build_batch(m, nf, 17 + bps)
    ref = m.rx_frames(rx, nf, want_h=True)
    ref = {k: (None if v is None else v.clone()) for k, v in ref.items()}
    for rep in range(REPS):
        assert same(m.tx_modulate(bits, nf), tx0), ("tx", rep)
        got = m.rx_frames(rx, nf, want_h=True)
        for k in ("delay", "peak", "locked", "h", "eq", "llr", "hard"):
            assert same(got[k], ref[k]), (k, rep)


@pytest.mark.parametrize("kw", [
    dict(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6),   # bulk rows + CP pattern
    dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=6, bps=4),                             # sparse epilogue, 16 points per thread
    dict(nfft=1024, cp_len=72, num_synch_bins=1022, num_data_bins=600, m0=2, d=3, bps=6),                               # two synch symbols
    dict(nfft=64, cp_len=16, num_synch_bins=64, num_data_bins=60, m0=1, d=3, bps=2),                                    # +-N/2 alias
    dict(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=240, m0=1, d=3, bps=4, pilot_spacing=6),                # comb pilots
    dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=6, pilot_spacing=4, pilot_offset=1),
], ids=["lte_cpp_bulk", "lte_sparse16", "n1024_m0_2", "n64_alias", "n256_pilots", "n64_pilots"])
def test_framed_rx_variants_are_bit_identical_across_launches(kw, monkeypatch):
    if kw["bps"] == 6 and kw["nfft"] >= 1024:
        monkeypatch.setenv("OFDM_BULK_MIN_N", "1024")        # the bulk-row epilogue (selectable, not the default any more)
    m = make(dict(kw, snr=100, gate=0.5, noise_var=0.02))
    nf = frames_for(kw["nfft"], kw["d"])
    bits, tx0, rx = build_batch(m, nf, 5)
    ref = {k: (None if v is None else v.clone()) for k, v in m.rx_frames(rx, nf).items()}
    for rep in range(REPS):
        assert same(m.tx_modulate(bits, nf), tx0), ("tx", rep)
        got = m.rx_frames(rx, nf)
        for k in ("delay", "peak", "locked", "eq", "llr", "hard"):
            assert same(got[k], ref[k]), (k, rep)
    # the output-set selection (no equalised symbols, hard bits only) writes the same LLRs / bits
    lean = m.rx_frames(rx, nf, want_eq=False)
    assert lean["eq"] is None and same(lean["llr"], ref["llr"]) and same(lean["hard"], ref["hard"])
    hard_only = m.rx_frames(rx, nf, want_eq=False, want_llr=False)
    assert same(hard_only["hard"], ref["hard"]) and same(hard_only["delay"], ref["delay"])


def test_hypothesis_bank_receiver_is_bit_identical_across_launches():
    m = make(dict(nfft=128, cp_len=32, num_synch_bins=126, num_data_bins=100, m0=2, d=2, bps=4, snr=100, gate=0.5, noise_var=0.02))
    m.set_cfo_bank([-2000.0, 0.0, 2000.0, 4000.0], 1.92e6)
    nf = frames_for(128, 2)
    bits, tx0, rx = build_batch(m, nf, 9)
    ref = {k: (None if v is None else v.clone()) for k, v in m.rx_frames(rx, nf).items()}
    for rep in range(REPS):
        got = m.rx_frames(rx, nf)
        for k in ("delay", "peak", "locked", "hyp", "eq", "llr", "hard"):
            assert same(got[k], ref[k]), (k, rep)


@pytest.mark.parametrize("nfft,cp,stride", [(64, 16, 1), (64, 16, 15), (512, 128, 1), (2048, 144, 1), (2048, 144, 143), (4096, 1024, 1)])
def test_search_estimate_and_symbol_list_kernels_are_bit_identical_across_launches(nfft, cp, stride, monkeypatch):
    ndb = int(0.9375 * nfft) // 4 * 4
    m = make(dict(nfft=nfft, cp_len=cp, num_synch_bins=nfft - 2, num_data_bins=ndb, m0=1, d=2, bps=4, snr=100, gate=0.5,
                  noise_var=0.02))
    cfg = m.cfg
    nf = max(24, (1 << 21) // cfg.frame_len)
    bits, tx0, rx = build_batch(m, nf, 3)
    n_trials = (rx.numel() - cfg.sym_len - nfft - cp - 1) // stride + 1
    for mode in (("3", "2", "1", "0") if stride == 1 else ("0",)):     # overlap-save / register-blocked / plain time-domain, per-trial transforms
        monkeypatch.setenv("OFDM_SYNC_TD", mode)
        pk0, dl0 = (t.clone() for t in m.sync_search(rx, cp, stride, n_trials))
        for rep in range(REPS // 3):
            pk, dl = m.sync_search(rx, cp, stride, n_trials)
            assert same(pk, pk0) and same(dl, dl0), (mode, rep)
    pos = torch.arange(nf, device="cuda", dtype=torch.int64) * cfg.frame_len + cp - 1
    delay = torch.full((nf,), 2, device="cuda", dtype=torch.int32)
    h0, ht0, se0 = (t.clone() for t in m.chanest(rx, pos, delay, want_aux=True))
    starts = (pos.repeat_interleave(cfg.d) + cfg.sym_len * (1 + torch.arange(cfg.d, device="cuda").repeat(nf))).contiguous()
    rows = torch.arange(nf, device="cuda", dtype=torch.int32).repeat_interleave(cfg.d).contiguous()
    dls = delay.repeat_interleave(cfg.d).contiguous()
    e0, l0, b0 = (t.clone() for t in m.demod(rx, starts, dls, rows, h0, want_llr=True, want_hard=True))
    for rep in range(REPS):
        h, ht, se = m.chanest(rx, pos, delay, want_aux=True)
        assert same(h, h0) and same(ht, ht0) and same(se, se0), rep
        e, l, b = m.demod(rx, starts, dls, rows, h0, want_llr=True, want_hard=True)
        assert same(e, e0) and same(l, l0) and same(b, b0), rep


@pytest.mark.parametrize("kw", [
    dict(nfft=1024, cp_len=72, num_synch_bins=1022, num_data_bins=600, m0=1, d=3, bps=4),
    dict(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6),
    dict(nfft=2048, cp_len=512, num_synch_bins=2046, num_data_bins=1920, m0=2, d=3, bps=2),
    dict(nfft=4096, cp_len=1024, num_synch_bins=4094, num_data_bins=3840, m0=1, d=3, bps=6),
], ids=["n1024", "lte_cpp", "n2048_m0_2", "n4096"])
def test_wide_transmitter_is_bit_identical_across_launches(kw, monkeypatch):
    """tx_wide.cuh: one exchange buffer per group, reused symbol after symbol behind two group barriers, and (N = 4096)
    alternating reduction slots between the two warps of a group."""
    monkeypatch.setenv("OFDM_TX_WIDE", "2")
    m = make(dict(kw, snr=100))
    nf = frames_for(kw["nfft"], kw["d"])
    g = torch.Generator(device="cuda")
    g.manual_seed(41)
    bits = torch.randint(0, 256, (nf * m.cfg.d * m.cfg.hard_stride,), dtype=torch.uint8, device="cuda", generator=g)
    tx0 = m.tx_modulate(bits, nf).clone()
    for rep in range(REPS):
        assert same(m.tx_modulate(bits, nf), tx0), rep


@pytest.mark.parametrize("m0,bank", [(2, False), (1, True), (2, True)], ids=["m0_2", "bank", "m0_2_bank"])
def test_general_time_domain_search_is_bit_identical_across_launches(m0, bank):
    """rx_sync_search_tdg_kernel: per-hypothesis tap tables rebuilt in shared memory between barriers, 64-bit shared
    atomicMax per trial offset (order-independent), running best kept in registers."""
    m = make(dict(nfft=256, cp_len=24, num_synch_bins=254, num_data_bins=200, m0=m0, d=2, bps=4, snr=100, gate=0.5,
                  noise_var=0.02, compat="utsa", zc_root=23))
    if bank:
        m.set_cfo_bank([-3000.0, 0.0, 1500.0, 3000.0], 1.92e6, phase_continuous=(m0 == 2))
    nf = max(24, (1 << 20) // m.cfg.frame_len)
    bits, tx0, rx = build_batch(m, nf, 3)
    n_trials = rx.numel() - m0 * m.cfg.sym_len - 256
    r0 = [t.clone() for t in m.sync_search(rx, 24, 1, n_trials)]
    for rep in range(REPS // 3):
        r = m.sync_search(rx, 24, 1, n_trials)
        assert all(same(a, b) for a, b in zip(r, r0)), rep


@pytest.mark.parametrize("nfft,cp,bps", [(512, 128, 4), (1024, 256, 6), (2048, 512, 2), (4096, 1024, 4), (64, 16, 6)])
def test_output_selection_on_dense_grids_writes_identical_values(nfft, cp, bps):
    """The dense receiver has two code forms for the optional outputs (csrc/kernels.cuh data_epilogue: the variant with
    the eq switch compiled out, launched for N = 512..2048 when the equalised symbols are wanted, and the per-output bin
    masks everywhere else): every output that both calls produce must be bit-identical, and nothing may be written
    through a pointer that was not passed."""
    ndb = int(0.9375 * nfft) // 4 * 4
    m = make(dict(nfft=nfft, cp_len=cp, num_synch_bins=nfft - 2, num_data_bins=ndb, m0=1, d=3, bps=bps, snr=100, gate=0.5,
                  noise_var=0.02))
    nf = frames_for(nfft, 3)
    bits, tx0, rx = build_batch(m, nf, 23)
    full = {k: (None if v is None else v.clone()) for k, v in m.rx_frames(rx, nf).items()}
    lean = m.rx_frames(rx, nf, want_eq=False)
    assert lean["eq"] is None and same(lean["llr"], full["llr"]) and same(lean["hard"], full["hard"])
    eq_only = m.rx_frames(rx, nf, want_llr=False, want_hard=False)
    assert eq_only["llr"] is None and eq_only["hard"] is None and same(eq_only["eq"], full["eq"])
    hard_only = m.rx_frames(rx, nf, want_eq=False, want_llr=False)
    assert same(hard_only["hard"], full["hard"]) and same(hard_only["delay"], full["delay"]) and same(hard_only["peak"], full["peak"])
