"""Every kernel on small ragged cases: odd payload alignment, frames cut off by the end of the buffer,
windows before/after the buffer, all modulations.  (compute-sanitizer is closed on the GPU pool, so the
bounds are exercised here and results compared between differently aligned copies of the same data.)"""
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu
pytest.importorskip("torch")


def test_edge_case_driver_runs_clean():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "tools", "sanitize_cases.py"), run_name="__main__")


GUARD_CASES = [
    dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2),      # 15-byte rows: aligned-word path
    dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=6),      # 45-byte rows + staged LLR rows
    dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=58, m0=1, d=2, bps=1),      # 8-byte rows holding 58 bits
    dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4),      # 30-byte rows: 16-bit path
    dict(nfft=128, cp_len=32, num_synch_bins=126, num_data_bins=118, m0=1, d=2, bps=2),   # 30-byte rows, 118 bins
    dict(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=250, m0=2, d=1, bps=6),   # 188-byte rows, bits end mid-byte
    dict(nfft=1024, cp_len=72, num_synch_bins=1022, num_data_bins=1002, m0=1, d=2, bps=2),  # 251-byte rows
]


@pytest.mark.parametrize("kw", GUARD_CASES, ids=lambda k: "N%d_b%d_n%d" % (k["nfft"], k["bps"], k["num_data_bins"]))
@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_outputs_stay_inside_their_buffers(kw, shift):
    """Every output of the framed receiver and of the symbol-list demod is written into a window of a larger
    sentinel-filled buffer that starts `shift` elements in; afterwards the window must equal a reference run into
    plain buffers and every sentinel outside it must be intact.  Covers the word-wide / staged store paths."""
    import numpy as np
    import torch
    from ofdm_b200 import Modem, ModemConfig
    cfg = ModemConfig(snr=100, gate=0.5, **kw)
    m = Modem(cfg)
    nf = 11
    bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
    tx = m.tx_modulate(bits, nf)
    rx = m.channel_awgn(tx, nf, torch.tensor([0.2, 1.0, 0.1j], dtype=torch.complex64, device="cuda"), 0.01, 5)
    ref = m.rx_frames(rx, nf)
    nd = nf * cfg.d
    G = 64

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G + 3,), fill, dtype=dtype, device="cuda")
        return buf, buf[G + shift:G + shift + n]

    raw, win = {}, {}
    for name, n, dt, fill in (("delay", nf, torch.int32, -77), ("peak", nf, torch.float32, -7.5),
                              ("locked", nf, torch.uint8, 0xA5), ("eq", nd * cfg.payload_bins * 2, torch.float32, -7.5),
                              ("llr", nd * cfg.payload_bins * cfg.bps, torch.float32, -7.5),
                              ("hard", nd * cfg.hard_stride, torch.uint8, 0xA5)):
        raw[name], win[name] = guarded(n, dt, fill)
    # complex64 / float4 stores need their natural alignment: eq and llr windows may only move by whole 16 bytes
    for name in ("eq", "llr"):
        raw[name], _ = guarded(win[name].numel(), torch.float32, -7.5)
        win[name] = raw[name][G + 4 * shift:G + 4 * shift + win[name].numel()]
    out = dict(delay=win["delay"], peak=win["peak"], locked=win["locked"],
               eq=torch.view_as_complex(win["eq"].view(nd, cfg.payload_bins, 2)),
               llr=win["llr"].view(nd, cfg.payload_bins * cfg.bps), hard=win["hard"].view(nd, cfg.hard_stride), h=None, hyp=None)
    m.rx_frames(rx, nf, out=out)
    torch.cuda.synchronize()
    for name in ("delay", "peak", "locked", "hard", "llr"):
        assert torch.equal(win[name].reshape(-1), ref[name].reshape(-1)), name
    assert torch.equal(torch.view_as_real(out["eq"]).reshape(-1), torch.view_as_real(ref["eq"]).reshape(-1))
    for name, fill in (("delay", -77), ("peak", -7.5), ("locked", 0xA5), ("eq", -7.5), ("llr", -7.5), ("hard", 0xA5)):
        b, w = raw[name], win[name]
        lo = w.data_ptr() - b.data_ptr()
        lo //= b.element_size()
        outside = torch.cat((b[:lo], b[lo + w.numel():]))
        assert bool((outside == fill).all()), "guard band of %s overwritten" % name


TX_GUARD_CASES = [
    dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4),
    dict(nfft=512, cp_len=36, num_synch_bins=510, num_data_bins=300, m0=2, d=2, bps=6),
    dict(nfft=1024, cp_len=256, num_synch_bins=1022, num_data_bins=960, m0=1, d=2, bps=2),
    dict(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6),
    dict(nfft=4096, cp_len=1024, num_synch_bins=4094, num_data_bins=3840, m0=1, d=2, bps=4),
]


@pytest.mark.parametrize("kw", TX_GUARD_CASES, ids=lambda k: "N%d_b%d_n%d" % (k["nfft"], k["bps"], k["num_data_bins"]))
@pytest.mark.parametrize("wide", ["0", "2"])
def test_transmitter_output_stays_inside_its_buffer(kw, wide, monkeypatch):
    """Both transmitter kernels (16 and 64 points per thread; OFDM_TX_WIDE) write the waveform -- body and cyclic-prefix
    copies, incl. the longer first prefix of a slot -- into a window of a sentinel-filled buffer: the window must equal a
    run into a plain buffer and every sentinel around it must be intact.  The grid is oversubscribed by a non-multiple
    of the group count, so the tail groups that redo the last symbol are exercised as well."""
    import numpy as np
    import torch
    from ofdm_b200 import Modem, ModemConfig
    monkeypatch.setenv("OFDM_TX_WIDE", wide)
    cfg = ModemConfig(snr=100, **kw)
    m = Modem(cfg)
    nf = 2 * 148 * 2 // (cfg.m0 + cfg.d) + 3
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda", generator=g)
    ref = m.tx_modulate(bits, nf).clone()
    n = nf * cfg.frame_len
    pad = 2 * cfg.nfft + 5
    sent = complex(12345.0, -54321.0)
    big = torch.full((n + 2 * pad,), sent, dtype=torch.complex64, device="cuda")
    m.tx_modulate(bits, nf, out=big[pad:pad + n])
    torch.cuda.synchronize()
    assert torch.equal(big[pad:pad + n], ref)
    assert bool((big[:pad] == sent).all()) and bool((big[pad + n:] == sent).all())
    assert float(ref.abs().max()) < 100.0 and bool(torch.isfinite(ref.abs()).all())
