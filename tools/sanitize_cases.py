"""Small cases of every kernel for compute-sanitizer (memcheck): short, ragged, edge-of-buffer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
from ofdm_b200 import Modem, ModemConfig
for kw in (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4),
           dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2),     # 15-byte hard-bit rows
           dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=6),     # 45-byte rows, staged LLRs
           dict(nfft=64, cp_len=16, num_synch_bins=64, num_data_bins=60, m0=2, d=2, bps=6),
           dict(nfft=128, cp_len=9, num_synch_bins=126, num_data_bins=118, m0=1, d=3, bps=1),
           dict(nfft=512, cp_len=36, num_synch_bins=510, num_data_bins=480, m0=1, d=2, bps=2, pilot_spacing=6),
           dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=3, bps=6),
           dict(nfft=4096, cp_len=288, num_synch_bins=4094, num_data_bins=3840, m0=1, d=1, bps=4)):
    cfg = ModemConfig(snr=100, gate=0.5, **kw)
    m = Modem(cfg)
    nf = 37
    # exactly-sized payload buffer (tail guard of the word loads) taken from the middle of a bigger one (odd alignment)
    # (every phase class of the transmitter's 16-byte row chunks)
    big = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride + 16,), dtype=torch.uint8, device="cuda")
    for off in (15, 8, 5, 1):
        bits = big[off:off + nf * cfg.d * cfg.hard_stride]
        tx = m.tx_modulate(bits.contiguous(), nf)
        tx2 = m.tx_modulate(bits.clone(), nf)
        assert torch.equal(tx, tx2), off
    rx = m.channel_awgn(tx, nf, torch.tensor([0.2, 1.0, 0.1j], dtype=torch.complex64, device="cuda"), 0.01, 5)
    cut = rx[: rx.numel() - 41].contiguous()                     # last frame runs off the buffer
    out = m.rx_frames(cut, nf, want_h=True)
    peak, delay = m.sync_search(cut, cfg.cp_len, cfg.stride, 50)
    pos = torch.tensor([cfg.cp_len, cfg.cp_len + cfg.frame_len], dtype=torch.int64, device="cuda")
    h, ht, se = m.chanest(cut, pos, torch.tensor([1, 1], dtype=torch.int32, device="cuda"), want_aux=True)
    st = torch.tensor([cfg.cp_len + cfg.sym_len, cut.numel() - 10, -5], dtype=torch.int64, device="cuda")
    eq, llr, hard = m.demod(cut, st, torch.zeros(3, dtype=torch.int32, device="cuda"),
                            torch.tensor([0, 1, 1], dtype=torch.int32, device="cuda"), h, want_llr=True, want_hard=True)
    cnt = m.counters(bits.contiguous(), nf, bits_rx=out["hard"], eq=out["eq"], locked=out["locked"])
    if cfg.bps == 2 and not cfg.pilot_spacing:
        m.bit_recovery(out["eq"].reshape(-1))
    if not cfg.pilot_spacing:                                    # frequency-offset hypothesis bank, both phase modes
        for cont in (False, True):
            m.set_cfo_bank([-1000.0, 0.0, 1500.0], 1e6, phase_continuous=cont)
            o2 = m.rx_frames(cut, nf)
            pk, dl, hy = m.sync_search(cut, cfg.cp_len, cfg.stride, 20)
            m.chanest(cut, pos, torch.tensor([1, 1], dtype=torch.int32, device="cuda"), want_aux=True,
                      hyp=torch.tensor([2, 0], dtype=torch.int32, device="cuda"))
            m.demod(cut, st, torch.zeros(3, dtype=torch.int32, device="cuda"), torch.tensor([0, 1, 1], dtype=torch.int32, device="cuda"),
                    h, want_llr=True, want_hard=True, hyp=torch.tensor([1, 2, 0], dtype=torch.int32, device="cuda"))
        m.set_cfo_bank([], 1e6)
    torch.cuda.synchronize()
    print("ok", kw["nfft"], kw["bps"], float(cnt[0]) / max(float(cnt[1]), 1))
