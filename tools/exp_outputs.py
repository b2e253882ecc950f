import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ofdm_b200 import Modem, ModemConfig
from quick_bench import time_it
for name, cfg, nf in [("n64", ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4, snr=100, noise_var=0.01), 333334),
                      ("lte", ModemConfig(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6, snr=100, noise_var=0.01), 4096)]:
    m = Modem(cfg)
    bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
    tx = m.tx_modulate(bits, nf)
    taps = torch.tensor([0.0, 1.0], dtype=torch.complex64, device="cuda")
    rx = m.channel_awgn(tx, nf, taps, 0.02, 1)
    nsym = nf * cfg.d
    for llr, hard in [(True, True), (True, False), (False, True), (False, False)]:
        out = m.rx_frames(rx, nf, want_llr=llr, want_hard=hard)
        t, tmin = time_it(lambda: m.rx_frames(rx, nf, out=out))
        b = nf * cfg.frame_len * 8 + nsym * cfg.num_data_bins * 8 + (nsym * cfg.num_data_bins * cfg.bps * 4 if llr else 0) + (nsym * cfg.hard_stride if hard else 0)
        print("%s llr=%d hard=%d: %.3f ms  %.0f GB/s (%.2f GB)" % (name, llr, hard, t, b / t / 1e6, b / 1e9))
    # pure copy of the same volume for reference
    a = torch.empty(int(1.16e9), dtype=torch.uint8, device="cuda"); c = torch.empty_like(a)
    t, _ = time_it(lambda: c.copy_(a))
    print("copy 1.16GB->1.16GB: %.3f ms %.0f GB/s" % (t, 2 * a.numel() / t / 1e6))
    a = torch.empty(int(0.69e9), dtype=torch.uint8, device="cuda"); c = torch.empty(int(1.42e9), dtype=torch.uint8, device="cuda")
    def rw():
        c[:a.numel()].copy_(a); c[a.numel():].zero_()
    t, _ = time_it(rw)
    print("read 0.69 + write 1.42 GB (2 kernels): %.3f ms %.0f GB/s" % (t, (a.numel() + c.numel()) / t / 1e6))
