"""Scratch timing of the device-resident kernels (not the contract bench; see bench.py)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
sys.path.insert(0, ROOT)
from ofdm_b200 import Modem, ModemConfig  # noqa: E402


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


def run(name, cfg, n_frames, llr=True, hard=True):
    from bench import ClockSampler                      # SM clock and throttle reasons sampled over the timed region
    m = Modem(cfg)
    bits = torch.randint(0, 256, (n_frames * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
    tx = m.tx_modulate(bits, n_frames)
    taps = torch.tensor([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398], dtype=torch.complex64, device="cuda")
    taps = taps / torch.linalg.norm(taps)
    rx = m.channel_awgn(tx, n_frames, taps, 0.05, 2024)
    out = m.rx_frames(rx, n_frames, want_llr=llr, want_hard=hard)
    sampler = ClockSampler(0)
    sampler.start()
    t_tx, t_tx_min = time_it(lambda: m.tx_modulate(bits, n_frames, out=tx))
    t_rx, t_rx_min = time_it(lambda: m.rx_frames(rx, n_frames, out=out))
    clocks = sampler.stop()
    cnt = m.counters(bits, n_frames, bits_rx=out["hard"], eq=out["eq"], locked=out["locked"]).cpu().numpy()
    L = cfg.sym_len
    nsym = n_frames * cfg.d
    rx_bytes = n_frames * cfg.frame_len * 8 + nsym * cfg.num_data_bins * 8
    if llr:
        rx_bytes += nsym * cfg.num_data_bins * cfg.bps * 4
    if hard:
        rx_bytes += nsym * cfg.hard_stride
    tx_bytes = nsym * cfg.hard_stride + n_frames * cfg.frame_len * 8
    res = dict(name=name, n_frames=n_frames, data_syms=nsym,
               rx_ms=t_rx, rx_GBs=rx_bytes / t_rx / 1e6, rx_Msym_s=nsym / t_rx / 1e3,
               tx_ms=t_tx, tx_GBs=tx_bytes / t_tx / 1e6, tx_Msym_s=nsym / t_tx / 1e3,
               ber=cnt[0] / max(cnt[1], 1), evm=float(np.sqrt(cnt[2] / max(cnt[3], 1e-30))),
               locked=cnt[4] / max(cnt[5], 1), clocks=clocks)
    print(json.dumps(res))
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("focus",):
        for bps, nm in ((4, "16qam"), (2, "qpsk"), (6, "64qam")):
            run("N64_" + nm, ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=bps,
                                         snr=100, noise_var=0.01), 333334)
        run("LTE_64qam", ModemConfig(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13,
                                     bps=6, snr=100, noise_var=0.01), 4096)
        for n in (1024, 2048, 4096):
            ndb = int(0.9375 * n) // 4 * 4
            run("N%d_16qam" % n, ModemConfig(nfft=n, cp_len=n // 4, num_synch_bins=n - 2, num_data_bins=ndb, m0=1,
                                             d=3, bps=4, snr=100, noise_var=0.01), max(64, (1 << 26) // (4 * n * 5 // 4)))
    if which in ("q64",):
        for n in (64, 128, 256, 512, 1024, 2048, 4096):
            ndb = int(0.9375 * n) // 4 * 4
            run("N%d_64qam" % n, ModemConfig(nfft=n, cp_len=n // 4, num_synch_bins=n - 2, num_data_bins=ndb, m0=1,
                                             d=3, bps=6, snr=100, noise_var=0.01), max(64, (1 << 26) // (4 * n * 5 // 4)))
    if which in ("large",):
        run("LTE_64qam", ModemConfig(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13,
                                     bps=6, snr=100, noise_var=0.01), 4096)
        for n in (1024, 2048, 4096):
            ndb = int(0.9375 * n) // 4 * 4
            run("N%d_16qam" % n, ModemConfig(nfft=n, cp_len=n // 4, num_synch_bins=n - 2, num_data_bins=ndb, m0=1,
                                             d=3, bps=4, snr=100, noise_var=0.01), max(64, (1 << 26) // (4 * n * 5 // 4)))
    if which in ("ab",):                                    # round-2 A/B set: LTE, the N=4096 sweep points, two HBM-bound sizes
        run("LTE_64qam", ModemConfig(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13,
                                     bps=6, snr=100, noise_var=0.01), 4096)
        for bps, nm in ((2, "qpsk"), (4, "16qam"), (6, "64qam")):
            run("N4096_" + nm, ModemConfig(nfft=4096, cp_len=1024, num_synch_bins=4094, num_data_bins=3840, m0=1, d=3,
                                           bps=bps, snr=100, noise_var=0.01), 6553)
        run("N2048_16qam", ModemConfig(nfft=2048, cp_len=512, num_synch_bins=2046, num_data_bins=1920, m0=1, d=3,
                                       bps=4, snr=100, noise_var=0.01), 13107)
        run("N512_16qam", ModemConfig(nfft=512, cp_len=128, num_synch_bins=510, num_data_bins=480, m0=1, d=3,
                                      bps=4, snr=100, noise_var=0.01), 52428)
        run("N64_qpsk", ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2,
                                    snr=100, noise_var=0.01), 419430)
    if which in ("all", "n64"):
        run("N64_16qam", ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4,
                                     snr=100, noise_var=0.01), 333334)
        run("N64_qpsk", ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2,
                                    snr=100, noise_var=0.01), 333334)
    if which in ("all", "lte"):
        run("LTE_64qam", ModemConfig(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13,
                                     bps=6, snr=100, noise_var=0.01), 4096)
    if which in ("ltex",):                                  # LTE grid (59 % of the bins used) on the per-lane epilogues
        for bps, nm in ((2, "qpsk"), (4, "16qam")):
            run("LTE_" + nm, ModemConfig(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13,
                                         bps=bps, snr=100, noise_var=0.01), 4096)
        run("N2048_16qam", ModemConfig(nfft=2048, cp_len=512, num_synch_bins=2046, num_data_bins=1920, m0=1, d=3,
                                       bps=4, snr=100, noise_var=0.01), 6553)
        run("N2048_qpsk", ModemConfig(nfft=2048, cp_len=512, num_synch_bins=2046, num_data_bins=1920, m0=1, d=3,
                                      bps=2, snr=100, noise_var=0.01), 6553)
    if which in ("all", "sweep"):
        for n in (128, 256, 512, 1024, 4096):
            ndb = int(0.9375 * n) // 4 * 4
            run("N%d_16qam" % n, ModemConfig(nfft=n, cp_len=n // 4, num_synch_bins=n - 2, num_data_bins=ndb, m0=1,
                                             d=3, bps=4, snr=100, noise_var=0.01), max(64, (1 << 26) // (4 * n * 5 // 4)))
