"""Acquisition benchmark: the sliding Zadoff-Chu time search over a long capture with unknown frame timing
(SURVEY.md 8a rows a8-a10 -- the loop the reference spends its time in: one window FFT, unit-power step and
(CP+1)-delay bank per trial, GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:141-166 at stride 1 and
GRR/gr-RXOFDM/python/synch_and_chan_est.py:146-168 at stride CP-1).

Prints one JSON line per case: trials/s of `rx_sync_search_kernel` (CUDA events, device-resident capture),
the DRAM-side algorithmic rate (every sample read once, 8 bytes written per trial) and the L2-side window rate
(every trial reads M0*N samples), next to the oracle's literal trial (dense diag matmuls, what the reference
executes) and its vectorised form on one host core.  Not the contract bench (see bench.py)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench import ClockSampler  # noqa: E402
from ofdm_b200 import Modem, ModemConfig  # noqa: E402
from quick_bench import time_it  # noqa: E402
from oracle import ofdm_oracle as oo  # noqa: E402  (CPU baseline leg only)

CASES = {
    # name: (config, stride, samples in the capture)
    "n64_utsa_stride1": (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, snr=100,
                              compat="utsa", zc_root=23), 1, 1 << 24),
    "n64_rxofdm_stride15": (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, snr=100,
                                 compat="rxofdm", zc_root=23), 15, 1 << 26),
    "n256_stride1": (dict(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=240, m0=1, d=3, bps=2, snr=100), 1, 1 << 23),
    "n1024_stride1": (dict(nfft=1024, cp_len=256, num_synch_bins=1022, num_data_bins=960, m0=1, d=3, bps=2, snr=100), 1, 1 << 22),
    "lte_stride1": (dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6, snr=100),
                    1, 1 << 22),
    # the general time-domain kernel (M0 = 2 and / or the frequency-offset hypothesis bank): a 4th tuple entry is the
    # bank in cycles per sample; these cases also time the per-trial transform kernel on the same capture (OFDM_SYNC_TDG=0)
    "n64_m2_stride1": (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=2, d=2, bps=2, snr=100,
                            compat="utsa", zc_root=23), 1, 1 << 23),
    "n64_bank7_stride1": (dict(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=2, snr=100,
                               compat="utsa", zc_root=23), 1, 1 << 23, [f / 960000.0 for f in range(-75, 100, 25)]),
    "lte_bank3_stride1": (dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6, snr=100),
                          1, 1 << 21, [-1e-5, 0.0, 1e-5]),
    "lte_stride143": (dict(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6, snr=100),
                      143, 1 << 27),
}


def cpu_rate(cfg, rx, stride, literal, seconds):
    prm = oo.ModemParams(nfft=cfg.nfft, cp_len=cfg.cp_len, num_synch_bins=cfg.num_synch_bins,
                         num_data_bins=cfg.num_data_bins, m0=cfg.m0, d=cfg.d, bps=cfg.bps, snr=cfg.snr, gate=cfg.gate,
                         compat=cfg.compat, zc_root=cfg.zc_root)
    zc, bank = oo.zc_for(prm), oo.delay_bank(prm)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        oo.sync_trial(rx, cfg.cp_len + n * stride, prm, zc, bank, literal=literal)
        n += 1
    return n / (time.perf_counter() - t0)


def run(name, cpu_seconds):
    kw, stride, n_samp = CASES[name][:3]
    nus = CASES[name][3] if len(CASES[name]) > 3 else None
    cfg = ModemConfig(**kw)
    m = Modem(cfg)
    if nus:
        m.set_cfo_bank(nus, 1.0)
    nf = n_samp // cfg.frame_len
    bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
    tx = m.tx_modulate(bits, nf)
    taps = torch.tensor([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398], dtype=torch.complex64, device="cuda")
    rx = m.channel_awgn(tx, nf, taps / torch.linalg.norm(taps), 0.02, 11)
    n = rx.numel()
    n_trials = (n - cfg.m0 * cfg.sym_len - cfg.nfft - cfg.cp_len - 1) // stride + 1
    peak, delay = m.sync_search(rx, cfg.cp_len, stride, n_trials)[:2]
    sampler = ClockSampler(0)
    sampler.start()
    t_ms, _ = time_it(lambda: m.sync_search(rx, cfg.cp_len, stride, n_trials), iters=12)
    clocks = sampler.stop()
    pk = peak.cpu().numpy()
    # every frame start must show up as a gated peak at the trial that lands on it
    hits = int((pk > cfg.gate * cfg.mm).sum())
    res = dict(case=name, nfft=cfg.nfft, cp_len=cfg.cp_len, stride=stride, samples=n, trials=n_trials, ms=t_ms,
               trials_per_s=n_trials / t_ms * 1e3, samples_per_s=n / t_ms * 1e3,
               dram_GBs=(n * 8 + n_trials * 8) / t_ms / 1e6, window_GBs=n_trials * cfg.m0 * cfg.nfft * 8 / t_ms / 1e6,
               gated_trials=hits, frames=nf, clocks=clocks)
    if nus or cfg.m0 > 1:
        os.environ["OFDM_SYNC_TDG"] = "0"
        pk0 = m.sync_search(rx, cfg.cp_len, stride, n_trials)[0].cpu().numpy()
        t0_ms, _ = time_it(lambda: m.sync_search(rx, cfg.cp_len, stride, n_trials), iters=4, warm=1)
        os.environ["OFDM_SYNC_TDG"] = "1"
        res.update(m0=cfg.m0, hypotheses=len(nus) if nus else 0, ms_per_trial_transforms=t0_ms, speedup_vs_transforms=t0_ms / t_ms,
                   peak_rel_err_vs_transforms=float(np.max(np.abs(pk - pk0)) / np.max(np.abs(pk0))))
    if cpu_seconds > 0 and not nus:
        h = rx[: min(n, 1 << 20)].cpu().numpy()
        res["cpu_literal_trials_per_s_1core"] = cpu_rate(cfg, h, stride, True, cpu_seconds)
        res["cpu_vectorised_trials_per_s_1core"] = cpu_rate(cfg, h, stride, False, cpu_seconds)
        res["speedup_vs_literal_1core"] = res["trials_per_s"] / res["cpu_literal_trials_per_s_1core"]
    print(json.dumps(res), flush=True)
    del m, rx, tx, bits
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or list(CASES)
    secs = 0.0 if "--no-cpu" in sys.argv else 3.0
    for c in which:
        run(c, secs)
