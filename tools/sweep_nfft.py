"""BASELINE config 5: NFFT sweep 64..4096 x QPSK/16-QAM/64-QAM, soft-LLR demap on, throughput vs the
HBM roofline.  CP = N/4 (reference convention, SDRScript.py:96-97) or, with the argument `lte`, the LTE normal-CP
ratio 9N/128 rounded to the nearest sample (SURVEY.md 8d config 5 names both); ndb = 0.9375*N rounded to a multiple
of 4 (SDRScript.py:57-58), frames of 1 synch + 3 data symbols, ~1 GiB of input samples per point."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
from bench import ClockSampler  # noqa: E402
from ofdm_b200 import Modem, ModemConfig  # noqa: E402
from quick_bench import time_it  # noqa: E402

PEAK = 6542.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

CP_LTE = len(sys.argv) > 1 and sys.argv[1] == "lte"
rows = []
for n in (64, 128, 256, 512, 1024, 2048, 4096):
    for bps in (2, 4, 6):
        ndb = int(0.9375 * n) // 4 * 4
        cfg = ModemConfig(nfft=n, cp_len=(9 * n + 64) // 128 if CP_LTE else n // 4, num_synch_bins=n - 2, num_data_bins=ndb, m0=1, d=3, bps=bps, snr=100,
                          gate=0.5, noise_var=0.01)
        nf = (1 << 27) // cfg.frame_len                      # 2^27 samples = 1 GiB of complex64
        m = Modem(cfg)
        bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
        tx = m.tx_modulate(bits, nf)
        taps = torch.tensor([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398], dtype=torch.complex64, device="cuda")
        rx = m.channel_awgn(tx, nf, taps / torch.linalg.norm(taps), 0.01, 3)
        out = m.rx_frames(rx, nf)
        sampler = ClockSampler(0)
        sampler.start()
        t_rx, _ = time_it(lambda: m.rx_frames(rx, nf, out=out), iters=12)
        t_tx, _ = time_it(lambda: m.tx_modulate(bits, nf, out=tx), iters=12)
        clocks = sampler.stop()
        cnt = m.counters(bits, nf, bits_rx=out["hard"], eq=out["eq"], locked=out["locked"]).cpu().numpy()
        nsym = nf * cfg.d
        rx_b = nf * cfg.frame_len * 8 + nsym * ndb * (8 + 4 * bps) + nsym * cfg.hard_stride + nf * 9
        tx_b = nsym * cfg.hard_stride + nf * cfg.frame_len * 8
        r = dict(nfft=n, cp_len=cfg.cp_len, bps=bps, ndb=ndb, frames=nf, rx_ms=t_rx, rx_GBs=rx_b / t_rx / 1e6, rx_frac=rx_b / t_rx / 1e6 / PEAK,
                 rx_Msym_s=nsym / t_rx / 1e3, tx_ms=t_tx, tx_GBs=tx_b / t_tx / 1e6, tx_frac=tx_b / t_tx / 1e6 / PEAK,
                 tx_Msym_s=nf * cfg.frame_syms / t_tx / 1e3, ber=cnt[0] / cnt[1], locked=cnt[4] / cnt[5], peak_GBs=PEAK,
                 clocks=clocks)
        rows.append(r)
        print(json.dumps(r), flush=True)
        del m, bits, tx, rx, out
        torch.cuda.empty_cache()
