"""Launch the hot kernels a few times each so ncu can capture them (see profiles/README.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
from ofdm_b200 import Modem, ModemConfig  # noqa: E402


def n64(bps):
    return (ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=bps, snr=100,
                        gate=0.5, noise_var=0.01), 333334)


CASES = {
    "n64": n64(4),
    "n64_qpsk": n64(2),
    "n64_64qam": n64(6),
    "lte": (ModemConfig(nfft=2048, cp_len=144, cp_first=160, slot_syms=7, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6, snr=100,
                        gate=0.5, noise_var=0.01), 4096),
    "n128": (ModemConfig(nfft=128, cp_len=32, num_synch_bins=126, num_data_bins=120, m0=1, d=3, bps=4, snr=100,
                         gate=0.5, noise_var=0.01), 104857),
    "n256": (ModemConfig(nfft=256, cp_len=64, num_synch_bins=254, num_data_bins=240, m0=1, d=3, bps=4, snr=100,
                         gate=0.5, noise_var=0.01), 52428),
    "n512": (ModemConfig(nfft=512, cp_len=128, num_synch_bins=510, num_data_bins=480, m0=1, d=3, bps=4, snr=100,
                         gate=0.5, noise_var=0.01), 26214),
    "n1024": (ModemConfig(nfft=1024, cp_len=256, num_synch_bins=1022, num_data_bins=960, m0=1, d=3, bps=4, snr=100,
                          gate=0.5, noise_var=0.01), 13107),
    "n4096": (ModemConfig(nfft=4096, cp_len=1024, num_synch_bins=4094, num_data_bins=3840, m0=1, d=3, bps=4, snr=100,
                          gate=0.5, noise_var=0.01), 3276),
    "n4096_64qam": (ModemConfig(nfft=4096, cp_len=1024, num_synch_bins=4094, num_data_bins=3840, m0=1, d=3, bps=6, snr=100,
                                gate=0.5, noise_var=0.01), 3276),
}

if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    reps = int(os.environ.get("PROF_REPS", "2"))
    for name in which:
        cfg, nf = CASES[name]
        m = Modem(cfg)
        bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
        tx = m.tx_modulate(bits, nf)
        taps = torch.tensor([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398], dtype=torch.complex64, device="cuda")
        rx = m.channel_awgn(tx, nf, taps / torch.linalg.norm(taps), 0.02, 1)
        out = m.rx_frames(rx, nf)
        for _ in range(reps):
            m.tx_modulate(bits, nf, out=tx)
        for _ in range(reps):
            m.rx_frames(rx, nf, out=out)
        torch.cuda.synchronize()
        print(name, "done")
