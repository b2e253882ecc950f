"""Time the host-buffer receiver (ofdm_rx_frames_host) on the default bench workload; chunk sizes come from the
environment (OFDM_HOST_CHUNK_MB, OFDM_HOST_FIRST_MB), so run one process per setting."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lte-gnu-radio-code_b200", "python"))
from ofdm_b200 import Modem, ModemConfig
which = sys.argv[1] if len(sys.argv) > 1 else "n64"
if which == "n64":
    cfg, nf = ModemConfig(nfft=64, cp_len=16, num_synch_bins=62, num_data_bins=60, m0=1, d=3, bps=4, snr=100, gate=0.5, noise_var=0.01), 333334
else:
    cfg, nf = ModemConfig(nfft=2048, cp_len=144, num_synch_bins=2046, num_data_bins=1200, m0=1, d=13, bps=6, snr=100, gate=0.5, noise_var=0.01), 4096
m = Modem(cfg)
bits = torch.randint(0, 256, (nf * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device="cuda")
tx = m.tx_modulate(bits, nf)
taps = torch.tensor([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398], dtype=torch.complex64, device="cuda")
rx = m.channel_awgn(tx, nf, taps / torch.linalg.norm(taps), 0.02, 1)
ref = m.rx_frames(rx, nf)
h_rx = torch.empty(rx.shape, dtype=torch.complex64, pin_memory=True); h_rx.copy_(rx)
nd = nf * cfg.d
h = {"delay": torch.empty(nf, dtype=torch.int32, pin_memory=True), "peak": torch.empty(nf, dtype=torch.float32, pin_memory=True),
     "locked": torch.empty(nf, dtype=torch.uint8, pin_memory=True), "eq": torch.empty((nd, cfg.num_data_bins), dtype=torch.complex64, pin_memory=True),
     "llr": torch.empty((nd, cfg.num_data_bins * cfg.bps), dtype=torch.float32, pin_memory=True),
     "hard": torch.empty((nd, cfg.hard_stride), dtype=torch.uint8, pin_memory=True)}
hn = {k: v.numpy() for k, v in h.items()}
for _ in range(2):
    m.rx_frames_host(h_rx.numpy(), nf, out=hn)
ts = []
for _ in range(8):
    t0 = time.perf_counter(); m.rx_frames_host(h_rx.numpy(), nf, out=hn); ts.append(time.perf_counter() - t0)
ok = all(np.array_equal(hn[k], ref[k].cpu().numpy()) for k in ("delay", "locked", "hard")) and np.array_equal(hn["eq"], ref["eq"].cpu().numpy()) and np.array_equal(hn["llr"], ref["llr"].cpu().numpy())
d2h = sum(v.nbytes for v in hn.values())
print("%s chunk=%s first=%s: median %.2f ms min %.2f ms  d2h %.1f GB/s  identical=%s" % (which, os.environ.get("OFDM_HOST_CHUNK_MB", "def"), os.environ.get("OFDM_HOST_FIRST_MB", "def"), 1e3 * np.median(ts), 1e3 * min(ts), d2h / min(ts) / 1e9, ok))
