"""TX -> channel -> RX -> BER/EVM entirely in device memory.

The reference builds its receiver inputs offline on the CPU (bits -> MultiAntennaSystem.multi_ant_binary_map /
multi_ant_symb_gen -> rx_signal_gen -> additive_noise, GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/SDRScript.py:89-153)
and pickles the waveform for the replay block.  Here the same chain is four kernel launches on one GPU and nothing
crosses PCIe: payload bits are drawn on the device, modulated (ofdm_tx_modulate), passed through a shared or
per-frame FIR channel with counter-based AWGN (ofdm_channel_awgn), received (ofdm_rx_frames) and scored
(ofdm_counters).  Frames are independent, so under torch.distributed every rank runs its own share and the
8-entry counter vector is the only thing reduced (shard.reduce_counters).
"""
import numpy as np

from . import shard
from .modem import Modem, ModemConfig

# reference channel taps, MultiAntennaSystem.py:64
REF_TAPS = np.array([0.3977, 0.7954 - 0.3977j, -0.1988, 0.0994, -0.0398])


class Loopback:
    def __init__(self, cfg: ModemConfig, device=None):
        import torch
        self.torch = torch
        self.cfg = cfg
        self.modem = Modem(cfg, device=device)
        self.device = self.modem.device

    def taps_reference(self):
        t = REF_TAPS / np.linalg.norm(REF_TAPS)
        return self.torch.tensor(t, dtype=self.torch.complex64, device=self.device)

    def taps_random(self, n_frames, pdp, seed):
        """One independent tap vector per frame, tap k ~ CN(0, pdp[k])."""
        torch = self.torch
        g = torch.Generator(device=self.device)
        g.manual_seed(int(seed))
        t = torch.randn((n_frames, len(pdp)), dtype=torch.complex64, device=self.device, generator=g)
        return (t * torch.tensor(np.sqrt(np.asarray(pdp, dtype=np.float64)), dtype=torch.float32, device=self.device)).contiguous()

    def run(self, n_frames, seed=0, taps=None, noise_sigma=0.0, bits=None, want_llr=True, reduce=True):
        """One pass over n_frames frames.  Returns dict(bits, tx, rx, out, counters, summary); every tensor stays on
        the device.  taps: complex64 [n_taps] (shared) or [n_frames, n_taps] (per frame); default = reference channel."""
        torch = self.torch
        cfg, m = self.cfg, self.modem
        if bits is None:
            g = torch.Generator(device=self.device)
            g.manual_seed(int(seed))
            bits = torch.randint(0, 256, (n_frames * cfg.d * cfg.hard_stride,), dtype=torch.uint8, device=self.device, generator=g)
            spare = cfg.hard_stride * 8 - cfg.payload_bins * cfg.bps          # pad bits of each row are not transmitted
            if spare:
                rows = bits.view(n_frames * cfg.d, cfg.hard_stride)
                rows[:, -1] &= (0xFF << spare) & 0xFF
        tx = m.tx_modulate(bits, n_frames)
        if taps is None:
            taps = self.taps_reference()
        rx = m.channel_awgn(tx, n_frames, taps, noise_sigma, int(seed) + 1)
        out = m.rx_frames(rx, n_frames, want_llr=want_llr)
        cnt = m.counters(bits, n_frames, bits_rx=out["hard"], eq=out["eq"], locked=out["locked"])
        if reduce:
            shard.reduce_counters(cnt)
        return dict(bits=bits, tx=tx, rx=rx, out=out, counters=cnt, summary=shard.summarize(cnt.cpu().numpy()))
