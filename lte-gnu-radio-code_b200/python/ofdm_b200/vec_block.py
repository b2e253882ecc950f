"""Vector-item form of the two live receiver blocks (additive; SURVEY.md 8f rank 3, "vector I/O with vlen = frame").

The reference's blocks are 1:1 stream blocks on single complex64 items (in_sig = out_sig = [np.complex64],
GRR/gr-RXOFDM/python/synch_and_chan_est.py:121-124, GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:112-115): the scheduler
hands work() a few thousand samples at arbitrary frame phase, so the block has to search, keep lock state across
calls and drop the symbols a buffer cuts in two (:147,:226).  When the upstream side already delimits frames (a
stream-to-vector block after a frame-synchronous source, a tagged-stream adapter, a file of frames) GNU Radio can carry
ONE FRAME PER ITEM instead:

    in_sig  = [(np.complex64, frame_len)]                 one frame of M0 synch + D data symbols, prefixes included
    out_sig = [(np.complex64, D * num_data_bins)]         its equalised data symbols, row-major [D, num_data_bins]
              (+ (np.float32, D * num_data_bins * bps)    max-log LLRs, when soft=True)

work() is then one call of the fused framed receiver (ofdm_rx_frames_host -> rx_frames_kernel) on the scheduler's
buffers, written straight into output_items: per frame it is the arithmetic of one accepted lock of the reference's
work() -- loop A's body at the frame's nominal position (window gather, transform, synch-bin pick, unit-power
normalisation, the delay bank: SynchAndChanEst.py:145-165), the LS channel estimate (:177-188) and loop B for the
frame's D data symbols (:225-248).  No state crosses calls, nothing is dropped at buffer edges, and an item is large
enough (N=64: 320 samples; LTE: 30 720) for the scheduler's per-item overhead to vanish.  Items whose peak fails the
gate (:166) are still equalised with their own estimate -- a sync block owes one output item per input item -- and are
flagged in `locked`; `delay`, `peak`, `locked` hold the last call's per-frame lock record."""
import numpy as np

from .gr_compat import gr
from .modem import Modem, ModemConfig


class RxVecBlockBase(gr.sync_block):
    compat = "utsa"

    def _setup_vec(self, name, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, gate, modulation, zc_root,
                   device, soft, noise_var):
        self.cfg = ModemConfig(nfft=int(nfft), cp_len=int(cp_len), num_synch_bins=int(num_synch_bins),
                               num_data_bins=int(num_data_bins), m0=int(synch_dat[0]), d=int(synch_dat[1]),
                               bps=modulation, snr=float(snr), gate=float(gate), compat=self.compat, zc_root=zc_root,
                               noise_var=float(noise_var))
        c = self.cfg
        self.frame_len = c.frame_len
        self.out_len = c.d * c.num_data_bins
        self.soft = bool(soft)
        out_sig = [(np.complex64, self.out_len)]
        if self.soft:
            out_sig.append((np.float32, self.out_len * c.bps))
        gr.sync_block.__init__(self, name=name, in_sig=[(np.complex64, self.frame_len)], out_sig=out_sig)
        self.modem = Modem(c, device=device)                # raises when the CUDA library is missing: no CPU fallback
        self.delay = np.zeros(0, np.int32)
        self.peak = np.zeros(0, np.float32)
        self.locked = np.zeros(0, np.uint8)
        self._cap = 0

    def work(self, input_items, output_items):
        c = self.cfg
        in0 = input_items[0]
        n = min(len(in0), *(len(o) for o in output_items))
        if n == 0:
            return 0
        if n > self._cap:
            self._cap = max(n, 2 * self._cap)
            self._rec = (np.empty(self._cap, np.int32), np.empty(self._cap, np.float32), np.empty(self._cap, np.uint8))
        rx = in0[:n]
        if not (isinstance(rx, np.ndarray) and rx.dtype == np.complex64 and rx.flags.c_contiguous):
            rx = np.ascontiguousarray(rx, dtype=np.complex64)
        if rx.size != n * self.frame_len:
            raise ValueError("items of %d samples expected, got %d" % (self.frame_len, rx.size // n))

        def direct(buf, dtype):
            v = buf[:n]
            return v if (isinstance(v, np.ndarray) and v.dtype == dtype and v.flags.c_contiguous) else None
        eq = direct(output_items[0], np.complex64)
        eq_t = eq if eq is not None else np.empty((n, self.out_len), np.complex64)
        llr_t = None
        if self.soft:
            llr = direct(output_items[1], np.float32)
            llr_t = llr if llr is not None else np.empty((n, self.out_len * c.bps), np.float32)
        o = {"delay": self._rec[0][:n], "peak": self._rec[1][:n], "locked": self._rec[2][:n],
             "eq": eq_t, "llr": llr_t, "hard": None}
        self.modem.rx_frames_host(rx, n, out=o)
        if eq is None:
            output_items[0][:n] = eq_t.reshape(np.shape(output_items[0][:n]))
        if self.soft and llr is None:
            output_items[1][:n] = llr_t.reshape(np.shape(output_items[1][:n]))
        self.delay, self.peak, self.locked = o["delay"], o["peak"], o["locked"]
        return n

    def close(self):
        m = getattr(self, "modem", None)
        if m is not None and hasattr(m, "close"):
            m.close()


class TxVecBlockBase(gr.sync_block):
    """The transmit side of the same idea: one frame's packed payload per input item, one frame of samples per output
    item.  The reference modulates OFFLINE (SDRScript.py:128-134 -> MultiAntennaSystem.multi_ant_binary_map /
    multi_ant_symb_gen, :113-219) and its live blocks replay the stored waveform (tx_signal_transmitter.py:24-27); this
    block runs that arithmetic -- bits -> points / Zadoff-Chu symbols -> IFFT -> prefix -> the two normalisations -- in
    the flowgraph, one ofdm_tx_modulate_host call per work().

        in_sig  = [(np.uint8, D * hard_stride)]      D payload rows, num_data_bins * bps bits each, MSB first, each row
                                                     padded to whole bytes (np.packbits of the reference's bit rows)
        out_sig = [(np.complex64, frame_len)]        M0 synch + D data symbols with their prefixes"""

    def _setup_vec(self, name, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, modulation, zc_root, compat, device):
        self.cfg = ModemConfig(nfft=int(nfft), cp_len=int(cp_len), num_synch_bins=int(num_synch_bins),
                               num_data_bins=int(num_data_bins), m0=int(synch_dat[0]), d=int(synch_dat[1]),
                               bps=modulation, compat=compat, zc_root=zc_root)
        c = self.cfg
        self.frame_len = c.frame_len
        self.in_len = c.d * c.hard_stride
        gr.sync_block.__init__(self, name=name, in_sig=[(np.uint8, self.in_len)], out_sig=[(np.complex64, self.frame_len)])
        self.modem = Modem(c, device=device)

    def work(self, input_items, output_items):
        in0, out0 = input_items[0], output_items[0]
        n = min(len(in0), len(out0))
        if n == 0:
            return 0
        bits = in0[:n]
        if not (isinstance(bits, np.ndarray) and bits.dtype == np.uint8 and bits.flags.c_contiguous):
            bits = np.ascontiguousarray(bits, dtype=np.uint8)
        if bits.size != n * self.in_len:
            raise ValueError("items of %d bytes expected, got %d" % (self.in_len, bits.size // n))
        v = out0[:n]
        if isinstance(v, np.ndarray) and v.dtype == np.complex64 and v.flags.c_contiguous:
            self.modem.tx_modulate_host(bits.reshape(-1), n, out=v.reshape(-1))
        else:
            out0[:n] = self.modem.tx_modulate_host(bits.reshape(-1), n).reshape(np.shape(v))
        return n

    def close(self):
        m = getattr(self, "modem", None)
        if m is not None:
            m.close()
