"""Host-side driver of the CUDA modem: numerology -> plan, and thin methods that hand device
buffers (PyTorch tensors, used only as memory) to the C ABI.

The numerology fields are the reference blocks' constructor arguments
(GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:17-19); ``compat`` picks which live variant's
constants apply (D10 Zadoff-Chu root, D13 regularisers, search stride, gate).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

_BPS_NAMES = {"bpsk": 1, "qpsk": 2, "16qam": 4, "qam16": 4, "64qam": 6, "qam64": 6}


@dataclass
class ModemConfig:
    nfft: int = 64
    cp_len: int = 16
    num_synch_bins: int = 62
    num_data_bins: int = 60
    m0: int = 1
    d: int = 3
    bps: int = 2
    snr: float = 100.0
    gate: float = 0.7
    compat: str = "utsa"                # "utsa" | "rxofdm"
    zc_root: Optional[int] = None
    equalizer: str = "mmse"             # "mmse" | "zf"
    sync_advance: int = 0
    noise_var: float = 1.0              # LLR scale = 1/noise_var
    pilot_spacing: int = 0              # comb pilots in data symbols (extension); 0 = none
    pilot_offset: int = -1              # -1 -> pilot_spacing // 2
    pilot_root: int = 25
    cp_first: int = 0                   # LTE normal-CP pattern (extension): prefix of symbols k % slot_syms == 0 ...
    slot_syms: int = 0                  # ... of every frame; 0 = uniform cp_len.  Needs m0 == 1.

    def __post_init__(self):
        if isinstance(self.bps, str):
            self.bps = _BPS_NAMES[self.bps.lower()]
        if self.compat not in ("utsa", "rxofdm"):
            raise ValueError("compat must be 'utsa' or 'rxofdm'")
        if self.zc_root is None:
            self.zc_root = 23 if self.compat == "utsa" else 37
        if self.compat == "rxofdm":
            self.gate = 0.4             # hard-coded: gr-RXOFDM/python/synch_and_chan_est.py:170
        if self.pilot_spacing and self.pilot_offset < 0:
            self.pilot_offset = self.pilot_spacing // 2

    # derived quantities (same names as the oracle's, computed independently)
    @property
    def sym_len(self):
        return self.nfft + self.cp_len

    @property
    def frame_syms(self):
        return self.m0 + self.d

    @property
    def cp_pattern(self):
        return self.cp_first > 0 and self.slot_syms > 0 and self.cp_first != self.cp_len

    @property
    def frame_len(self):
        n = self.frame_syms * self.sym_len
        if self.cp_pattern:             # one longer prefix per started slot
            n += (self.cp_first - self.cp_len) * ((self.frame_syms + self.slot_syms - 1) // self.slot_syms)
        return n

    @property
    def mm(self):
        return self.m0 * self.num_synch_bins

    @property
    def stride(self):
        return 1 if self.compat == "utsa" else self.cp_len - 1

    @property
    def n_pilots(self):
        if not self.pilot_spacing:
            return 0
        return len(range(self.pilot_offset, self.num_data_bins, self.pilot_spacing))

    @property
    def payload_bins(self):
        """bins per data symbol that carry payload (num_data_bins minus comb pilots)"""
        return self.num_data_bins - self.n_pilots

    @property
    def hard_stride(self):
        return (self.payload_bins * self.bps + 7) // 8

    def regularisers(self):
        """(inv_reg_chanest, inv_reg_eq, inv_reg_eq_synch): SynchAndChanEst.py:99,180,214,245 vs
        synch_and_chan_est.py:184,217,247."""
        if self.compat == "utsa":
            lin = 10.0 ** (self.snr / 20.0)
            return 1.0 / lin, 1.0 / lin, 1.0 / float(self.snr)
        return 1.0 / float(self.snr), 1.0 / float(self.snr), 1.0 / float(self.snr)

    def to_struct(self):
        rc, re_, rs = self.regularisers()
        p = _lib.OfdmParams()
        p.nfft, p.cp_len = self.nfft, self.cp_len
        p.num_synch_bins, p.num_data_bins = self.num_synch_bins, self.num_data_bins
        p.m0, p.d, p.bps = self.m0, self.d, self.bps
        p.zc_root = self.zc_root
        p.zc_parity_of = self.mm if self.compat == "utsa" else self.num_synch_bins
        p.equalizer = 1 if self.equalizer == "zf" else 0
        p.sync_advance = self.sync_advance
        p.gate = self.gate
        p.inv_reg_chanest, p.inv_reg_eq, p.inv_reg_eq_synch = rc, re_, rs
        p.llr_scale = 1.0 / self.noise_var
        p.pilot_spacing, p.pilot_offset, p.pilot_root = self.pilot_spacing, max(self.pilot_offset, 0), self.pilot_root
        p.cp_first, p.slot_syms = (self.cp_first, self.slot_syms) if self.cp_pattern else (0, 0)
        return p


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream(device=None):
    return C.c_void_p(_torch().cuda.current_stream(device).cuda_stream)


class Modem:
    """One plan on the current CUDA device.  Every method launches on torch's current stream and
    returns device tensors; nothing here computes on the CPU."""

    def __init__(self, cfg: ModemConfig, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.OfdmError("ofdm_b200 needs a CUDA device (there is no CPU fallback)")
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p()
        self.n_cfo = 0
        with torch.cuda.device(self.device):
            s = cfg.to_struct()
            _lib.check(_lib.lib().ofdm_plan_create(C.byref(s), C.byref(self._h)))
        assert _lib.lib().ofdm_plan_hard_stride(self._h) == cfg.hard_stride
        assert _lib.lib().ofdm_plan_frame_len(self._h) == cfg.frame_len
        assert _lib.lib().ofdm_plan_payload_bins(self._h) == cfg.payload_bins

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().ofdm_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- frequency-offset hypothesis bank ------------------------------------------------------
    def set_cfo_bank(self, fo_range, fs, phase_continuous=False):
        """Install (or with an empty list remove) the bank of GRR/LEGACY/gr-ofdm-rx/python/SynchEstAndFO.py:196:
        hypothesis h multiplies sample n of every FFT window by exp(j*2pi*fo_range[h]/fs*n).  From then on
        sync_search / rx_frames report the winning hypothesis and chanest / demod accept one per item.
        phase_continuous=True (extension) continues the phasor from the buffer origin instead of restarting it at
        every window, which is what a real carrier offset needs."""
        nu = np.asarray([float(f) / float(fs) for f in fo_range], dtype=np.float64)
        with _torch().cuda.device(self.device):
            _lib.check(_lib.lib().ofdm_plan_set_cfo_bank(self._h, C.c_void_p(nu.ctypes.data if nu.size else 0), nu.size,
                                                         1 if phase_continuous else 0))
        self.n_cfo = int(nu.size)                        # only once the plan has accepted the bank

    # ---- helpers ---------------------------------------------------------------------------
    def _stream(self):
        """The plan's device's current stream (the library itself makes that device current for the launch, so a
        Modem on cuda:1 works while cuda:0 is torch's current device)."""
        return _stream(self.device)

    def _c64(self, x):
        torch = _torch()
        assert x.dtype == torch.complex64 and x.is_cuda and x.is_contiguous()
        assert x.device == self.device, "tensor on %s handed to a plan on %s" % (x.device, self.device)
        return x

    def empty(self, shape, dtype):
        return _torch().empty(shape, dtype=dtype, device=self.device)

    # ---- TX ----------------------------------------------------------------------------------
    def tx_modulate(self, bits, n_frames, out=None):
        """bits: uint8 device tensor, n_frames*D*hard_stride packed bytes -> complex64 [n_frames*frame_len]."""
        torch = _torch()
        cfg = self.cfg
        assert bits.dtype == torch.uint8 and bits.is_cuda and bits.is_contiguous()
        assert bits.numel() >= n_frames * cfg.d * cfg.hard_stride
        if out is None:
            out = self.empty((n_frames * cfg.frame_len,), torch.complex64)
        _lib.check(_lib.lib().ofdm_tx_modulate(self._h, _ptr(bits), _ptr(out), n_frames, self._stream()))
        return out

    # ---- RX, framed (throughput path) --------------------------------------------------------
    def rx_frames(self, rx, n_frames, first=0, frame_stride=None, want_llr=True, want_hard=True, want_h=False,
                  out=None, want_eq=True):
        torch = _torch()
        cfg = self.cfg
        rx = self._c64(rx)
        fs = cfg.frame_len if frame_stride is None else frame_stride
        nd = n_frames * cfg.d
        o = out if out is not None else {}
        if "delay" not in o:
            o["delay"] = self.empty((n_frames,), torch.int32)
            o["peak"] = self.empty((n_frames,), torch.float32)
            o["locked"] = self.empty((n_frames,), torch.uint8)
            o["eq"] = self.empty((nd, cfg.payload_bins), torch.complex64) if want_eq else None
            o["llr"] = self.empty((nd, cfg.payload_bins * cfg.bps), torch.float32) if want_llr else None
            o["hard"] = self.empty((nd, cfg.hard_stride), torch.uint8) if want_hard else None
            o["h"] = self.empty((n_frames, cfg.nfft), torch.complex64) if want_h else None
            o["hyp"] = self.empty((n_frames,), torch.int32) if self.n_cfo else None
        _lib.check(_lib.lib().ofdm_rx_frames_cfo(
            self._h, _ptr(rx), rx.numel(), first, fs, n_frames, _ptr(o["delay"]), _ptr(o["peak"]),
            _ptr(o["locked"]), _ptr(o.get("hyp")), _ptr(o.get("h")), _ptr(o.get("eq")), _ptr(o.get("llr")),
            _ptr(o.get("hard")), self._stream()))
        return o

    # ---- RX, modular (what the stream blocks' work() drives) ---------------------------------
    def sync_search(self, rx, first_start, stride, n_trials, packed=None):
        """-> (peak f32, delay i32[, hyp i32]) per trial.  `packed`: an int32 device tensor of 2 (3 with a CFO bank) *
        n_trials words that the results are written into back to back, so a caller can fetch them with one copy."""
        torch = _torch()
        rx = self._c64(rx)
        if packed is None:
            packed = self.empty(((3 if self.n_cfo else 2) * n_trials,), torch.int32)
        assert packed.dtype == torch.int32 and packed.numel() >= (3 if self.n_cfo else 2) * n_trials
        peak = packed[:n_trials].view(torch.float32)
        delay = packed[n_trials:2 * n_trials]
        if not self.n_cfo:
            _lib.check(_lib.lib().ofdm_rx_sync_search(self._h, _ptr(rx), rx.numel(), first_start, stride, n_trials,
                                                      _ptr(peak), _ptr(delay), self._stream()))
            return peak, delay
        hyp = packed[2 * n_trials:3 * n_trials]
        _lib.check(_lib.lib().ofdm_rx_sync_search_cfo(self._h, _ptr(rx), rx.numel(), first_start, stride, n_trials,
                                                      _ptr(peak), _ptr(delay), _ptr(hyp), self._stream()))
        return peak, delay, hyp

    def chanest(self, rx, pos, delay, want_aux=False, hyp=None):
        torch = _torch()
        rx = self._c64(rx)
        n = pos.numel()
        assert pos.dtype == torch.int64 and delay.dtype == torch.int32
        h = self.empty((n, self.cfg.nfft), torch.complex64)
        ht = se = None
        if want_aux:                                     # one buffer, so both come back with one copy (.aux_flat)
            aux = self.empty((n * (self.cfg.nfft + self.cfg.mm),), torch.complex64)
            ht = aux[:n * self.cfg.nfft].view(n, self.cfg.nfft)
            se = aux[n * self.cfg.nfft:].view(n, self.cfg.mm)
            self.aux_flat = aux
        assert hyp is None or (hyp.dtype == torch.int32 and hyp.numel() == n)
        _lib.check(_lib.lib().ofdm_rx_chanest_cfo(self._h, _ptr(rx), rx.numel(), _ptr(pos), _ptr(delay), _ptr(hyp), n,
                                                  _ptr(h), _ptr(ht), _ptr(se), self._stream()))
        return h, ht, se

    def demod(self, rx, start, delay, hrow, h, want_llr=False, want_hard=False, hyp=None):
        torch = _torch()
        cfg = self.cfg
        rx = self._c64(rx)
        n = start.numel()
        assert start.dtype == torch.int64 and delay.dtype == torch.int32 and hrow.dtype == torch.int32
        eq = self.empty((n, cfg.payload_bins), torch.complex64)
        llr = self.empty((n, cfg.payload_bins * cfg.bps), torch.float32) if want_llr else None
        hard = self.empty((n, cfg.hard_stride), torch.uint8) if want_hard else None
        assert hyp is None or (hyp.dtype == torch.int32 and hyp.numel() == n)
        _lib.check(_lib.lib().ofdm_rx_demod_cfo(self._h, _ptr(rx), rx.numel(), _ptr(start), _ptr(delay), _ptr(hrow),
                                                _ptr(hyp), _ptr(self._c64(h)), n, _ptr(eq), _ptr(llr), _ptr(hard),
                                                self._stream()))
        return eq, llr, hard

    # ---- DSSS despreading (legacy receiver) ------------------------------------------------------
    def dsss_despread(self, eq, code):
        """eq: complex64 [n_rows, row_len]; code: complex64 [sf] -> [n_rows, row_len // sf]
        (GRR/LEGACY/gr-ofdm-rx/python/SynchEstFOAndDSSS.py:392-397)."""
        torch = _torch()
        eq = self._c64(eq)
        code = self._c64(code)
        n_rows, row_len = eq.shape
        sf = code.numel()
        out = self.empty((n_rows, row_len // sf), torch.complex64)
        _lib.check(_lib.lib().ofdm_dsss_despread(self._h, _ptr(eq), n_rows, row_len, _ptr(code), sf, _ptr(out), self._stream()))
        return out

    # ---- legacy demapper, counters, test-bench channel ----------------------------------------
    def bit_recovery(self, x):
        torch = _torch()
        x = self._c64(x)
        n = x.numel()
        llr0 = self.empty((2 * n,), torch.float32)
        llr1 = self.empty((2 * n,), torch.float32)
        hard = self.empty((2 * n,), torch.uint8)
        ws = self.empty((2,), torch.float64)
        _lib.check(_lib.lib().ofdm_bit_recovery(self._h, _ptr(x), n, _ptr(llr0), _ptr(llr1), _ptr(hard), _ptr(ws),
                                                self._stream()))
        return llr0, llr1, hard

    def counters(self, bits_tx, n_frames, bits_rx=None, eq=None, locked=None, cnt=None):
        torch = _torch()
        if cnt is None:
            cnt = torch.zeros(8, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().ofdm_counters(self._h, _ptr(bits_rx), _ptr(bits_tx), _ptr(eq), _ptr(locked), n_frames,
                                            _ptr(cnt), self._stream()))
        return cnt

    def channel_awgn(self, tx, n_frames, taps, noise_sigma, seed, out=None):
        """taps: complex64 [n_taps] (shared) or [n_frames, n_taps] (per frame)."""
        torch = _torch()
        tx = self._c64(tx)
        taps = self._c64(taps)
        per_frame = 1 if taps.dim() == 2 else 0
        n_taps = taps.shape[-1]
        if out is None:
            out = torch.empty_like(tx)
        _lib.check(_lib.lib().ofdm_channel_awgn(self._h, _ptr(tx), _ptr(out), n_frames, _ptr(taps), n_taps,
                                                per_frame, float(noise_sigma), int(seed), self._stream()))
        return out

    # ---- host-buffer entry points (numpy in, numpy out; copies inside the call) -----------------
    def rx_frames_host(self, rx_np, n_frames, first=0, frame_stride=None, out=None, want_llr=True, want_hard=True,
                       want_eq=True):
        """Outputs that are not asked for (want_* False, or a None entry in a caller-supplied `out`) are neither
        computed nor copied back; at least one of eq / llr / hard is needed."""
        cfg = self.cfg
        assert rx_np.dtype == np.complex64 and rx_np.flags.c_contiguous
        fs = cfg.frame_len if frame_stride is None else frame_stride
        nd = n_frames * cfg.d
        o = out if out is not None else {}
        if "delay" not in o:
            o["delay"] = np.empty(n_frames, np.int32)
            o["peak"] = np.empty(n_frames, np.float32)
            o["locked"] = np.empty(n_frames, np.uint8)
            o["eq"] = np.empty((nd, cfg.payload_bins), np.complex64) if want_eq else None
            o["llr"] = np.empty((nd, cfg.payload_bins * cfg.bps), np.float32) if want_llr else None
            o["hard"] = np.empty((nd, cfg.hard_stride), np.uint8) if want_hard else None

        def p(a):
            return C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)
        with _torch().cuda.device(self.device):
            _lib.check(_lib.lib().ofdm_rx_frames_host(self._h, p(rx_np), rx_np.size, first, fs, n_frames, p(o["delay"]),
                                                      p(o["peak"]), p(o["locked"]), p(o.get("eq")), p(o.get("llr")),
                                                      p(o.get("hard"))))
        return o

    def tx_modulate_host(self, bits_np, n_frames, out=None):
        cfg = self.cfg
        assert bits_np.dtype == np.uint8 and bits_np.flags.c_contiguous
        if out is None:
            out = np.empty(n_frames * cfg.frame_len, np.complex64)
        with _torch().cuda.device(self.device):
            _lib.check(_lib.lib().ofdm_tx_modulate_host(self._h, C.c_void_p(bits_np.ctypes.data),
                                                        C.c_void_p(out.ctypes.data), n_frames))
        return out


def launch_count():
    return int(_lib.lib().ofdm_launch_count())
