"""Shared machinery of the two receiver blocks (RXOFDM.synch_and_chan_est and
utsa_ofdm.SynchAndChanEst): host<->device staging for one work() call and the lock bookkeeping,
which stays on the host because it is a handful of scalar decisions per call
(GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:166-175, GRR/gr-RXOFDM/python/synch_and_chan_est.py:170-179).
All arithmetic on samples runs in the CUDA kernels."""
import ctypes as C
import pickle

import numpy as np

from . import _lib
from .gr_compat import gr
from .modem import Modem, ModemConfig


class RxBlockBase(gr.sync_block):
    compat = "utsa"

    def _setup(self, num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, gate,
               directory_name, file_name_cest, diagnostics, genie, modulation, zc_root, rows, device, fused=None):
        import torch
        self._torch = torch
        self.num_ofdm_symb = int(num_ofdm_symb)
        self.nfft = int(nfft)
        self.cp_len = int(cp_len)
        self.num_synch_bins = int(num_synch_bins)
        self.synch_dat = list(synch_dat)
        self.num_data_bins = int(num_data_bins)
        self.SNR = snr
        self.genie = genie
        self.directory_name = directory_name
        self.file_name_cest = file_name_cest
        self.cfg = ModemConfig(nfft=self.nfft, cp_len=self.cp_len, num_synch_bins=self.num_synch_bins,
                               num_data_bins=self.num_data_bins, m0=int(synch_dat[0]), d=int(synch_dat[1]),
                               bps=modulation, snr=float(snr), gate=float(gate), compat=self.compat, zc_root=zc_root)
        self.modem = Modem(self.cfg, device=device)
        c = self.cfg
        self.M = [c.m0, c.num_synch_bins]
        self.MM = c.mm
        self.L_synch = c.num_synch_bins
        self.stride_val = c.stride
        self.start_samp = c.cp_len
        self.rx_b_len = c.sym_len
        self.max_num_corr = 100
        self.rows = rows
        self._d_H = torch.zeros((rows, c.nfft), dtype=torch.complex64, device=self.modem.device)
        self.est_chan_time = np.zeros((rows, c.nfft), dtype=complex)
        self.est_synch_freq = np.zeros((rows, c.mm), dtype=complex)
        self.est_data_freq = np.zeros((rows, c.num_data_bins), dtype=complex)
        self.count = 0
        self.last_written = 0
        # One C-ABI call per work() (ofdm_rx_block_work: upload, search, lock rule on the device, estimate, demodulate,
        # one synchronisation) for the two live blocks; `fused=False` keeps the call-by-call path that the legacy
        # blocks with a hypothesis bank use (and that tools/bench_blocks.py times beside it).
        self._blk = None
        if fused is None:
            fused = self.compat in ("utsa", "rxofdm") and type(self).work_fused is not None
        if fused:
            h = C.c_void_p()
            with torch.cuda.device(self.modem.device):
                _lib.check(_lib.lib().ofdm_rx_block_create(self.modem._h, 0 if self.compat == "utsa" else 1, rows, C.byref(h)))
            self._blk = h
            self._res = _lib.RxBlockResult()
            self._res_ref = C.byref(self._res)
            self._work_fn = _lib.lib().ofdm_rx_block_work
            self._views = None

    work_fused = None          # set by the subclasses that have a fused path

    def close(self):
        if getattr(self, "_blk", None):
            _lib.lib().ofdm_rx_block_destroy(self._blk)
            self._blk = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # state the reference keeps as a NumPy array; here it lives on the device and is read on demand
    @property
    def est_chan_freq_P(self):
        if self._blk:
            out = np.empty((self.rows, self.cfg.nfft), np.complex64)
            _lib.check(_lib.lib().ofdm_rx_block_chan_freq(self._blk, C.c_void_p(out.ctypes.data), self.rows))
            return out.astype(complex)
        return self._d_H.cpu().numpy().astype(complex)

    def _fused_call(self, in0):
        """ofdm_rx_block_work on the scheduler's buffer; mirrors the per-call results into the NumPy state the
        reference exposes (est_chan_time, est_synch_freq, est_data_freq) and returns the result struct."""
        c = self.cfg
        a = in0 if (isinstance(in0, np.ndarray) and in0.dtype == np.complex64 and in0.flags.c_contiguous) \
            else np.ascontiguousarray(in0, dtype=np.complex64)
        r = self._res
        _lib.check(self._work_fn(self._blk, a.ctypes.data, len(a), self._res_ref))
        if self._views is None:
            # the result pointers address pinned memory owned by the block object and never move: wrap them once
            ml, ms = (1, self.rows) if self.compat == "utsa" else (100, 100)
            as_arr = np.ctypeslib.as_array
            self._views = (as_arr(r.new_rows, (ml,)), as_arr(r.chan_time, (ml, 2 * c.nfft)).view(np.complex64),
                           as_arr(r.synch_freq, (ml, 2 * c.mm)).view(np.complex64), as_arr(r.demod_rows, (ms,)),
                           as_arr(r.data_freq, (ms, 2 * c.num_data_bins)).view(np.complex64),
                           as_arr(r.time_synch_ref, (r.n_state_rows, 3)))
        new_rows, ht, se, dm_rows, eq, self._tsr_view = self._views
        k = r.n_new_locks
        if k:
            self.est_chan_time[new_rows[:k]] = ht[:k]
            self.est_synch_freq[new_rows[:k]] = se[:k]
            self._last_chan_time = ht[:k]
        k = r.n_demod
        if k:
            self.est_data_freq[dm_rows[:k]] = eq[:k]
        return r

    # ---- device-buffer form of work() (additive) --------------------------------------------
    def work_device(self, rx, out=None):
        """work() for a block whose scheduler buffers already live in device memory -- what a GNU Radio 3.10 block
        built on custom (device) buffers is handed instead of the host views of
        GRR/gr-RXOFDM/python/synch_and_chan_est.py:121-124,:140-142.  One C-ABI call (ofdm_rx_block_work_device): the
        samples are read in place, nothing is staged or uploaded, the equalised data rows and est_data_freq stay in
        device memory, and the output stream is packed there.

        rx: complex64 CUDA tensor [n] on the block's device.  Loops A and B, the lock rule and the block's state
        (lock counter, time_synch_ref, est_chan_freq_P, est_chan_time, est_synch_freq) behave exactly as in work(),
        with which it may be mixed.  out (optional): complex64 CUDA tensor [m], the output buffer of work(): receives
        out[0:last_written] exactly as work() would write it -- nothing on the block's first call, then the packed rows
        of est_data_freq (SynchAndChanEst.py:249-259 / synch_and_chan_est.py:252-263), where est_data_freq is the
        device-resident copy that work_device calls maintain.  Returns (rows, eq): eq is a CUDA tensor
        [n_demod, num_data_bins] -- a view of a buffer the block owns, valid until its next call -- and rows (NumPy
        int32 [n_demod]) names the est_data_freq row each one is.  The HOST mirror est_data_freq is not refreshed (that
        copy is what this entry point saves)."""
        torch = self._torch
        if not self._blk:
            raise _lib.OfdmError("work_device needs the one-call block path (fused=True)")
        if not (torch.is_tensor(rx) and rx.is_cuda and rx.dtype == torch.complex64 and rx.dim() == 1 and rx.is_contiguous()):
            raise ValueError("work_device expects a contiguous 1-D complex64 CUDA tensor")
        if rx.device != self.modem.device:
            raise ValueError("input lives on %s, the block on %s" % (rx.device, self.modem.device))
        c = self.cfg
        if rx.numel() == 0:
            if out is not None:
                self.last_written = 0
            return np.zeros(0, np.int32), torch.empty((0, c.num_data_bins), dtype=torch.complex64, device=rx.device)
        if getattr(self, "_d_eq", None) is None:
            max_rows = _lib.lib().ofdm_rx_block_max_rows(self._blk)
            self._d_eq = torch.empty((max_rows, c.num_data_bins), dtype=torch.complex64, device=self.modem.device)
        r = self._res
        d_stream, cap, pack_rows = None, 0, 0
        if out is not None:
            if not (torch.is_tensor(out) and out.is_cuda and out.dtype == torch.complex64 and out.dim() == 1
                    and out.is_contiguous() and out.device == self.modem.device):
                raise ValueError("work_device expects `out` as a contiguous 1-D complex64 CUDA tensor on the block's device")
            self.last_written = 0
            if self.count > 0 and out.numel() > 0:           # nothing on the first call (:257 / :261)
                d_stream, cap = out.data_ptr(), out.numel()
                if self.compat == "utsa":                    # n_data_symb, SynchAndChanEst.py:136-138
                    pack_rows = int((rx.numel() // c.sym_len) * (c.d / c.frame_syms))
                else:                                        # corr_size, synch_and_chan_est.py:253 (made integral)
                    pack_rows = min(int(self.num_ofdm_symb / c.frame_syms), self.rows)
        n_written = C.c_int64(0)
        with torch.cuda.device(self.modem.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().ofdm_rx_block_work_device(self._blk, rx.data_ptr(), rx.numel(), self._d_eq.data_ptr(),
                                                            d_stream, pack_rows, cap, st, self._res_ref,
                                                            C.byref(n_written)))
        if out is not None:
            self.last_written = int(n_written.value)
        ml, ms = (1, self.rows) if self.compat == "utsa" else (100, 100)
        as_arr = np.ctypeslib.as_array
        if getattr(self, "_views_dev", None) is None:
            # host-side result pointers address pinned memory owned by the block object and never move
            self._views_dev = (as_arr(r.new_rows, (ml,)), as_arr(r.chan_time, (ml, 2 * c.nfft)).view(np.complex64),
                               as_arr(r.synch_freq, (ml, 2 * c.mm)).view(np.complex64), as_arr(r.demod_rows, (ms,)),
                               as_arr(r.time_synch_ref, (r.n_state_rows, 3)))
        new_rows, ht, se, dm_rows, self._tsr_view = self._views_dev
        k = r.n_new_locks
        if k:
            self.est_chan_time[new_rows[:k]] = ht[:k]
            self.est_synch_freq[new_rows[:k]] = se[:k]
            self._last_chan_time = ht[:k]
        if self.compat == "utsa":
            self.time_synch_ref[:] = self._tsr_view[0]
            self.corr_obs = 0                                # SynchAndChanEst.py:261
        else:
            k = r.obs_after_search + 1
            if k > 0:
                self.time_synch_ref[:k] = self._tsr_view[:k]
            self.cor_obs = 0                                 # synch_and_chan_est.py:265
        self.count += 1
        return dm_rows[:r.n_demod].copy(), self._d_eq[:r.n_demod]

    def _upload(self, in0):
        torch = self._torch
        a = np.ascontiguousarray(in0, dtype=np.complex64)
        return torch.from_numpy(a).to(self.modem.device, non_blocking=False)

    def _n_valid_trials(self, n):
        """Trials P = 0.. that pass the window guard M0*L + P*stride + N + CP < len(in0)
        (SynchAndChanEst.py:139,144)."""
        c = self.cfg
        n_trials = int(np.around(n / c.stride))
        room = n - c.m0 * c.sym_len - c.nfft - c.cp_len      # need P*stride < room
        if room <= 0:
            return 0
        return min(n_trials, (room - 1) // c.stride + 1)

    def _search(self, rx, n):
        nv = self._n_valid_trials(n)
        if nv == 0:
            return np.zeros(0, np.float32), np.zeros(0, np.int32)
        k = 3 if self.modem.n_cfo else 2
        packed = self.modem.empty((k * nv,), self._torch.int32)
        self.modem.sync_search(rx, self.cfg.cp_len, self.cfg.stride, nv, packed=packed)
        host = packed.cpu().numpy()                         # one copy: (peak, delay) or, with a CFO bank, (peak, delay, hyp)
        return (host[:nv].view(np.float32),) + tuple(host[i * nv:(i + 1) * nv] for i in range(1, k))

    def _to_device(self, i64=(), i32=()):
        """Small per-call index lists -> device tensors with ONE host-to-device copy (every copy of a work() call is a
        synchronisation, and the call is latency-bound): int64 lists first, then int32 lists, all of one length."""
        torch = self._torch
        n = len(i64[0]) if i64 else len(i32[0])
        raw = np.empty(8 * n * len(i64) + 4 * n * len(i32), np.uint8)
        at = 0
        for v in i64:
            raw[at:at + 8 * n].view(np.int64)[:] = v
            at += 8 * n
        for v in i32:
            raw[at:at + 4 * n].view(np.int32)[:] = v
            at += 4 * n
        d = torch.from_numpy(raw).to(self.modem.device)
        out, at = [], 0
        for _ in i64:
            out.append(d[at:at + 8 * n].view(torch.int64))
            at += 8 * n
        for _ in i32:
            out.append(d[at:at + 4 * n].view(torch.int32))
            at += 4 * n
        return out

    def _estimate_into(self, rx, rows, pos, delay, hyp=None):
        """Channel estimates for new locks -> state rows (+ time-domain channel and equalised preamble, which the
        reference keeps as block state)."""
        dv = self._to_device(i64=(pos, rows), i32=(delay,) if hyp is None else (delay, hyp))
        h, ht, se = self.modem.chanest(rx, dv[0], dv[2], want_aux=True, hyp=None if hyp is None else dv[3])
        self._d_H.index_copy_(0, dv[1], h)
        n, c = len(rows), self.cfg
        aux = self.modem.aux_flat.cpu().numpy()              # both auxiliary outputs in one copy
        ht_h, se_h = aux[:n * c.nfft].reshape(n, c.nfft), aux[n * c.nfft:].reshape(n, c.mm)
        for i, r in enumerate(rows):
            self.est_chan_time[r] = ht_h[i]
            self.est_synch_freq[r] = se_h[i]
        return ht_h

    # ---- throughput API (additive): device-resident, no scheduler buffers ---------------
    def process_batch(self, rx, first=0, want_llr=True, want_hard=True):
        """Framed receiver over a batch of independent streams that stay in device memory.

        rx: complex64 CUDA tensor [B, T] (B streams) or [T]; every stream holds whole frames of M0 synch + D data
        symbols starting at sample `first`.  Per frame this is the arithmetic of one accepted lock of work() --
        loop A's body at the frame's nominal position, the LS channel estimate, and loop B for the frame's D data
        symbols -- in one fused kernel.  Returns a dict of device tensors: delay / peak / locked [B, F],
        eq [B, F*D, ndb] (+ llr [B, F*D, ndb*bps], hard [B, F*D, hard_stride])."""
        c = self.cfg
        flat = rx.reshape(-1) if rx.dim() == 2 else rx
        B = rx.shape[0] if rx.dim() == 2 else 1
        T = flat.numel() // B
        F = (T - first) // c.frame_len
        if F <= 0:
            raise ValueError("streams of %d samples hold no whole frame of %d samples" % (T, c.frame_len))
        if B == 1 or T == F * c.frame_len + first and first == 0:
            out = self.modem.rx_frames(flat, B * F, first=first, want_llr=want_llr, want_hard=want_hard)
            parts = [out]
        else:                                    # ragged streams: one launch per stream, each with its own origin
            parts = [self.modem.rx_frames(flat, F, first=b * T + first, want_llr=want_llr, want_hard=want_hard)
                     for b in range(B)]
        torch = self._torch
        res = {}
        for k in ("delay", "peak", "locked", "eq", "llr", "hard"):
            vals = [p_[k] for p_ in parts]
            if vals[0] is None:
                res[k] = None
                continue
            v = vals[0] if len(vals) == 1 else torch.cat(vals)
            res[k] = v.reshape((B, F) if k in ("delay", "peak", "locked") else (B, F * c.d, -1))
        return res

    def _dump_chan(self, chan_est_tim, path):
        with open(path, "wb") as f:
            pickle.dump(np.asarray(chan_est_tim, dtype=complex).reshape(1, -1), f, protocol=2)

    def _demod_into(self, rx, rows, start, delay, hrow, hyp=None):
        if not rows:
            return
        torch = self._torch
        dev = self.modem.device
        dv = self._to_device(i64=(start,), i32=(delay, hrow) if hyp is None else (delay, hrow, hyp))
        eq, _, _ = self.modem.demod(rx, dv[0], dv[1], dv[2], self._d_H, hyp=None if hyp is None else dv[3])
        self.est_data_freq[np.asarray(rows)] = eq.cpu().numpy()

    # ---- streaming mode (additive; SURVEY.md 5, "partial frames") ------------------------------------------------
    # The reference looks at each scheduler buffer on its own: a trial needs its whole window inside the buffer
    # (SynchAndChanEst.py:144, synch_and_chan_est.py:147) and a data symbol likewise (:223 / :226), so a frame that
    # straddles two work() calls is lost, and frames are re-demodulated when buffers overlap.  With the keyword-only
    # constructor argument carry_over=True the block treats successive buffers as ONE stream instead: it keeps the
    # samples it could not finish (from the first incomplete frame, or from the first trial whose window did not fit),
    # prepends them to the next buffer, and emits every frame exactly once -- loop A's trial arithmetic and lock rule
    # (every acceptable lock, each more than 2CP+N samples after the one before), the LS estimate of that lock, and
    # loop B's equaliser for the frame's D data symbols with that lock's estimate.  Output: the equalised data symbols
    # of the frames completed in this call, in stream order, at out[0:last_written].  The result does not depend on
    # how the stream is cut into buffers (tests/test_gpu_blocks.py: chunked == whole, bit for bit).
    def _init_stream(self):
        self._tail = np.zeros(0, np.complex64)
        self._stream_prev = None           # window start of the last emitted lock, relative to the current tail
        self.stream_frames = 0             # frames emitted so far
        self.stream_locks = []             # (absolute window start, delay, int(peak)) of every emitted frame
        self._stream_origin = 0            # absolute sample index of _tail[0]

    def work_stream(self, in0):
        """-> complex64 [frames * D, ndb]: equalised data symbols of the frames completed by this buffer."""
        c = self.cfg
        N, CP, L = c.nfft, c.cp_len, c.sym_len
        buf = np.concatenate((self._tail, np.ascontiguousarray(in0, dtype=np.complex64)))
        n = len(buf)
        nv = self._n_valid_trials(n)
        locks = []
        cut = None                                               # where the kept tail starts
        if nv > 0:
            rx = self._upload(buf)
            peak, delay = self._search(rx, n)[:2]
            prev = self._stream_prev
            for P in np.nonzero(peak > c.gate * c.mm)[0]:
                start = int(P) * c.stride + CP
                if prev is not None and start - prev <= 2 * CP + N:
                    continue
                if start + (c.m0 + c.d - 1) * L + N > n:         # the frame's last data window is not here yet
                    cut = start - CP
                    break
                locks.append((start, int(delay[P]), int(peak[P])))
                prev = start
            self._stream_prev = prev
        first_untried = nv * c.stride                            # start - CP of the first trial that did not fit
        if cut is None:
            cut = first_untried
        if locks:
            cut = max(cut, 0)
        cut = max(0, min(cut, n))
        out = np.zeros((0, c.num_data_bins), np.complex64)
        if locks:
            k = len(locks)
            pos = [s for s, _, _ in locks]
            dv = self._to_device(i64=(pos,), i32=([d for _, d, _ in locks],))
            h, _, _ = self.modem.chanest(rx, dv[0], dv[1])
            starts = [s + (c.m0 + j) * L for s in pos for j in range(c.d)]
            dls = [d for _, d, _ in locks for _ in range(c.d)]
            rows = [i for i in range(k) for _ in range(c.d)]
            dv = self._to_device(i64=(starts,), i32=(dls, rows))
            eq, _, _ = self.modem.demod(rx, dv[0], dv[1], dv[2], h)
            out = eq.cpu().numpy()
            self.stream_locks += [(self._stream_origin + s, d, p) for s, d, p in locks]
            self.stream_frames += k
        self._tail = buf[cut:].copy()
        self._stream_origin += cut
        if self._stream_prev is not None:
            self._stream_prev -= cut
        return out

    def _work_carry_over(self, input_items, output_items):
        out = output_items[0]
        eq = self.work_stream(input_items[0]).reshape(-1)
        m = min(len(out), eq.size)
        out[0:m] = eq[:m]
        self.last_written = m
        self.count += 1
        return len(out)
