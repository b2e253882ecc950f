"""utsa_ofdm.SynchAndChanEst -- CUDA-backed drop-in for
GRR/gr-utsa_ofdm/python/SynchAndChanEst.py (ctor :17-116, work :135-262).

Same class name, positional constructor arguments, complex64 stream in/out and work() contract.
Differences from the reference are the ones SURVEY.md 2.3 lists as defects that stop it running
(np.complex, plt.show() in the constructor, 12-vs-13 argument GRC template -> `channel` now has a
default); its arithmetic, lock rule and packing -- including "nothing written on the first call"
and "counter reset to 0 after each call" -- are kept.
"""
import numpy as np

from ofdm_b200.gr_compat import gr
from ofdm_b200.rx_block import RxBlockBase


class SynchAndChanEst(RxBlockBase):
    compat = "utsa"

    def __init__(self, num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr,
                 scale_factor_gate, directory_name, file_name_cest, diagnostics, genie, channel="Fading", *,
                 modulation=2, zc_root=None, device=None, fused=None, carry_over=False):
        gr.sync_block.__init__(self, name="SynchAndChanEst", in_sig=[np.complex64], out_sig=[np.complex64])
        self.channel = channel
        self.scale_factor_gate = scale_factor_gate
        self.diagnostic = diagnostics
        self._setup(num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, scale_factor_gate,
                    directory_name, file_name_cest, diagnostics, genie, modulation, zc_root,
                    rows=int(num_ofdm_symb), device=device, fused=fused)
        self.carry_over = bool(carry_over)
        self._init_stream()
        self._keep_rows = None
        self.SNR_lin = 10 ** (self.SNR / 20)
        self.corr_obs = -1
        self.time_synch_ref = np.zeros(3)

    def work_fused(self, in0):
        """Loops A and B in one C-ABI call (ofdm_rx_block_work, mode 0); the lock record comes back with it."""
        r = self._fused_call(in0)
        self.time_synch_ref[:] = self._tsr_view[0]
        if r.n_new_locks and self.diagnostic == 1:
            self._dump_chan(self._last_chan_time[0], str(self.directory_name) + "_" + str(self.file_name_cest))

    def work(self, input_items, output_items):
        if self.carry_over:                                  # additive streaming mode, see ofdm_b200/rx_block.py
            return self._work_carry_over(input_items, output_items)
        in0 = input_items[0]
        out = output_items[0]
        c = self.cfg
        n = len(in0)
        N, CP, L, fsyms = c.nfft, c.cp_len, c.sym_len, c.frame_syms
        n_unique_symb = n // L
        n_data_symb = int(n_unique_symb * (c.d / fsyms))
        if self._blk:
            self.work_fused(in0)
            return self._pack(out, output_items, n_data_symb)
        rx = self._upload(in0)

        # loop A: all trials on the GPU, then the first acceptable lock (:143-219)
        peak, delay = self._search(rx, n)
        cand = np.nonzero(peak > c.gate * c.mm)[0]
        if cand.size:
            start = cand * c.stride + CP
            ok = (start - self.time_synch_ref[0] > 2 * CP + N) | (self.corr_obs == -1)
            hit = np.nonzero(ok)[0]
            if hit.size:
                P = int(cand[hit[0]])
                self.corr_obs += 1
                if self.corr_obs < self.rows:
                    self.time_synch_ref[:] = (P * c.stride + CP, int(delay[P]), int(peak[P]))
                    ht = self._estimate_into(rx, [self.corr_obs], [P * c.stride + CP], [int(delay[P])])
                    if self.diagnostic == 1:
                        self._dump_chan(ht[0], str(self.directory_name) + "_" + str(self.file_name_cest))

        # loop B: every data symbol of every frame, channel row 0 (:221-248)
        pos, d = int(self.time_synch_ref[0]), int(self.time_synch_ref[1])
        rows, starts = [], []
        for P in range(0, n_unique_symb, fsyms):
            base = pos + c.m0 * L * (P + 1)
            if base + N - 1 <= n and P + c.d <= self.rows:
                for k in range(c.d):
                    rows.append(P + k)
                    starts.append(base + L * k)
        self._demod_into(rx, rows, starts, [d] * len(rows), [0] * len(rows))

        return self._pack(out, output_items, n_data_symb)

    def _pack(self, out, output_items, n_data_symb):
        c, fsyms = self.cfg, self.cfg.frame_syms
        # packing (:249-262)
        if self._keep_rows is None:                         # np.delete(est_data_freq, range(3, rows, M0+D), axis=0), :249
            self._keep_rows = np.setdiff1d(np.arange(self.rows), np.arange(3, self.rows, fsyms))
        data_demod = self.est_data_freq[self._keep_rows]
        data_out = data_demod.reshape(-1)[: n_data_symb * c.num_data_bins]
        self.last_written = 0
        if self.count > 0:
            m = min(len(out), data_out.size)
            out[0:m] = data_out[:m]
            self.last_written = m
        self.count += 1
        self.corr_obs = 0
        return len(output_items[0])
