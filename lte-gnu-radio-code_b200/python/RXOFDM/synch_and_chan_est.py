"""RXOFDM.synch_and_chan_est -- CUDA-backed drop-in for
GRR/gr-RXOFDM/python/synch_and_chan_est.py (ctor :17-124, work :140-266).

Variant constants kept: Zadoff-Chu root 37 keyed on the parity of num_synch_bins (:54-64), search
stride CP-1 (:81), gate 0.4 (:170), `snr` used as a linear regulariser (:184,:247), every lock
recorded (no break), only the first data symbol after each lock demodulated (:224-250), the first
num_ofdm_symb/sum(synch_dat) rows written from the second call on (:253-263), counter reset to 0
(:265).  Pass zc_root=23 to receive what the repository's transmitter actually sends.
"""
import datetime

import numpy as np

from ofdm_b200.gr_compat import gr
from ofdm_b200.rx_block import RxBlockBase


class synch_and_chan_est(RxBlockBase):
    compat = "rxofdm"

    def __init__(self, num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr,
                 directory_name, file_name_cest, diagnostics, genie, *, modulation=2, zc_root=None, device=None,
                 fused=None, carry_over=False):
        gr.sync_block.__init__(self, name="SynchAndChanEst", in_sig=[np.complex64], out_sig=[np.complex64])
        self.diagnostics = diagnostics
        self._setup(num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr, 0.4,
                    directory_name, file_name_cest, diagnostics, genie, modulation, zc_root, rows=100,
                    device=device, fused=fused)
        self.carry_over = bool(carry_over)
        self._init_stream()
        self.cor_obs = -1
        self.time_synch_ref = np.zeros((self.max_num_corr, 3))

    def work_fused(self, in0):
        """Loops A and B in one C-ABI call (ofdm_rx_block_work, mode 1)."""
        r = self._fused_call(in0)
        k = r.obs_after_search + 1
        if k > 0:
            self.time_synch_ref[:k] = self._tsr_view[:k]
        if r.n_new_locks and self.diagnostics == 1:
            stamp = datetime.datetime.now().strftime("%Y_%m_%d_%Hh_%Mm")
            self._dump_chan(self._last_chan_time[-1], str(self.directory_name) + stamp + str(self.file_name_cest) + ".pckl")

    def work(self, input_items, output_items):
        if self.carry_over:                                  # additive streaming mode, see ofdm_b200/rx_block.py
            return self._work_carry_over(input_items, output_items)
        in0 = input_items[0]
        out = output_items[0]
        c = self.cfg
        n = len(in0)
        N, CP, L = c.nfft, c.cp_len, c.sym_len
        if self._blk:
            self.work_fused(in0)
            return self._pack(out, output_items)
        rx = self._upload(in0)

        # loop A (:146-221): every trial on the GPU; the accept rule is sequential but tiny
        peak, delay = self._search(rx, n)
        cand = np.nonzero(peak > 0.4 * c.mm)[0]
        new_rows, new_pos, new_delay = [], [], []
        for P in cand:
            start = int(P) * c.stride + CP
            prev = self.time_synch_ref[max(self.cor_obs, 0)][0]
            if (start - prev > 2 * CP + N) or self.cor_obs == -1:
                if self.cor_obs + 1 >= self.max_num_corr:
                    break                      # the reference would raise IndexError here
                self.cor_obs += 1
                self.time_synch_ref[self.cor_obs] = (start, int(delay[P]), int(peak[P]))
                new_rows.append(self.cor_obs)
                new_pos.append(start)
                new_delay.append(int(delay[P]))
        if new_rows:
            ht = self._estimate_into(rx, new_rows, new_pos, new_delay)
            if self.diagnostics == 1:
                stamp = datetime.datetime.now().strftime("%Y_%m_%d_%Hh_%Mm")
                self._dump_chan(ht[-1], str(self.directory_name) + stamp + str(self.file_name_cest) + ".pckl")

        # loop B (:224-250): first data symbol after each lock, with that lock's channel row
        rows, starts, delays = [], [], []
        for r in range(self.cor_obs + 1):
            pos = self.time_synch_ref[r][0]
            if pos + c.m0 * L + N - 1 <= n:
                rows.append(r)
                starts.append(int(pos + c.m0 * L))
                delays.append(int(self.time_synch_ref[r][1]))
        self._demod_into(rx, rows, starts, delays, rows)

        return self._pack(out, output_items)

    def _pack(self, out, output_items):
        c = self.cfg
        corr_size = min(int(self.num_ofdm_symb / c.frame_syms), self.rows)      # :253 made integral
        data_out = self.est_data_freq[0:corr_size].reshape(-1)
        self.last_written = 0
        if self.count > 0:
            m = min(len(out), data_out.size)
            out[0:m] = data_out[:m]
            self.last_written = m
        self.count += 1
        self.cor_obs = 0
        return len(output_items[0])
