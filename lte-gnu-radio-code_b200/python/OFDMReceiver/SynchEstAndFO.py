"""OFDMReceiver.SynchEstAndFO -- CUDA-backed drop-in for the legacy receiver with a
frequency-offset hypothesis bank, GRR/LEGACY/gr-ofdm-rx/python/SynchEstAndFO.py
(ctor :31-227, work :230-364; GRC template OFDMReceiver.SynchEstAndFO(case, fo_range,
directory_name, file_name_cest, diagnostics), grc/OFDMReceiver_SynchEstAndFO.block.yml:33-35).

It is the gr-RXOFDM receiver (Zadoff-Chu root 37, stride CP-1, gate 0.4, linear `SNR`, first data
symbol after each lock) with the numerology taken from the hard-coded case table (:36-134) and,
per trial, one FFT + delay bank per row of exp(j*2pi*fo/fs*n) (:196, :247-277).  Two behaviours of
the reference are kept because they change its output (oracle/blocks_oracle.py, SynchEstAndFOOracle):
the channel estimate of a lock is taken from the LAST hypothesis' spectrum (:262-266 vs :293), and
every data window of a call is de-rotated with the winner of the last trial evaluated (:274, :332).
`cp_len = nfft/4` (:40, a float on Python 3) is made integral.
"""
import datetime

import numpy as np

from ofdm_b200.gr_compat import gr
from ofdm_b200.rx_block import RxBlockBase

# case -> (num_ofdm_symb, fs, nfft, synch_dat, num_data_bins)     SynchEstAndFO.py:36-134
CASES = {
    0: (48, 960000, 64, [1, 1], 12),
    1: (48, 960000, 64, [1, 1], 36),
    2: (48, 960000, 64, [2, 1], 180),
    3: (48, 960000, 64, [2, 1], 48),
    4: (48, 960000, 64, [3, 1], 24),
    5: (48, 960000, 64, [2, 1], 24),
    6: (24, 1920000, 128, [3, 1], 24),
    7: (24, 1920000, 128, [5, 1], 100),
    8: (12, 3840000, 256, [5, 1], 36),
    9: (12, 3840000, 256, [2, 1], 180),
}


class SynchEstAndFO(RxBlockBase):
    compat = "rxofdm"

    def __init__(self, case, fo_range, directory_name, file_name_cest, diagnostics, *, modulation=2, zc_root=None,
                 device=None):
        gr.sync_block.__init__(self, name="SynchEstAndFO", in_sig=[np.complex64], out_sig=[np.complex64])
        if case not in CASES:
            raise ValueError("Error: Case Out of Bounds")            # the reference prints this and then fails (:133)
        num_ofdm_symb, fs, nfft, synch_dat, num_data_bins = CASES[case]
        if num_data_bins >= nfft:
            raise ValueError("case %d asks for %d data bins on a %d-point grid (the bins alias); no transmitter in "
                             "the reference produces such a signal" % (case, num_data_bins, nfft))
        self.case = case
        self.fs = fs
        self.fo_range = list(fo_range)
        self.diagnostics = diagnostics
        self._setup(num_ofdm_symb, nfft, nfft // 4, nfft - 2, synch_dat, num_data_bins, 100000000, 0.4,
                    directory_name, file_name_cest, diagnostics, 0, modulation, zc_root, rows=100, device=device)
        self.modem.set_cfo_bank(self.fo_range, fs)
        self.cor_obs = -1
        self.dmax_tmp_ind = 0
        self.time_synch_ref = np.zeros((self.max_num_corr, 3))

    def work(self, input_items, output_items):
        in0 = input_items[0]
        out = output_items[0]
        c = self.cfg
        n = len(in0)
        N, CP, L = c.nfft, c.cp_len, c.sym_len
        rx = self._upload(in0)
        last_h = len(self.fo_range) - 1

        # loop A (:235-325): every trial and every hypothesis on the GPU; the accept rule is sequential but tiny
        res = self._search(rx, n)
        new_rows, new_pos, new_delay = [], [], []
        if len(res[0]):
            peak, delay, hyp = res
            self.dmax_tmp_ind = int(hyp[-1])                       # winner of the last trial evaluated (:274)
            for P in np.nonzero(peak > 0.4 * c.mm)[0]:
                start = int(P) * c.stride + CP
                prev = self.time_synch_ref[max(self.cor_obs, 0)][0]
                if (start - prev > 2 * CP + N) or self.cor_obs == -1:
                    if self.cor_obs + 1 >= self.max_num_corr:
                        break                                      # the reference would raise IndexError here
                    self.cor_obs += 1
                    self.time_synch_ref[self.cor_obs] = (start, int(delay[P]), int(peak[P]))
                    new_rows.append(self.cor_obs)
                    new_pos.append(start)
                    new_delay.append(int(delay[P]))
        if new_rows:
            # synchdat is whatever the hypothesis loop left behind: the last row of the bank (:262-266, :293)
            ht = self._estimate_into(rx, new_rows, new_pos, new_delay, hyp=[last_h] * len(new_rows))
            if self.diagnostics == 1:
                stamp = datetime.datetime.now().strftime("%Y_%m_%d_%Hh_%Mm")
                self._dump_chan(ht[-1], str(self.directory_name) + str(self.file_name_cest) + stamp + ".pckl")

        # loop B (:326-352): first data symbol after each lock, de-rotated with self.dmax_tmp_ind
        rows, starts, delays = [], [], []
        for r in range(self.cor_obs + 1):
            pos = self.time_synch_ref[r][0]
            if pos + c.m0 * L + N - 1 <= n:
                rows.append(r)
                starts.append(int(pos + c.m0 * L))
                delays.append(int(self.time_synch_ref[r][1]))
        self._demod_into(rx, rows, starts, delays, rows, hyp=[self.dmax_tmp_ind] * len(rows))

        corr_size = min(int(self.num_ofdm_symb / c.frame_syms), self.rows)      # :356 made integral
        data_out = self.est_data_freq[0:corr_size].reshape(-1)
        self.last_written = 0
        if self.count > 0:
            m = min(len(out), data_out.size)
            out[0:m] = data_out[:m]
            self.last_written = m
        self.count += 1
        self.cor_obs = 0
        return len(output_items[0])
