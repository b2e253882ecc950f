"""Stateful restatement of the two live RX blocks' work() (TEST INFRASTRUCTURE ONLY).

These classes reproduce, call for call, what a fresh instance of the reference block computes and
writes -- including the behaviours SURVEY.md 2.3 lists as D7 (nothing written on the first call,
full length always returned), D8 (correlation counter reset to 0, row 0 persists) and D13 (which
SNR regulariser goes where).  Defects that merely stop the reference from running on Python 3 /
NumPy 2 (np.complex, float slice bounds, undefined attribute names) are fixed, not reproduced.

  UtsaBlockOracle   <- GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:17-116 (ctor), :135-262 (work)
  RxofdmBlockOracle <- GRR/gr-RXOFDM/python/synch_and_chan_est.py:17-124 (ctor), :140-266 (work)
"""
import numpy as np

from . import ofdm_oracle as oo


class UtsaBlockOracle:
    def __init__(self, num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr,
                 scale_factor_gate, literal=False, zc_root=None):
        self.prm = oo.ModemParams(nfft=nfft, cp_len=cp_len, num_synch_bins=num_synch_bins,
                                  num_data_bins=num_data_bins, m0=synch_dat[0], d=synch_dat[1],
                                  snr=snr, gate=scale_factor_gate, compat="utsa", zc_root=zc_root)
        self.num_ofdm_symb = num_ofdm_symb
        self.literal = literal
        p = self.prm
        self.zc = oo.zc_for(p)
        self.bank = oo.delay_bank(p)
        self.corr_obs = -1                                         # :72
        self.time_synch_ref = np.zeros(3)                          # :83
        self.est_chan_time = np.zeros((num_ofdm_symb, nfft), dtype=complex)
        self.est_synch_freq = np.zeros((num_ofdm_symb, p.mm), dtype=complex)
        self.est_chan_freq_P = np.zeros((num_ofdm_symb, nfft), dtype=complex)
        self.est_data_freq = np.zeros((num_ofdm_symb, num_data_bins), dtype=complex)
        self.count = 0
        self.last_written = 0

    def work(self, in0, out):
        p = self.prm
        N, CP, L = p.nfft, p.cp_len, p.sym_len
        fsyms = p.frame_syms
        n_trials = int(np.around(len(in0) / p.stride))             # :139
        n_unique = int(np.floor(len(in0) / L))                     # :140
        n_data = int(n_unique * (p.d / fsyms))                     # :141

        # loop A: sliding search, stops at the first accepted lock (:143-219)
        for P in range(n_trials):
            if p.m0 * L + P * p.stride + N + CP < len(in0):        # :144
                start = P * p.stride + CP
                d_i, d_v, sd = oo.sync_trial(in0, start, p, self.zc, self.bank, self.literal)
                if d_v > p.gate * p.mm:                            # :166
                    if (start - self.time_synch_ref[0] > 2 * CP + N) or self.corr_obs == -1:
                        self.corr_obs += 1
                        self.time_synch_ref[:] = (start, d_i, int(d_v))
                        ce, ce1, dr = oo.chan_est(sd, d_i, p, self.zc, self.bank, self.literal)
                        self.est_chan_freq_P[self.corr_obs] = ce1
                        tim, esf = oo.chan_est_aux(ce, ce1, dr, p)
                        self.est_chan_time[self.corr_obs] = tim
                        self.est_synch_freq[self.corr_obs] = esf
                        break                                      # :219

        # loop B: every data symbol of every frame, with row 0's channel (:221-248)
        pos, delay = self.time_synch_ref[0], self.time_synch_ref[1]
        for P in range(0, n_unique, fsyms):
            base = pos + p.m0 * L * (P + 1)                        # :222
            if base + N - 1 <= len(in0):                           # :223
                ptr = int(base)
                for n in range(p.d):
                    self.est_data_freq[P + n] = oo.demod_symbol(
                        in0, ptr + L * n, delay, self.est_chan_freq_P[0], p, self.literal)

        # packing (:249-262): drop the never-written row of each frame, flatten
        rows = self.est_data_freq.shape[0]
        data_demod = np.delete(self.est_data_freq, list(range(3, rows, fsyms)), axis=0)
        data_out = np.reshape(data_demod, (1, n_data * p.num_data_bins))
        self.last_written = 0
        if self.count > 0:                                         # :257 (D7)
            out[0:data_out.shape[1]] = data_out[0]
            self.last_written = data_out.shape[1]
        self.count += 1
        self.corr_obs = 0                                          # :261 (D8)
        return len(out)


class RxofdmBlockOracle:
    def __init__(self, num_ofdm_symb, nfft, cp_len, num_synch_bins, synch_dat, num_data_bins, snr,
                 literal=False, zc_root=None):
        self.prm = oo.ModemParams(nfft=nfft, cp_len=cp_len, num_synch_bins=num_synch_bins,
                                  num_data_bins=num_data_bins, m0=synch_dat[0], d=synch_dat[1],
                                  snr=snr, compat="rxofdm", zc_root=zc_root)
        self.num_ofdm_symb = num_ofdm_symb
        self.literal = literal
        p = self.prm
        self.zc = oo.zc_for(p)
        self.bank = oo.delay_bank(p)
        self.max_num_corr = 100                                    # :85
        self.cor_obs = -1                                          # :76
        self.time_synch_ref = np.zeros((self.max_num_corr, 3))
        self.est_chan_time = np.zeros((self.max_num_corr, nfft), dtype=complex)
        self.est_synch_freq = np.zeros((self.max_num_corr, p.mm), dtype=complex)
        self.est_chan_freq_P = np.zeros((self.max_num_corr, nfft), dtype=complex)
        self.est_data_freq = np.zeros((self.max_num_corr, num_data_bins), dtype=complex)
        self.count = 0
        self.last_written = 0

    def work(self, in0, out):
        p = self.prm
        N, CP, L = p.nfft, p.cp_len, p.sym_len
        n_trials = int(np.around(len(in0) / p.stride))             # :144

        # loop A: every accepted lock is recorded (no break) (:146-221)
        for P in range(n_trials):
            if p.m0 * L + P * p.stride + N + CP < len(in0):        # :147
                start = P * p.stride + CP
                d_i, d_v, sd = oo.sync_trial(in0, start, p, self.zc, self.bank, self.literal)
                if d_v > 0.4 * p.mm:                               # :170
                    prev = self.time_synch_ref[max(self.cor_obs, 0)][0]
                    if (start - prev > 2 * CP + N) or self.cor_obs == -1:
                        self.cor_obs += 1
                        r = self.cor_obs
                        self.time_synch_ref[r] = (start, d_i, int(d_v))
                        ce, ce1, dr = oo.chan_est(sd, d_i, p, self.zc, self.bank, self.literal)
                        self.est_chan_freq_P[r] = ce1
                        tim, esf = oo.chan_est_aux(ce, ce1, dr, p)
                        self.est_chan_time[r] = tim
                        self.est_synch_freq[r] = esf

        # loop B: the first data symbol after each lock (:224-250)
        for r in range(self.cor_obs + 1):
            pos = self.time_synch_ref[r][0]
            if pos + p.m0 * L + N - 1 <= len(in0):                 # :226
                ptr = int(pos + p.m0 * L)
                self.est_data_freq[r] = oo.demod_symbol(
                    in0, ptr, self.time_synch_ref[r][1], self.est_chan_freq_P[r], p, self.literal)

        corr_size = int(self.num_ofdm_symb / p.frame_syms)         # :253 (D6: made integral)
        data_out = np.reshape(self.est_data_freq[0:corr_size], (1, corr_size * p.num_data_bins))
        self.last_written = 0
        if self.count > 0:                                         # :260 (D7)
            out[0:data_out.shape[1]] = data_out[0]
            self.last_written = data_out.shape[1]
        self.count += 1
        self.cor_obs = 0                                           # :265 (D8)
        return len(out)
