"""OFDMReceiver.SynchEstFOAndDSSS -- CUDA-backed drop-in for
GRR/LEGACY/gr-ofdm-rx/python/SynchEstFOAndDSSS.py (ctor :33-266, work :268-410; GRC template
OFDMReceiver.SynchEstFOAndDSSS(case, fo_range, directory_name, file_name_cest, diagnostics)).

SynchEstAndFO (same search, hypothesis bank, channel estimate and first-data-symbol demod, with the same
two quirks) followed by DSSS despreading of the equalised data bins with a length-DSSS Zadoff-Chu code
(:253-261, :392-397).  Kept as written in the reference: a lock whose data window does not fit the buffer
first receives a copy of the most recent demodulated row (:392 sits outside the guard of :366), and the
despread rows are written on every call, the first included (:404-406).
"""
import datetime

import numpy as np

from ofdm_b200.gr_compat import gr
from ofdm_b200.rx_block import RxBlockBase

# case -> (num_ofdm_symb, fs, nfft, synch_dat, num_data_bins, DSSS)     SynchEstFOAndDSSS.py:37-157
CASES = {
    0: (48, 960000, 64, [2, 1], 12, 1),
    1: (48, 960000, 64, [3, 1], 36, 6),
    2: (45, 960000, 64, [4, 1], 48, 6),
    3: (45, 960000, 64, [4, 1], 48, 12),
    4: (24, 1920000, 128, [3, 1], 32, 8),
    5: (20, 1920000, 128, [4, 1], 84, 12),
    6: (20, 1920000, 128, [4, 1], 96, 16),
    7: (24, 1920000, 128, [5, 1], 120, 24),
    8: (12, 3840000, 256, [3, 1], 168, 12),
    9: (10, 3840000, 256, [4, 1], 192, 16),
    10: (10, 3840000, 256, [4, 1], 240, 24),
}


def spreading_code(sf, root):
    """Zadoff-Chu of length sf: exp(-j*pi*root*k^2/sf) (even sf) or k(k+1) (odd).  SynchEstFOAndDSSS.py:253-261."""
    k = np.arange(sf, dtype=np.float64)
    xx = k * k if sf % 2 == 0 else k * (k + 1)
    return np.exp((-1j * (2 * np.pi / sf) * root / 2.0) * xx)


class SynchEstFOAndDSSS(RxBlockBase):
    compat = "rxofdm"

    def __init__(self, case, fo_range, directory_name, file_name_cest, diagnostics, *, modulation=2, zc_root=None,
                 device=None):
        gr.sync_block.__init__(self, name="SynchEstFOAndDSSS", in_sig=[np.complex64], out_sig=[np.complex64])
        if case not in CASES:
            raise ValueError("Error: Case Out of Bounds")
        num_ofdm_symb, fs, nfft, synch_dat, num_data_bins, dsss = CASES[case]
        self.case = case
        self.fs = fs
        self.fo_range = list(fo_range)
        self.diagnostics = diagnostics
        self.DSSS = dsss
        self._setup(num_ofdm_symb, nfft, nfft // 4, nfft - 2, synch_dat, num_data_bins, 100000000, 0.4,
                    directory_name, file_name_cest, diagnostics, 0, modulation, zc_root, rows=100, device=device)
        self.modem.set_cfo_bank(self.fo_range, fs)
        self.SC = spreading_code(dsss, self.cfg.zc_root)
        self._d_SC = self._torch.from_numpy(self.SC.astype(np.complex64)).to(self.modem.device)
        self.est_data_freq_d = np.zeros((self.max_num_corr, num_data_bins // dsss), dtype=complex)
        self.cor_obs = -1
        self.dmax_tmp_ind = 0
        self._have_row = False
        self.time_synch_ref = np.zeros((self.max_num_corr, 3))

    def work(self, input_items, output_items):
        in0 = input_items[0]
        out = output_items[0]
        c = self.cfg
        n = len(in0)
        N, CP, L = c.nfft, c.cp_len, c.sym_len
        torch = self._torch
        rx = self._upload(in0)
        last_h = len(self.fo_range) - 1

        # loop A (:273-357): as in SynchEstAndFO
        res = self._search(rx, n)
        new_rows, new_pos, new_delay = [], [], []
        if len(res[0]):
            peak, delay, hyp = res
            self.dmax_tmp_ind = int(hyp[-1])
            for P in np.nonzero(peak > 0.4 * c.mm)[0]:
                start = int(P) * c.stride + CP
                prev = self.time_synch_ref[max(self.cor_obs, 0)][0]
                if (start - prev > 2 * CP + N) or self.cor_obs == -1:
                    if self.cor_obs + 1 >= self.max_num_corr:
                        break
                    self.cor_obs += 1
                    self.time_synch_ref[self.cor_obs] = (start, int(delay[P]), int(peak[P]))
                    new_rows.append(self.cor_obs)
                    new_pos.append(start)
                    new_delay.append(int(delay[P]))
        if new_rows:
            ht = self._estimate_into(rx, new_rows, new_pos, new_delay, hyp=[last_h] * len(new_rows))
            if self.diagnostics == 1:
                stamp = datetime.datetime.now().strftime("%Y_%m_%d_%Hh_%Mm")
                self._dump_chan(ht[-1], str(self.directory_name) + str(self.file_name_cest) + stamp + ".pckl")

        # loop B (:360-397): demod where the window fits; the other locks inherit the most recent demodulated row
        rows, starts, delays = [], [], []
        for r in range(self.cor_obs + 1):
            pos = self.time_synch_ref[r][0]
            if pos + c.m0 * L + N - 1 <= n:
                rows.append(r)
                starts.append(int(pos + c.m0 * L))
                delays.append(int(self.time_synch_ref[r][1]))
        self._demod_into(rx, rows, starts, delays, rows, hyp=[self.dmax_tmp_ind] * len(rows))
        fitted = set(rows)
        last = None
        for r in range(self.cor_obs + 1):
            if r in fitted:
                last = self.est_data_freq[r]
            elif last is not None:
                self.est_data_freq[r] = last
        nrows = self.cor_obs + 1
        if nrows > 0:
            d_eq = torch.from_numpy(self.est_data_freq[:nrows].astype(np.complex64)).to(self.modem.device)
            self.est_data_freq_d[:nrows] = self.modem.dsss_despread(d_eq, self._d_SC).cpu().numpy()

        corr_size = min(int(self.num_ofdm_symb / c.frame_syms), self.rows)
        data_out = self.est_data_freq_d[0:corr_size].reshape(-1)
        m = min(len(out), data_out.size)
        out[0:m] = data_out[:m]                                    # :406: no count guard in this block
        self.last_written = m
        self.count += 1
        self.cor_obs = 0
        return len(output_items[0])
