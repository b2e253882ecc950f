"""OFDMReceiver.SynchronizeAndEstimate -- CUDA-backed drop-in for the pointer-tracking receiver,
GRR/LEGACY/gr-ofdm-rx/python/SynchronizeAndEstimate.py (ctor :31-180, work :186-441; GRC template
OFDMReceiver.SynchronizeAndEstimate(case), grc/OFDMReceiver_SynchronizeAndEstimate.block.yml).

The reference steers one FFT window through the buffer: it strides by ceil(CP/2) until the first lock, then
jumps a frame at a time and, from the sixth lock on, predicts the next frame start from a straight-line fit
through its last five (pointer + delay) observations (:231-236, :333-351).  Every step costs it one window
transform and a dense (CP+1)-delay bank, and where the window goes next depends on what the last one found, so
the loop cannot be batched as it stands.  Here the GPU evaluates the synchronisation metric -- peak and arg-max
over the delay bank -- at EVERY sample offset of the buffer in one launch (the stride-1 time-domain search,
ofdm_rx_sync_search), after which the reference's control flow, quirks included (oracle/blocks_oracle.py:
SynchronizeAndEstimateOracle lists them), is a walk over that table on the host: a few dozen scalar decisions per
call.  Channel estimates of the accepted locks and all data symbols are then two more launches.
Only case 0 (SISO, NFFT 64) exists; case 1 is the reference's unfinished 2x2 profile.
"""
import numpy as np

from ofdm_b200.gr_compat import gr
from ofdm_b200.rx_block import RxBlockBase


class SynchronizeAndEstimate(RxBlockBase):
    compat = "rxofdm"          # `SNR` enters as a linear regulariser (:362, :389, :429), parity keyed like gr-RXOFDM

    def __init__(self, case, *, modulation=2, device=None):
        gr.sync_block.__init__(self, name="SynchronizeAndEstimate", in_sig=[np.complex64], out_sig=[np.complex64])
        if case != 0:
            raise ValueError("case %r: only the SISO profile (case 0) is a complete receiver in the reference" % (case,))
        self.case = case
        # :84-103 for case 0: NFFT = 2^ceil(log2(round(0.97*960e3/15e3))) = 64, CP = NFFT/4, 60 data bins, 62 synch bins
        self._setup(48, 64, 16, 62, [1, 3], 60, 100, 0.5, "", "", 0, 0, modulation, 23, rows=12, device=device, fused=False)
        self.lmax_s, self.lmax_d = 12, 36
        self.corr_obs = -1
        self.time_synch_ref = np.zeros((250, 3))
        self.est_chan_freq_p = np.zeros((self.lmax_s, self.nfft), dtype=complex)
        self.est_chan_impulse = np.zeros((self.lmax_s, self.nfft), dtype=complex)
        self.est_synch_freq = np.zeros((self.lmax_s, self.cfg.mm), dtype=complex)
        self.est_data_freq = np.zeros((self.lmax_d, self.num_data_bins), dtype=complex)

    def work(self, input_items, output_items):
        in0 = input_items[0]
        out = output_items[0]
        c = self.cfg
        n = len(in0)
        N, CP, L, fsum = c.nfft, c.cp_len, c.sym_len, c.frame_syms
        torch = self._torch
        rx = self._upload(in0)
        # the metric at every window position the guard (M0-1)*L + N + ptr < n admits (:239)
        n_pos = n - N - (c.m0 - 1) * L
        if n_pos > 0:
            packed = self.modem.empty((2 * n_pos,), torch.int32)
            self.modem.sync_search(rx, 0, 1, n_pos, packed=packed)
            host = packed.cpu().numpy()
            peak, delay0 = host[:n_pos].view(np.float32), host[n_pos:]
        stride = int(np.ceil(CP / 2))
        start_samp = (CP - 4) - 1
        total_loops = int(np.ceil(n / stride))
        self.corr_obs = -1
        ptr_frame, b, xp = 0, 0, []
        ptr_adj, loop_count, sym_count = 0, 0, 0
        tap_delay = 5
        x = np.zeros(tap_delay)
        ptr_synch0 = np.zeros(1000)
        locks = []                                               # (row, window the estimate is taken from, delay-bank column)
        while loop_count <= total_loops:
            if self.corr_obs == -1:
                ptr_frame = loop_count * stride + start_samp + ptr_adj
            elif self.corr_obs < 5:
                ptr_frame += fsum * L
            else:
                ptr_frame = (np.ceil(np.dot(xp[-1:], b) - CP / 4))[0]
            if (c.m0 - 1) * L + N + ptr_frame < n:
                win = int(ptr_frame)
                dmax, dmax_ind = float(peak[win]), int(delay0[win]) - 1                    # :278
                if dmax > 0.5 * c.mm or self.corr_obs > -1:
                    if dmax_ind > np.ceil(0.75 * CP):
                        if self.corr_obs == -1:
                            ptr_adj += np.ceil(0.5 * CP)
                            ptr_frame = loop_count * stride + start_samp + ptr_adj
                        elif self.corr_obs < 5:
                            ptr_frame += np.ceil(0.5 * CP)
                    time_synch_ind = self.time_synch_ref[max(self.corr_obs, 1), 0]
                    if ptr_frame - time_synch_ind > (2 * CP + N) or self.corr_obs == -1:
                        if self.corr_obs + 1 < self.lmax_s:
                            self.corr_obs += 1
                            r = self.corr_obs
                            self.time_synch_ref[r] = (ptr_frame, dmax_ind, dmax)
                            ptr_synch0[sym_count % tap_delay] = ptr_frame + dmax_ind
                            x[sym_count % tap_delay] = sym_count * fsum
                            sym_count += 1
                            x2 = x[0:min(r, tap_delay)]
                            x_plus = np.concatenate((x2, np.atleast_1d(sym_count * fsum)))
                            xp = np.zeros((len(x_plus), 2))
                            xp[:, 0] = 1.0
                            xp[:, 1] = x_plus
                            if r > 3:
                                y = ptr_synch0[0:min(tap_delay, r)]
                                X = np.zeros((len(x2), 2))
                                X[:, 0] = 1.0
                                X[:, 1] = x2
                                b = np.linalg.lstsq(X, y, rcond=None)[0]
                            # p_mat[:, dmax_ind] with dmax_ind = -1 is the LAST column of the bank (:354)
                            locks.append((r, win, dmax_ind if dmax_ind >= 0 else CP))
            loop_count += 1
        if locks:
            rows = [k[0] for k in locks]
            dv = self._to_device(i64=([k[1] for k in locks], rows), i32=([k[2] for k in locks],))
            h, ht, se = self.modem.chanest(rx, dv[0], dv[2], want_aux=True)
            self._d_H.index_copy_(0, dv[1], h)
            aux = self.modem.aux_flat.cpu().numpy()
            k = len(rows)
            self.est_chan_impulse[rows] = aux[:k * N].reshape(k, N)
            self.est_synch_freq[rows] = aux[k * N:].reshape(k, c.mm)
            self.est_chan_freq_p[rows] = h.cpu().numpy()
        # data symbols (:398-441)
        sym_rows, starts, delays, hrows = [], [], [], []
        for r in range(self.corr_obs + 1):
            for k in range(c.d):
                if np.sum(self.time_synch_ref[r]) + N < n:
                    sym_rows.append(r * c.d + k)
                    starts.append(int(self.time_synch_ref[r, 0] + (k + 1) * L))
                    delays.append(int(self.time_synch_ref[r, 1]))
                    hrows.append(r)
        self.last_written = 0
        if sym_rows:
            dv = self._to_device(i64=(starts,), i32=(delays, hrows))
            eq, _, _ = self.modem.demod(rx, dv[0], dv[1], dv[2], self._d_H)
            eq = eq.cpu().numpy().astype(complex)
            for i, row in enumerate(sym_rows):                   # the rescaling by row p's power is sequential (:434-437)
                self.est_data_freq[row] = eq[i]
                data = self.est_data_freq[hrows[i]]
                self.est_data_freq[row] = self.est_data_freq[row] / np.sqrt(np.sum(data * np.conj(data)) / len(data))
            m = min(len(out), c.num_data_bins)
            out[0:m] = self.est_data_freq[sym_rows[-1]][:m]     # every symbol is written to out[0:60]; the last one stays
            self.last_written = m
        return len(output_items[0])
