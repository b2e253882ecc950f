/*
 * ofdm_b200.h -- C ABI of libofdm_b200.so, the B200 (sm_100a) implementation of the OFDM modem
 * hot path of tayloreisman16/LTE-GNU-Radio-Code.
 *
 * The reference has no native interface for this path: its blocks are pure-Python
 * gr.sync_block subclasses whose work() bodies call NumPy (the pybind11 modules are empty stubs,
 * GNU-Radio-Repositories/gr-RXOFDM/python/bindings/python_bindings.cc:37-53).  The entry points
 * below are therefore the functions a maintainer would bind from those work() bodies with ctypes
 * (see INTEGRATION.md); each one names the reference statements it replaces.  Paths are relative
 * to GNU-Radio-Repositories/ (GRR/).
 *
 * Conventions
 *   - plain C types only; complex samples are interleaved float pairs (numpy complex64 / float2);
 *   - every function returns 0 on success or a negative OFDM_E* code and never throws;
 *     ofdm_last_error() returns a thread-local description of the last failure;
 *   - pointers named d_* are DEVICE pointers owned by the caller, h_* are HOST pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all d_* entry
 *     points are asynchronous on it and never synchronise or allocate;
 *   - the library owns only the plan (twiddles, Zadoff-Chu, bin tables, optional staging buffers
 *     used by the h_* convenience entry points).
 *   - bits are packed MSB-first (numpy.packbits order); each OFDM data symbol owns
 *     ofdm_plan_hard_stride() bytes (= ceil(payload_bins*bps/8), payload_bins = ofdm_plan_payload_bins()
 *     = num_data_bins minus the comb pilots, if any); wherever a size below says "ndb" it means payload_bins.
 */
#ifndef OFDM_B200_H
#define OFDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDM_B200_ABI_VERSION 5

enum {
    OFDM_OK = 0,
    OFDM_EINVAL = -1,      /* bad argument / unsupported numerology */
    OFDM_ECUDA = -2,       /* a CUDA runtime call failed; see ofdm_last_error() */
    OFDM_ENOMEM = -3
};

enum { OFDM_EQ_MMSE = 0, OFDM_EQ_ZF = 1 };

/* Numerology + constants of one modem instance.  The first six fields are the reference block's
 * constructor arguments (GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:17-19,
 * GRR/gr-RXOFDM/python/synch_and_chan_est.py:17-18). */
typedef struct ofdm_params {
    int32_t nfft;            /* power of two, 64..4096 */
    int32_t cp_len;          /* 0 <= cp_len < nfft */
    int32_t num_synch_bins;  /* even, <= nfft (== nfft reproduces the +-N/2 alias of ofdm_chain.py:83) */
    int32_t num_data_bins;   /* even, < nfft */
    int32_t m0;              /* synch symbols per frame, synch_dat[0] */
    int32_t d;               /* data symbols per frame,  synch_dat[1] */
    int32_t bps;             /* bits per data bin: 1 BPSK, 2 QPSK, 4 16-QAM, 6 64-QAM */
    int32_t zc_root;         /* 23 (gr-utsa_ofdm, TX simulator) or 37 (gr-RXOFDM) */
    int32_t zc_parity_of;    /* value whose parity picks the ZC formula: M0*nsb (utsa) or nsb (rxofdm); 0 = M0*nsb */
    int32_t equalizer;       /* OFDM_EQ_MMSE | OFDM_EQ_ZF */
    int32_t sync_advance;    /* framed mode: samples before the CP end at which the synch window opens */
    float gate;              /* lock when peak > gate * M0*nsb  (SynchAndChanEst.py:166; 0.4 in gr-RXOFDM :170) */
    float inv_reg_chanest;   /* 1/x in H = r*conj(zc)/(1 + 1/x)          (SynchAndChanEst.py:180-181) */
    float inv_reg_eq;        /* 1/x in g = conj(H)/(|H|^2 + 1/x), data   (SynchAndChanEst.py:244-246) */
    float inv_reg_eq_synch;  /* same for the synch-symbol equaliser       (SynchAndChanEst.py:213-218) */
    float llr_scale;         /* 1/noise variance multiplying the max-log LLRs */
    /* Comb pilots inside the data symbols (extension; the reference runs with ref_sigs = 0,
     * GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/SystemModel.py:30).  pilot_spacing = 0 disables them.  Otherwise
     * positions pilot_offset, pilot_offset+pilot_spacing, ... of the used-bin list carry a unit-modulus
     * Zadoff-Chu pilot (root pilot_root); the receiver equalises each data symbol with the LS estimate
     * on its pilots, linearly interpolated across the list, and every per-symbol row holds
     * num_data_bins - n_pilots payload bins. */
    int32_t pilot_spacing;
    int32_t pilot_offset;
    int32_t pilot_root;
    /* LTE "normal CP" pattern (extension; the reference assumes one prefix length, rx_b_len = nfft + cp_len,
     * GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:79).  With cp_first > cp_len and slot_syms >= 1, symbol k of a frame
     * carries a prefix of cp_first samples when k % slot_syms == 0 and of cp_len samples otherwise (TS 36.211 5.6:
     * 160 / 144 samples at nfft 2048, 7 symbols per slot), each symbol is normalised over its own length, and
     * ofdm_plan_frame_len() is the sum of the symbol lengths.  Requires m0 == 1; not combined with comb pilots or the
     * frequency-offset bank.  The delay bank still spans
     * 0..cp_len.  cp_first = 0 (or == cp_len) means a uniform prefix.  The symbol-list entry points
     * (ofdm_rx_sync_search / chanest / demod) take explicit window positions and are unaffected. */
    int32_t cp_first;
    int32_t slot_syms;
    int32_t reserved[3];     /* must be zero */
} ofdm_params;

typedef struct ofdm_plan ofdm_plan;

int ofdm_abi_version(void);
const char *ofdm_last_error(void);

/* Builds the device tables on the CURRENT CUDA device.  Replaces the constructor-time precompute
 * of the blocks: bin maps, Zadoff-Chu, delay bank (SynchAndChanEst.py:36-99). */
int ofdm_plan_create(const ofdm_params *params, ofdm_plan **plan_out);
int ofdm_plan_destroy(ofdm_plan *plan);
int ofdm_plan_hard_stride(const ofdm_plan *plan);          /* bytes of packed bits per data symbol */
int ofdm_plan_frame_len(const ofdm_plan *plan);            /* (M0+D)*(nfft+cp_len) samples (+ the longer prefixes of a CP pattern) */
int ofdm_plan_payload_bins(const ofdm_plan *plan);         /* payload bins per data symbol */

/* TX: bits -> QAM + Zadoff-Chu/resource-element insertion -> IFFT -> CP add -> power normalise.
 * Replaces MultiAntennaSystem.multi_ant_binary_map + multi_ant_symb_gen
 * (GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/MultiAntennaSystem.py:113-219).
 * d_bits: n_frames*D*hard_stride bytes.  d_time: n_frames*frame_len complex64. */
int ofdm_tx_modulate(ofdm_plan *plan, const uint8_t *d_bits, float *d_time, int64_t n_frames, void *stream);

/* RX sliding search: trial t opens its FFT window at sample first_start + t*stride.  For each
 * trial: M0 FFTs, synch-bin pick, unit-power normalise, (CP+1)-delay correlation with the ZC,
 * max / argmax.  Replaces the body of loop A up to the gate test (SynchAndChanEst.py:145-164,
 * synch_and_chan_est.py:148-168).  d_peak[t] = max|c|, d_delay[t] = argmax.
 * With stride == 1, one synch symbol on every bin but DC and the band edge (the gr-utsa_ofdm block's configuration)
 * and no hypothesis bank, the same values come from a time-domain kernel that needs no transform per trial
 * (environment OFDM_SYNC_TD=0 forces the per-trial transforms, 1 the plain time-domain kernel). */
int ofdm_rx_sync_search(ofdm_plan *plan, const float *d_time, int64_t n_samples, int64_t first_start,
                        int64_t stride, int64_t n_trials, float *d_peak, int32_t *d_delay, void *stream);

/* RX LS channel estimate for n_locks accepted locks (window start d_pos[i], delay d_delay[i]).
 * Replaces SynchAndChanEst.py:177-188 (d_H, nfft per lock) and, when the pointers are non-NULL,
 * :202-218 (d_htime = ifft(H), nfft per lock; d_synch_eq = equalised synch symbols, M0*nsb per lock). */
int ofdm_rx_chanest(ofdm_plan *plan, const float *d_time, int64_t n_samples, const int64_t *d_pos,
                    const int32_t *d_delay, int64_t n_locks, float *d_H, float *d_htime,
                    float *d_synch_eq, void *stream);

/* RX data symbols: FFT with the CP skipped by the start pointer, data-bin pick, unit-power
 * normalise, delay de-rotation, one-tap MMSE/ZF, optional max-log LLR + packed hard bits.
 * Symbol i opens at d_start[i], uses delay d_delay[i] and channel row d_hrow[i] of d_H.
 * Windows past n_samples are zero-padded (np.fft.fft semantics).  Replaces loop B
 * (SynchAndChanEst.py:225-248, synch_and_chan_est.py:230-250).
 * d_eq: n_sym*ndb complex64.  d_llr (nullable): n_sym*ndb*bps float.  d_hard (nullable):
 * n_sym*hard_stride bytes. */
int ofdm_rx_demod(ofdm_plan *plan, const float *d_time, int64_t n_samples, const int64_t *d_start,
                  const int32_t *d_delay, const int32_t *d_hrow, const float *d_H, int64_t n_sym,
                  float *d_eq, float *d_llr, uint8_t *d_hard, void *stream);

/* RX framed mode (throughput path): loop A body once per frame at the nominal position plus
 * loop B for the frame's D data symbols, fused in one kernel -- the channel estimate never
 * leaves the SM.  Frame f's synch window opens at first + f*frame_stride + cp_len - sync_advance
 * (cp_first instead of cp_len with a CP pattern).
 * Outputs per frame: d_delay, d_peak, d_locked (peak > gate*M0*nsb); d_H nullable (nfft per frame).
 * Per data symbol (n_frames*D rows): d_eq, d_llr, d_hard -- each nullable, at least one given: a caller that feeds a
 * decoder asks for LLRs (or hard bits) only and the equalised symbols (SynchAndChanEst.py:246, est_data_freq) are
 * then never written. */
int ofdm_rx_frames(ofdm_plan *plan, const float *d_time, int64_t n_samples, int64_t first,
                   int64_t frame_stride, int64_t n_frames, int32_t *d_delay, float *d_peak,
                   uint8_t *d_locked, float *d_H, float *d_eq, float *d_llr, uint8_t *d_hard,
                   void *stream);

/* Frequency-offset hypothesis bank (GRR/LEGACY/gr-ofdm-rx/python/SynchEstAndFO.py:196): hypothesis h
 * multiplies sample n of every FFT window by exp(j*2pi*cycles_per_sample[h]*n), n = 0..nfft-1 (the
 * reference's fo_range[h]/fs).  n_hyp = 0 removes the bank; at most 64 rows; not combined with comb pilots.
 * With a bank installed
 *   - the search evaluates every hypothesis per trial and reports the one with the largest peak (first on
 *     ties) in d_hyp (nullable)                                            (SynchEstAndFO.py:247-277);
 *   - ofdm_rx_chanest_cfo / ofdm_rx_demod_cfo de-rotate each lock / symbol with row d_hyp[i] (NULL = no
 *     rotation)                                                            (:252-256, :331-333);
 *   - the framed receiver picks the hypothesis per frame on the preamble, applies it to the frame's data
 *     windows and reports it in d_hyp (nullable).
 * phase_continuous = 0 restarts the phasor at every window like the reference, which leaves windows L samples
 * apart with a relative phase of 2*pi*nu*L (harmless only for offsets far below the symbol rate).
 * phase_continuous = 1 (extension, oracle-defined) continues it from the buffer origin: a window that opens at
 * absolute sample s is de-rotated by exp(j*2pi*nu*(s+n)) -- what undoing a real carrier offset needs.
 * The entry points without the _cfo suffix are the same calls with d_hyp = NULL. */
int ofdm_plan_set_cfo_bank(ofdm_plan *plan, const double *cycles_per_sample, int32_t n_hyp, int32_t phase_continuous);
int ofdm_rx_sync_search_cfo(ofdm_plan *plan, const float *d_time, int64_t n_samples, int64_t first_start,
                            int64_t stride, int64_t n_trials, float *d_peak, int32_t *d_delay, int32_t *d_hyp,
                            void *stream);
int ofdm_rx_chanest_cfo(ofdm_plan *plan, const float *d_time, int64_t n_samples, const int64_t *d_pos,
                        const int32_t *d_delay, const int32_t *d_hyp, int64_t n_locks, float *d_H, float *d_htime,
                        float *d_synch_eq, void *stream);
int ofdm_rx_demod_cfo(ofdm_plan *plan, const float *d_time, int64_t n_samples, const int64_t *d_start,
                      const int32_t *d_delay, const int32_t *d_hrow, const int32_t *d_hyp, const float *d_H,
                      int64_t n_sym, float *d_eq, float *d_llr, uint8_t *d_hard, void *stream);
int ofdm_rx_frames_cfo(ofdm_plan *plan, const float *d_time, int64_t n_samples, int64_t first,
                       int64_t frame_stride, int64_t n_frames, int32_t *d_delay, float *d_peak,
                       uint8_t *d_locked, int32_t *d_hyp, float *d_H, float *d_eq, float *d_llr, uint8_t *d_hard,
                       void *stream);

/* Legacy QPSK soft/hard demapper with a buffer-wide noise estimate.  Replaces BitRecovery.work
 * (GRR/LEGACY/gr-ofdm-rx/python/BitRecovery.py:77-157).  Two launches: the first reduces
 * sum(dmin) into d_ws[0] (double, zeroed by the call), the second writes d_llr0/d_llr1 (2n float
 * each) and d_hard (2n bytes, one 0/1 per byte like the reference's int array). */
int ofdm_bit_recovery(ofdm_plan *plan, const float *d_in, int64_t n, float *d_llr0, float *d_llr1,
                      uint8_t *d_hard, double *d_ws, void *stream);

/* DSSS despreading of equalised rows (legacy receiver): every group of sf consecutive data bins of a row is
 * multiplied by conj(code) and averaged; bins past the last whole group are ignored.
 * d_eq: n_rows x row_len complex64, d_code: sf complex64, d_out: n_rows x (row_len / sf) complex64.
 * Replaces GRR/LEGACY/gr-ofdm-rx/python/SynchEstFOAndDSSS.py:392-397 (code: :253-261). */
int ofdm_dsss_despread(ofdm_plan *plan, const float *d_eq, int64_t n_rows, int32_t row_len, const float *d_code,
                       int32_t sf, float *d_out, void *stream);

/* BER/EVM counters, ACCUMULATED into d_cnt[8] (double): [0] bit errors, [1] bits compared,
 * [2] sum|eq-ref|^2, [3] sum|ref|^2, [4] frames locked, [5] frames; [6],[7] spare.
 * ref symbols are rebuilt from d_bits_tx.  d_locked nullable.  This vector is the only thing
 * ever reduced across GPUs (NCCL all-reduce on the host side). */
int ofdm_counters(ofdm_plan *plan, const uint8_t *d_bits_rx, const uint8_t *d_bits_tx, const float *d_eq,
                  const uint8_t *d_locked, int64_t n_frames, double *d_cnt, void *stream);

/* Test-bench channel (NOT on the modem path): per-frame n_taps-tap FIR + AWGN, counter-based
 * Philox noise.  d_taps: n_frames*n_taps complex64 (or one set when taps_per_frame == 0).
 * Replaces MultiAntennaSystem.rx_signal_gen / additive_noise (:221-260) for synthetic inputs.
 * Frames are filtered independently (no inter-frame tail). */
int ofdm_channel_awgn(ofdm_plan *plan, const float *d_tx, float *d_rx, int64_t n_frames, const float *d_taps,
                      int32_t n_taps, int32_t taps_per_frame, float noise_sigma, uint64_t seed, void *stream);

/* Host-buffer convenience entry points (what a GNU Radio work() sees): stage through plan-owned
 * device buffers, copies inside the call, chunked and double-buffered on internal streams.
 * They synchronise before returning.  Pinned host memory makes the copies asynchronous.
 * ofdm_rx_frames_host: h_eq, h_llr, h_hard are each nullable (at least one given); an output that is not asked for is
 * neither computed into device memory nor copied back. */
int ofdm_rx_frames_host(ofdm_plan *plan, const float *h_time, int64_t n_samples, int64_t first,
                        int64_t frame_stride, int64_t n_frames, int32_t *h_delay, float *h_peak,
                        uint8_t *h_locked, float *h_eq, float *h_llr, uint8_t *h_hard);
int ofdm_tx_modulate_host(ofdm_plan *plan, const uint8_t *h_bits, float *h_time, int64_t n_frames);

/* One call per work() of the two live receiver blocks (the scheduler boundary,
 * GRR/gr-RXOFDM/python/synch_and_chan_est.py:121-124,:140-142).  A block object owns the state the reference keeps in
 * NumPy arrays between calls -- the lock counter, time_synch_ref, est_chan_freq_P -- plus pinned staging memory and a
 * stream.  ofdm_rx_block_work copies the scheduler's samples in, runs loop A (sliding search, the lock rule and its
 * bookkeeping, channel estimate of the new locks: SynchAndChanEst.py:143-219 / synch_and_chan_est.py:146-221) and
 * loop B (data symbols: :221-248 / :224-250) on the device and synchronises ONCE; the result pointers refer to pinned
 * memory owned by the block and stay valid until its next call.
 *   mode 0 = gr-utsa_ofdm SynchAndChanEst: search stride 1, the first acceptable lock only, every data symbol of the
 *            buffer equalised with channel row 0; `rows` = num_ofdm_symb state rows.
 *   mode 1 = gr-RXOFDM synch_and_chan_est: stride cp_len-1, every acceptable lock (up to 100), the first data symbol
 *            after each lock with that lock's row.
 * The counter reset at the end of work() (:261 / :265) is applied; what the block writes to its output stream (and
 * the "nothing on the first call" rule, :257) stays with the caller, which owns the output buffer -- for HOST buffers;
 * ofdm_rx_block_work_device below writes the stream too. */
typedef struct ofdm_rx_block ofdm_rx_block;
typedef struct {
    int32_t n_new_locks;            /* locks accepted by this call */
    int32_t n_demod;                /* data symbols demodulated by this call */
    int32_t obs_after_search;       /* corr_obs / cor_obs after loop A (before the reset) */
    int32_t n_state_rows;           /* rows of time_synch_ref (1 in mode 0, 100 in mode 1) */
    const int64_t *time_synch_ref;  /* [n_state_rows][3]: window start, delay, int(peak) */
    const int32_t *new_rows;        /* [n_new_locks]: state row each new lock was stored in */
    const float *chan_time;         /* [n_new_locks][nfft] complex64: ifft of the new estimates (:202) */
    const float *synch_freq;        /* [n_new_locks][M0*nsb] complex64: equalised synch symbols (:213-218) */
    const int32_t *demod_rows;      /* [n_demod]: row of est_data_freq each demodulated symbol belongs to */
    const float *data_freq;         /* [n_demod][num_data_bins] complex64 */
} ofdm_rx_block_result;
int ofdm_rx_block_create(ofdm_plan *plan, int32_t mode, int32_t rows, ofdm_rx_block **block_out);
int ofdm_rx_block_destroy(ofdm_rx_block *block);
int ofdm_rx_block_work(ofdm_rx_block *block, const float *h_in, int64_t n_samples, ofdm_rx_block_result *result);
/* The same call for a block whose scheduler buffers already live in DEVICE memory -- what a GNU Radio 3.10 block built on
 * custom (device) buffers is handed by the scheduler instead of host views (the boundary of
 * GRR/gr-RXOFDM/python/synch_and_chan_est.py:121-124,:140-142 and GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:135-137).
 * d_in (n_samples complex64, n_samples >= 1, 8-byte aligned) is read in place: no staging copy, no host-to-device copy.
 * Work queued on `stream` before the call (the upstream block's kernel, the previous reader of the output buffers) is
 * ordered before the block's own stream; the call returns after its one synchronisation, so d_in may be reused and the
 * outputs read on any stream afterwards.
 *   d_rows (nullable, 16-byte aligned, ofdm_rx_block_max_rows() x num_data_bins complex64): receives the equalised data
 *     rows of this call's list, in list order; without it they go to a device buffer of the block's.  Either way
 *     result->data_freq is a DEVICE pointer after this call.  Everything else in `result` is host memory as above.
 *   est_data_freq (SynchAndChanEst.py:93 / synch_and_chan_est.py:97), from which the reference packs its output stream
 *     and whose rows persist from call to call, is kept in device memory by this entry point (rows listed by
 *     ofdm_rx_block_work calls go to the caller's host mirror instead and are not seen by it).
 *   d_stream_out (nullable, 8-byte aligned, out_capacity complex64): the block's output stream, out[0:m] of
 *     SynchAndChanEst.py:249-259 (mode 0: the rows np.delete(est_data_freq, range(3, rows, M0+D)) keeps, laid end to end)
 *     and synch_and_chan_est.py:252-263 (mode 1: est_data_freq[0:corr_size]).  pack_rows = how many packed rows the
 *     stream has: n_data_symb of SynchAndChanEst.py:138 in mode 0, corr_size of synch_and_chan_est.py:253 in mode 1 --
 *     one line of arithmetic on the caller's constructor arguments; it is clamped to the rows that exist.
 *     *n_written (nullable) = m = min(out_capacity, packed rows x num_data_bins).  The reference writes nothing on its
 *     first call (:257 / :261): the caller passes d_stream_out = NULL then.
 * State (lock counter, time_synch_ref, est_chan_freq_P) is shared with ofdm_rx_block_work: the two calls may be mixed. */
int ofdm_rx_block_work_device(ofdm_rx_block *block, const float *d_in, int64_t n_samples, float *d_rows,
                              float *d_stream_out, int64_t pack_rows, int64_t out_capacity, void *stream,
                              ofdm_rx_block_result *result, int64_t *n_written);
/* rows a call can list at most: `rows` in mode 0, 100 in mode 1 (the size of d_rows above and of result->demod_rows) */
int ofdm_rx_block_max_rows(const ofdm_rx_block *block);
/* est_chan_freq_P: copies the first `rows` channel rows (nfft complex64 each) to h_H. */
int ofdm_rx_block_chan_freq(ofdm_rx_block *block, float *h_H, int32_t rows);

/* Introspection for benchmarks: kernels launched by this library since load. */
uint64_t ofdm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OFDM_B200_H */
