// kernels.cuh -- the OFDM modem kernels (templated on log2(NFFT) and bits per bin).
//
// One "symbol group" of T = N/8 threads owns one OFDM symbol at a time; groups are persistent and
// stride over the work list.  All per-bin work (ZC correlation, channel estimate, equaliser gain,
// de-rotation, demapping) is done on the registers the FFT leaves behind, so a received sample is
// read from HBM exactly once and every output is written exactly once.
#pragma once
#include "fft_core.cuh"

namespace ofdm {

struct DevParams {
    int nfft, cp, nsb, ndb, m0, d, bps, mm;
    int sync_advance, hard_stride, dup_bin, equalizer;
    float gate_thresh;        // gate * M0*nsb
    float chanest_scale;      // 1 / ((1 + inv_reg_chanest) * M0)
    float inv_reg_eq, inv_reg_eq_synch, llr_scale;
    const float2 *tw;         // [N]      exp(-2*pi*i*k/N)
    const float2 *zc_bin;     // [M0][N]  sum of conj(zc[m*nsb+idx]) over the synch indices that map to FFT bin k
    const float *w_bin;       // [N]      how many synch indices map to bin k (0, 1, or 2 for the +-N/2 alias)
    const float2 *zc_dup;     // [M0][2]  conj(zc) of the first / last synch index (alias handling)
    const float2 *zc_tx;      // [M0][N]  ZC value the transmitter puts in bin k of synch symbol m (0 if unused)
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// output position of FFT bin k in the reference's bin list  [-n/2..-1, 1..n/2]  (-1: unused bin)
__device__ __forceinline__ int used_index(int k, int N, int n) {
    const int h = n >> 1;
    if (k >= 1 && k <= h) return k - 1 + h;
    if (k >= N - h) return k - (N - h);
    return -1;
}

__device__ __forceinline__ float2 ldg_sample(const float2 *__restrict__ x, long long i, long long n) {
    return (i >= 0 && i < n) ? __ldg(x + i) : make_float2(0.f, 0.f);
}

template <int T, int GROUPS>
__device__ __forceinline__ float3 group_sum3(float3 x, float *red, int j, int gid) {
    constexpr int W = T < 32 ? T : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        x.x += __shfl_xor_sync(0xffffffffu, x.x, o);
        x.y += __shfl_xor_sync(0xffffffffu, x.y, o);
        x.z += __shfl_xor_sync(0xffffffffu, x.z, o);
    }
    if (T > 32) {
        if ((j & 31) == 0) {
            red[3 * (j >> 5) + 0] = x.x;
            red[3 * (j >> 5) + 1] = x.y;
            red[3 * (j >> 5) + 2] = x.z;
        }
        group_sync<T, GROUPS>(gid);
        float3 s = make_float3(0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < T / 32; ++w) {
            s.x += red[3 * w + 0];
            s.y += red[3 * w + 1];
            s.z += red[3 * w + 2];
        }
        x = s;
    }
    return x;
}

// ------------------------------------------------------------------------------------------------
// demapping: exact max-log LLR per bit for Gray PAM per axis (closed forms; the oracle brute-forces
// the same definition).  L > 0 <=> bit 1.  Bit order MSB first: I0 Q0 I1 Q1 I2 Q2.
// ------------------------------------------------------------------------------------------------
template <int BPA> __device__ __forceinline__ void pam_llr(float x, float a, float (&A)[3]) {
    const float t = fabsf(x);
    if (BPA == 1) {
        A[0] = -4.f * a * x;
    } else if (BPA == 2) {
        A[0] = -copysignf(4.f * a * (t + fmaxf(0.f, t - 2.f * a)), x);
        A[1] = 4.f * a * (t - 2.f * a);
    } else {
        A[0] = -copysignf(4.f * a * (t + fmaxf(0.f, t - 2.f * a) + fmaxf(0.f, t - 4.f * a) + fmaxf(0.f, t - 6.f * a)), x);
        A[1] = 4.f * a * ((t - 4.f * a) + fminf(0.f, t - 2.f * a) + fmaxf(0.f, t - 6.f * a));
        A[2] = 4.f * a * fabsf(t - 4.f * a) - 8.f * a * a;
    }
}

template <int BPS> __device__ __forceinline__ void demap(float2 z, float scale, float (&L)[BPS], unsigned &bits) {
    if (BPS == 1) {
        L[0] = 4.f * z.x * scale;                       // levels 2b-1: bit 1 -> +1
        bits = L[0] > 0.f;
        return;
    }
    constexpr int BPA = BPS / 2 > 0 ? BPS / 2 : 1;
    const float a = BPS == 2 ? 0.70710678118654752440f : (BPS == 4 ? 0.31622776601683793320f : 0.15430334996209191026f);
    float AI[3], AQ[3];
    pam_llr<BPA>(z.x, a, AI);
    pam_llr<BPA>(z.y, a, AQ);
    bits = 0;
#pragma unroll
    for (int i = 0; i < BPA; ++i) {
        L[2 * i] = AI[i] * scale;
        L[2 * i + 1] = AQ[i] * scale;
        bits = (bits << 2) | ((AI[i] > 0.f) << 1) | (AQ[i] > 0.f);
    }
}

// bits (MSB-first value of one bin) -> constellation point; inverse of demap's hard decision.
// QPSK points follow MultiAntennaSystem.py:171-178, BPSK :156-157, 16/64-QAM TS 36.211 7.1.
template <int BPS> __device__ __forceinline__ float2 map_bits(unsigned v) {
    if (BPS == 1) return make_float2(v ? 1.f : -1.f, 0.f);
    if (BPS == 2) {
        const float a = 0.70710678118654752440f;
        return make_float2((v & 2) ? -a : a, (v & 1) ? -a : a);
    }
    if (BPS == 4) {
        const float a = 0.31622776601683793320f;
        const float i = ((v & 8) ? -1.f : 1.f) * ((v & 2) ? 3.f : 1.f);
        const float q = ((v & 4) ? -1.f : 1.f) * ((v & 1) ? 3.f : 1.f);
        return make_float2(i * a, q * a);
    }
    const float a = 0.15430334996209191026f;
    // magnitude by (b1,b2): 00->3 01->1 10->5 11->7
    const int mi = ((v >> 3) & 1) * 2 + ((v >> 1) & 1), mq = ((v >> 2) & 1) * 2 + (v & 1);
    const float tab[4] = {3.f, 1.f, 5.f, 7.f};
    const float i = ((v & 32) ? -1.f : 1.f) * tab[mi];
    const float q = ((v & 16) ? -1.f : 1.f) * tab[mq];
    return make_float2(i * a, q * a);
}

// read BPS bits starting at bit offset `bo` of a packed MSB-first byte string
template <int BPS> __device__ __forceinline__ unsigned read_bits(const uint8_t *__restrict__ p, int bo, int nbytes) {
    const int by = bo >> 3, sh = bo & 7;
    unsigned w = (unsigned)__ldg(p + by) << 8;
    if (BPS + sh > 8 && by + 1 < nbytes) w |= (unsigned)__ldg(p + by + 1);
    return (w >> (16 - sh - BPS)) & ((1u << BPS) - 1u);
}

// pack the per-bin values staged in hb[0..ndb) into MSB-first bytes; thread i makes bytes i, i+T, ...
template <int BPS, int T>
__device__ __forceinline__ void pack_bits(const uint8_t *hb, uint8_t *__restrict__ out, int ndb, int nbytes, int j) {
    for (int b = j; b < nbytes; b += T) {
        unsigned acc = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = 8 * b + i;
            const int bin = p / BPS, bit = p - bin * BPS;
            const unsigned v = bin < ndb ? hb[bin] : 0u;
            acc = (acc << 1) | ((v >> (BPS - 1 - bit)) & 1u);
        }
        out[b] = (uint8_t)acc;
    }
}

// shared-memory carve-up per CTA
template <int LOG2N, int GROUPS> struct Smem {
    using C = FftCfg<LOG2N>;
    static constexpr int GSTRIDE = (LOG2N == 6) ? (2 * C::BUF + 8) : 2 * C::BUF;   // complex slots per group
    static constexpr int FFT_BYTES = GROUPS * GSTRIDE * 8;
    static constexpr int RED_BYTES = GROUPS * 64 * 4 + GROUPS * 16 * 8;             // 64 floats + 16 u64 per group
    static constexpr int HB_BYTES = GROUPS * C::N;
    static constexpr int BYTES = FFT_BYTES + RED_BYTES + HB_BYTES;
    __device__ static float2 *buf0(unsigned char *s, int gid) { return reinterpret_cast<float2 *>(s) + gid * GSTRIDE; }
    __device__ static float *redf(unsigned char *s, int gid) {
        return reinterpret_cast<float *>(s + FFT_BYTES + GROUPS * 16 * 8) + gid * 64;
    }
    __device__ static unsigned long long *redu(unsigned char *s, int gid) {
        return reinterpret_cast<unsigned long long *>(s + FFT_BYTES) + gid * 16;
    }
    __device__ static uint8_t *hb(unsigned char *s, int gid) { return s + FFT_BYTES + RED_BYTES + gid * C::N; }
};

// ------------------------------------------------------------------------------------------------
// synchronisation on one preamble: M0 FFTs -> ZC correlation in the frequency domain -> the
// (CP+1)-delay bank evaluated as ONE inverse FFT (c[d] = sum_k y_k e^{+j2pi dk/N} is an N-point
// inverse DFT sampled at d = 0..CP), instead of the reference's dense (CP+1) x MM matmul.
// SynchAndChanEst.py:145-164.  On return ys[m] holds sum_m0 X[k]*conj(zc) for k = j+m*T,
// p_est the unit-power factor; key encodes (max |c|^2, lowest d).
// ------------------------------------------------------------------------------------------------
template <int LOG2N, int GROUPS>
__device__ __forceinline__ void sync_core(const DevParams &P, const float2 *__restrict__ x, long long n_samples,
                                          long long pos, bool valid, const float2 *twr, float2 *b0, float2 *b1,
                                          float *redf, unsigned long long *redu, int j, int gid, int &pp,
                                          float2 (&ys)[8], float2 (&ydup)[2], float &p_est, float &peak, int &delay) {
    using C = FftCfg<LOG2N>;
    constexpr int T = C::T, N = C::N;
    const int L = N + P.cp;
    float pw = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) ys[m] = make_float2(0.f, 0.f);
    ydup[0] = ydup[1] = make_float2(0.f, 0.f);
    for (int s = 0; s < P.m0; ++s) {
        float2 v[8];
        const long long base = pos + (long long)s * L + j;
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = valid ? ldg_sample(x, base + m * T, n_samples) : make_float2(0.f, 0.f);
        fft_group<LOG2N, false, GROUPS>(v, twr, b0, b1, j, gid, pp);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = j + m * T;
            pw = fmaf(__ldg(P.w_bin + k), cabs2(v[m]), pw);
            const float2 z = __ldg(P.zc_bin + s * N + k);
            ys[m] = cadd(ys[m], cmul(v[m], z));
            if (k == P.dup_bin) {
                ydup[0] = cadd(ydup[0], cmul(v[m], __ldg(P.zc_dup + 2 * s)));
                ydup[1] = cadd(ydup[1], cmul(v[m], __ldg(P.zc_dup + 2 * s + 1)));
            }
        }
    }
    pw = group_sum<T, GROUPS>(pw, redf, j, gid);
    p_est = pw > 0.f ? sqrtf((float)P.mm / pw) : 0.f;
    float2 c[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) c[m] = ys[m];
    fft_group<LOG2N, true, GROUPS>(c, twr, b0, b1, j, gid, pp);
    unsigned long long key = 0ull;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int dd = j + m * T;
        if (dd <= P.cp) {
            const unsigned long long kk =
                ((unsigned long long)__float_as_uint(cabs2(c[m])) << 32) | (unsigned)(0x7fffffff - dd);
            key = kk > key ? kk : key;
        }
    }
    key = group_max_u64<T, GROUPS>(key, redu, j, gid);
    delay = 0x7fffffff - (int)(key & 0xffffffffu);
    peak = p_est * sqrtf(__uint_as_float((unsigned)(key >> 32)));
}

// channel estimate + equaliser gain for this thread's 8 bins from the correlation sums.
// H_k = rot_k * p_est * ys_k / ((1+1/x) * M0)        SynchAndChanEst.py:177-188
// G_k = conj(H_k) / (|H_k|^2 + inv_reg) * rot_k      :237-246 (de-rotation folded into the gain)
template <int LOG2N>
__device__ __forceinline__ void chan_core(const DevParams &P, const float2 (&ys)[8], const float2 (&ydup)[2],
                                          float p_est, int delay, float inv_reg, int j, float2 (&H)[8],
                                          float2 (&G)[8]) {
    using C = FftCfg<LOG2N>;
    constexpr int T = C::T, N = C::N;
    const float sc = p_est * P.chanest_scale;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int k = j + m * T;
        const float2 w = __ldg(P.tw + ((delay * k) & (N - 1)));   // exp(-j*2pi*d*k/N); rot = conj(w)
        float2 y = ys[m];
        if (k == P.dup_bin) y = ydup[1];                          // numpy fancy assignment: last index wins
        const float2 h = cscale(cmulc(y, w), sc);
        H[m] = h;
        const float den = cabs2(h) + inv_reg;
        const float inv = den > 0.f ? 1.f / den : 0.f;
        G[m] = cscale(cmulc(make_float2(h.x, -h.y), w), inv);
    }
}

// equalise + demap + store one data symbol held in v[] (FFT output).  SynchAndChanEst.py:232-248.
template <int LOG2N, int BPS, int GROUPS>
__device__ __forceinline__ void data_epilogue(const DevParams &P, const float2 (&v)[8], const float2 (&G)[8],
                                              long long row, bool valid, float *redf, uint8_t *hb, int j, int gid,
                                              float2 *__restrict__ eq_out, float *__restrict__ llr_out,
                                              uint8_t *__restrict__ hard_out) {
    using C = FftCfg<LOG2N>;
    constexpr int T = C::T, N = C::N;
    int jo[8];
    float pw = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        jo[m] = used_index(j + m * T, N, P.ndb);
        pw += jo[m] >= 0 ? cabs2(v[m]) : 0.f;
    }
    pw = group_sum<T, GROUPS>(pw, redf, j, gid);
    const float p0 = pw > 0.f ? sqrtf((float)P.ndb / pw) : 0.f;
    const bool soft = (llr_out != nullptr), hard = (hard_out != nullptr);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (jo[m] >= 0) {
            const float2 z = cscale(cmul(G[m], v[m]), p0);
            const long long o = row * P.ndb + jo[m];
            if (valid) eq_out[o] = z;
            if (soft || hard) {
                float Lb[BPS];
                unsigned bits;
                demap<BPS>(z, P.llr_scale, Lb, bits);
                if (soft && valid) {
                    float *lp = llr_out + o * BPS;
                    if (BPS == 1) {
                        lp[0] = Lb[0];
                    } else if (BPS == 4) {
                        *reinterpret_cast<float4 *>(lp) = make_float4(Lb[0], Lb[1], Lb[2], Lb[3]);
                    } else {
#pragma unroll
                        for (int b = 0; b < BPS; b += 2)
                            *reinterpret_cast<float2 *>(lp + b) = make_float2(Lb[b], Lb[b + 1 < BPS ? b + 1 : b]);
                    }
                }
                if (hard) hb[jo[m]] = (uint8_t)bits;
            }
        }
    }
    if (hard) {
        group_sync<T, GROUPS>(gid);
        if (valid) pack_bits<BPS, T>(hb, hard_out + row * P.hard_stride, P.ndb, P.hard_stride, j);
    }
}

#define OFDM_KERNEL_PROLOGUE(LOG2N, GROUPS)                                            \
    using C = FftCfg<LOG2N>;                                                           \
    using S = Smem<LOG2N, GROUPS>;                                                     \
    constexpr int T = C::T;                                                            \
    extern __shared__ __align__(16) unsigned char smem_raw[];                          \
    const int gid = threadIdx.x / T, j = threadIdx.x % T;                              \
    float2 *b0 = S::buf0(smem_raw, gid), *b1 = b0 + C::BUF;                            \
    float *redf = S::redf(smem_raw, gid);                                              \
    unsigned long long *redu = S::redu(smem_raw, gid);                                 \
    uint8_t *hb = S::hb(smem_raw, gid);                                                \
    (void)redu; (void)hb; (void)redf;                                                  \
    float2 twr[C::NTW > 0 ? C::NTW : 1];                                               \
    load_twiddles<LOG2N>(twr, P.tw, j);                                                \
    const long long total_groups = (long long)gridDim.x * GROUPS;                      \
    const long long gg = (long long)blockIdx.x * GROUPS + gid;                         \
    int pp = 0;

// ------------------------------------------------------------------------------------------------
// K1  framed receiver: sync + channel estimate + D data symbols per frame, fused.
// ------------------------------------------------------------------------------------------------
template <int LOG2N, int BPS, int GROUPS>
__global__ void __launch_bounds__(FftCfg<LOG2N>::T *GROUPS)
rx_frames_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first,
                 long long frame_stride, long long n_frames, int iters, int *__restrict__ delay_out,
                 float *__restrict__ peak_out, uint8_t *__restrict__ locked_out, float2 *__restrict__ H_out,
                 float2 *__restrict__ eq_out, float *__restrict__ llr_out, uint8_t *__restrict__ hard_out) {
    OFDM_KERNEL_PROLOGUE(LOG2N, GROUPS)
    constexpr int N = C::N;
    const int L = N + P.cp;
    for (int it = 0; it < iters; ++it) {
        const long long f = gg + (long long)it * total_groups;
        const bool valid = f < n_frames;
        const long long pos = first + f * frame_stride + P.cp - P.sync_advance;
        float2 ys[8], ydup[2], H[8], G[8];
        float p_est, peak;
        int delay;
        sync_core<LOG2N, GROUPS>(P, x, n_samples, pos, valid, twr, b0, b1, redf, redu, j, gid, pp, ys, ydup, p_est,
                                 peak, delay);
        chan_core<LOG2N>(P, ys, ydup, p_est, delay, P.inv_reg_eq, j, H, G);
        if (valid) {
            if (j == 0) {
                delay_out[f] = delay;
                peak_out[f] = peak;
                locked_out[f] = peak > P.gate_thresh ? 1 : 0;
            }
            if (H_out != nullptr) {
#pragma unroll
                for (int m = 0; m < 8; ++m) H_out[f * N + j + m * T] = H[m];
            }
        }
        const long long dbase = pos + (long long)P.m0 * L + j;
        for (int n = 0; n < P.d; ++n) {
            float2 v[8];
            const long long sb = dbase + (long long)n * L;
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = valid ? ldg_sample(x, sb + m * T, n_samples) : make_float2(0.f, 0.f);
            fft_group<LOG2N, false, GROUPS>(v, twr, b0, b1, j, gid, pp);
            data_epilogue<LOG2N, BPS, GROUPS>(P, v, G, f * P.d + n, valid, redf, hb, j, gid, eq_out, llr_out,
                                              hard_out);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2  sliding search: one trial per group iteration.  Outputs max|c| and argmax per trial.
// ------------------------------------------------------------------------------------------------
template <int LOG2N, int GROUPS>
__global__ void __launch_bounds__(FftCfg<LOG2N>::T *GROUPS)
rx_sync_search_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first_start,
                      long long stride, long long n_trials, int iters, float *__restrict__ peak_out,
                      int *__restrict__ delay_out) {
    OFDM_KERNEL_PROLOGUE(LOG2N, GROUPS)
    for (int it = 0; it < iters; ++it) {
        const long long t = gg + (long long)it * total_groups;
        const bool valid = t < n_trials;
        float2 ys[8], ydup[2];
        float p_est, peak;
        int delay;
        sync_core<LOG2N, GROUPS>(P, x, n_samples, first_start + t * stride, valid, twr, b0, b1, redf, redu, j, gid,
                                 pp, ys, ydup, p_est, peak, delay);
        if (valid && j == 0) {
            peak_out[t] = peak;
            delay_out[t] = delay;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3  channel estimate for accepted locks (+ optional time-domain channel and equalised preamble).
// ------------------------------------------------------------------------------------------------
template <int LOG2N, int GROUPS>
__global__ void __launch_bounds__(FftCfg<LOG2N>::T *GROUPS)
rx_chanest_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples,
                  const long long *__restrict__ pos_in, const int *__restrict__ delay_in, long long n_locks, int iters,
                  float2 *__restrict__ H_out, float2 *__restrict__ htime_out, float2 *__restrict__ synch_eq_out) {
    OFDM_KERNEL_PROLOGUE(LOG2N, GROUPS)
    constexpr int N = C::N;
    const int L = N + P.cp;
    for (int it = 0; it < iters; ++it) {
        const long long i = gg + (long long)it * total_groups;
        const bool valid = i < n_locks;
        const long long pos = valid ? pos_in[i] : 0;
        const int delay = valid ? delay_in[i] : 0;
        float2 ys[8], ydup[2], H[8], G[8];
        float p_est, peak;
        int dfound;
        sync_core<LOG2N, GROUPS>(P, x, n_samples, pos, valid, twr, b0, b1, redf, redu, j, gid, pp, ys, ydup, p_est,
                                 peak, dfound);
        chan_core<LOG2N>(P, ys, ydup, p_est, delay, P.inv_reg_eq_synch, j, H, G);
        if (valid) {
#pragma unroll
            for (int m = 0; m < 8; ++m) H_out[i * N + j + m * T] = H[m];
        }
        if (htime_out != nullptr) {                              // ifft(chan_est1): SynchAndChanEst.py:202
            float2 c[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) c[m] = H[m];
            fft_group<LOG2N, true, GROUPS>(c, twr, b0, b1, j, gid, pp);
            if (valid) {
#pragma unroll
                for (int m = 0; m < 8; ++m) htime_out[i * N + j + m * T] = cscale(c[m], 1.f / (float)N);
            }
        }
        if (synch_eq_out != nullptr) {                           // eq_gain_ext * data_recov: :213-218
            for (int s = 0; s < P.m0; ++s) {
                float2 v[8];
                const long long base = pos + (long long)s * L + j;
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    v[m] = valid ? ldg_sample(x, base + m * T, n_samples) : make_float2(0.f, 0.f);
                fft_group<LOG2N, false, GROUPS>(v, twr, b0, b1, j, gid, pp);
                if (valid) {
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int k = j + m * T;
                        const int idx = used_index(k, N, P.nsb);
                        if (idx >= 0) {
                            // G already carries conj(H)/(|H|^2+1/snr) * rot ; data_recov = rot*p_est*X
                            const float2 z = cscale(cmul(G[m], v[m]), p_est);
                            synch_eq_out[i * P.mm + s * P.nsb + idx] = z;
                            if (k == P.dup_bin) {
                                // the aliased bin also appears as synch index 0, with the gain of index 0
                                const float2 w = __ldg(P.tw + ((delay * k) & (N - 1)));
                                const float2 h0 = cscale(cmulc(ydup[0], w), p_est * P.chanest_scale);
                                const float den = cabs2(h0) + P.inv_reg_eq_synch;
                                const float2 g0 = cscale(cmulc(make_float2(h0.x, -h0.y), w), den > 0.f ? 1.f / den : 0.f);
                                synch_eq_out[i * P.mm + s * P.nsb + 0] = cscale(cmul(g0, v[m]), p_est);
                            }
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4  data symbols from an explicit symbol list (used by the stream blocks' work()).
// ------------------------------------------------------------------------------------------------
template <int LOG2N, int BPS, int GROUPS>
__global__ void __launch_bounds__(FftCfg<LOG2N>::T *GROUPS)
rx_demod_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples,
                const long long *__restrict__ start_in, const int *__restrict__ delay_in,
                const int *__restrict__ hrow_in, const float2 *__restrict__ Hmat, long long n_sym, int iters,
                float2 *__restrict__ eq_out, float *__restrict__ llr_out, uint8_t *__restrict__ hard_out) {
    OFDM_KERNEL_PROLOGUE(LOG2N, GROUPS)
    constexpr int N = C::N;
    for (int it = 0; it < iters; ++it) {
        const long long i = gg + (long long)it * total_groups;
        const bool valid = i < n_sym;
        const long long start = valid ? start_in[i] : 0;
        const int delay = valid ? delay_in[i] : 0;
        const long long hrow = valid ? hrow_in[i] : 0;
        float2 v[8], G[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = j + m * T;
            v[m] = valid ? ldg_sample(x, start + k, n_samples) : make_float2(0.f, 0.f);
            const float2 h = valid ? __ldg(Hmat + hrow * N + k) : make_float2(0.f, 0.f);
            const float2 w = __ldg(P.tw + ((delay * k) & (N - 1)));
            const float den = cabs2(h) + P.inv_reg_eq;
            G[m] = cscale(cmulc(make_float2(h.x, -h.y), w), den > 0.f ? 1.f / den : 0.f);
        }
        fft_group<LOG2N, false, GROUPS>(v, twr, b0, b1, j, gid, pp);
        data_epilogue<LOG2N, BPS, GROUPS>(P, v, G, i, valid, redf, hb, j, gid, eq_out, llr_out, hard_out);
    }
}

// ------------------------------------------------------------------------------------------------
// K5  transmitter: one OFDM symbol per group iteration.
//   bits -> QAM into data bins / ZC into synch bins (MultiAntennaSystem.py:135-183)
//   -> IFFT -> CP -> unit energy per sample -> divide by sqrt(var)   (:189-219)
// Both normalisations are scale factors derived from two sums over the L = N+CP output samples
// (CP samples count twice), so the symbol is written to HBM once, already normalised.
// ------------------------------------------------------------------------------------------------
template <int LOG2N, int BPS, int GROUPS>
__global__ void __launch_bounds__(FftCfg<LOG2N>::T *GROUPS)
tx_kernel(const DevParams P, const uint8_t *__restrict__ bits, float2 *__restrict__ out, long long n_sym_total,
          int iters) {
    OFDM_KERNEL_PROLOGUE(LOG2N, GROUPS)
    constexpr int N = C::N;
    const int L = N + P.cp, fs = P.m0 + P.d;
    for (int it = 0; it < iters; ++it) {
        const long long si = gg + (long long)it * total_groups;
        const bool valid = si < n_sym_total;
        const long long f = si / fs;
        const int s = (int)(si - f * fs);
        float2 v[8];
        if (!valid) {
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = make_float2(0.f, 0.f);
        } else if (s < P.m0) {
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = __ldg(P.zc_tx + s * N + j + m * T);
        } else {
            const uint8_t *bp = bits + (f * P.d + (s - P.m0)) * P.hard_stride;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int jo = used_index(j + m * T, N, P.ndb);
                v[m] = jo >= 0 ? map_bits<BPS>(read_bits<BPS>(bp, jo * BPS, P.hard_stride)) : make_float2(0.f, 0.f);
            }
        }
        fft_group<LOG2N, true, GROUPS>(v, twr, b0, b1, j, gid, pp);
        float3 acc = make_float3(0.f, 0.f, 0.f);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float w = (j + m * T >= N - P.cp) ? 2.f : 1.f;
            acc.x = fmaf(w, cabs2(v[m]), acc.x);
            acc.y = fmaf(w, v[m].x, acc.y);
            acc.z = fmaf(w, v[m].y, acc.z);
        }
        acc = group_sum3<T, GROUPS>(acc, redf, j, gid);
        // v holds N * ifft(X).  Reference: t/N -> *sqrt(L/E) -> /sqrt(var)
        const float invN = 1.f / (float)N;
        const float E = acc.x * invN * invN;
        const float s1 = E > 1e-30f ? sqrtf((float)L / E) * invN : invN;
        const float mx = s1 * acc.y / (float)L, my = s1 * acc.z / (float)L;
        const float var = s1 * s1 * acc.x / (float)L - (mx * mx + my * my);
        const float g = var > 0.f ? s1 * rsqrtf(var) : 0.f;
        if (valid) {
            float2 *o = out + si * L;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int n = j + m * T;
                const float2 z = cscale(v[m], g);
                o[P.cp + n] = z;
                if (n >= N - P.cp) o[n - (N - P.cp)] = z;
            }
        }
    }
}

}  // namespace ofdm
