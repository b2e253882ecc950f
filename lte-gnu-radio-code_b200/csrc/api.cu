// api.cu -- the C ABI of libofdm_b200.so (include/ofdm_b200.h), the plan, and the small
// non-FFT kernels (legacy demapper, BER/EVM counters, test-bench channel).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "plan.cuh"

namespace ofdm {

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";
int fail(int code, const char *fmt, const char *a, long long b) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
int cuda_fail(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return OFDM_ECUDA;
}

// ------------------------------------------------------------------------------------------------
// misc kernels
// ------------------------------------------------------------------------------------------------
// Legacy QPSK demapper, phase 1: sum over the buffer of the distance to the nearest point.
// BitRecovery.py:86-88,:102
__global__ void bitrec_sum_kernel(const float2 *__restrict__ x, long long n, double *__restrict__ ws) {
    const float a = 0.70710678118654752440f;
    float acc = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 z = __ldg(x + i);
        const float ex = z.x - (z.x >= 0.f ? a : -a), ey = z.y - (z.y >= 0.f ? a : -a);
        acc += sqrtf(ex * ex + ey * ey);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)part[w];
        atomicAdd(ws, s);
    }
}

// phase 2: llrp0 / llrp1 / hard bit per axis.  BitRecovery.py:102-125,:155-156
__global__ void bitrec_llr_kernel(const float2 *__restrict__ x, long long n, const double *__restrict__ ws,
                                  float *__restrict__ llr0, float *__restrict__ llr1, uint8_t *__restrict__ hard) {
    const float a = 0.70710678118654752440f, K = 1.414213562373095f;
    const double sigma = 0.7071067811865476 * (ws[0] / (double)n);
    const float hd = (float)(-0.5 / (sigma * sigma));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 z = __ldg(x + i);
        const float er = fabsf(z.x - (z.x >= 0.f ? a : -a)), ei = fabsf(z.y - (z.y >= 0.f ? a : -a));
        // quadrant tests in the reference's first-match order reduce to these two predicates
        const bool re_pos = (z.x > 0.f) || (z.x == 0.f && z.y >= 0.f);
        const bool im_pos = (z.y >= 0.f);
        const float sr = hd * er, orr = hd * (K - er), si = hd * ei, oi = hd * (K - ei);
        const float l0r = re_pos ? sr : orr, l1r = re_pos ? orr : sr;
        const float l0i = im_pos ? si : oi, l1i = im_pos ? oi : si;
        reinterpret_cast<float2 *>(llr0)[i] = make_float2(l0r, l0i);
        reinterpret_cast<float2 *>(llr1)[i] = make_float2(l1r, l1i);
        uchar2 h;
        h.x = (l1r - l0r) > 0.f;
        h.y = (l1i - l0i) > 0.f;
        reinterpret_cast<uchar2 *>(hard)[i] = h;
    }
}

template <int BPS>
__device__ __forceinline__ void cnt_symbol(const DevParams &P, const uint8_t *tx, const float2 *eq, int ndb, float &e2,
                                           float &r2, int lane0, int step) {
    for (int b = lane0; b < ndb; b += step) {
        const float2 r = map_bits<BPS>(read_bits<BPS>(tx, b * BPS, P.hard_stride));
        const float2 z = __ldg(eq + b);
        e2 += (z.x - r.x) * (z.x - r.x) + (z.y - r.y) * (z.y - r.y);
        r2 += r.x * r.x + r.y * r.y;
    }
}

// BER / EVM counters: one warp per data symbol.
__global__ void counters_kernel(const DevParams P, const uint8_t *__restrict__ rx, const uint8_t *__restrict__ tx,
                                const float2 *__restrict__ eq, const uint8_t *__restrict__ locked, long long n_frames,
                                double *__restrict__ cnt) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const long long n_sym = n_frames * P.d;
    float e2 = 0.f, r2 = 0.f;
    unsigned long long errs = 0;
    for (long long s = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); s < n_sym; s += (long long)gridDim.x * wpb) {
        const uint8_t *t = tx + s * P.hard_stride;
        if (rx != nullptr) {
            const uint8_t *r = rx + s * P.hard_stride;
            for (int b = lane; b < P.hard_stride; b += 32) errs += __popc((unsigned)(__ldg(r + b) ^ __ldg(t + b)));
        }
        if (eq != nullptr) {
            const float2 *q = eq + s * P.ndd;
            switch (P.bps) {
                case 1: cnt_symbol<1>(P, t, q, P.ndd, e2, r2, lane, 32); break;
                case 2: cnt_symbol<2>(P, t, q, P.ndd, e2, r2, lane, 32); break;
                case 4: cnt_symbol<4>(P, t, q, P.ndd, e2, r2, lane, 32); break;
                default: cnt_symbol<6>(P, t, q, P.ndd, e2, r2, lane, 32); break;
            }
        }
    }
    double de = (double)e2, dr = (double)r2, dq = (double)errs;
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_xor_sync(0xffffffffu, de, o);
        dr += __shfl_xor_sync(0xffffffffu, dr, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    __shared__ double part[3][32];
    if (lane == 0) {
        part[0][threadIdx.x >> 5] = dq;
        part[1][threadIdx.x >> 5] = de;
        part[2][threadIdx.x >> 5] = dr;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int w = 0; w < wpb; ++w) {
            a += part[0][w];
            b += part[1][w];
            c += part[2][w];
        }
        if (rx != nullptr) atomicAdd(cnt + 0, a);
        if (eq != nullptr) {
            atomicAdd(cnt + 2, b);
            atomicAdd(cnt + 3, c);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // frames / locked frames / bits compared: counted once by the first warp
        double lk = 0;
        if (locked != nullptr)
            for (long long f = lane; f < n_frames; f += 32) lk += locked[f] ? 1.0 : 0.0;
        for (int o = 16; o > 0; o >>= 1) lk += __shfl_xor_sync(0xffffffffu, lk, o);
        if (lane == 0) {
            if (rx != nullptr) atomicAdd(cnt + 1, (double)n_sym * (double)(P.ndd * P.bps));
            atomicAdd(cnt + 4, lk);
            atomicAdd(cnt + 5, (double)n_frames);
        }
    }
}

// DSSS despreading of equalised rows: out[r][g] = mean_i eq[r][g*sf + i] * conj(code[i]).
// SynchEstFOAndDSSS.py:392-397.  One thread per output; a warp's reads sweep one contiguous stretch of the row.
__global__ void dsss_despread_kernel(const float2 *__restrict__ eq, long long n_rows, int row_len, int n_groups,
                                     const float2 *__restrict__ code, int sf, float2 *__restrict__ out) {
    const long long total = n_rows * n_groups;
    const float inv = 1.f / (float)sf;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n_groups;
        const int g = (int)(i - r * n_groups);
        const float2 *src = eq + r * row_len + (long long)g * sf;
        float2 acc = make_float2(0.f, 0.f);
        for (int c = 0; c < sf; ++c) acc = cadd(acc, cmulc(__ldg(src + c), __ldg(code + c)));
        out[i] = cscale(acc, inv);
    }
}

// Philox4x32-10 counter-based generator (Salmon et al., SC'11), one call = 4 x 32 random bits.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// test-bench channel: y[f,n] = sum_t h[f,t] x[f,n-t] + sigma*(g1 + j g2), frames independent
// (MultiAntennaSystem.py:221-231 rx_signal_gen + :257-260 additive_noise; oracle: channel_awgn_frames.  Sample i of
// the batch draws its noise from Philox counter (i, 0, 0) under key `seed`: u1, u2 from the first two words, Box-Muller.)
__global__ void channel_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, long long n_frames, int frame_len,
                               const float2 *__restrict__ taps, int n_taps, int taps_per_frame, float sigma,
                               unsigned long long seed) {
    const long long total = n_frames * frame_len;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / frame_len;
        const int n = (int)(i - f * frame_len);
        const float2 *h = taps + (taps_per_frame ? f * n_taps : 0);
        float2 acc = make_float2(0.f, 0.f);
        for (int t = 0; t < n_taps && t <= n; ++t) acc = cadd(acc, cmul(__ldg(h + t), __ldg(x + i - t)));
        if (sigma > 0.f) {
            const uint4 r = philox4x32_10(make_uint4((unsigned)i, (unsigned)(i >> 32), 0u, 0u),
                                          make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
            const float u1 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f;
            const float u2 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
            // full-precision logf / sincospif (not the fast intrinsics): the oracle restates this stream in double
            // precision (oracle/ofdm_oracle.py: philox_awgn) and the parity test holds the two to 1e-5 of sigma
            const float rad = sigma * sqrtf(-2.f * logf(u1));
            float sn, cs;
            sincospif(2.f * u2, &sn, &cs);
            acc.x += rad * cs;
            acc.y += rad * sn;
        }
        y[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Acquisition without a transform per trial (stride-1 sliding search, one synch symbol, synch bins = all but DC and
// the band edge -- the gr-utsa_ofdm block's configuration).  Same values as rx_sync_search_kernel up to rounding:
//   c[d; p] / p_est = r(p+d) + sum_{i=1..d} (x[p+d-i] - x[p+d-i+N]) h[N-i],   r(q) = sum_m x[q+m] h[m],
//   h[m] = sum_{k in S} conj(zc_k) e^{-j2pi km/N}      (oracle/ofdm_oracle.py: sync_search_time_domain has the derivation)
// so every position q needs one N-tap correlation and a (CP+1)-long PREFIX sum, and the unit-power factor follows from
// sum_{k in S} |X_k|^2 = N sum_n |x[p+n]|^2 - |X_0|^2 - |X_{N/2}|^2, all three accumulated in the same pass over the
// window.  A CTA takes TP = 128*R - CP consecutive trials: thread t owns positions q0 + t + 128 j, j < R (lane-contiguous
// shared-memory reads, the tap h[m] is a broadcast), and the delay maximum of trial p = q - i is kept in shared memory
// with a packed (|c|^2, -d) 64-bit atomicMax, which resolves ties to the lowest delay like np.argmax.
// ------------------------------------------------------------------------------------------------
constexpr int TD_R = 8, TD_THREADS = 128, TD_TQ = TD_R * TD_THREADS;
#ifndef OFDM_SYNC_TD_DEFAULT
#define OFDM_SYNC_TD_DEFAULT 3
#endif

__global__ void __launch_bounds__(TD_THREADS)
rx_sync_search_td_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first_start,
                         long long n_trials, float *__restrict__ peak_out, int *__restrict__ delay_out) {
    extern __shared__ __align__(16) unsigned char td_smem[];
    const int N = P.nfft, CP = P.cp, t = threadIdx.x;
    const int TP = TD_TQ - CP;                                          // trials per CTA
    const int XW = TD_TQ + N + CP;                                      // samples staged: q0 - CP .. q0 + TQ + N - 1
    float2 *X = reinterpret_cast<float2 *>(td_smem);
    float2 *H = X + XW;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(H + N);
    float *pw = reinterpret_cast<float *>(best + TD_TQ);
    const long long tile0 = (long long)blockIdx.x * TP;                 // first trial of the tile
    const long long q0 = first_start + tile0;                           // its window start = first position of the tile
    const long long last_trial = n_trials - 1;
    for (int i = t; i < XW; i += TD_THREADS) {
        const long long g = q0 - CP + i;
        X[i] = (g >= 0 && g < n_samples) ? __ldg(x + g) : make_float2(0.f, 0.f);
    }
    for (int i = t; i < N; i += TD_THREADS) H[i] = __ldg(P.h_td + i);
    for (int i = t; i < TD_TQ; i += TD_THREADS) best[i] = 0ull;
    __syncthreads();
    float2 acc[TD_R], s0[TD_R], s1[TD_R];
    float en[TD_R];
#pragma unroll
    for (int j = 0; j < TD_R; ++j) {
        acc[j] = s0[j] = s1[j] = make_float2(0.f, 0.f);
        en[j] = 0.f;
    }
    const float2 *xb = X + CP + t;                                      // position q0 + t + 128 j reads xb[128 j + m]
    for (int m = 0; m < N; m += 2) {
        const float2 h0 = H[m], h1 = H[m + 1];
#pragma unroll
        for (int j = 0; j < TD_R; ++j) {
            const float2 a = xb[TD_THREADS * j + m], b = xb[TD_THREADS * j + m + 1];
            acc[j].x = fmaf(a.x, h0.x, fmaf(-a.y, h0.y, acc[j].x));
            acc[j].y = fmaf(a.x, h0.y, fmaf(a.y, h0.x, acc[j].y));
            acc[j].x = fmaf(b.x, h1.x, fmaf(-b.y, h1.y, acc[j].x));
            acc[j].y = fmaf(b.x, h1.y, fmaf(b.y, h1.x, acc[j].y));
            en[j] = fmaf(a.x, a.x, fmaf(a.y, a.y, en[j]));
            en[j] = fmaf(b.x, b.x, fmaf(b.y, b.y, en[j]));
            s0[j].x += a.x + b.x;                                       // X_0
            s0[j].y += a.y + b.y;
            s1[j].x += a.x - b.x;                                       // X_{N/2}: alternating signs
            s1[j].y += a.y - b.y;
        }
    }
#pragma unroll
    for (int j = 0; j < TD_R; ++j)                                      // power in the synch bins of the window at q
        pw[t + TD_THREADS * j] = (float)N * en[j] - cabs2(s0[j]) - cabs2(s1[j]);
    auto offer = [&](int j, int i) {                                    // c[i; q - i] for position j of this thread
        const int pl = t + TD_THREADS * j - i;                          // trial index inside the tile
        if (pl >= 0 && pl < TP && tile0 + pl <= last_trial) {
            const unsigned long long key =
                ((unsigned long long)__float_as_uint(cabs2(acc[j])) << 32) | (unsigned)(0x7fffffff - i);
            atomicMax(best + pl, key);
        }
    };
#pragma unroll
    for (int j = 0; j < TD_R; ++j) offer(j, 0);
    for (int i = 1; i <= CP; ++i) {
        const float2 hh = H[N - i];
#pragma unroll
        for (int j = 0; j < TD_R; ++j) {
            const float2 a = xb[TD_THREADS * j - i], b = xb[TD_THREADS * j - i + N];
            const float2 dl = make_float2(a.x - b.x, a.y - b.y);
            acc[j].x = fmaf(dl.x, hh.x, fmaf(-dl.y, hh.y, acc[j].x));
            acc[j].y = fmaf(dl.x, hh.y, fmaf(dl.y, hh.x, acc[j].y));
            offer(j, i);
        }
    }
    __syncthreads();
    for (int pl = t; pl < TP; pl += TD_THREADS) {
        const long long tr = tile0 + pl;
        if (tr > last_trial) break;
        const unsigned long long key = best[pl];
        const float p = pw[pl];
        const float p_est = p > 0.f ? sqrtf((float)P.mm / p) : 0.f;
        peak_out[tr] = p_est * sqrtf(__uint_as_float((unsigned)(key >> 32)));
        delay_out[tr] = 0x7fffffff - (int)(key & 0xffffffffu);
    }
}

// Register-blocked form of the same search: a thread owns EIGHT CONSECUTIVE positions q0 + 8t + j and slides an
// eight-sample register window over its part of the capture, so one shared-memory load feeds eight complex MACs
// instead of one.  The staged samples are stored transposed, Xs[(i & 7) * S + (i >> 3)], which makes lane t's load of
// sample 8t + c hit consecutive words; the per-trial maxima and window powers use the same index map.  The window
// sums of positions 1..7 follow from position 0's by sliding (E(q+1) = E(q) - |x[q]|^2 + |x[q+N]|^2, likewise X_0, and
// X_{N/2} with a sign flip).  N is a multiple of 8.  (The 64-bit shared-memory atomicMax compiles to a
// compare-and-swap loop; two native 32-bit passes -- maximum, then lowest delay attaining it, with the delay
// recursion repeated -- were measured slower: N=64 0.67 -> 1.14 ms, N=2048 2.86 -> 3.60 ms.)
__global__ void __launch_bounds__(TD_THREADS)
rx_sync_search_td8_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first_start,
                          long long n_trials, float *__restrict__ peak_out, int *__restrict__ delay_out) {
    extern __shared__ __align__(16) unsigned char td_smem[];
    const int N = P.nfft, CP = P.cp, t = threadIdx.x;
    const int A = (CP + 7) & ~7;                                        // staged samples before position q0
    const int TP = TD_TQ - CP;                                          // trials per CTA
    const int XW = TD_TQ + N + A;                                       // a multiple of 8
    const int S = XW / 8 + 1, SB = TD_TQ / 8 + 1;
    float2 *Xs = reinterpret_cast<float2 *>(td_smem);
    float2 *H = Xs + 8 * S;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(H + N);
    float *pw = reinterpret_cast<float *>(best + 8 * SB);
    const long long tile0 = (long long)blockIdx.x * TP;
    const long long q0 = first_start + tile0;
    const long long last_trial = n_trials - 1;
    for (int i = t; i < XW; i += TD_THREADS) {
        const long long g = q0 - A + i;
        Xs[(i & 7) * S + (i >> 3)] = (g >= 0 && g < n_samples) ? __ldg(x + g) : make_float2(0.f, 0.f);
    }
    for (int i = t; i < N; i += TD_THREADS) H[i] = __ldg(P.h_td + i);
    for (int i = t; i < 8 * SB; i += TD_THREADS) best[i] = 0ull;
    __syncthreads();
    // sample base + c of this thread, base = A + 8t (a multiple of 8): row c & 7, column A/8 + t + (c >> 3)
    const float2 *xt = Xs + A / 8 + t;
    auto LD = [&](int c) -> float2 { return xt[(c & 7) * S + (c >> 3)]; };      // c >> 3 floors, also for c < 0
    float2 acc[8], W[8];
    float en = 0.f;
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] = make_float2(0.f, 0.f);
        W[j] = LD(j);
    }
    for (int M = 0; M < N; M += 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {                                   // tap m = M + r; position j reads slot (r + j) & 7
            const float2 h = H[M + r];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[j] = cmul_acc(acc[j], W[(r + j) & 7], h);          // two packed FMAs per complex multiply-add
            }
            const float2 a0 = W[r];                                     // position 0's own sample n = m
            en = fmaf(a0.x, a0.x, fmaf(a0.y, a0.y, en));
            s0.x += a0.x;
            s0.y += a0.y;
            s1.x += (r & 1) ? -a0.x : a0.x;
            s1.y += (r & 1) ? -a0.y : a0.y;
            W[r] = LD(M + r + 8);                                       // slot of sample m is free: sample m + 8
        }
    }
    // W[j] = x[q_0 + N + j].  Window powers of the eight positions by sliding, then the delay-0 offers.
    float2 U[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) U[j] = LD(j);
    auto put_pw = [&](int j) {
        const int pq = 8 * t + j;
        pw[(pq & 7) * SB + (pq >> 3)] = (float)N * en - cabs2(s0) - cabs2(s1);
    };
    // Offers: position j at delay i belongs to trial 8t + c, c = j - i.  A trial offset c receives this thread's offers
    // at the eight consecutive delays i = -c .. 7-c and never again, so the running maximum lives in a circular
    // register file K[c & 7] and goes to shared memory with ONE atomic when the offset is complete (the 64-bit
    // shared-memory atomicMax is a compare-and-swap loop: eight times fewer of them).
    unsigned long long K[8];
    auto key_of = [&](int j, int i) -> unsigned long long {
        return ((unsigned long long)__float_as_uint(cabs2(acc[j])) << 32) | (unsigned)(0x7fffffff - i);
    };
    auto flush = [&](unsigned long long key, int c) {                   // trial offset c of this thread is complete
        const int pl = 8 * t + c;
        if (pl >= 0 && pl < TP && tile0 + pl <= last_trial) atomicMax(best + (pl & 7) * SB + (pl >> 3), key);
    };
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        put_pw(j);
        K[j] = key_of(j, 0);                                            // offset c = j starts at delay 0
        const float2 a = U[j], b = W[j];                                // leave x[q_j], enter x[q_j + N]
        en += cabs2(b) - cabs2(a);
        s0.x += b.x - a.x;
        s0.y += b.y - a.y;
        s1.x = -(s1.x - a.x + b.x);
        s1.y = -(s1.y - a.y + b.y);
    }
    flush(K[7], 7);                                                     // offset 7 only ever sees delay 0
    // delays 1..CP: position j needs x[q_j - i] - x[q_j - i + N]; slot (j - i) & 7 of the windows U, W and of K
    for (int I = 0; I < CP; I += 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = I + r + 1;
            if (i <= CP) {
                const float2 hh = H[N - i];
                U[7 - r] = LD(-i);                                      // sample -i replaces sample 8 - i
                W[7 - r] = LD(N - i);
                K[7 - r] = 0ull;                                        // offset -i starts (its slot was flushed one step ago)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 a = U[(j - r - 1) & 7], b = W[(j - r - 1) & 7];
                    acc[j] = cmul_acc(acc[j], csub(a, b), hh);
                    const unsigned long long key = key_of(j, i);
                    K[(j - r - 1) & 7] = key > K[(j - r - 1) & 7] ? key : K[(j - r - 1) & 7];
                }
                flush(K[(6 - r) & 7], 7 - i);                           // offset 7 - i got its last offer (from j = 7)
            }
        }
    }
    // the seven offsets still open: c = -CP .. 6-CP, slot c & 7
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = ((u - (-CP & 7)) & 7) - CP;                       // the offset in [-CP, 7-CP] that lives in slot u
        if (c <= 6 - CP) flush(K[u], c);
    }
    __syncthreads();
    for (int pl = t; pl < TP; pl += TD_THREADS) {
        const long long tr = tile0 + pl;
        if (tr > last_trial) break;
        const unsigned long long key = best[(pl & 7) * SB + (pl >> 3)];
        const float p = pw[(pl & 7) * SB + (pl >> 3)];
        const float p_est = p > 0.f ? sqrtf((float)P.mm / p) : 0.f;
        peak_out[tr] = p_est * sqrtf(__uint_as_float((unsigned)(key >> 32)));
        delay_out[tr] = 0x7fffffff - (int)(key & 0xffffffffu);
    }
}


// General form of the register-blocked search: M0 synch symbols (template, 1 or 2) and / or the frequency-offset
// hypothesis bank (SynchEstAndFO.py:247-277), still without a transform per trial.  Derivation and CPU restatement:
// oracle/ofdm_oracle.py sync_search_time_domain(..., nus=...).  With u = e^{j2pi nu} (row h of the bank is u^n):
//   c_h[d; p] / p_est = u^d sum_s phi_s [ r_s(q_s) + sum_{i=1..d} ( x[q_s-i] Hp_s[i] - x[q_s-i+N] Hm_s[N-i] ) ],   q_s = p + sL + d,
//   r_s(q) = sum_m x[q+m] Hm_s[m],   Hm_s[m] = phi_s h_s[m] u^m,   Hp_s[i] = phi_s h_s[N-i] u^{-i},
// |u^d| = 1 drops out of the magnitude; phi_s = 1 for the reference's window-local phasor and u^{sL} when the phasor
// is continued from the buffer origin (the common factor u^p drops out as well).  The power of the rotated window in
// the synch bins is N sum|x|^2 - |sum_n x[p+n] u^n|^2 - |sum_n x[p+n] (-u)^n|^2 per synch symbol, the two sums slid from
// position to position as X(q+1) = (X(q) - x[q]) / u' + x[q+N] u'^{N-1}, u' = +-u.  Hypotheses are evaluated one after
// the other (their taps are rebuilt in shared memory from h_s and the bank row); the first maximum of the peak wins,
// like np.argmax over the hypotheses.  Needs L = N + CP to be a multiple of 8 (the transposed staging layout).
template <int M0>
__global__ void __launch_bounds__(TD_THREADS)
rx_sync_search_tdg_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first_start,
                          long long n_trials, float *__restrict__ peak_out, int *__restrict__ delay_out,
                          int *__restrict__ hyp_out) {
    extern __shared__ __align__(16) unsigned char td_smem[];
    const int N = P.nfft, CP = P.cp, L = N + CP, t = threadIdx.x;
    const int A = (CP + 7) & ~7;
    const int TP = TD_TQ - CP;
    const int XW = TD_TQ + N + A + (M0 - 1) * L;                        // a multiple of 8
    const int S = XW / 8 + 1, SB = TD_TQ / 8 + 1, CPA = CP + 1;
    float2 *Xs = reinterpret_cast<float2 *>(td_smem);
    float2 *Hm = Xs + 8 * S;                                            // [M0][N]
    float2 *Hp = Hm + M0 * N;                                           // [M0][CPA], entry 0 unused
    float2 *Er = Hp + M0 * CPA;                                         // [N] u^n
    unsigned long long *best = reinterpret_cast<unsigned long long *>(Er + N);
    float *pw = reinterpret_cast<float *>(best + 8 * SB);
    float2 *phi = reinterpret_cast<float2 *>(pw + 8 * SB);              // [M0]
    const long long tile0 = (long long)blockIdx.x * TP;
    const long long q0 = first_start + tile0;
    const long long last_trial = n_trials - 1;
    for (int i = t; i < XW; i += TD_THREADS) {
        const long long g = q0 - A + i;
        Xs[(i & 7) * S + (i >> 3)] = (g >= 0 && g < n_samples) ? __ldg(x + g) : make_float2(0.f, 0.f);
    }
    const int n_hyp = P.n_cfo > 0 ? P.n_cfo : 1;
    float bpk[8];
    int bdl[8], bhy[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        bpk[u] = -1.f;
        bdl[u] = bhy[u] = 0;
    }
    for (int h = 0; h < n_hyp; ++h) {
        __syncthreads();                                                // previous hypothesis is done with the tables
        const float2 *row = P.n_cfo > 0 ? P.cfo + (long long)h * N : nullptr;
        if (t < M0) {
            float2 f = make_float2(1.f, 0.f);
            if (P.n_cfo > 0 && P.cfo_continuous) {
                double a = P.cfo_nu[h] * (double)(t * L);
                a -= floor(a);
                double sn, cs;
                sincospi(2.0 * a, &sn, &cs);
                f = make_float2((float)cs, (float)sn);
            }
            phi[t] = f;
        }
        for (int i = t; i < 8 * SB; i += TD_THREADS) best[i] = 0ull;
        __syncthreads();
        for (int i = t; i < M0 * N; i += TD_THREADS) {
            const int m = i & (N - 1);
            float2 v = cmul(__ldg(P.h_td + i), phi[i / N]);
            if (row) v = cmul(v, __ldg(row + m));
            Hm[i] = v;
        }
        for (int i = t; i < M0 * CPA; i += TD_THREADS) {
            const int s = i / CPA, d = i - s * CPA;
            float2 v = make_float2(0.f, 0.f);
            if (d > 0) {
                v = cmul(__ldg(P.h_td + s * N + N - d), phi[s]);
                if (row) v = cmulc(v, __ldg(row + d));                   // u^{-d}
            }
            Hp[i] = v;
        }
        for (int i = t; i < N; i += TD_THREADS) Er[i] = row ? __ldg(row + i) : make_float2(1.f, 0.f);
        __syncthreads();
        const float2 uinv = row ? make_float2(Er[1].x, -Er[1].y) : make_float2(1.f, 0.f);
        const float2 un1 = row ? Er[N - 1] : make_float2(1.f, 0.f);
        const float2 *xt = Xs + A / 8 + t;
        float2 acc[8], W[M0][8], U[M0][8], s0[M0], s1[M0];
        float en = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < M0; ++s) {
            const float2 *xs = xt + s * (L / 8);
            auto LD = [&](int c) -> float2 { return xs[(c & 7) * S + (c >> 3)]; };
            const float2 *hm = Hm + s * N;
            s0[s] = s1[s] = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) W[s][j] = LD(j);
            for (int M = 0; M < N; M += 8) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 hh = hm[M + r];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = cmul_acc(acc[j], W[s][(r + j) & 7], hh);
                    const float2 a0 = W[s][r];
                    en = fmaf(a0.x, a0.x, fmaf(a0.y, a0.y, en));
                    const float2 ae = row ? cmul(a0, Er[M + r]) : a0;
                    s0[s] = cadd(s0[s], ae);
                    s1[s] = (r & 1) ? csub(s1[s], ae) : cadd(s1[s], ae);
                    W[s][r] = LD(M + r + 8);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) U[s][j] = LD(j);
        }
        unsigned long long K[8];
        auto key_of = [&](int j, int i) -> unsigned long long {
            return ((unsigned long long)__float_as_uint(cabs2(acc[j])) << 32) | (unsigned)(0x7fffffff - i);
        };
        auto flush = [&](unsigned long long key, int c) {
            const int pl = 8 * t + c;
            if (pl >= 0 && pl < TP && tile0 + pl <= last_trial) atomicMax(best + (pl & 7) * SB + (pl >> 3), key);
        };
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float p = (float)N * en;
#pragma unroll
            for (int s = 0; s < M0; ++s) p -= cabs2(s0[s]) + cabs2(s1[s]);
            const int pq = 8 * t + j;
            pw[(pq & 7) * SB + (pq >> 3)] = p;
            K[j] = key_of(j, 0);
#pragma unroll
            for (int s = 0; s < M0; ++s) {                              // slide the window sums to the next position
                const float2 a = U[s][j], b = W[s][j];
                en += cabs2(b) - cabs2(a);
                const float2 bu = cmul(b, un1);
                s0[s] = cadd(cmul(csub(s0[s], a), uinv), bu);
                const float2 m1 = cadd(cmul(csub(s1[s], a), uinv), bu);
                s1[s] = make_float2(-m1.x, -m1.y);
            }
        }
        flush(K[7], 7);
        for (int I = 0; I < CP; I += 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = I + r + 1;
                if (i <= CP) {
                    K[7 - r] = 0ull;
#pragma unroll
                    for (int s = 0; s < M0; ++s) {
                        const float2 *xs = xt + s * (L / 8);
                        auto LD = [&](int c) -> float2 { return xs[(c & 7) * S + (c >> 3)]; };
                        const float2 hp = Hp[s * CPA + i], hm = Hm[s * N + N - i];
                        const float2 nhm = make_float2(-hm.x, -hm.y);
                        U[s][7 - r] = LD(-i);
                        W[s][7 - r] = LD(N - i);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            acc[j] = cmul_acc(cmul_acc(acc[j], U[s][(j - r - 1) & 7], hp), W[s][(j - r - 1) & 7], nhm);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const unsigned long long key = key_of(j, i);
                        K[(j - r - 1) & 7] = key > K[(j - r - 1) & 7] ? key : K[(j - r - 1) & 7];
                    }
                    flush(K[(6 - r) & 7], 7 - i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = ((u - (-CP & 7)) & 7) - CP;
            if (c <= 6 - CP) flush(K[u], c);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int pl = t + TD_THREADS * u;
            if (pl < TP && tile0 + pl <= last_trial) {
                const unsigned long long key = best[(pl & 7) * SB + (pl >> 3)];
                const float p = pw[(pl & 7) * SB + (pl >> 3)];
                const float p_est = p > 0.f ? sqrtf((float)P.mm / p) : 0.f;
                const float pk = p_est * sqrtf(__uint_as_float((unsigned)(key >> 32)));
                if (pk > bpk[u]) {                                      // first maximum over the hypotheses
                    bpk[u] = pk;
                    bdl[u] = 0x7fffffff - (int)(key & 0xffffffffu);
                    bhy[u] = h;
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int pl = t + TD_THREADS * u;
        if (pl < TP && tile0 + pl <= last_trial) {
            peak_out[tile0 + pl] = bpk[u];
            delay_out[tile0 + pl] = bdl[u];
            if (hyp_out != nullptr) hyp_out[tile0 + pl] = bhy[u];
        }
    }
}


// Overlap-save form of the stride-1 search (one synch symbol, no hypothesis bank): the N-tap sliding correlation r(q) is
// a FIR filter over the capture, so a block of NB >= 2N samples is filtered in the frequency domain -- forward transform,
// multiply by the taps' spectrum G[k] = (1/NB) sum_m h[m] e^{+2 pi i k m / NB}, inverse transform -- and the NB - N + 1
// outputs that did not wrap are r(q0 .. q0 + NB - N): about 2 NB log2 NB operations per block instead of N per
// position (N = 2048: 2 x 4096-point transforms per 2048 positions).  Everything after r(q) is the register-blocked
// kernel's second half unchanged: the CP+1 delays as a prefix sum per position, the circular register file of running
// maxima, one shared-memory atomic per completed trial offset.  The window sums behind the unit-power factor (energy,
// DC bin, band-edge bin) come from an exclusive prefix over 8-sample chunks of the staged stretch -- a window starting at
// a multiple of 8 is a whole number of chunks -- and slide from there.  One CTA = one block; grid = tiles of
// PB - CP trials.  Specification: oracle/ofdm_oracle.py sync_search_time_domain (same numbers up to rounding).
template <int LOG2NB>
__global__ void __launch_bounds__(CfgOs<LOG2NB>::T)
rx_sync_search_os_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first_start,
                         long long n_trials, float *__restrict__ peak_out, int *__restrict__ delay_out,
                         const float2 *__restrict__ gos, int PB, int alias_all) {
    using C = CfgOs<LOG2NB>;
    constexpr int GROUPS = 1;
    OFDM_KERNEL_PROLOGUE_(GROUPS, false, false, false, false, 0)
    (void)total_groups; (void)gg;
    const int N = P.nfft, CP = P.cp, t = tid;
    const int A = (CP + 7) & ~7;
    const int TP = PB - CP;                                             // trials per tile
    const int XW = A + PB + N;                                          // staged samples q0 - A .. (a multiple of 8)
    const int SX = XW / 8 + 1, SB = PB / 8 + 1, NC = XW / 8;
    float2 *Xs = reinterpret_cast<float2 *>(smem_raw + S::BYTES);
    float2 *Hh = Xs + 8 * SX;                                           // Hh[i] = h[N - i], i = 1..CP
    // r(q) and the per-trial maxima live in the exchange buffers once the transforms are done; the window powers and the
    // chunk prefixes as well where they fit (alias_all: NB = 4096, which brings the CTA under half an SM's shared memory)
    float2 *Rs = b0;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(Rs + 8 * SB);
    float *pw = alias_all ? reinterpret_cast<float *>(best + 8 * SB) : reinterpret_cast<float *>(Hh + ((CP + 2) & ~1));
    float *cs = pw + 8 * SB;                                            // chunk prefixes: 5 arrays of NC + 1
    const long long tile0 = (long long)blockIdx.x * TP;
    const long long q0 = first_start + tile0;
    const long long last_trial = n_trials - 1;
    for (int i = t; i < XW; i += T) {
        const long long g = q0 - A + i;
        Xs[(i & 7) * SX + (i >> 3)] = (g >= 0 && g < n_samples) ? __ldg(x + g) : make_float2(0.f, 0.f);
    }
    for (int i = t; i <= CP; i += T) Hh[i] = i ? __ldg(P.h_td + N - i) : make_float2(0.f, 0.f);
    {
        float2 v[PPT];
        load_pts<T, PPT>(v, x, q0, n_samples, j);
        fft_group<C, false, GROUPS>(v, tws, b0, b1, j, gid, pp);
#pragma unroll
        for (int m = 0; m < PPT; ++m) v[m] = cmul(v[m], __ldg(gos + j + m * T));
        fft_group<C, true, GROUPS>(v, tws, b0, b1, j, gid, pp);
        __syncthreads();                                                // every read of the last exchange is done
#pragma unroll
        for (int m = 0; m < PPT; ++m) {
            const int n = j + m * T;
            if (n < PB) Rs[(n & 7) * SB + (n >> 3)] = v[m];
        }
    }
    for (int i = t; i < 8 * SB; i += T) best[i] = 0ull;
    // chunk sums: chunk c = staged samples 8c .. 8c+7 (sign of the band-edge sum alternates from the chunk's first sample,
    // which has an even staged index, like every window start this kernel looks at)
    for (int c = t; c < NC; c += T) {
        float e = 0.f;
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 a = Xs[r * SX + c];
            e = fmaf(a.x, a.x, fmaf(a.y, a.y, e));
            s0.x += a.x;
            s0.y += a.y;
            s1.x += (r & 1) ? -a.x : a.x;
            s1.y += (r & 1) ? -a.y : a.y;
        }
        cs[c] = e;
        cs[(NC + 1) + c] = s0.x;
        cs[2 * (NC + 1) + c] = s0.y;
        cs[3 * (NC + 1) + c] = s1.x;
        cs[4 * (NC + 1) + c] = s1.y;
    }
    __syncthreads();
    if (t < 32) {                                                       // exclusive prefix over the chunks, one warp
        const int per = (NC + 1 + 31) / 32, lo = t * per, hi = min(lo + per, NC + 1);
#pragma unroll 1
        for (int a = 0; a < 5; ++a) {
            float *arr = cs + a * (NC + 1);
            float run = 0.f;
            for (int c = lo; c < hi; ++c) {                             // local exclusive prefix, run = lane total
                const float val = c < NC ? arr[c] : 0.f;
                arr[c] = run;
                run += val;
            }
            float incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float up = __shfl_up_sync(0xffffffffu, incl, o);
                if (t >= o) incl += up;
            }
            const float off = incl - run;                               // total of the lanes before this one
            for (int c = lo; c < hi; ++c) arr[c] += off;
        }
    }
    __syncthreads();
    if (8 * t < PB) {
        const float2 *xt = Xs + A / 8 + t;
        auto LD = [&](int c) -> float2 { return xt[(c & 7) * SX + (c >> 3)]; };
        float2 acc[8], W[8], U[8];
        const int c0 = A / 8 + t, c1 = c0 + N / 8;
        float en = cs[c1] - cs[c0];
        float2 s0 = make_float2(cs[(NC + 1) + c1] - cs[(NC + 1) + c0], cs[2 * (NC + 1) + c1] - cs[2 * (NC + 1) + c0]);
        float2 s1 = make_float2(cs[3 * (NC + 1) + c1] - cs[3 * (NC + 1) + c0], cs[4 * (NC + 1) + c1] - cs[4 * (NC + 1) + c0]);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            acc[jj] = Rs[jj * SB + t];
            U[jj] = LD(jj);
            W[jj] = LD(N + jj);
        }
        auto put_pw = [&](int jj) {
            pw[jj * SB + t] = (float)N * en - cabs2(s0) - cabs2(s1);
        };
        unsigned long long K[8];
        auto key_of = [&](int jj, int i) -> unsigned long long {
            return ((unsigned long long)__float_as_uint(cabs2(acc[jj])) << 32) | (unsigned)(0x7fffffff - i);
        };
        auto flush = [&](unsigned long long key, int c) {
            const int pl = 8 * t + c;
            if (pl >= 0 && pl < TP && tile0 + pl <= last_trial) atomicMax(best + (pl & 7) * SB + (pl >> 3), key);
        };
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            put_pw(jj);
            K[jj] = key_of(jj, 0);
            const float2 a = U[jj], b = W[jj];
            en += cabs2(b) - cabs2(a);
            s0.x += b.x - a.x;
            s0.y += b.y - a.y;
            s1.x = -(s1.x - a.x + b.x);
            s1.y = -(s1.y - a.y + b.y);
        }
        flush(K[7], 7);
        for (int I = 0; I < CP; I += 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = I + r + 1;
                if (i <= CP) {
                    const float2 hh = Hh[i];
                    U[7 - r] = LD(-i);
                    W[7 - r] = LD(N - i);
                    K[7 - r] = 0ull;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        acc[jj] = cmul_acc(acc[jj], csub(U[(jj - r - 1) & 7], W[(jj - r - 1) & 7]), hh);
                        const unsigned long long key = key_of(jj, i);
                        K[(jj - r - 1) & 7] = key > K[(jj - r - 1) & 7] ? key : K[(jj - r - 1) & 7];
                    }
                    flush(K[(6 - r) & 7], 7 - i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = ((u - (-CP & 7)) & 7) - CP;
            if (c <= 6 - CP) flush(K[u], c);
        }
    }
    __syncthreads();
    for (int pl = t; pl < TP; pl += T) {
        const long long tr = tile0 + pl;
        if (tr > last_trial) break;
        const unsigned long long key = best[(pl & 7) * SB + (pl >> 3)];
        const float pv = pw[(pl & 7) * SB + (pl >> 3)];
        const float p_est = pv > 0.f ? sqrtf((float)P.mm / pv) : 0.f;
        peak_out[tr] = p_est * sqrtf(__uint_as_float((unsigned)(key >> 32)));
        delay_out[tr] = 0x7fffffff - (int)(key & 0xffffffffu);
    }
}

}  // namespace ofdm

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
using namespace ofdm;

// Round 2: with the butterflies on the packed FP32 pipe and the power-table twiddles, the per-lane kernel with 16 points
// per thread overtook the bulk-row kernel on the only grid where the rows fit beside the transform (LTE 20 MHz, 1200 of
// 2048 bins, 64-QAM: 0.643 ms per-lane vs 0.677 ms bulk rows at prefetch depth 2, 0.672 at depth 1, same box; round 1,
// scalar butterflies: 0.760 vs 0.708).  For 16-QAM / QPSK the rows never won (LTE grid 16-QAM 0.524 -> 0.610 ms, N=4096
// 0.403 -> 0.406 ms).  The variant stays selectable (OFDM_BULK_MIN_N / OFDM_BULK16_MIN_N / OFDM_BULK4_MIN_N) and
// covered by the parity and stress tests, but no plan asks for it by default.
#ifndef OFDM_BULK_MIN_N_DEFAULT
#define OFDM_BULK_MIN_N_DEFAULT 8192
#endif
#ifndef OFDM_BULK16_MIN_N_DEFAULT
#define OFDM_BULK16_MIN_N_DEFAULT 8192
#endif
#ifndef OFDM_BULK4_MIN_N_DEFAULT
#define OFDM_BULK4_MIN_N_DEFAULT 8192
#endif

static const KernelSet *kset_for(int log2n) {
    switch (log2n) {
        case 6: return &kset_6;
        case 7: return &kset_7;
        case 8: return &kset_8;
        case 9: return &kset_9;
        case 10: return &kset_10;
        case 11: return &kset_11;
        case 12: return &kset_12;
    }
    return nullptr;
}

extern "C" {

int ofdm_abi_version(void) { return OFDM_B200_ABI_VERSION; }
const char *ofdm_last_error(void) { return g_err; }
uint64_t ofdm_launch_count(void) { return g_launches.load(); }

int ofdm_plan_create(const ofdm_params *p, ofdm_plan **out) {
    if (!p || !out) return fail(OFDM_EINVAL, "null argument%s");
    *out = nullptr;
    int log2n = 0;
    while ((1 << log2n) < p->nfft) ++log2n;
    if ((1 << log2n) != p->nfft || log2n < 6 || log2n > 12) return fail(OFDM_EINVAL, "nfft must be a power of two in 64..4096%s");
    const int N = p->nfft;
    if (p->cp_len < 0 || p->cp_len >= N) return fail(OFDM_EINVAL, "cp_len out of range%s");
    if (p->num_synch_bins < 2 || p->num_synch_bins > N || (p->num_synch_bins & 1))
        return fail(OFDM_EINVAL, "num_synch_bins must be even and <= nfft%s");
    if (p->num_data_bins < 2 || p->num_data_bins >= N || (p->num_data_bins & 1))
        return fail(OFDM_EINVAL, "num_data_bins must be even and < nfft%s");
    if (p->m0 < 1 || p->m0 > 16 || p->d < 0) return fail(OFDM_EINVAL, "synch_dat out of range%s");
    if (p->bps != 1 && p->bps != 2 && p->bps != 4 && p->bps != 6) return fail(OFDM_EINVAL, "bps must be 1, 2, 4 or 6%s");
    if (p->sync_advance < 0 || p->sync_advance > p->cp_len) return fail(OFDM_EINVAL, "sync_advance must be in 0..cp_len%s");
    for (int i = 0; i < 3; ++i)
        if (p->reserved[i] != 0) return fail(OFDM_EINVAL, "reserved fields must be zero%s");
    const bool cp_pattern = p->cp_first != 0 && p->cp_first != p->cp_len;
    if (cp_pattern && (p->cp_first < p->cp_len || p->cp_first >= N || p->slot_syms < 1 || p->m0 != 1 || p->pilot_spacing))
        return fail(OFDM_EINVAL, "CP pattern needs cp_len < cp_first < nfft, slot_syms >= 1, m0 == 1 and no comb pilots%s");
    if (p->pilot_spacing < 0 || p->pilot_spacing == 1 || p->pilot_spacing > p->num_data_bins / 2)
        return fail(OFDM_EINVAL, "pilot_spacing must be 0 or in 2..num_data_bins/2%s");
    if (p->pilot_spacing && (p->pilot_offset < 0 || p->pilot_offset >= p->pilot_spacing))
        return fail(OFDM_EINVAL, "pilot_offset must be in 0..pilot_spacing-1%s");

    ofdm_plan *pl = new (std::nothrow) ofdm_plan();
    if (!pl) return fail(OFDM_ENOMEM, "out of host memory%s");
    memset(pl, 0, sizeof(*pl));
    pl->prm = *p;
    pl->log2n = log2n;
    pl->ks = kset_for(log2n);
    cudaError_t e = cudaGetDevice(&pl->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->sms, cudaDevAttrMultiProcessorCount, pl->device);
    if (e != cudaSuccess) {
        delete pl;
        return cuda_fail(e, "ofdm_plan_create(device query)");
    }

    const int nsb = p->num_synch_bins, m0 = p->m0, mm = m0 * nsb;
    // host tables in double precision, rounded once to fp32
    std::vector<float2> tw(N), zc_bin((size_t)m0 * N, make_float2(0.f, 0.f)), zc_dup((size_t)m0 * 2), zc_tx((size_t)m0 * N, make_float2(0.f, 0.f));
    std::vector<float> w_bin(N, 0.f);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < N; ++k) {
        const double a = -two_pi * (double)k / (double)N;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    // Zadoff-Chu: SynchAndChanEst.py:54-59 / synch_and_chan_est.py:54-64 / SynchSignal.py:21-30
    std::vector<double> zr(mm), zi(mm);
    const int par = p->zc_parity_of > 0 ? p->zc_parity_of : mm;
    for (int k = 0; k < mm; ++k) {
        const double kk = (double)k;
        const double q = (par % 2 == 0) ? kk * kk / 2.0 : kk * (kk + 1.0) / 2.0;
        const double a = -(two_pi / (double)mm) * (double)p->zc_root * q;
        zr[k] = cos(a);
        zi[k] = sin(a);
    }
    // TX side keys the parity on MM (SynchSignal.py:24); only differs from `par` for gr-RXOFDM with even M0, odd nsb.
    // Synch symbol m of a frame carries chunk m of the length-MM sequence.  (The reference never advances
    // synch_state, MultiAntennaSystem.py:143-146, so with M0 > 1 its transmitter repeats chunk 0 while its
    // receiver correlates against chunk m; M0 = 1, the only pinned case, is unaffected.  DESIGN.md D14.)
    std::vector<double> tzr(mm), tzi(mm);
    for (int k = 0; k < mm; ++k) {
        const double kk = (double)k;
        const double q = (mm % 2 == 0) ? kk * kk / 2.0 : kk * (kk + 1.0) / 2.0;
        const double a = -(two_pi / (double)mm) * (double)p->zc_root * q;
        tzr[k] = cos(a);
        tzi[k] = sin(a);
    }
    int dup_bin = -1;
    const int h = nsb / 2;
    for (int idx = 0; idx < nsb; ++idx) {
        const int sb = idx < h ? idx - h : idx - h + 1;              // signed bin
        const int k = ((sb % N) + N) % N;
        if (w_bin[k] > 0.f) dup_bin = k;
        w_bin[k] += 1.f;
        for (int s = 0; s < m0; ++s) {
            zc_bin[(size_t)s * N + k].x += (float)zr[s * nsb + idx];
            zc_bin[(size_t)s * N + k].y += (float)(-zi[s * nsb + idx]);
        }
        for (int s = 0; s < m0; ++s)   // fancy assignment: last index wins
            zc_tx[(size_t)s * N + k] = make_float2((float)tzr[s * nsb + idx], (float)tzi[s * nsb + idx]);
    }
    for (int s = 0; s < m0; ++s) {
        zc_dup[2 * s + 0] = make_float2((float)zr[s * nsb + 0], (float)(-zi[s * nsb + 0]));
        zc_dup[2 * s + 1] = make_float2((float)zr[s * nsb + nsb - 1], (float)(-zi[s * nsb + nsb - 1]));
    }

    // comb pilots (extension): unit-modulus Zadoff-Chu of length n_pil placed on used-bin positions off + i*K
    const int ndb = p->num_data_bins, pk = p->pilot_spacing, poff = p->pilot_offset;
    const int n_pil = pk ? (ndb - poff + pk - 1) / pk : 0;
    std::vector<float2> pilot_tx(N, make_float2(0.f, 0.f));
    for (int i = 0; i < n_pil; ++i) {
        const int jo = poff + i * pk, hh = ndb / 2;
        const int sb = jo < hh ? jo - hh : jo - hh + 1;
        const int k = ((sb % N) + N) % N;
        const double ii = (double)i;
        const double q = (n_pil % 2 == 0) ? ii * ii / 2.0 : ii * (ii + 1.0) / 2.0;
        const double a = -(two_pi / (double)n_pil) * (double)p->pilot_root * q;
        pilot_tx[k] = make_float2((float)cos(a), (float)sin(a));
    }
    // taps of the time-domain acquisition kernel (rx_sync_search_td_kernel): h[m] = sum_idx conj(zc[idx]) e^{-j2pi k m/N},
    // only for the configuration it serves (one synch symbol on every bin but DC and the band edge)
    const bool td_ok = (m0 <= 2 && nsb == N - 2);
    std::vector<float2> h_td((size_t)m0 * N, make_float2(0.f, 0.f));
    std::vector<double> hd_r(N, 0.0), hd_i(N, 0.0);
    if (td_ok) {
        std::vector<double> cs(N), sn(N);
        for (int k = 0; k < N; ++k) {
            cs[k] = cos(two_pi * (double)k / (double)N);
            sn[k] = sin(two_pi * (double)k / (double)N);
        }
        for (int s = 0; s < m0; ++s)
            for (int m = 0; m < N; ++m) {
                double ar = 0.0, ai = 0.0;
                for (int idx = 0; idx < nsb; ++idx) {
                    const int sbin = idx < h ? idx - h : idx - h + 1;
                    const int k = ((sbin % N) + N) % N;
                    const int e2 = (int)(((long long)k * m) % N);            // e^{-j2pi km/N} = cs - j sn
                    ar += zr[s * nsb + idx] * cs[e2] - zi[s * nsb + idx] * sn[e2];     // conj(zc) = zr - j zi
                    ai += -zr[s * nsb + idx] * sn[e2] - zi[s * nsb + idx] * cs[e2];
                }
                h_td[(size_t)s * N + m] = make_float2((float)ar, (float)ai);
                if (s == 0) {
                    hd_r[m] = ar;
                    hd_i[m] = ai;
                }
            }
    }
    // overlap-save search (one synch symbol, N <= 2048): block of NB = max(512, 4N) <= 4096 samples, its forward twiddles
    // and G[k] = (1/NB) sum_m h[m] e^{+2 pi i k m / NB}
    int os_log2nb = 0;
    std::vector<float2> tw_os, gos;
    // (B200, stride-1 search, ms per launch, N-tap kernel -> overlap-save: N=64 0.385 -> 0.502, N=256 0.703 -> 0.600, N=1024
    // 1.67 -> 1.13, N=2048 2.05 -> 0.68: from N=256 on.  Its rounding error is relative to the largest sample of the BLOCK, not
    // of the window -- 1e-6 of the strongest peak -- so trials on noise next to a burst agree with the N-tap kernel to 1e-3 only.)
    if (td_ok && m0 == 1 && N >= 256 && N <= 2048) {
        int NB = 4 * N < 512 ? 512 : 4 * N;
        if (NB > 4096) NB = 4096;
        while ((1 << os_log2nb) < NB) ++os_log2nb;
        tw_os.resize(NB);
        gos.resize(NB);
        std::vector<double> cb(NB), sb2(NB);
        for (int k = 0; k < NB; ++k) {
            cb[k] = cos(two_pi * (double)k / (double)NB);
            sb2[k] = sin(two_pi * (double)k / (double)NB);
            tw_os[k] = make_float2((float)cb[k], (float)(-sb2[k]));
        }
        for (int k = 0; k < NB; ++k) {
            double gr = 0.0, gi = 0.0;
            for (int m = 0; m < N; ++m) {
                const int e2 = (int)(((long long)k * m) % NB);                // e^{+j2pi km/NB} = cb + j sb2
                gr += hd_r[m] * cb[e2] - hd_i[m] * sb2[e2];
                gi += hd_r[m] * sb2[e2] + hd_i[m] * cb[e2];
            }
            gos[k] = make_float2((float)(gr / (double)NB), (float)(gi / (double)NB));
        }
    }
    const size_t b_tw = sizeof(float2) * N, b_zb = sizeof(float2) * m0 * N, b_w = sizeof(float) * N,
                 b_zd = sizeof(float2) * m0 * 2, b_zt = sizeof(float2) * m0 * N, b_pt = sizeof(float2) * N,
                 b_ht = sizeof(float2) * m0 * N, b_os = sizeof(float2) * tw_os.size();
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t total = al(b_tw) + al(b_zb) + al(b_w) + al(b_zd) + al(b_zt) + al(b_pt) + al(b_ht) + 2 * al(b_os);
    e = cudaMalloc(&pl->d_tables, total);
    if (e != cudaSuccess) {
        delete pl;
        return cuda_fail(e, "ofdm_plan_create(cudaMalloc)");
    }
    char *base = (char *)pl->d_tables;
    size_t o = 0;
    auto put = [&](const void *src, size_t n) -> void * {
        void *dst = base + o;
        if (e == cudaSuccess) e = cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
        o += al(n);
        return dst;
    };
    DevParams &D = pl->dev;
    D.tw = (const float2 *)put(tw.data(), b_tw);
    D.zc_bin = (const float2 *)put(zc_bin.data(), b_zb);
    D.w_bin = (const float *)put(w_bin.data(), b_w);
    D.zc_dup = (const float2 *)put(zc_dup.data(), b_zd);
    D.zc_tx = (const float2 *)put(zc_tx.data(), b_zt);
    D.pilot_tx = (const float2 *)put(pilot_tx.data(), b_pt);
    D.h_td = (const float2 *)put(h_td.data(), b_ht);
    if (!td_ok) D.h_td = nullptr;
    pl->os_log2nb = os_log2nb;
    if (os_log2nb) {
        pl->d_tw_os = (const float2 *)put(tw_os.data(), b_os);
        pl->d_gos = (const float2 *)put(gos.data(), b_os);
    }
    if (e != cudaSuccess) {
        cudaFree(pl->d_tables);
        delete pl;
        return cuda_fail(e, "ofdm_plan_create(cudaMemcpy)");
    }
    D.nfft = N;
    D.cp = p->cp_len;
    D.cp_extra = cp_pattern ? p->cp_first - p->cp_len : 0;
    D.slot = cp_pattern ? p->slot_syms : 0x7fffffff;           // a slot that never ends: the uniform prefix
    D.nsb = nsb;
    D.ndb = p->num_data_bins;
    D.m0 = m0;
    D.d = p->d;
    D.bps = p->bps;
    D.mm = mm;
    D.sync_advance = p->sync_advance;
    D.pil_k = pk;
    D.pil_off = poff;
    D.n_pil = n_pil;
    D.ndd = ndb - n_pil;
    D.hard_stride = (D.ndd * p->bps + 7) / 8;
    // 64-QAM LLR rows are staged in shared memory (kernels.cuh, data_epilogue); group stride = row length
    // rounded so that neighbouring groups start 16 banks apart.  Only for transforms up to
    // OFDM_LLR_STAGE_MAX_N points (environment override for A/B runs; 0 disables): from N = 128 on the extra
    // shared-memory traffic costs more than the partial-sector stores it removes (B200, 64-QAM: N=64 0.53 -> 0.39 ms, N=128 0.37 -> 0.42 ms, N=2048 0.76 -> 0.81 ms).
    // (Staging the equalised-symbol rows and the LLR rows of the other modulations as well was tried: 7-12 % slower
    // at every size, and even the dormant run-time switch cost the fast path 2-7 %, so it is not in the code.)
    {
        const char *env = getenv("OFDM_LLR_STAGE_MAX_N");
        const int max_n = env ? atoi(env) : 64;
        D.ls_stride = (p->bps == 6 && pk == 0 && N <= max_n) ? (((ndb * 6 + 31) & ~31) + 16) : 0;
        D.es_stride = 0;
        D.bulk_rows = 0;
        // Bulk-store variant of the framed receiver (kernels.cuh, data_epilogue_bulk): rows of eq and LLR leave shared
        // memory as cp.async.bulk copies.  Used for 64-QAM (whose per-lane LLR stores write every sector three times)
        // from OFDM_BULK_MIN_N points on, when the rows fit next to the transform without costing a resident CTA
        // (checked at launch).  B200, LTE 20 MHz numerology: 0.760 -> 0.708 ms.  At N <= 512 the per-row copies are too
        // small to pay for the extra barrier (N=64: 0.45 -> 0.49 ms), and 16-QAM / QPSK rows are already stored with
        // full sectors (N=2048, 1920 bins, 16-QAM: 0.353 -> 0.367 ms), so only the 64-QAM kernels have this variant.
        // Environment override for A/B runs: OFDM_BULK_MIN_N (a value above 4096 disables it).
        // Below 64-QAM the per-lane stores already write full sectors, but the registers a store reads stay reserved
        // until the memory system has taken the data, and the largest transforms (one CTA of 222-register threads per
        // SM) stall on exactly that (ncu, N=4096 16-QAM: 26 % long-scoreboard on the demapper's register reuse).
        // OFDM_BULK16_MIN_N / OFDM_BULK4_MIN_N: smallest N whose 16-QAM / QPSK-BPSK rows leave through the async proxy.
        const char *eb = getenv("OFDM_BULK_MIN_N"), *eb16 = getenv("OFDM_BULK16_MIN_N"), *eb4 = getenv("OFDM_BULK4_MIN_N");
        const int bulk_min = p->bps == 6 ? (eb ? atoi(eb) : OFDM_BULK_MIN_N_DEFAULT)
                                         : (p->bps == 4 ? (eb16 ? atoi(eb16) : OFDM_BULK16_MIN_N_DEFAULT)
                                                        : (eb4 ? atoi(eb4) : OFDM_BULK4_MIN_N_DEFAULT));
        if (pk == 0 && N >= bulk_min && ((ndb * p->bps) & 3) == 0 && (ndb & 1) == 0) {
            D.bulk_rows = 1;
            D.ls_stride = ((ndb * p->bps + 31) & ~31) + 16;
            D.es_stride = ((ndb + 15) & ~15) + 8;
        }
    }
    D.dup_bin = dup_bin;
    D.equalizer = p->equalizer;
    D.gate_thresh = p->gate * (float)mm;
    D.chanest_scale = (float)(1.0 / ((1.0 + (double)p->inv_reg_chanest) * (double)m0));
    D.inv_reg_eq = p->equalizer == OFDM_EQ_ZF ? 0.f : p->inv_reg_eq;
    D.inv_reg_eq_synch = p->inv_reg_eq_synch;
    D.llr_scale = p->llr_scale;
    *out = pl;
    return OFDM_OK;
}

int ofdm_plan_destroy(ofdm_plan *pl) {
    DevGuard dev_guard(pl);
    if (!pl) return OFDM_OK;
    cudaFree(pl->d_tables);
    if (pl->d_cfo) cudaFree(pl->d_cfo);
    if (pl->small) cudaFree(pl->small);
    for (int i = 0; i < ofdm_plan::NSLOT; ++i) {
        if (pl->stage[i]) cudaFree(pl->stage[i]);
        if (pl->streams_made) {
            cudaStreamDestroy(pl->st[i]);
            cudaEventDestroy(pl->ev[i]);
        }
    }
    delete pl;
    return OFDM_OK;
}

int ofdm_plan_hard_stride(const ofdm_plan *pl) { return pl ? pl->dev.hard_stride : OFDM_EINVAL; }
int ofdm_plan_payload_bins(const ofdm_plan *pl) { return pl ? pl->dev.ndd : OFDM_EINVAL; }
int ofdm_plan_frame_len(const ofdm_plan *pl) {
    if (!pl) return OFDM_EINVAL;
    const int fs = pl->prm.m0 + pl->prm.d;
    int n = fs * (pl->prm.nfft + pl->prm.cp_len);
    if (pl->dev.cp_extra) n += pl->dev.cp_extra * ((fs + pl->dev.slot - 1) / pl->dev.slot);   // one longer prefix per slot
    return n;
}

int ofdm_tx_modulate(ofdm_plan *pl, const uint8_t *d_bits, float *d_time, int64_t n_frames, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_frames == 0) return OFDM_OK;
    if (!pl || !d_time || (!d_bits && pl->prm.d > 0) || n_frames < 0) return fail(OFDM_EINVAL, "ofdm_tx_modulate: bad argument%s");
    CU(pl->ks->tx(pl->dev, d_bits, (float2 *)d_time, n_frames * (pl->prm.m0 + pl->prm.d), (cudaStream_t)stream, pl->sms),
       "ofdm_tx_modulate");
    return OFDM_OK;
}

int ofdm_plan_set_cfo_bank(ofdm_plan *pl, const double *cycles_per_sample, int32_t n_hyp, int32_t phase_continuous) {
    DevGuard dev_guard(pl);
    if (!pl || n_hyp < 0 || n_hyp > 64 || (n_hyp > 0 && !cycles_per_sample))
        return fail(OFDM_EINVAL, "ofdm_plan_set_cfo_bank: bad argument (0..64 hypotheses)%s");
    if (n_hyp > 0 && pl->prm.pilot_spacing) return fail(OFDM_EINVAL, "ofdm_plan_set_cfo_bank: not available with comb pilots%s");
    if (n_hyp > 0 && pl->dev.cp_extra) return fail(OFDM_EINVAL, "ofdm_plan_set_cfo_bank: not available with a CP pattern%s");
    if (pl->d_cfo) cudaFree(pl->d_cfo);
    pl->d_cfo = nullptr;
    pl->dev.n_cfo = 0;
    pl->dev.cfo = nullptr;
    pl->dev.cfo_nu = nullptr;
    pl->dev.cfo_continuous = 0;
    if (n_hyp == 0) return OFDM_OK;
    const int N = pl->prm.nfft;
    std::vector<float2> tab((size_t)n_hyp * N);
    const double two_pi = 6.283185307179586476925286766559;
    for (int h = 0; h < n_hyp; ++h)
        for (int n = 0; n < N; ++n) {
            const double a = two_pi * cycles_per_sample[h] * (double)n;          // SynchEstAndFO.py:196
            tab[(size_t)h * N + n] = make_float2((float)cos(a), (float)sin(a));
        }
    // one allocation: [n_hyp] double frequencies (8-byte aligned first), then the phasor rows
    const size_t b_nu = ((size_t)n_hyp * sizeof(double) + 255) & ~(size_t)255;
    CU(cudaMalloc(&pl->d_cfo, b_nu + tab.size() * sizeof(float2)), "ofdm_plan_set_cfo_bank(cudaMalloc)");
    CU(cudaMemcpy(pl->d_cfo, cycles_per_sample, (size_t)n_hyp * sizeof(double), cudaMemcpyHostToDevice), "ofdm_plan_set_cfo_bank(cudaMemcpy)");
    CU(cudaMemcpy((char *)pl->d_cfo + b_nu, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice), "ofdm_plan_set_cfo_bank(cudaMemcpy)");
    pl->dev.cfo_nu = (const double *)pl->d_cfo;
    pl->dev.cfo = (const float2 *)((char *)pl->d_cfo + b_nu);
    pl->dev.n_cfo = n_hyp;
    pl->dev.cfo_continuous = phase_continuous ? 1 : 0;
    return OFDM_OK;
}

int ofdm_rx_sync_search_cfo(ofdm_plan *pl, const float *d_time, int64_t n_samples, int64_t first_start, int64_t stride,
                            int64_t n_trials, float *d_peak, int32_t *d_delay, int32_t *d_hyp, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_trials == 0) return OFDM_OK;
    if (!pl || !d_time || !d_peak || !d_delay || n_trials < 0 || stride < 1) return fail(OFDM_EINVAL, "ofdm_rx_sync_search: bad argument%s");
    // stride-1 search of the utsa configuration: time-domain kernel, no transform per trial (OFDM_SYNC_TD=0 disables it)
    // (OFDM_SYNC_TD: 0 = per-trial transforms, 1 = plain time-domain kernel, 2 = register-blocked time-domain kernel)
    const char *td_env = getenv("OFDM_SYNC_TD");
    const int td_mode = td_env ? atoi(td_env) : OFDM_SYNC_TD_DEFAULT;
    // (a tile yields TD_TQ - CP trials for TD_TQ N-tap correlations: only worth it -- and only well-defined -- while
    // the prefix is at most half a tile; longer prefixes fall through to the per-trial transforms)
    // M0 = 2 and / or a hypothesis bank: the general time-domain kernel (OFDM_SYNC_TDG=0 keeps the per-trial transforms)
    const char *tdg_env = getenv("OFDM_SYNC_TDG");
    if (td_mode != 0 && (!tdg_env || atoi(tdg_env) != 0) && stride == 1 && pl->dev.h_td && (pl->dev.n_cfo || pl->prm.m0 == 2) &&
        first_start >= 0 && 2 * pl->prm.cp_len <= TD_TQ && (pl->prm.nfft + pl->prm.cp_len) % 8 == 0) {
        const int N = pl->prm.nfft, CP = pl->prm.cp_len, M0 = pl->prm.m0, L = N + CP, TP = TD_TQ - CP, A = (CP + 7) & ~7;
        const int XW = TD_TQ + N + A + (M0 - 1) * L, SB = TD_TQ / 8 + 1, CPA = CP + 1;
        const int smem = 8 * (XW / 8 + 1) * 8 + M0 * N * 8 + M0 * CPA * 8 + (N + 1) * 8 + 8 * SB * 12 + M0 * 8 + 16;
        if (smem <= 200 * 1024) {
            if (smem > pl->tdg_smem_set) {
                CU(cudaFuncSetAttribute(rx_sync_search_tdg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "ofdm_rx_sync_search(attr)");
                CU(cudaFuncSetAttribute(rx_sync_search_tdg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "ofdm_rx_sync_search(attr)");
                pl->tdg_smem_set = smem;
            }
            const long long tiles = (n_trials + TP - 1) / TP;
            if (M0 == 1)
                rx_sync_search_tdg_kernel<1><<<(unsigned)tiles, TD_THREADS, smem, (cudaStream_t)stream>>>(
                    pl->dev, (const float2 *)d_time, n_samples, first_start, n_trials, d_peak, d_delay, d_hyp);
            else
                rx_sync_search_tdg_kernel<2><<<(unsigned)tiles, TD_THREADS, smem, (cudaStream_t)stream>>>(
                    pl->dev, (const float2 *)d_time, n_samples, first_start, n_trials, d_peak, d_delay, d_hyp);
            count_launch();
            CU(cudaGetLastError(), "ofdm_rx_sync_search(tdg)");
            return OFDM_OK;
        }
    }
    // one synch symbol, no bank, N <= 2048: the overlap-save kernel (OFDM_SYNC_TD=3, the default; 2 keeps the N-tap kernel)
    if (td_mode == 3 && stride == 1 && pl->os_log2nb && pl->dev.h_td && pl->prm.m0 == 1 && !pl->dev.n_cfo && first_start >= 0 &&
        pl->prm.nfft % 8 == 0) {
        const int N = pl->prm.nfft, CP = pl->prm.cp_len, A = (CP + 7) & ~7, NB = 1 << pl->os_log2nb;
        int T = 0, sbytes = 0;
        switch (pl->os_log2nb) {
            case 9: T = CfgOs<9>::T; sbytes = Smem<CfgOs<9>, 1, false, false, false, false, 0>::BYTES; break;
            case 10: T = CfgOs<10>::T; sbytes = Smem<CfgOs<10>, 1, false, false, false, false, 0>::BYTES; break;
            case 11: T = CfgOs<11>::T; sbytes = Smem<CfgOs<11>, 1, false, false, false, false, 0>::BYTES; break;
            default: T = CfgOs<12>::T; sbytes = Smem<CfgOs<12>, 1, false, false, false, false, 0>::BYTES; break;
        }
        int PB = (NB - N + 1) & ~7;
        if (PB > 8 * T) PB = 8 * T;
        if (PB + 8 > NB) PB = NB - 8;
        const int TP = PB - CP, XW = A + PB + N, SX = XW / 8 + 1, SB = PB / 8 + 1, NC = XW / 8;
        const int tail = 8 * SB * 4 + 5 * (NC + 1) * 4;              // window powers + chunk prefixes
        const int alias_all = 16 * (PB + 8) + tail <= 16 * NB ? 1 : 0;
        const int smem = sbytes + 8 * SX * 8 + ((CP + 2) & ~1) * 8 + (alias_all ? 0 : tail) + 16;
        if (TP > 0 && smem <= 220 * 1024) {
            DevParams Q = pl->dev;
            Q.tw = pl->d_tw_os;                                 // the block transform's own twiddles
            const long long tiles = (n_trials + TP - 1) / TP;
#define OFDM_OS_LAUNCH(L2)                                                                                                   \
    {                                                                                                                        \
        if (smem > pl->os_smem_set) {                                                                                        \
            CU(cudaFuncSetAttribute(rx_sync_search_os_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),        \
               "ofdm_rx_sync_search(attr)");                                                                                 \
            pl->os_smem_set = smem;                                                                                          \
        }                                                                                                                    \
        rx_sync_search_os_kernel<L2><<<(unsigned)tiles, T, smem, (cudaStream_t)stream>>>(                                    \
            Q, (const float2 *)d_time, n_samples, first_start, n_trials, d_peak, d_delay, pl->d_gos, PB, alias_all);         \
    }
            switch (pl->os_log2nb) {
                case 9: OFDM_OS_LAUNCH(9) break;
                case 10: OFDM_OS_LAUNCH(10) break;
                case 11: OFDM_OS_LAUNCH(11) break;
                default: OFDM_OS_LAUNCH(12) break;
            }
#undef OFDM_OS_LAUNCH
            count_launch();
            CU(cudaGetLastError(), "ofdm_rx_sync_search(os)");
            if (d_hyp) CU(cudaMemsetAsync(d_hyp, 0, (size_t)n_trials * sizeof(int32_t), (cudaStream_t)stream), "ofdm_rx_sync_search(hyp)");
            return OFDM_OK;
        }
    }
    if (td_mode != 0 && stride == 1 && pl->dev.h_td && pl->prm.m0 == 1 && !pl->dev.n_cfo && first_start >= 0 && 2 * pl->prm.cp_len <= TD_TQ) {
        const int N = pl->prm.nfft, CP = pl->prm.cp_len, TP = TD_TQ - CP, A = (CP + 7) & ~7;
        const int smem = td_mode >= 2 ? 8 * ((TD_TQ + N + A) / 8 + 1) * 8 + N * 8 + 8 * (TD_TQ / 8 + 1) * 12
                                      : (TD_TQ + N + CP) * 8 + N * 8 + TD_TQ * 8 + TD_TQ * 4;
        if (smem > pl->td_smem_set) {                          // per plan, i.e. per device: the attribute belongs to the context
            CU(cudaFuncSetAttribute(rx_sync_search_td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "ofdm_rx_sync_search(attr)");
            CU(cudaFuncSetAttribute(rx_sync_search_td8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "ofdm_rx_sync_search(attr)");
            pl->td_smem_set = smem;
        }
        const long long tiles = (n_trials + TP - 1) / TP;
        if (td_mode >= 2)
            rx_sync_search_td8_kernel<<<(unsigned)tiles, TD_THREADS, smem, (cudaStream_t)stream>>>(
                pl->dev, (const float2 *)d_time, n_samples, first_start, n_trials, d_peak, d_delay);
        else
            rx_sync_search_td_kernel<<<(unsigned)tiles, TD_THREADS, smem, (cudaStream_t)stream>>>(
                pl->dev, (const float2 *)d_time, n_samples, first_start, n_trials, d_peak, d_delay);
        count_launch();
        CU(cudaGetLastError(), "ofdm_rx_sync_search(td)");
        // no hypothesis bank on this path: the winning hypothesis is 0, as the per-trial kernel reports it
        if (d_hyp) CU(cudaMemsetAsync(d_hyp, 0, (size_t)n_trials * sizeof(int32_t), (cudaStream_t)stream), "ofdm_rx_sync_search(hyp)");
        return OFDM_OK;
    }
    CU(pl->ks->sync_search(pl->dev, (const float2 *)d_time, n_samples, first_start, stride, n_trials, d_peak, d_delay,
                           d_hyp, (cudaStream_t)stream, pl->sms),
       "ofdm_rx_sync_search");
    return OFDM_OK;
}
int ofdm_rx_sync_search(ofdm_plan *pl, const float *d_time, int64_t n_samples, int64_t first_start, int64_t stride,
                        int64_t n_trials, float *d_peak, int32_t *d_delay, void *stream) {
    return ofdm_rx_sync_search_cfo(pl, d_time, n_samples, first_start, stride, n_trials, d_peak, d_delay, nullptr, stream);
}

int ofdm_rx_chanest_cfo(ofdm_plan *pl, const float *d_time, int64_t n_samples, const int64_t *d_pos,
                        const int32_t *d_delay, const int32_t *d_hyp, int64_t n_locks, float *d_H, float *d_htime,
                        float *d_synch_eq, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_locks == 0) return OFDM_OK;
    if (!pl || !d_time || !d_pos || !d_delay || !d_H || n_locks < 0) return fail(OFDM_EINVAL, "ofdm_rx_chanest: bad argument%s");
    CU(pl->ks->chanest(pl->dev, (const float2 *)d_time, n_samples, (const long long *)d_pos, d_delay, d_hyp, n_locks,
                       (float2 *)d_H, (float2 *)d_htime, (float2 *)d_synch_eq, (cudaStream_t)stream, pl->sms, nullptr, nullptr),
       "ofdm_rx_chanest");
    return OFDM_OK;
}
int ofdm_rx_chanest(ofdm_plan *pl, const float *d_time, int64_t n_samples, const int64_t *d_pos, const int32_t *d_delay,
                    int64_t n_locks, float *d_H, float *d_htime, float *d_synch_eq, void *stream) {
    return ofdm_rx_chanest_cfo(pl, d_time, n_samples, d_pos, d_delay, nullptr, n_locks, d_H, d_htime, d_synch_eq, stream);
}

int ofdm_rx_demod_cfo(ofdm_plan *pl, const float *d_time, int64_t n_samples, const int64_t *d_start,
                      const int32_t *d_delay, const int32_t *d_hrow, const int32_t *d_hyp, const float *d_H, int64_t n_sym,
                      float *d_eq, float *d_llr, uint8_t *d_hard, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_sym == 0) return OFDM_OK;
    if (!pl || !d_time || !d_start || !d_delay || !d_hrow || !d_H || !d_eq || n_sym < 0)
        return fail(OFDM_EINVAL, "ofdm_rx_demod: bad argument%s");
    CU(pl->ks->rx_demod(pl->dev, (const float2 *)d_time, n_samples, (const long long *)d_start, d_delay, d_hrow, d_hyp,
                        (const float2 *)d_H, n_sym, (float2 *)d_eq, d_llr, d_hard, (cudaStream_t)stream, pl->sms, nullptr),
       "ofdm_rx_demod");
    return OFDM_OK;
}
int ofdm_rx_demod(ofdm_plan *pl, const float *d_time, int64_t n_samples, const int64_t *d_start, const int32_t *d_delay,
                  const int32_t *d_hrow, const float *d_H, int64_t n_sym, float *d_eq, float *d_llr, uint8_t *d_hard,
                  void *stream) {
    return ofdm_rx_demod_cfo(pl, d_time, n_samples, d_start, d_delay, d_hrow, nullptr, d_H, n_sym, d_eq, d_llr, d_hard, stream);
}

int ofdm_rx_frames_cfo(ofdm_plan *pl, const float *d_time, int64_t n_samples, int64_t first, int64_t frame_stride,
                       int64_t n_frames, int32_t *d_delay, float *d_peak, uint8_t *d_locked, int32_t *d_hyp, float *d_H,
                       float *d_eq, float *d_llr, uint8_t *d_hard, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_frames == 0) return OFDM_OK;
    if (!pl || !d_time || !d_delay || !d_peak || !d_locked || n_frames < 0 || (!d_eq && !d_llr && !d_hard))
        return fail(OFDM_EINVAL, "ofdm_rx_frames: bad argument (at least one of eq / llr / hard is needed)%s");
    CU(pl->ks->rx_frames(pl->dev, (const float2 *)d_time, n_samples, first, frame_stride, n_frames, d_delay, d_peak,
                         d_locked, d_hyp, (float2 *)d_H, (float2 *)d_eq, d_llr, d_hard, (cudaStream_t)stream, pl->sms),
       "ofdm_rx_frames");
    return OFDM_OK;
}
int ofdm_rx_frames(ofdm_plan *pl, const float *d_time, int64_t n_samples, int64_t first, int64_t frame_stride,
                   int64_t n_frames, int32_t *d_delay, float *d_peak, uint8_t *d_locked, float *d_H, float *d_eq,
                   float *d_llr, uint8_t *d_hard, void *stream) {
    return ofdm_rx_frames_cfo(pl, d_time, n_samples, first, frame_stride, n_frames, d_delay, d_peak, d_locked, nullptr, d_H,
                              d_eq, d_llr, d_hard, stream);
}

int ofdm_bit_recovery(ofdm_plan *pl, const float *d_in, int64_t n, float *d_llr0, float *d_llr1, uint8_t *d_hard,
                      double *d_ws, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n == 0) return OFDM_OK;
    if (!pl || !d_in || !d_llr0 || !d_llr1 || !d_hard || !d_ws || n < 0) return fail(OFDM_EINVAL, "ofdm_bit_recovery: bad argument%s");
    if (n == 0) return OFDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemsetAsync(d_ws, 0, sizeof(double), st), "ofdm_bit_recovery(memset)");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)pl->sms * 8;
    if (blocks > cap) blocks = cap;
    bitrec_sum_kernel<<<(int)blocks, 256, 0, st>>>((const float2 *)d_in, n, d_ws);
    count_launch();
    bitrec_llr_kernel<<<(int)blocks, 256, 0, st>>>((const float2 *)d_in, n, d_ws, d_llr0, d_llr1, d_hard);
    count_launch();
    CU(cudaGetLastError(), "ofdm_bit_recovery");
    return OFDM_OK;
}

int ofdm_dsss_despread(ofdm_plan *pl, const float *d_eq, int64_t n_rows, int32_t row_len, const float *d_code, int32_t sf,
                       float *d_out, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_rows == 0) return OFDM_OK;
    if (!pl || !d_eq || !d_code || !d_out || n_rows < 0 || sf < 1 || row_len < sf)
        return fail(OFDM_EINVAL, "ofdm_dsss_despread: bad argument%s");
    const int n_groups = row_len / sf;
    long long blocks = (n_rows * n_groups + 255) / 256;
    const long long cap = (long long)pl->sms * 16;
    if (blocks > cap) blocks = cap;
    dsss_despread_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float2 *)d_eq, n_rows, row_len, n_groups,
                                                                        (const float2 *)d_code, sf, (float2 *)d_out);
    count_launch();
    CU(cudaGetLastError(), "ofdm_dsss_despread");
    return OFDM_OK;
}

int ofdm_counters(ofdm_plan *pl, const uint8_t *d_bits_rx, const uint8_t *d_bits_tx, const float *d_eq,
                  const uint8_t *d_locked, int64_t n_frames, double *d_cnt, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_frames == 0) return OFDM_OK;
    if (!pl || !d_bits_tx || !d_cnt || n_frames < 0) return fail(OFDM_EINVAL, "ofdm_counters: bad argument%s");
    if (n_frames == 0) return OFDM_OK;
    long long blocks = (n_frames * pl->prm.d + 7) / 8;
    const long long cap = (long long)pl->sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    counters_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pl->dev, d_bits_rx, d_bits_tx, (const float2 *)d_eq,
                                                                   d_locked, n_frames, d_cnt);
    count_launch();
    CU(cudaGetLastError(), "ofdm_counters");
    return OFDM_OK;
}

int ofdm_channel_awgn(ofdm_plan *pl, const float *d_tx, float *d_rx, int64_t n_frames, const float *d_taps,
                      int32_t n_taps, int32_t taps_per_frame, float noise_sigma, uint64_t seed, void *stream) {
    DevGuard dev_guard(pl);
    if (pl && n_frames == 0) return OFDM_OK;
    if (!pl || !d_tx || !d_rx || !d_taps || n_taps < 1 || n_frames < 0) return fail(OFDM_EINVAL, "ofdm_channel_awgn: bad argument%s");
    if (n_frames == 0) return OFDM_OK;
    const int fl = ofdm_plan_frame_len(pl);
    long long blocks = (n_frames * fl + 255) / 256;
    const long long cap = (long long)pl->sms * 16;
    if (blocks > cap) blocks = cap;
    channel_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float2 *)d_tx, (float2 *)d_rx, n_frames, fl,
                                                                  (const float2 *)d_taps, n_taps, taps_per_frame,
                                                                  noise_sigma, seed);
    count_launch();
    CU(cudaGetLastError(), "ofdm_channel_awgn");
    return OFDM_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points: chunked, double-buffered H2D -> kernel -> D2H on two internal streams
// ------------------------------------------------------------------------------------------------
static int ensure_streams(ofdm_plan *pl) {
    if (pl->streams_made) return OFDM_OK;
    for (int i = 0; i < ofdm_plan::NSLOT; ++i) {
        CU(cudaStreamCreateWithFlags(&pl->st[i], cudaStreamNonBlocking), "cudaStreamCreate");
        CU(cudaEventCreateWithFlags(&pl->ev[i], cudaEventDisableTiming), "cudaEventCreate");
    }
    pl->streams_made = true;
    return OFDM_OK;
}
static int ensure_stage(ofdm_plan *pl, int slot, size_t bytes) {
    if (pl->stage_bytes[slot] >= bytes) return OFDM_OK;
    if (pl->stage[slot]) cudaFree(pl->stage[slot]);
    pl->stage[slot] = nullptr;
    pl->stage_bytes[slot] = 0;
    CU(cudaMalloc(&pl->stage[slot], bytes), "staging cudaMalloc");
    pl->stage_bytes[slot] = bytes;
    return OFDM_OK;
}
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int ofdm_rx_frames_host(ofdm_plan *pl, const float *h_time, int64_t n_samples, int64_t first, int64_t frame_stride,
                        int64_t n_frames, int32_t *h_delay, float *h_peak, uint8_t *h_locked, float *h_eq, float *h_llr,
                        uint8_t *h_hard) {
    DevGuard dev_guard(pl);
    if (pl && n_frames == 0) return OFDM_OK;
    if (!pl || !h_time || (!h_eq && !h_llr && !h_hard) || n_frames < 0 || frame_stride < 1)
        return fail(OFDM_EINVAL, "ofdm_rx_frames_host: bad argument (at least one of eq / llr / hard is needed)%s");
    if (n_frames == 0) return OFDM_OK;
    int rc = ensure_streams(pl);
    if (rc) return rc;
    const ofdm_params &p = pl->prm;
    const int64_t fl = ofdm_plan_frame_len(pl), cpmax = p.cp_len + pl->dev.cp_extra;
    const int64_t ndb = pl->dev.ndd, hs = pl->dev.hard_stride;
    // Chunk schedule: the link back to the host is the bottleneck (D2H bytes > H2D bytes).  The first chunk is small
    // so that the first D2H starts early, and chunks then grow by 1.5x (slower than the D2H : H2D byte ratio, so the
    // upload of chunk i+1 always hides behind the download of chunk i) up to the steady size.
    // OFDM_HOST_CHUNK_MB / OFDM_HOST_FIRST_MB (environment) override the two sizes.
    static const int big_mb = getenv("OFDM_HOST_CHUNK_MB") ? atoi(getenv("OFDM_HOST_CHUNK_MB")) : 32;
    static const int first_mb = getenv("OFDM_HOST_FIRST_MB") ? atoi(getenv("OFDM_HOST_FIRST_MB")) : 4;
    auto frames_for = [&](int mb) {
        int64_t n = ((int64_t)(mb > 0 ? mb : 1) << 20) / (frame_stride * 8);
        if (n < 1) n = 1;
        return n < n_frames ? n : n_frames;
    };
    const int64_t per_big = frames_for(big_mb), per_first = frames_for(first_mb < big_mb ? first_mb : big_mb);
    const int64_t per = per_big;
    const size_t b_in = al256((size_t)((per - 1) * frame_stride + fl + 2 * cpmax) * 8);
    const size_t b_eq = al256(per * p.d * ndb * 8), b_ll = al256(per * p.d * ndb * p.bps * 4), b_hd = al256(per * p.d * hs);
    const size_t tot = b_in + b_eq + b_ll + b_hd;
    for (int s = 0; s < ofdm_plan::NSLOT; ++s) {
        rc = ensure_stage(pl, s, tot);
        if (rc) return rc;
    }
    // per-frame results (9 bytes per frame) are kept for the whole call and copied back once at the end
    const size_t b_dl = al256(n_frames * 4), b_pk = al256(n_frames * 4), b_lk = al256(n_frames);
    if (pl->small_bytes < b_dl + b_pk + b_lk) {
        if (pl->small) cudaFree(pl->small);
        pl->small = nullptr;
        pl->small_bytes = 0;
        CU(cudaMalloc(&pl->small, b_dl + b_pk + b_lk), "per-frame staging cudaMalloc");
        pl->small_bytes = b_dl + b_pk + b_lk;
    }
    int32_t *ddl = (int32_t *)pl->small;
    float *dpk = (float *)((char *)pl->small + b_dl);
    uint8_t *dlk = (uint8_t *)((char *)pl->small + b_dl + b_pk);
    int slot = 0;
    int64_t want = per_first;
    for (int64_t f0 = 0; f0 < n_frames; slot = (slot + 1) % ofdm_plan::NSLOT) {
        const int64_t nf = (n_frames - f0 < want) ? (n_frames - f0) : want;
        want = want + want / 2 + 1;
        if (want > per_big) want = per_big;
        cudaStream_t st = pl->st[slot];
        char *b = (char *)pl->stage[slot];
        float *din = (float *)b;
        float *deq = (float *)(b + b_in);
        float *dll = (float *)(b + b_in + b_eq);
        uint8_t *dhd = (uint8_t *)(b + b_in + b_eq + b_ll);
        // samples needed by this chunk: [lo, hi)
        int64_t lo = first + f0 * frame_stride - p.sync_advance;
        if (lo < 0) lo = 0;
        int64_t hi = first + (f0 + nf - 1) * frame_stride + fl + cpmax;
        if (hi > n_samples) hi = n_samples;
        if (hi < lo) hi = lo;
        CU(cudaMemcpyAsync(din, h_time + 2 * lo, (size_t)(hi - lo) * 8, cudaMemcpyHostToDevice, st), "H2D");
        CU(pl->ks->rx_frames(pl->dev, (const float2 *)din, hi - lo, first + f0 * frame_stride - lo, frame_stride, nf, ddl + f0,
                             dpk + f0, dlk + f0, nullptr, nullptr, h_eq ? (float2 *)deq : nullptr, h_llr ? dll : nullptr,
                             h_hard ? dhd : nullptr, st, pl->sms),
           "ofdm_rx_frames_host(kernel)");
        if (h_llr)
            CU(cudaMemcpyAsync(h_llr + f0 * p.d * ndb * p.bps, dll, (size_t)nf * p.d * ndb * p.bps * 4, cudaMemcpyDeviceToHost, st), "D2H");
        if (h_eq) CU(cudaMemcpyAsync(h_eq + 2 * f0 * p.d * ndb, deq, (size_t)nf * p.d * ndb * 8, cudaMemcpyDeviceToHost, st), "D2H");
        if (h_hard) CU(cudaMemcpyAsync(h_hard + f0 * p.d * hs, dhd, (size_t)nf * p.d * hs, cudaMemcpyDeviceToHost, st), "D2H");
        f0 += nf;
    }
    for (int s = 0; s < ofdm_plan::NSLOT; ++s) CU(cudaStreamSynchronize(pl->st[s]), "ofdm_rx_frames_host(sync)");
    if (h_delay) CU(cudaMemcpyAsync(h_delay, ddl, n_frames * 4, cudaMemcpyDeviceToHost, pl->st[0]), "D2H");
    if (h_peak) CU(cudaMemcpyAsync(h_peak, dpk, n_frames * 4, cudaMemcpyDeviceToHost, pl->st[0]), "D2H");
    if (h_locked) CU(cudaMemcpyAsync(h_locked, dlk, n_frames, cudaMemcpyDeviceToHost, pl->st[0]), "D2H");
    CU(cudaStreamSynchronize(pl->st[0]), "ofdm_rx_frames_host(sync)");
    return OFDM_OK;
}

int ofdm_tx_modulate_host(ofdm_plan *pl, const uint8_t *h_bits, float *h_time, int64_t n_frames) {
    DevGuard dev_guard(pl);
    if (pl && n_frames == 0) return OFDM_OK;
    if (!pl || !h_time || (!h_bits && pl->prm.d > 0) || n_frames < 0) return fail(OFDM_EINVAL, "ofdm_tx_modulate_host: bad argument%s");
    if (n_frames == 0) return OFDM_OK;
    int rc = ensure_streams(pl);
    if (rc) return rc;
    const ofdm_params &p = pl->prm;
    const int64_t fl = ofdm_plan_frame_len(pl), hs = pl->dev.hard_stride;
    int64_t per = (32ll << 20) / (fl * 8);
    if (per < 1) per = 1;
    if (per > n_frames) per = n_frames;
    const size_t b_bits = al256((size_t)per * p.d * hs + 16), b_out = al256((size_t)per * fl * 8);
    for (int s = 0; s < ofdm_plan::NSLOT; ++s) {
        rc = ensure_stage(pl, s, b_bits + b_out);
        if (rc) return rc;
    }
    int slot = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += per, slot = (slot + 1) % ofdm_plan::NSLOT) {
        const int64_t nf = (n_frames - f0 < per) ? (n_frames - f0) : per;
        cudaStream_t st = pl->st[slot];
        char *b = (char *)pl->stage[slot];
        if (p.d > 0) CU(cudaMemcpyAsync(b, h_bits + f0 * p.d * hs, (size_t)nf * p.d * hs, cudaMemcpyHostToDevice, st), "H2D");
        CU(pl->ks->tx(pl->dev, (const uint8_t *)b, (float2 *)(b + b_bits), nf * (p.m0 + p.d), st, pl->sms), "ofdm_tx_modulate_host(kernel)");
        CU(cudaMemcpyAsync(h_time + 2 * f0 * fl, b + b_bits, (size_t)nf * fl * 8, cudaMemcpyDeviceToHost, st), "D2H");
    }
    for (int s = 0; s < ofdm_plan::NSLOT; ++s) CU(cudaStreamSynchronize(pl->st[s]), "ofdm_tx_modulate_host(sync)");
    return OFDM_OK;
}

}  // extern "C"
