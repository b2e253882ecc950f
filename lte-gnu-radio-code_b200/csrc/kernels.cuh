// kernels.cuh -- the OFDM modem kernels (templated on log2(NFFT) and bits per bin).
//
// One "symbol group" of T = N/8 threads owns one OFDM symbol at a time; groups are persistent and
// stride over the work list.  All per-bin work (ZC correlation, channel estimate, equaliser gain,
// de-rotation, demapping) is done on the registers the FFT leaves behind, so a received sample is
// read from HBM exactly once and every output is written exactly once.
//
// Instruction-count rules followed throughout (the N=64 case is issue-bound before it is HBM-bound):
//   * per-thread bin bookkeeping (which of my 8 bins are used, where they land) is computed once;
//     thread j's bins are j + m*T, so within each spectrum half its outputs sit at a fixed stride
//     and every store is `base + compile-time offset`;
//   * groups past the end of the work list redo the last item (identical duplicate writes) instead
//     of carrying a `valid` predicate through every load and store;
//   * the next symbol's samples are requested with cp.async into per-thread shared-memory slots
//     before the current one is transformed, so DRAM latency overlaps the arithmetic without
//     holding registers.
#pragma once
#include <cuda_fp16.h>
#include <type_traits>
#include "fft_core.cuh"

namespace ofdm {

struct DevParams {
    int nfft, cp, nsb, ndb, m0, d, bps, mm;
    // CP pattern (extension): symbols k % slot == 0 of a frame carry cp + cp_extra prefix samples; cp_extra = 0 and
    // slot = INT_MAX for the reference's uniform prefix, so the counters below never fire
    int cp_extra, slot;
    int sync_advance, hard_stride, dup_bin, equalizer;
    float gate_thresh;        // gate * M0*nsb
    float chanest_scale;      // 1 / ((1 + inv_reg_chanest) * M0)
    float inv_reg_eq, inv_reg_eq_synch, llr_scale;
    const float2 *tw;         // [N]      exp(-2*pi*i*k/N)
    const float2 *zc_bin;     // [M0][N]  sum of conj(zc[m*nsb+idx]) over the synch indices that map to FFT bin k
    const float *w_bin;       // [N]      how many synch indices map to bin k (0, 1, or 2 for the +-N/2 alias)
    const float2 *zc_dup;     // [M0][2]  conj(zc) of the first / last synch index (alias handling)
    const float2 *zc_tx;      // [M0][N]  ZC value the transmitter puts in bin k of synch symbol m (0 if unused)
    // comb pilots in data symbols (extension): positions pil_off + i*pil_k of the used-bin list
    int pil_k, pil_off, n_pil, ndd;   // ndd = payload bins per data symbol = ndb - n_pil
    int ls_stride;                    // floats per group of the LLR row staging area (0: none)
    int es_stride;                    // float2 per group of the equalised-symbol row staging area (bulk-store variant)
    int bulk_rows;                    // 1: the plan asks for the bulk-store variant of the framed receiver
    // frequency-offset hypothesis bank (SynchEstAndFO.py:196): cfo[h][n] = exp(j*2pi*nu_h*n), n = 0..N-1
    int n_cfo;
    const float2 *cfo;
    // extension: continue every hypothesis' phasor from the buffer origin, exp(j*2pi*nu_h*(start+n)), instead of
    // restarting it at each window like the reference (needed to undo a real carrier offset)
    int cfo_continuous;
    const double *cfo_nu;             // [n_cfo] cycles per sample
    const float2 *pilot_tx;   // [N]      pilot value in FFT bin k (0 elsewhere)
    const float2 *h_td;       // [N]      taps of the time-domain acquisition kernel (api.cu), or null when it does not apply
};

// position of a used-bin-list index among pilots / payload bins
struct PilotPos {
    bool is_pilot;
    int ip;        // pilot index (is_pilot) or index of the pilot at or below jo (clamped to 0)
    int dj;        // payload index (jo minus the pilots before it)
    float frac;    // interpolation weight towards pilot ip+1 (0 when held constant at an edge)
};
__device__ __forceinline__ PilotPos pilot_pos(const DevParams &P, int jo) {
    PilotPos r;
    const int rel = jo - P.pil_off;
    const int q = rel >= 0 ? rel / P.pil_k : 0, rem = rel >= 0 ? rel - q * P.pil_k : 0;
    r.is_pilot = rel >= 0 && rem == 0 && q < P.n_pil;
    int before = rel <= 0 ? 0 : q + (rem != 0 ? 1 : 0);            // pilots strictly below jo
    before = before < P.n_pil ? before : P.n_pil;
    r.dj = jo - before;
    if (rel < 0) {
        r.ip = 0;
        r.frac = 0.f;
    } else if (q >= P.n_pil - 1) {
        r.ip = P.n_pil - 1;
        r.frac = 0.f;
    } else {
        r.ip = q;
        r.frac = (float)rem / (float)P.pil_k;
    }
    return r;
}

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

// Where thread j's PPT bins (j + m*T) land in a used-bin list of length n.  m < PPT/2 is the positive
// half of the spectrum (incl. DC), m >= PPT/2 the negative half; inside a half consecutive m are T apart
// in the output as well, so a store is  half_base + m*T  with m*T a compile-time constant.
struct BinMap {
    int base_p;        // output index of bin j (m = 0) if it were used:            j - 1 + n/2
    int base_n;        // output index of bin j + N/2 (m = PPT/2) if it were used:  j + N/2 - (N - n/2)
    unsigned mask;     // bit m set <=> bin j + m*T is in the list
};
template <int T, int PPT> __device__ __forceinline__ BinMap make_bin_map(int n, int j) {
    constexpr int N = PPT * T, H = PPT / 2;
    const int h = n >> 1;
    BinMap b;
    b.base_p = j - 1 + h;
    b.base_n = j + H * T - (N - h);
    b.mask = 0;
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        const int k = j + m * T;
        const bool used = (m < H) ? (k >= 1 && k <= h) : (k >= N - h);
        b.mask |= used ? (1u << m) : 0u;
    }
    return b;
}
template <int T, int PPT> __device__ __forceinline__ int bin_out(const BinMap &b, int m) {
    return (m < PPT / 2 ? b.base_p : b.base_n - (PPT / 2) * T) + m * T;
}

// The framed receiver's epilogue is branch-free (see data_epilogue, DENSE) when at least this many 64ths of the FFT bins
// carry data; 65 = never.  B200: +3..10 % at a 94 % fill for every size, slower at the LTE fill of 59 %.
// tuning aid: 1 compiles the receiver's optional equalised-symbol output as always present (A/B of the run-time switch)
#ifndef OFDM_EPI_MASKS
#define OFDM_EPI_MASKS 3          // 0: `used && wanted` per bin; 1: the eq output through its own mask; 3: all three outputs
#endif
#ifndef OFDM_EQ_ALWAYS
#define OFDM_EQ_ALWAYS 0
#endif
#ifndef OFDM_EPI_CHUNK_SPARSE
#define OFDM_EPI_CHUNK_SPARSE 0
#endif
#ifndef OFDM_EPI_CHUNK_BULK
#define OFDM_EPI_CHUNK_BULK 0
#endif
#ifndef OFDM_DENSE_MIN_FILL64
#define OFDM_DENSE_MIN_FILL64 48
#endif

// Output stores: every output byte is written once and never read again by the kernel.  CS selects
// st.global.cs (streaming, evict first) over a plain write-back store.  Measured on B200: streaming stores are
// worth 4 % in the transmitter and 1-2 % in the receiver at N >= 1024, plain stores 2 % in the receiver at N = 64
// (OFDM_RX_CS_MIN_LOG2N: smallest transform whose receiver uses the streaming form).
#ifndef OFDM_RX_CS_MIN_LOG2N
#define OFDM_RX_CS_MIN_LOG2N 7
#endif
template <bool CS = true, class V> __device__ __forceinline__ void st_out(V *p, V v) {
    if constexpr (std::is_same<V, float2>::value) v = pin(v);
    if (CS) __stcs(p, v); else *p = v;
}

// Bulk (TMA) row stores: a finished output row is assembled in shared memory and handed to the async proxy as
// ONE shared->global copy (cp.async.bulk), so the LSU sees 16-byte shared-memory stores instead of global stores
// and no register stays pinned behind a store queue.  Sizes and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ void bulk_store_row(void *gdst, const void *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all earlier bulk copies of this thread have finished READING shared memory (the rows may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have completed altogether (their global writes are done)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's shared-memory writes visible to the async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float2 ld_stream(const float2 *p) {      // read-once data: do not allocate in L1
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// this thread's PPT samples of the N-sample window that opens at `start` (zero beyond the buffer)
template <int T, int PPT>
__device__ __forceinline__ void load_pts(float2 (&v)[PPT], const float2 *__restrict__ x, long long start, long long n,
                                         int j) {
    constexpr int N = PPT * T;
    if (start >= 0 && start + N <= n) {
        const float2 *p = x + start + j;
#pragma unroll
        for (int m = 0; m < PPT; ++m) v[m] = ld_stream(p + m * T);
    } else {
#pragma unroll
        for (int m = 0; m < PPT; ++m) {
            const long long i = start + j + m * T;
            v[m] = (i >= 0 && i < n) ? __ldg(x + i) : make_float2(0.f, 0.f);
        }
    }
}

// CFO de-rotation of the window that opens at absolute sample `start`: sample n times row[n]
// (window-local, SynchEstAndFO.py:254,:332), times exp(j*2pi*nu*start) in the phase-continuous extension.
__device__ __forceinline__ float2 cfo_window_phase(const DevParams &P, int h, long long start) {
    if (!P.cfo_continuous) return make_float2(1.f, 0.f);
    double t = P.cfo_nu[h] * (double)start;
    t -= floor(t);
    double sn, cs;
    sincospi(2.0 * t, &sn, &cs);
    return make_float2((float)cs, (float)sn);
}
template <int T, int PPT>
__device__ __forceinline__ void rotate_pts(const DevParams &P, float2 (&v)[PPT], int h, long long start, int j) {
    const float2 *row = P.cfo + (long long)h * (PPT * T);
    const float2 ph = cfo_window_phase(P, h, start);
#pragma unroll
    for (int m = 0; m < PPT; ++m) v[m] = cmul(v[m], cmul(__ldg(row + j + m * T), ph));
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
//   L_b = scale * ( min_{s:b=0} (x-s)^2 - min_{s:b=1} (x-s)^2 ),  c = 4*a*scale, t = |x|
//   sign bit  : -sgn(x) * c * (t + max(t-2a,0) [+ max(t-4a,0) + max(t-6a,0)])
//   16-QAM b1 : c * (t - 2a)
//   64-QAM b1 : c * ((t-4a) + min(t-2a,0) + max(t-6a,0));   b2 : c * (|t-4a| - 2a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float flip_sign_of(float mag, float x) {       // mag >= 0 -> -sgn(x)*mag
    return __uint_as_float(__float_as_uint(mag) | (~__float_as_uint(x) & 0x80000000u));
}
__device__ __forceinline__ unsigned neg_bit(float x) { return __float_as_uint(x) >> 31; }

// append the sign bit of x to an MSB-first bit accumulator: one funnel shift per bit
__device__ __forceinline__ unsigned push_sign(unsigned acc, float x) { return __funnelshift_l(__float_as_uint(x), acc, 1); }

#ifndef OFDM_DEMAP64_PACKED
#define OFDM_DEMAP64_PACKED 0    // B200: LTE RX 0.643 -> 0.640 ms, N=4096 64-QAM 0.922 -> 0.955 ms: instruction count is not what bounds them
#endif
template <int BPS> struct Demapper {
    static constexpr float A = BPS == 2 ? 0.70710678118654752440f
                                        : (BPS == 4 ? 0.31622776601683793320f : 0.15430334996209191026f);
    float c;
    __device__ __forceinline__ explicit Demapper(float scale) : c(BPS == 1 ? 4.f * scale : 4.f * A * scale) {}

    __device__ __forceinline__ void run(float2 z, float (&L)[BPS], unsigned &bits) const {
        if (BPS == 1) {
            L[0] = c * z.x;                              // levels 2b-1: bit 1 -> +1
            bits = neg_bit(z.x) ^ 1u;
        } else if (BPS == 2) {
            L[0] = -c * z.x;
            L[1] = -c * z.y;
            bits = push_sign(push_sign(0u, z.x), z.y);
        } else if (BPS == 4) {
            const float ux = fabsf(z.x) - 2.f * A, uy = fabsf(z.y) - 2.f * A;
            L[0] = flip_sign_of(c * (fabsf(z.x) + fmaxf(ux, 0.f)), z.x);
            L[1] = flip_sign_of(c * (fabsf(z.y) + fmaxf(uy, 0.f)), z.y);
            L[2] = c * ux;
            L[3] = c * uy;
            // hard bits = [x<0, y<0, |x|>2a, |y|>2a]: the last two are the inverted signs of ux, uy
            bits = push_sign(push_sign(push_sign(push_sign(0u, z.x), z.y), ux), uy) ^ 3u;
        } else {
            // with m2 = max(t-2a,0), m4 = max(t-4a,0), m6 = max(t-6a,0):
            //   b0: -sgn(x) c (t + m2 + m4 + m6)    b1: c (2t - 6a - m2 + m6)    b2: c (2 m4 - t + 2a)
            float a0[2], a1[2], a2[2];
            const float c2 = 2.f * c, k1 = -6.f * A * c, k2 = 2.f * A * c;
#if OFDM_F32X2 && OFDM_DEMAP64_PACKED
            // both axes at once on the packed pipe: the same IEEE operations in the same order as the scalar form below
            // (bit-identical results), 22 instead of 32 instructions per bin; only |x| and max(., 0) stay scalar
            const float2 t = make_float2(fabsf(z.x), fabsf(z.y));
            const float2 u2 = __fadd2_rn(t, make_float2(-2.f * A, -2.f * A)), u4 = __fadd2_rn(t, make_float2(-4.f * A, -4.f * A)),
                         u6 = __fadd2_rn(t, make_float2(-6.f * A, -6.f * A));
            const float2 m2 = make_float2(fmaxf(u2.x, 0.f), fmaxf(u2.y, 0.f)), m4 = make_float2(fmaxf(u4.x, 0.f), fmaxf(u4.y, 0.f)),
                         m6 = make_float2(fmaxf(u6.x, 0.f), fmaxf(u6.y, 0.f));
            const float2 s0 = __fmul2_rn(make_float2(c, c), __fadd2_rn(__fadd2_rn(t, m2), __fadd2_rn(m4, m6)));
            a0[0] = flip_sign_of(s0.x, z.x);
            a0[1] = flip_sign_of(s0.y, z.y);
            const float2 b1 = __ffma2_rn(make_float2(c, c), __fadd2_rn(m6, make_float2(-m2.x, -m2.y)),
                                         __ffma2_rn(make_float2(c2, c2), t, make_float2(k1, k1)));
            const float2 b2 = __ffma2_rn(make_float2(c2, c2), m4, __ffma2_rn(make_float2(-c, -c), t, make_float2(k2, k2)));
            a1[0] = b1.x; a1[1] = b1.y; a2[0] = b2.x; a2[1] = b2.y;
#else
            const float xs[2] = {z.x, z.y};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float t = fabsf(xs[i]);
                const float m2 = fmaxf(t - 2.f * A, 0.f), m4 = fmaxf(t - 4.f * A, 0.f), m6 = fmaxf(t - 6.f * A, 0.f);
                a0[i] = flip_sign_of(c * ((t + m2) + (m4 + m6)), xs[i]);
                a1[i] = fmaf(c, m6 - m2, fmaf(c2, t, k1));
                a2[i] = fmaf(c2, m4, fmaf(-c, t, k2));
            }
#endif
            L[0] = a0[0]; L[1] = a0[1]; L[2] = a1[0]; L[3] = a1[1]; L[4] = a2[0]; L[5] = a2[1];
            bits = push_sign(push_sign(push_sign(push_sign(push_sign(push_sign(0u, z.x), z.y), a1[0]), a1[1]), a2[0]), a2[1]) ^ 15u;
        }
    }
};

// bits (MSB-first value of one bin) -> constellation point; inverse of the demapper's hard decision.
// QPSK points follow MultiAntennaSystem.py:171-178, BPSK :156-157, 16/64-QAM TS 36.211 7.1.
template <int BPS> __device__ __forceinline__ float2 map_bits(unsigned v) {
    if (BPS == 1) return make_float2(v ? 1.f : -1.f, 0.f);
    if (BPS == 2) {
        const float a = 0.70710678118654752440f;
        return make_float2((v & 2) ? -a : a, (v & 1) ? -a : a);
    }
    if (BPS == 4) {
        const float a = 0.31622776601683793320f;
        const float i = ((v & 8) ? -a : a) * ((v & 2) ? 3.f : 1.f);
        const float q = ((v & 4) ? -a : a) * ((v & 1) ? 3.f : 1.f);
        return make_float2(i, q);
    }
    const float a = 0.15430334996209191026f;
    // magnitude by (b1,b2): 00->3 01->1 10->5 11->7   ==  4 + (2*b1-1)*(2 + (2*b2-1)) ... as a 4-entry nibble table
    const unsigned mi = ((v >> 2) & 2) | ((v >> 1) & 1), mq = ((v >> 1) & 2) | (v & 1);
    const float i = (float)((0x7513u >> (4 * mi)) & 15u), q = (float)((0x7513u >> (4 * mq)) & 15u);
    return make_float2((v & 32) ? -a * i : a * i, (v & 16) ? -a * q : a * q);
}

// read BPS bits starting at bit offset `bo` of a packed MSB-first byte string
template <int BPS> __device__ __forceinline__ unsigned read_bits(const uint8_t *__restrict__ p, int bo, int nbytes) {
    const int by = bo >> 3, sh = bo & 7;
    unsigned w = (unsigned)__ldg(p + by) << 8;
    if (BPS == 6 && by + 1 < nbytes) w |= (unsigned)__ldg(p + by + 1);
    return (w >> (16 - sh - BPS)) & ((1u << BPS) - 1u);
}

// Pack the per-bin values staged in hb[] (one byte per data bin, zero-padded past ndb) into
// MSB-first bytes.  16-QAM / QPSK rows are built four output bytes at a time from one wide
// shared-memory read and stored with the widest store the row's alignment allows (a 30-byte row is
// 7 words + 2 bytes instead of 30 single-byte stores); thread i produces chunks i, i+T, ...
__device__ __forceinline__ void store_chunk(uint8_t *__restrict__ p, unsigned word, int rem) {
    const unsigned al = (unsigned)(reinterpret_cast<unsigned long long>(p) & 3ull);
    if (al == 0 && rem >= 4) {
        *reinterpret_cast<unsigned *>(p) = word;
    } else if ((al & 1u) == 0) {
        if (rem >= 2) *reinterpret_cast<uint16_t *>(p) = (uint16_t)word; else p[0] = (uint8_t)word;
        if (rem >= 4) *reinterpret_cast<uint16_t *>(p + 2) = (uint16_t)(word >> 16);
        else if (rem == 3) p[2] = (uint8_t)(word >> 16);
    } else {
        for (int i = 0; i < 4 && i < rem; ++i) p[i] = (uint8_t)(word >> (8 * i));
    }
}

// Packed row word `c` (row bytes 4c..4c+3, little-endian) from the per-bin values in hb[]; `c3` selects the
// word inside a 64-QAM 12-byte chunk.
template <int BPS> __device__ __forceinline__ unsigned pack_word(const uint8_t *hb, int c) {
    if (BPS == 4) {
        const uint2 w = *reinterpret_cast<const uint2 *>(hb + 8 * c);              // 8 bins, one nibble each
        const unsigned ux = (w.x << 4) | (w.x >> 8), uy = (w.y << 4) | (w.y >> 8);
        return __byte_perm(ux, uy, 0x6420);
    } else if (BPS == 2) {
        const uint4 w = *reinterpret_cast<const uint4 *>(hb + 16 * c);             // 16 bins, two bits each
        // (n0 | n1<<8 | n2<<16 | n3<<24) * 0x40100401 has n0<<6|n1<<4|n2<<2|n3 in its top byte
        return ((w.x * 0x40100401u) >> 24) | (((w.y * 0x40100401u) >> 24) << 8) |
               (((w.z * 0x40100401u) >> 24) << 16) | ((w.w * 0x40100401u) & 0xff000000u);
    } else {                                                                       // BPS == 1: 32 bins, one bit each
        const uint4 a = *reinterpret_cast<const uint4 *>(hb + 32 * c), b = *reinterpret_cast<const uint4 *>(hb + 32 * c + 16);
        auto nib = [](unsigned x) { return (x * 0x08040201u) >> 24; };
        return ((nib(a.x) << 4) | nib(a.y)) | (((nib(a.z) << 4) | nib(a.w)) << 8) | (((nib(b.x) << 4) | nib(b.y)) << 16) |
               (((nib(b.z) << 4) | nib(b.w)) << 24);
    }
}
// 64-QAM: 16 bins (one 128-bit read) -> four 24-bit groups -> three 32-bit words
__device__ __forceinline__ void pack_words6(const uint8_t *hb, int c, unsigned (&o)[3]) {
    const uint4 w = *reinterpret_cast<const uint4 *>(hb + 16 * c);
    auto g24 = [](unsigned x) { return ((x & 63u) << 18) | (((x >> 8) & 63u) << 12) | (((x >> 16) & 63u) << 6) | (x >> 24); };
    const unsigned v0 = g24(w.x), v1 = g24(w.y), v2 = g24(w.z), v3 = g24(w.w);           // bytes (msb..lsb) = b0 b1 b2
    o[0] = __byte_perm(v0, v1, 0x6012);
    o[1] = __byte_perm(v1, v2, 0x5601);
    o[2] = __byte_perm(v2, v3, 0x4560);
}

// Pack one symbol's hard decisions and store the row.  Rows whose length is even are stored straight
// from the packing step (32-bit or 16-bit stores depending on the row's phase).  Odd-length rows
// (15 and 45 bytes in the reference numerology) start at every byte phase, so they are packed into
// shared memory first and then written as ALIGNED 32-bit words assembled from two neighbouring packed
// words with a funnel shift; only the partial first / last word of a row falls back to byte stores.
template <int BPS, int T, int GROUPS>
__device__ __forceinline__ void pack_bits(const uint8_t *hb, unsigned *pk, uint8_t *__restrict__ out, int nbytes, int j,
                                          int gid) {
    const int nwords = (nbytes + 3) >> 2;
    if ((nbytes & 1) == 0) {
        if (BPS == 6) {
            for (int c = j; 3 * c < nwords; c += T) {
                unsigned o[3];
                pack_words6(hb, c, o);
                const int rem = nbytes - 12 * c;
                store_chunk(out + 12 * c, o[0], rem);
                if (rem > 4) store_chunk(out + 12 * c + 4, o[1], rem - 4);
                if (rem > 8) store_chunk(out + 12 * c + 8, o[2], rem - 8);
            }
        } else {
            for (int c = j; c < nwords; c += T) store_chunk(out + 4 * c, pack_word<BPS>(hb, c), nbytes - 4 * c);
        }
        return;
    }
    if (BPS == 6) {
        for (int c = j; 3 * c < nwords; c += T) {
            unsigned o[3];
            pack_words6(hb, c, o);
            pk[3 * c] = o[0];
            pk[3 * c + 1] = o[1];
            pk[3 * c + 2] = o[2];
        }
    } else {
        for (int c = j; c < nwords; c += T) pk[c] = pack_word<BPS>(hb, c);
    }
    group_sync<T, GROUPS>(gid);
    const int a = (int)(reinterpret_cast<unsigned long long>(out) & 3ull);     // row phase inside its first aligned word
    unsigned *ow = reinterpret_cast<unsigned *>(out - a);
    const int end = a + nbytes, nout = (end + 3) >> 2;
    for (int i = j; i < nout; i += T) {
        const unsigned lo = i > 0 ? pk[i - 1] : 0u, hi = i < nwords ? pk[i] : 0u;
        const unsigned word = __funnelshift_rc(lo, hi, 8 * (4 - a));           // row bytes 4i-a .. 4i-a+3
        const int t0 = a - 4 * i, t1 = end - 4 * i;                               // valid bytes of this word: [t0, t1)
        if (t0 <= 0 && t1 >= 4) {
            ow[i] = word;
        } else {
            uint8_t *pb = reinterpret_cast<uint8_t *>(ow + i);
            for (int t = (t0 > 0 ? t0 : 0); t < 4 && t < t1; ++t) pb[t] = (uint8_t)(word >> (8 * t));
        }
    }
}

// How many symbols ahead the framed receiver requests its samples (cp.async ring of that many slots).
#ifndef OFDM_PF_DEPTH
#define OFDM_PF_DEPTH 1
#endif
// Framed receiver: transforms from this size on exchange through split-phase (mbarrier) barriers; 13 = never.
#ifndef OFDM_SPLIT_MIN_LOG2N
#define OFDM_SPLIT_MIN_LOG2N 13
#endif
// ... and the same for the kernels below 64-QAM only.  B200, N=4096, same box, two alternating runs
// (profiles/r2_ab_split_phase_n4096_final.txt): QPSK 0.602 -> 0.594 ms, 16-QAM 0.674 -> 0.670 ms, 64-QAM 0.805 -> 0.810 ms.
#ifndef OFDM_SPLIT_LOWBPS_MIN_LOG2N
#define OFDM_SPLIT_LOWBPS_MIN_LOG2N 12
#endif
// ... and for the bulk-row variant (64-QAM, N >= 1024), whose 2 CTAs x 8 warps per SM leave DRAM latency exposed: three
// symbols ahead.  B200, LTE 20 MHz 64-QAM, one box, ms per launch (depth 1 / 2 / 3): 0.672 / 0.677 / 0.648 with the
// power-table twiddles, 0.697 / 0.652 / -- with chained rows; the per-lane kernels lose 2-3 % at depth 2 and stay at 1.
#ifndef OFDM_PF_DEPTH_BULK
#define OFDM_PF_DEPTH_BULK 2
#endif

// shared-memory carve-up per CTA.  TXK: transmitter (no staging, no packed-row area); PIL: room for the
// per-symbol pilot estimates; GSM: per-frame equaliser gains in thread-private slots; SB: one FFT
// exchange buffer instead of two.
template <class C, int GROUPS, bool TXK = false, bool PIL = true, bool GSM = false, bool SB = false, int PFD = OFDM_PF_DEPTH>
struct Smem {
    static constexpr int THREADS = C::T * GROUPS;
    static constexpr int NBUF = SB ? 1 : 2;
    static constexpr int GSTRIDE = (C::LOG2N == 6) ? (NBUF * C::BUF + 8) : NBUF * C::BUF;   // complex slots per group
    static constexpr int FFT_BYTES = GROUPS * GSTRIDE * 8;
    static constexpr int TW_BYTES = C::TW_ENTRIES * 8;                              // pass twiddles, shared by the CTA
    static constexpr int STAGE_BYTES = TXK ? 0 : PFD * THREADS * C::PPT * 8;   // symbols in flight: PPT samples per thread each (RX)
    static constexpr int RED_BYTES = GROUPS * 64 * 4 + GROUPS * 16 * 8 + GROUPS * 16; // 64 floats + 16 u64 + one mbarrier (16-byte slot) per group
    static constexpr int PK_BYTES = TXK ? 0 : 3 * C::N / 4 + 16;                    // packed row staged for aligned stores
    static constexpr int TX_ROW_BYTES = 3 * C::N / 4 + 16;                          // one packed payload row at any 16-byte phase (TX)
    static_assert((3 * C::N / 4 + 30) / 16 <= C::T, "one 16-byte chunk per thread must cover a payload row");
    // TX at N = 64: two 64-byte row buffers per group would put the four groups of a warp exactly 128 bytes (one bank
    // sweep) apart, a 4-way conflict on every row copy and bit fetch (ncu: 3.1 wavefronts per LDS.U8) -> skew by 16 bytes
    // (the receiver at N = 128 has two groups per warp 256 bytes apart, 2 wavefronts per hard-bit staging store, but is
    // HBM-bound: skewing it changed nothing)
    static constexpr int HB_RAW = TXK ? 2 * TX_ROW_BYTES : C::N + 16 + PK_BYTES;    // +16: zero pad, keeps 16-byte alignment
    static constexpr int HB_STRIDE = HB_RAW + ((TXK && C::T < 32 && HB_RAW % 128 == 0) ? 16 : 0);
    static constexpr int HB_BYTES = GROUPS * HB_STRIDE;
    static constexpr int LUT_BYTES = TXK ? 72 * 8 : 0;                              // constellation table (TX) + its zero entry
    static constexpr int HP_BYTES = (TXK || !PIL) ? 0 : GROUPS * (C::N / 2) * 8;    // per-symbol pilot estimates (RX, pilot mode)
    static constexpr int GS_BYTES = GSM ? GROUPS * C::N * 8 : 0;                    // equaliser gains, slot m*T + j of the group
    static constexpr int OFF_TW = FFT_BYTES;
    static constexpr int OFF_STAGE = OFF_TW + TW_BYTES;
    static constexpr int OFF_RED = OFF_STAGE + STAGE_BYTES;
    static constexpr int OFF_HB = OFF_RED + RED_BYTES;
    static constexpr int OFF_LUT = OFF_HB + ((HB_BYTES + 7) & ~7);
    static constexpr int OFF_HP = OFF_LUT + LUT_BYTES;
    static constexpr int OFF_GS = OFF_HP + HP_BYTES;
    static constexpr int BYTES = OFF_GS + GS_BYTES;
    static_assert(BYTES % 16 == 0, "the dynamic tail (LLR row staging) needs 16-byte alignment");
    static_assert(!TXK || (OFF_HB % 16 == 0 && HB_STRIDE % 16 == 0), "cp.async row buffers need 16-byte alignment");
    __device__ static float2 *buf0(unsigned char *s, int gid) { return reinterpret_cast<float2 *>(s) + gid * GSTRIDE; }
    __device__ static float2 *tws(unsigned char *s) { return reinterpret_cast<float2 *>(s + OFF_TW); }
    // slot m of thread t: stage[m * THREADS + t] -> lanes hit consecutive 8-byte words
    __device__ static float2 *stage(unsigned char *s) { return reinterpret_cast<float2 *>(s + OFF_STAGE); }
    __device__ static unsigned long long *redu(unsigned char *s, int gid) {
        return reinterpret_cast<unsigned long long *>(s + OFF_RED) + gid * 16;
    }
    __device__ static float *redf(unsigned char *s, int gid) {
        return reinterpret_cast<float *>(s + OFF_RED + GROUPS * 16 * 8) + gid * 64;
    }
    __device__ static unsigned long long *mbar(unsigned char *s, int gid) {
        return reinterpret_cast<unsigned long long *>(s + OFF_RED + GROUPS * 16 * 8 + GROUPS * 64 * 4) + 2 * gid;
    }
    __device__ static uint8_t *hb(unsigned char *s, int gid) { return s + OFF_HB + gid * HB_STRIDE; }
    __device__ static float2 *lut(unsigned char *s) { return reinterpret_cast<float2 *>(s + OFF_LUT); }
    __device__ static float2 *hp(unsigned char *s, int gid) { return reinterpret_cast<float2 *>(s + OFF_HP) + gid * (C::N / 2); }
    __device__ static float2 *gs(unsigned char *s, int gid) { return reinterpret_cast<float2 *>(s + OFF_GS) + gid * C::N; }
};

// The frame's equaliser gains: PPT registers per thread, or (GSM) the thread's own shared-memory slots
// m*T + j -- written and read by the same thread, so no barrier is involved.
template <class C, bool GSM> struct Gains {
    float2 r[GSM ? 1 : C::PPT];
    float2 *s;
    __device__ __forceinline__ float2 get(int m) const { return GSM ? s[m * C::T] : r[GSM ? 0 : m]; }
    __device__ __forceinline__ void set(int m, float2 g) {
        if (GSM) s[m * C::T] = g; else r[GSM ? 0 : m] = g;
    }
};

// Asynchronous prefetch of this thread's PPT samples of the window that opens at `start` into its
// private shared-memory slots (LDGSTS, no registers held while the data is in flight).  Samples
// outside [0, n) are zero-filled by the hardware (src-size 0).  The thread that issued the copies
// is the only one that reads them back, so cp.async.wait_group is the only synchronisation needed.
template <int T, int PPT, int THREADS>
__device__ __forceinline__ void prefetch_pts(float2 *stage, const float2 *__restrict__ x, long long start, long long n,
                                             int j, int tid) {
    constexpr int N = PPT * T;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + tid);
    if (start >= 0 && start + N <= n) {
        const float2 *p = x + start + j;
#pragma unroll
        for (int m = 0; m < PPT; ++m)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + m * THREADS * 8), "l"(p + m * T) : "memory");
    } else {
#pragma unroll
        for (int m = 0; m < PPT; ++m) {
            const long long i = start + j + m * T;
            const bool ok = (i >= 0 && i < n);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + m * THREADS * 8),
                         "l"(x + (ok ? i : 0)), "r"(ok ? 8 : 0) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int PPT, int THREADS, int PENDING = 0>
__device__ __forceinline__ void take_pts(float2 (&v)[PPT], const float2 *stage, int tid) {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#pragma unroll
    for (int m = 0; m < PPT; ++m) v[m] = stage[m * THREADS + tid];
}

// ------------------------------------------------------------------------------------------------
// Preamble handling, split in two so the framed kernel can prefetch between the pieces.
//
// sync_accumulate: after the forward FFT of synch symbol s, add this thread's bins into the
//   correlation sums  ys_k += X_k * conj(zc_k)  and the power sum (SynchAndChanEst.py:153-157).
// sync_finish: unit-power factor, then the (CP+1)-delay bank evaluated as ONE inverse FFT
//   (c[d] = sum_k ys_k e^{+j2pi dk/N} is an N-point inverse DFT sampled at d = 0..CP) instead of
//   the reference's dense (CP+1) x MM matmul, then max / first argmax over d (:160-164).
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void sync_accumulate(const DevParams &P, const float2 (&v)[C::PPT], int s, int j,
                                                float2 (&ys)[C::PPT], float2 (&ydup)[2], float &pw) {
    constexpr int T = C::T, N = C::N, PPT = C::PPT;
    const float2 *zb = P.zc_bin + s * N + j;
    const float *wb = P.w_bin + j;
#pragma unroll
    float2 pw2 = make_float2(pw, 0.f);                        // weighted power as two partial sums (re^2, im^2)
    for (int m = 0; m < PPT; ++m) {
        pw2 = cpow_acc_w(pw2, v[m], __ldg(wb + m * T));
        ys[m] = cmul_acc(ys[m], v[m], __ldg(zb + m * T));
    }
    pw = pw2.x + pw2.y;
    if (P.dup_bin >= 0) {
#pragma unroll
        for (int m = 0; m < PPT; ++m)
            if (j + m * T == P.dup_bin) {
                ydup[0] = cadd(ydup[0], cmul(v[m], __ldg(P.zc_dup + 2 * s)));
                ydup[1] = cadd(ydup[1], cmul(v[m], __ldg(P.zc_dup + 2 * s + 1)));
            }
    }
}

// c[] holds the correlation sums on entry and is overwritten by the delay bank.
template <class C, int GROUPS, bool SB = false, bool SPLIT = false>
__device__ __forceinline__ void sync_finish(const DevParams &P, float2 (&c)[C::PPT], float pw, const float2 *tws,
                                            float2 *b0, float2 *b1, float *redf, unsigned long long *redu, int j,
                                            int gid, int &pp, float &p_est, float &peak, int &delay, unsigned mbar = 0) {
    constexpr int T = C::T, PPT = C::PPT;
    pw = group_sum<T, GROUPS>(pw, redf, j, gid);
    p_est = pw > 0.f ? sqrtf((float)P.mm / pw) : 0.f;
    fft_group<C, true, GROUPS, SB, false, false, SPLIT>(c, tws, b0, b1, j, gid, pp, nullptr, mbar);
    unsigned long long key = 0ull;
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
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

// channel estimate + equaliser gain for one bin from its correlation sum.
// H_k = rot_k * p_est * ys_k / ((1+1/x) * M0)        SynchAndChanEst.py:177-188
// G_k = conj(H_k) / (|H_k|^2 + inv_reg) * rot_k      :237-246 (de-rotation folded into the gain)
template <int N>
__device__ __forceinline__ void chan_bin(const DevParams &P, float2 y, int k, float sc, int delay, float inv_reg, float2 &h,
                                         float2 &g) {
    const float2 w = __ldg(P.tw + ((delay * k) & (N - 1)));       // exp(-j*2pi*d*k/N); rot = conj(w)
    h = cscale(cmulc(y, w), sc);
    const float den = cabs2(h) + inv_reg;
    const float inv = den > 0.f ? __fdividef(1.f, den) : 0.f;
    g = cscale(cmulc(make_float2(h.x, -h.y), w), inv);
}
template <class C>
__device__ __forceinline__ void chan_core(const DevParams &P, const float2 (&ys)[C::PPT], const float2 (&ydup)[2],
                                          float p_est, int delay, float inv_reg, int j, float2 (&H)[C::PPT],
                                          float2 (&G)[C::PPT]) {
    constexpr int T = C::T, N = C::N, PPT = C::PPT;
    const float sc = p_est * P.chanest_scale;
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        const int k = j + m * T;
        chan_bin<N>(P, k == P.dup_bin ? ydup[1] : ys[m], k, sc, delay, inv_reg, H[m], G[m]);   // fancy assignment: last index wins
    }
}

// 64-QAM: a bin's six LLRs are 24 bytes at a 24-byte pitch, so three 8-byte stores per lane leave every 32-byte sector
// written by three instructions.  WIDE form: two stores per bin instead -- one 8-byte and one 16-byte, ordered by the
// parity of the bin's 8-byte slot so that the 16-byte one is aligned (even slot: {L0..L3} at q, {L4,L5} at q+4; odd
// slot: {L0,L1} at q, {L2..L5} at q+2); neighbouring lanes' 8-byte halves then share a sector.  Six selects per bin buy
// one store instruction less: B200, 64-QAM, same box (profiles/r2_ab_llr6_wide.txt): N=4096 0.470 -> 0.416 ms (one CTA
// of 192-register threads per SM, whose registers wait in the store queue), but N=64..1024 +4..9 % slower, N=2048
// +0.6 %, LTE grid 0.636 -> 0.667 ms -- so from N = 2^OFDM_LLR6_WIDE_MIN_LOG2N only.
#ifndef OFDM_LLR6_WIDE_MIN_LOG2N
#define OFDM_LLR6_WIDE_MIN_LOG2N 12
#endif
template <bool CS> __device__ __forceinline__ void store_llr6(float *q, const float (&L)[6], bool odd) {
    const float2 a = odd ? make_float2(L[0], L[1]) : make_float2(L[4], L[5]);
    const float4 b = odd ? make_float4(L[2], L[3], L[4], L[5]) : make_float4(L[0], L[1], L[2], L[3]);
    st_out<CS>(reinterpret_cast<float2 *>(q + (odd ? 0 : 4)), a);
    st_out<CS>(reinterpret_cast<float4 *>(q + (odd ? 2 : 0)), b);
}
template <bool CS, int BPS, bool WIDE> __device__ __forceinline__ void store_llr_pairs(float *q, const float (&L)[BPS], bool odd) {
    if constexpr (BPS == 6 && WIDE) {
        store_llr6<CS>(q, L, odd);
    } else {
#pragma unroll
        for (int b = 0; b + 1 < BPS; b += 2) st_out<CS>(reinterpret_cast<float2 *>(q + b), make_float2(L[b], L[b + 1]));
    }
}

// Everything a data symbol's epilogue needs that does not change from symbol to symbol.
template <class C, int BPS> struct DataCtx {
    BinMap map;
    Demapper<BPS> dem;
    float sqrt_ndb;
    __device__ __forceinline__ DataCtx(const DevParams &P, int j)
        : map(make_bin_map<C::T, C::PPT>(P.ndb, j)), dem(P.llr_scale), sqrt_ndb(sqrtf((float)P.ndb)) {}
};

// equalise + demap + store one data symbol held in v[] (FFT output).  SynchAndChanEst.py:232-248.
template <class C, int BPS, int GROUPS, bool GSM, bool DENSE = false, bool EQC = false>
__device__ __forceinline__ void data_epilogue(const DevParams &P, const DataCtx<C, BPS> &D, const float2 (&v)[C::PPT],
                                              const Gains<C, GSM> &G, long long row, float *redf, uint8_t *hb, float *ls,
                                              int j, int gid, float2 *__restrict__ eq_out, float *__restrict__ llr_out,
                                              uint8_t *__restrict__ hard_out) {
    constexpr int T = C::T, PPT = C::PPT;
    constexpr bool CS = C::LOG2N >= OFDM_RX_CS_MIN_LOG2N;
    // DENSE: every bin is equalised and demapped unconditionally and only the stores are predicated, so the PPT
    // per-bin dependency chains are one basic block the scheduler can interleave.  For grids where most bins are used
    // (chosen at launch from the fill ratio, launch_rx_frames).
    const unsigned mask = D.map.mask;
    float2 pw2 = make_float2(0.f, 0.f);                       // sum |Y|^2 as two partial sums (re^2, im^2)
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        if (DENSE) pw2 = cpow_acc_w(pw2, v[m], (mask & (1u << m)) ? 1.f : 0.f);
        else if (mask & (1u << m)) pw2 = cpow_acc(pw2, v[m]);
    }
    const float pw = group_sum<T, GROUPS>(pw2.x + pw2.y, redf, j, gid);
    const float p0 = pw > 0.f ? D.sqrt_ndb * rsqrtf(pw) : 0.f;        // sqrt(ndb / sum|Y|^2), :233
    // every output is optional: a caller that only decodes wants LLRs or hard bits, not the equalised symbols
    const bool soft = (llr_out != nullptr), hard = (hard_out != nullptr), weq = (OFDM_EQ_ALWAYS || EQC) ? true : (eq_out != nullptr);
    // 64-QAM: a bin's six LLRs are 24 bytes, so per-lane stores would leave every 32-byte sector
    // half-written by each instruction (2.9x the sector writes at L2, ncu).  The row is assembled in
    // shared memory instead and written with full 16-byte lanes.
    const bool staged = (BPS == 6) && soft && (ls != nullptr);
    float2 *eqp = eq_out + row * P.ndb + D.map.base_p;
    float2 *eqn = eq_out + row * P.ndb + (D.map.base_n - (PPT / 2) * T);
    float *lp = (staged ? ls : llr_out + row * P.ndb * BPS) + D.map.base_p * BPS;
    float *ln = (staged ? ls : llr_out + row * P.ndb * BPS) + (D.map.base_n - (PPT / 2) * T) * BPS;
    uint8_t *hp = hb + D.map.base_p, *hn = hb + (D.map.base_n - (PPT / 2) * T);
    // parity of the 8-byte slot of this thread's LLRs in each spectrum half (a bin's pitch is m*T*BPS floats, T even)
    static_assert(T % 2 == 0, "LLR slot parity is per thread");
    const bool odd_p = (reinterpret_cast<unsigned long long>(lp) & 8ull) != 0, odd_n = (reinterpret_cast<unsigned long long>(ln) & 8ull) != 0;
    (void)odd_p; (void)odd_n;
    // CH bins at a time: the chunk is skipped when no lane of the warp uses any of its bins (a warp-uniform vote, so no
    // divergence), otherwise all CH bins are computed branch-free and only the stores are predicated.  CH = PPT is the
    // fully dense form; CH = 0 keeps the per-lane branches of the loop further down.
    constexpr int CH = DENSE ? PPT : OFDM_EPI_CHUNK_SPARSE;
    // The optional outputs are folded into per-output copies of the bin mask, so a bin's store keeps ONE bit test as its
    // predicate: testing `used && wanted` per bin cost the branch-free epilogue 6-11 % (B200: N=2048 16-QAM 0.601 -> 0.540 ms,
    // N=4096 0.777 -> 0.701 ms with the equalised-symbol switch compiled out; OFDM_EPI_MASKS=0 rebuilds that form).
    const unsigned emask = weq ? mask : 0u, smask = soft ? mask : 0u, hmask = hard ? mask : 0u;
    (void)emask; (void)smask; (void)hmask;
    // EQC: the kernel variant launched when the equalised symbols ARE wanted (the default call) has that switch compiled
    // out and tests `used && wanted` for the other two -- the form ptxas schedules best at N = 512..2048 (see launch.cuh).
    // (Two copies of the block behind a uniform "all outputs wanted" branch were slower than either: N=2048 0.578 -> 0.625 ms.)
    if constexpr (CH > 0) {
#pragma unroll
        for (int c = 0; c < PPT; c += (CH > 0 ? CH : PPT)) {
            if (CH == PPT || __any_sync(0xffffffffu, ((mask >> c) & ((1u << CH) - 1u)) != 0u)) {
#pragma unroll
                for (int m = c; m < c + CH; ++m) {
                    const bool used = (mask & (1u << m)) != 0;
                    const float2 z = cscale(cmul(G.get(m), v[m]), p0);
                    if ((OFDM_EPI_MASKS >= 1 && !EQC) ? (emask & (1u << m)) != 0 : (used && weq)) st_out<CS>((m < PPT / 2 ? eqp : eqn) + m * T, z);
                    float Lb[BPS];
                    unsigned bits;
                    D.dem.run(z, Lb, bits);
                    float *q = (m < PPT / 2 ? lp : ln) + m * T * BPS;
                    if ((OFDM_EPI_MASKS >= 3 && !EQC) ? (smask & (1u << m)) != 0 : (used && soft)) {
                        if (BPS == 1) {
                            st_out<CS>(q, Lb[0]);
                        } else if (BPS == 4) {
                            st_out<CS>(reinterpret_cast<float4 *>(q), make_float4(Lb[0], Lb[1], Lb[2], Lb[3]));
                        } else if (staged) {
#pragma unroll
                            for (int b = 0; b + 1 < BPS; b += 2) *reinterpret_cast<float2 *>(q + b) = make_float2(Lb[b], Lb[b + 1]);
                        } else {
                            store_llr_pairs<CS, BPS, (C::LOG2N >= OFDM_LLR6_WIDE_MIN_LOG2N)>(q, Lb, m < PPT / 2 ? odd_p : odd_n);
                        }
                    }
                    if ((OFDM_EPI_MASKS >= 3 && !EQC) ? (hmask & (1u << m)) != 0 : (used && hard)) (m < PPT / 2 ? hp : hn)[m * T] = (uint8_t)bits;
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        if (CH == 0 && (mask & (1u << m))) {
            const float2 z = cscale(cmul(G.get(m), v[m]), p0);
            if (weq) st_out<CS>((m < PPT / 2 ? eqp : eqn) + m * T, z);
            if (soft || hard) {
                float Lb[BPS];
                unsigned bits;
                D.dem.run(z, Lb, bits);
                if (soft) {
                    float *q = (m < PPT / 2 ? lp : ln) + m * T * BPS;
                    if (BPS == 1) {
                        st_out<CS>(q, Lb[0]);
                    } else if (BPS == 4) {
                        st_out<CS>(reinterpret_cast<float4 *>(q), make_float4(Lb[0], Lb[1], Lb[2], Lb[3]));
                    } else if (staged) {
#pragma unroll
                        for (int b = 0; b + 1 < BPS; b += 2) *reinterpret_cast<float2 *>(q + b) = make_float2(Lb[b], Lb[b + 1]);
                    } else {
                        store_llr_pairs<CS, BPS, (C::LOG2N >= OFDM_LLR6_WIDE_MIN_LOG2N)>(q, Lb, m < PPT / 2 ? odd_p : odd_n);
                    }
                }
                if (hard) (m < PPT / 2 ? hp : hn)[m * T] = (uint8_t)bits;
            }
        }
    }
    if (hard || staged) group_sync<T, GROUPS>(gid);
    if (staged) {
        float4 *dst = reinterpret_cast<float4 *>(llr_out + row * P.ndb * BPS);
        const float4 *src = reinterpret_cast<const float4 *>(ls);
        const int nq = (P.ndb * BPS) >> 2;
        for (int i = j; i < nq; i += T) st_out<CS>(dst + i, src[i]);
    }
    if (hard)
        pack_bits<BPS, T, GROUPS>(hb, reinterpret_cast<unsigned *>(hb + C::N + 16), hard_out + row * P.hard_stride, P.hard_stride, j, gid);
}

// Bulk-store variant of data_epilogue: equalised symbols and LLRs go to the group's shared-memory rows (es, ls)
// and leave as two cp.async.bulk copies issued by lane 0 of the group.  The row buffers are reused every symbol:
// lane 0 first waits until the previous copies have read them, and the power reduction that follows is the
// group-wide synchronisation that keeps every other lane behind that wait.
template <class C, int BPS, int GROUPS, bool GSM>
__device__ __forceinline__ void data_epilogue_bulk(const DevParams &P, const DataCtx<C, BPS> &D, const float2 (&v)[C::PPT],
                                                   const Gains<C, GSM> &G, long long row, float *redf, uint8_t *hb,
                                                   float *ls, float2 *es, int j, int gid, float2 *__restrict__ eq_out,
                                                   float *__restrict__ llr_out, uint8_t *__restrict__ hard_out) {
    constexpr int T = C::T, PPT = C::PPT;
    const unsigned mask = D.map.mask;
    if (j == 0) bulk_wait_read();
    float2 pw2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < PPT; ++m)
        if (mask & (1u << m)) pw2 = cpow_acc(pw2, v[m]);
    const float pw = group_sum<T, GROUPS>(pw2.x + pw2.y, redf, j, gid);
    const float p0 = pw > 0.f ? D.sqrt_ndb * rsqrtf(pw) : 0.f;        // sqrt(ndb / sum|Y|^2), :233
    const bool soft = (llr_out != nullptr), hard = (hard_out != nullptr);
    float2 *eqp = es + D.map.base_p, *eqn = es + (D.map.base_n - (PPT / 2) * T);
    float *lp = ls + D.map.base_p * BPS, *ln = ls + (D.map.base_n - (PPT / 2) * T) * BPS;
    uint8_t *hp = hb + D.map.base_p, *hn = hb + (D.map.base_n - (PPT / 2) * T);
    constexpr int CH = OFDM_EPI_CHUNK_BULK;                 // as in data_epilogue
    if (CH > 0) {
#pragma unroll
        for (int c = 0; c < PPT; c += (CH > 0 ? CH : PPT)) {
            if (CH == PPT || __any_sync(0xffffffffu, ((mask >> c) & ((1u << CH) - 1u)) != 0u)) {
#pragma unroll
                for (int m = c; m < c + CH; ++m) {
                    const bool used = (mask & (1u << m)) != 0;
                    const float2 z = cscale(cmul(G.get(m), v[m]), p0);
                    if (used) (m < PPT / 2 ? eqp : eqn)[m * T] = pin(z);
                    float Lb[BPS];
                    unsigned bits;
                    D.dem.run(z, Lb, bits);
                    float *q = (m < PPT / 2 ? lp : ln) + m * T * BPS;
                    if (used && soft) {
                        if (BPS == 1) {
                            *q = Lb[0];
                        } else if (BPS == 4) {
                            *reinterpret_cast<float4 *>(q) = make_float4(Lb[0], Lb[1], Lb[2], Lb[3]);
                        } else {
#pragma unroll
                            for (int b = 0; b + 1 < BPS; b += 2) *reinterpret_cast<float2 *>(q + b) = make_float2(Lb[b], Lb[b + 1]);
                        }
                    }
                    if (used && hard) (m < PPT / 2 ? hp : hn)[m * T] = (uint8_t)bits;
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        if (CH == 0 && (mask & (1u << m))) {
            const float2 z = cscale(cmul(G.get(m), v[m]), p0);
            (m < PPT / 2 ? eqp : eqn)[m * T] = pin(z);
            if (soft || hard) {
                float Lb[BPS];
                unsigned bits;
                D.dem.run(z, Lb, bits);
                if (soft) {
                    float *q = (m < PPT / 2 ? lp : ln) + m * T * BPS;
                    if (BPS == 1) {
                        *q = Lb[0];
                    } else if (BPS == 4) {
                        *reinterpret_cast<float4 *>(q) = make_float4(Lb[0], Lb[1], Lb[2], Lb[3]);
                    } else {
#pragma unroll
                        for (int b = 0; b + 1 < BPS; b += 2) *reinterpret_cast<float2 *>(q + b) = make_float2(Lb[b], Lb[b + 1]);
                    }
                }
                if (hard) (m < PPT / 2 ? hp : hn)[m * T] = (uint8_t)bits;
            }
        }
    }
    fence_async_smem();
    group_sync<T, GROUPS>(gid);
    if (j == 0) {
        if (eq_out != nullptr) bulk_store_row(eq_out + row * P.ndb, es, (unsigned)P.ndb * 8u);
        if (soft) bulk_store_row(llr_out + row * P.ndb * BPS, ls, (unsigned)(P.ndb * BPS) * 4u);
        bulk_commit();
    }
    if (hard)
        pack_bits<BPS, T, GROUPS>(hb, reinterpret_cast<unsigned *>(hb + C::N + 16), hard_out + row * P.hard_stride, P.hard_stride, j, gid);
}

// Pilot mode (extension): G[] holds the de-rotation only.  Each data symbol is equalised with the LS
// estimate on its own comb pilots, linearly interpolated across the used-bin list and held constant
// outside the outermost pilots (the oracle's np.interp).  Off the fast path: positions are recomputed
// per bin because payload indices are no longer a fixed stride apart.
template <class C, int BPS, int GROUPS, bool GSM>
__device__ __forceinline__ void data_epilogue_pilot(const DevParams &P, const DataCtx<C, BPS> &D,
                                                    const float2 (&v)[C::PPT], const Gains<C, GSM> &R, long long row,
                                                    float *redf, uint8_t *hb, float2 *hp, int j, int gid,
                                                    float2 *__restrict__ eq_out, float *__restrict__ llr_out,
                                                    uint8_t *__restrict__ hard_out) {
    constexpr int T = C::T, PPT = C::PPT;
    const unsigned mask = D.map.mask;
    float2 pw2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < PPT; ++m)
        if (mask & (1u << m)) pw2 = cpow_acc(pw2, v[m]);
    const float pw = group_sum<T, GROUPS>(pw2.x + pw2.y, redf, j, gid);
    const float p0 = pw > 0.f ? D.sqrt_ndb * rsqrtf(pw) : 0.f;
    float2 z[C::PPT];
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        z[m] = cscale(cmul(R.get(m), v[m]), p0);
        if (mask & (1u << m)) {
            const PilotPos pp = pilot_pos(P, bin_out<T, PPT>(D.map, m));
            if (pp.is_pilot) hp[pp.ip] = cmulc(z[m], __ldg(P.pilot_tx + j + m * T));   // LS: Y * conj(p), |p| = 1
        }
    }
    group_sync<T, GROUPS>(gid);
    const bool soft = (llr_out != nullptr), hard = (hard_out != nullptr);
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        if (mask & (1u << m)) {
            const PilotPos pp = pilot_pos(P, bin_out<T, PPT>(D.map, m));
            if (!pp.is_pilot) {
                const float2 ha = hp[pp.ip], hb2 = hp[pp.ip + (pp.frac > 0.f ? 1 : 0)];
                const float2 h = make_float2(fmaf(pp.frac, hb2.x - ha.x, ha.x), fmaf(pp.frac, hb2.y - ha.y, ha.y));
                const float den = cabs2(h) + P.inv_reg_eq;
                const float inv = den > 0.f ? __fdividef(1.f, den) : 0.f;
                const float2 e = cscale(cmul(make_float2(h.x, -h.y), z[m]), inv);
                const long long o = row * P.ndd + pp.dj;
                if (eq_out != nullptr) st_out(eq_out + o, e);
                if (soft || hard) {
                    float Lb[BPS];
                    unsigned bits;
                    D.dem.run(e, Lb, bits);
                    if (soft) {
#pragma unroll
                        for (int b = 0; b < BPS; ++b) st_out(llr_out + o * BPS + b, Lb[b]);
                    }
                    if (hard) hb[pp.dj] = (uint8_t)bits;
                }
            }
        }
    }
    if (hard) {
        group_sync<T, GROUPS>(gid);
        pack_bits<BPS, T, GROUPS>(hb, reinterpret_cast<unsigned *>(hb + C::N + 16), hard_out + row * P.hard_stride, P.hard_stride, j, gid);
    }
}

#define OFDM_KERNEL_PROLOGUE_(GROUPS, ...)                                            \
    using S = Smem<C, GROUPS, __VA_ARGS__>;                                            \
    constexpr int T = C::T, PPT = C::PPT;                                              \
    extern __shared__ __align__(16) unsigned char smem_raw[];                          \
    const int tid = threadIdx.x, gid = tid / T, j = tid % T;                           \
    float2 *b0 = S::buf0(smem_raw, gid), *b1 = b0 + (S::NBUF - 1) * C::BUF;                            \
    const float2 *tws = S::tws(smem_raw);                                              \
    float2 *stage = S::stage(smem_raw);                                                \
    float *redf = S::redf(smem_raw, gid);                                              \
    unsigned long long *redu = S::redu(smem_raw, gid);                                 \
    uint8_t *hb = S::hb(smem_raw, gid);                                                \
    float2 *hp = S::hp(smem_raw, gid);                                                 \
    float *ls = P.ls_stride ? reinterpret_cast<float *>(smem_raw + S::BYTES) + gid * P.ls_stride : nullptr; \
    float2 *es = reinterpret_cast<float2 *>(smem_raw + S::BYTES + GROUPS * P.ls_stride * 4) + gid * P.es_stride; \
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(S::mbar(smem_raw, gid));   \
    if (j == 0) mbar_init(mbar, T);                      /* split-phase exchanges (fft_core.cuh, Pass SPLIT) */ \
    (void)redu; (void)redf; (void)stage; (void)hp; (void)ls; (void)es; (void)mbar;      \
    for (int i = j; i < S::HB_STRIDE; i += T) hb[i] = 0;                               \
    fill_twiddles<C>(S::tws(smem_raw), P.tw, tid, S::THREADS);                              \
    __syncthreads();                                                                   \
    const long long total_groups = (long long)gridDim.x * GROUPS;                      \
    const long long gg = (long long)blockIdx.x * GROUPS + gid;                         \
    int pp = 0;

// Resident CTAs per SM the register allocator should aim for (threads per CTA = max(T,128)).
// Receiver kernels hold the equaliser gains in registers and want ~128 (PPT 8) / ~255 (PPT 16)
// registers per thread; the transmitter fits 128 with 16 points per thread.  Overridable at build
// time for tuning sweeps, e.g.  make EXTRA="-DOFDM_MINB_SMALL=6".
#ifndef OFDM_MINB_SMALL
#define OFDM_MINB_SMALL 4
#endif
#define OFDM_KERNEL_PROLOGUE(GROUPS) OFDM_KERNEL_PROLOGUE_(GROUPS, false)

template <class C> struct MinBlocks {
    static constexpr int THREADS = C::T >= 128 ? C::T : 128;
    static constexpr int REGS = C::PPT == 16 ? 255 : 128;
    static constexpr int byregs = 65536 / (THREADS * REGS);
    static constexpr int value = C::LOG2N <= 9 ? OFDM_MINB_SMALL : (byregs < 1 ? 1 : byregs);
};
#ifndef OFDM_MINB_TX_SMALL_MAX_LOG2N
#define OFDM_MINB_TX_SMALL_MAX_LOG2N 5        // transmitter sizes up to this one ask for 5 resident CTAs (<= 102 registers); B200:
                                              // N=64 0.185 -> 0.193 ms with 5 x 89 registers, N=512 0.139 -> 0.145 ms: 4 CTAs stay
#endif
template <class C> struct MinBlocksTx {
    static constexpr int THREADS = C::T >= 128 ? C::T : 128;
    static constexpr int byregs = 65536 / (THREADS * 128);
    static constexpr int value = C::LOG2N <= OFDM_MINB_TX_SMALL_MAX_LOG2N ? 5
                                 : (C::LOG2N <= 9 ? OFDM_MINB_SMALL : (byregs < 1 ? 1 : byregs));
};

// ------------------------------------------------------------------------------------------------
// K1  framed receiver: sync + channel estimate + D data symbols per frame, fused.
//     The group walks its frames symbol by symbol; while symbol q is transformed the samples of
//     symbol q+1 (next symbol of the frame, or the preamble of the group's next frame) are in flight.
// ------------------------------------------------------------------------------------------------
// Register target of the framed receiver: gains in registers want ~128 (PPT 8) / ~255 (PPT 16) per thread;
// with the gains in shared memory 16 points per thread fit a 128..168-register budget.
#ifndef OFDM_MINB_RXF_1024
#define OFDM_MINB_RXF_1024 0
#endif
#ifndef OFDM_MINB_RXF_2048
#define OFDM_MINB_RXF_2048 0
#endif
#ifndef OFDM_MINB_RXF_4096
#define OFDM_MINB_RXF_4096 0
#endif
// The branch-free (dense) epilogue at N = 2048 with 16 points per thread gains from a third resident CTA (168 registers
// per thread, a few spilled words): B200, 1920-bin grid 16-QAM 0.627 -> 0.596 ms; the sparse LTE-grid kernel loses
// with it (64-QAM 0.645 -> 0.654 ms) and keeps two.
#ifndef OFDM_MINB_RXFD_2048
#define OFDM_MINB_RXFD_2048 3
#endif
template <class C, bool GSM, bool DENSE = false> struct MinBlocksRxF {
    static constexpr int THREADS = C::T >= 128 ? C::T : 128;
    static constexpr int REGS = (C::PPT == 16 && !GSM) ? 255 : 128;
    static constexpr int byregs = 65536 / (THREADS * REGS);
    static constexpr int forced = C::LOG2N == 10 ? OFDM_MINB_RXF_1024
                                  : (C::LOG2N == 11 ? ((DENSE && C::PPT == 16 && OFDM_MINB_RXF_2048 == 0) ? OFDM_MINB_RXFD_2048 : OFDM_MINB_RXF_2048)
                                                    : (C::LOG2N == 12 ? OFDM_MINB_RXF_4096 : 0));
    static constexpr int value = C::LOG2N <= 9 ? OFDM_MINB_SMALL : (forced > 0 ? forced : (byregs < 1 ? 1 : byregs));
};

// CFO: the plan carries a frequency-offset hypothesis bank (SynchEstAndFO.py:196,:247-277,:331-333).  The
// preamble is then evaluated once per hypothesis (window re-read through L1/L2, de-rotated, transformed,
// correlated), the hypothesis with the largest peak wins (first on ties) and its phasor row multiplies every
// data window of the frame before the FFT.
// EPI selects the data-symbol epilogue: EPI_SPARSE per-lane stores with unused bins skipped, EPI_BULK rows through the
// async proxy (data_epilogue_bulk), EPI_DENSE per-lane stores with branch-free arithmetic (data_epilogue, DENSE).
enum { EPI_SPARSE = 0, EPI_BULK = 1, EPI_DENSE = 2, EPI_DENSE_EQ = 3 };   // EPI_DENSE_EQ: dense, eq output compiled in (data_epilogue, EQC)
// CPP: the plan has a CP pattern (DevParams::cp_extra); instantiated without comb pilots and hypothesis bank only, so
// that the reference's uniform prefix pays nothing for it.
template <class C, int BPS, int GROUPS, bool PIL, bool GSM, bool SB, bool CFO, int EPI = EPI_SPARSE, bool CPP = false>
__global__ void __launch_bounds__(C::T *GROUPS, MinBlocksRxF<C, GSM, EPI >= 2>::value)
rx_frames_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first,
                 long long frame_stride, long long n_frames, int iters, int *__restrict__ delay_out,
                 float *__restrict__ peak_out, uint8_t *__restrict__ locked_out, int *__restrict__ hyp_out,
                 float2 *__restrict__ H_out, float2 *__restrict__ eq_out, float *__restrict__ llr_out,
                 uint8_t *__restrict__ hard_out) {
    constexpr int DEPTH = EPI == EPI_BULK ? OFDM_PF_DEPTH_BULK : OFDM_PF_DEPTH;
    // exchanges that wait on an mbarrier after preparing the next pass' twiddles instead of bar.sync (fft_core.cuh)
    constexpr bool SPLIT = (C::LOG2N >= OFDM_SPLIT_MIN_LOG2N || (BPS <= 4 && C::LOG2N >= OFDM_SPLIT_LOWBPS_MIN_LOG2N)) && !SB;
    OFDM_KERNEL_PROLOGUE_(GROUPS, false, PIL, GSM, SB, DEPTH)
    constexpr int N = C::N;
    const int L = N + P.cp, fs = P.m0 + P.d;
    const DataCtx<C, BPS> D(P, j);
    Gains<C, GSM> G;
    G.s = S::gs(smem_raw, gid) + j;
    const long long last_frame = n_frames - 1;
    const long long off0 = first + P.cp + (CPP ? P.cp_extra : 0) - P.sync_advance;   // symbol 0 opens a slot: the longer prefix
    long long f = gg < last_frame ? gg : last_frame;
    long long pos = off0 + f * frame_stride;
    // request cursor: runs OFDM_PF_DEPTH symbols ahead of the symbol being processed, through the same
    // (frame, symbol) sequence; frames past the end of the list repeat the last one (valid windows)
    constexpr int SLOT = S::THREADS * PPT;
    // Window of symbol s relative to the frame's first window: s*L, plus cp_extra for every slot boundary crossed (CP
    // pattern).  Kept as running offsets (wo: symbol in hand, pf_wo: request cursor) with a position-in-slot counter.
    long long pf_pos = pos;
    int pf_s = 0, pf_it = 0, pf_slot = 0, slot = 0, pf_wo = 0, pf_k = 0;
    (void)slot;
    auto request = [&]() {
        prefetch_pts<T, PPT, S::THREADS>(stage + pf_slot * SLOT, x, pf_pos + (CPP ? (long long)pf_wo : (long long)pf_s * L),
                                         n_samples, j, tid);
        pf_slot = pf_slot + 1 == DEPTH ? 0 : pf_slot + 1;
        if (CPP) {
            pf_wo += L;
            if (++pf_k == P.slot) { pf_k = 0; pf_wo += P.cp_extra; }
        }
        if (++pf_s == fs) {
            pf_s = pf_wo = pf_k = 0;
            long long fq = gg + (long long)(++pf_it) * total_groups;
            fq = fq < last_frame ? fq : last_frame;
            pf_pos = off0 + fq * frame_stride;
        }
    };
    if constexpr (DEPTH == 1) {
        prefetch_pts<T, PPT, S::THREADS>(stage, x, pos, n_samples, j, tid);
    } else {
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) request();
    }
    for (int it = 0; it < iters; ++it) {
        long long fn = gg + (long long)(it + 1) * total_groups;
        fn = fn < last_frame ? fn : last_frame;
        const long long npos = off0 + fn * frame_stride;
        float2 ys[C::PPT], ydup[2];
        float pw = 0.f;
        int hyp = 0;
#pragma unroll
        for (int m = 0; m < PPT; ++m) ys[m] = make_float2(0.f, 0.f);
        ydup[0] = ydup[1] = make_float2(0.f, 0.f);
        int wo = 0, kslot = 0;
        (void)kslot;
        for (int s = 0; s < fs; ++s) {
            float2 cur[C::PPT];
            const long long wcur = CPP ? (long long)wo : (long long)s * L;
            if (CPP) {
                wo += L;
                if (++kslot == P.slot) { kslot = 0; wo += P.cp_extra; }
            }
            if constexpr (DEPTH == 1) {
                take_pts<PPT, S::THREADS>(cur, stage, tid);
                // request the next symbol (the last request of the kernel re-reads a valid window)
                prefetch_pts<T, PPT, S::THREADS>(stage, x, (s + 1 < fs) ? pos + (CPP ? (long long)wo : (long long)(s + 1) * L) : npos,
                                                 n_samples, j, tid);
            } else {
                take_pts<PPT, S::THREADS, DEPTH - 1>(cur, stage + slot * SLOT, tid);
                slot = slot + 1 == DEPTH ? 0 : slot + 1;
                request();                               // refills the slot just emptied
            }
            if (!CFO || s >= P.m0) {                 // with a CFO bank the preamble is transformed per hypothesis below
                if (CFO) rotate_pts<T, PPT>(P, cur, hyp, pos + wcur, j);
                fft_group<C, false, GROUPS, SB, false, false, SPLIT>(cur, tws, b0, b1, j, gid, pp, nullptr, mbar);
            }
            if (s < P.m0) {
                if (!CFO) sync_accumulate<C>(P, cur, s, j, ys, ydup, pw);
                if (s == P.m0 - 1) {
                    float p_est, peak;
                    int delay;
                    if (CFO) {
                        p_est = 0.f;
                        peak = -1.f;
                        delay = 0;
                        // pass h < n_cfo scores hypothesis h; the final pass rebuilds the winner's correlation sums
                        for (int h = 0; h <= P.n_cfo; ++h) {
                            const int hh = h < P.n_cfo ? h : hyp;
#pragma unroll
                            for (int m = 0; m < PPT; ++m) ys[m] = make_float2(0.f, 0.f);
                            ydup[0] = ydup[1] = make_float2(0.f, 0.f);
                            pw = 0.f;
                            for (int s2 = 0; s2 < P.m0; ++s2) {
                                float2 v[C::PPT];
                                load_pts<T, PPT>(v, x, pos + (long long)s2 * L, n_samples, j);
                                rotate_pts<T, PPT>(P, v, hh, pos + (long long)s2 * L, j);
                                fft_group<C, false, GROUPS, SB, false, false, SPLIT>(v, tws, b0, b1, j, gid, pp, nullptr, mbar);
                                sync_accumulate<C>(P, v, s2, j, ys, ydup, pw);
                            }
                            if (h < P.n_cfo) {
                                float pe, pk;
                                int dl;
                                sync_finish<C, GROUPS, SB, SPLIT>(P, ys, pw, tws, b0, b1, redf, redu, j, gid, pp, pe, pk, dl, mbar);
                                if (pk > peak) {
                                    peak = pk;
                                    p_est = pe;
                                    delay = dl;
                                    hyp = h;
                                }
                            }
                        }
                    }
                    // the correlation sums are needed again after the delay bank has consumed them: park them
                    // in the gain slots (GSM) or keep a register copy
                    float2 keep[GSM ? 1 : C::PPT];
#pragma unroll
                    for (int m = 0; m < PPT; ++m) {
                        if (GSM) G.s[m * T] = ys[m]; else keep[GSM ? 0 : m] = ys[m];
                    }
                    if (!CFO) sync_finish<C, GROUPS, SB, SPLIT>(P, ys, pw, tws, b0, b1, redf, redu, j, gid, pp, p_est, peak, delay, mbar);
                    const float sc = p_est * P.chanest_scale;
#pragma unroll
                    for (int m = 0; m < PPT; ++m) {
                        const int k = j + m * T;
                        float2 y = GSM ? G.s[m * T] : keep[GSM ? 0 : m];
                        if (k == P.dup_bin) y = ydup[1];                      // numpy fancy assignment: last index wins
                        float2 h, g;
                        chan_bin<N>(P, y, k, sc, delay, P.inv_reg_eq, h, g);
                        if (PIL) {                               // pilot mode equalises per symbol: keep only the rotation
                            const float2 w = __ldg(P.tw + ((delay * k) & (N - 1)));
                            g = make_float2(w.x, -w.y);
                        }
                        G.set(m, g);
                        if (H_out != nullptr) H_out[f * N + k] = h;
                    }
                    if (j == 0) {
                        delay_out[f] = delay;
                        peak_out[f] = peak;
                        locked_out[f] = peak > P.gate_thresh ? 1 : 0;
                        if (CFO && hyp_out != nullptr) hyp_out[f] = hyp;
                    }
                }
            } else if (PIL) {
                data_epilogue_pilot<C, BPS, GROUPS, GSM>(P, D, cur, G, f * P.d + (s - P.m0), redf, hb, hp, j, gid, eq_out,
                                                        llr_out, hard_out);
            } else if (EPI == EPI_BULK) {
                data_epilogue_bulk<C, BPS, GROUPS, GSM>(P, D, cur, G, f * P.d + (s - P.m0), redf, hb, ls, es, j, gid, eq_out,
                                                       llr_out, hard_out);
            } else {
                data_epilogue<C, BPS, GROUPS, GSM, EPI >= EPI_DENSE, EPI == EPI_DENSE_EQ>(P, D, cur, G, f * P.d + (s - P.m0), redf, hb, ls, j, gid,
                                                                     eq_out, llr_out, hard_out);
            }
        }
        f = fn;
        pos = npos;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (EPI == EPI_BULK && j == 0) bulk_wait_all();             // the last rows must have left shared memory before the CTA retires
}

// ------------------------------------------------------------------------------------------------
// K2  sliding search: one trial per group iteration.  Outputs max|c| and argmax per trial.
// ------------------------------------------------------------------------------------------------
template <class C, int GROUPS>
__global__ void __launch_bounds__(C::T *GROUPS, MinBlocks<C>::value)
rx_sync_search_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples, long long first_start,
                      long long stride, long long n_trials, int iters, float *__restrict__ peak_out,
                      int *__restrict__ delay_out, int *__restrict__ hyp_out) {
    OFDM_KERNEL_PROLOGUE(GROUPS)
    constexpr int N = C::N;
    const int L = N + P.cp;
    const int n_hyp = P.n_cfo > 0 ? P.n_cfo : 1;                 // with a CFO bank: every hypothesis, best peak wins
    for (int it = 0; it < iters; ++it) {
        long long t = gg + (long long)it * total_groups;
        t = t < n_trials - 1 ? t : n_trials - 1;
        const long long pos = first_start + t * stride;
        float best_peak = -1.f;
        int best_delay = 0, best_hyp = 0;
        for (int h = 0; h < n_hyp; ++h) {
            float2 ys[C::PPT], ydup[2];
            float pw = 0.f, p_est, peak;
            int delay;
#pragma unroll
            for (int m = 0; m < PPT; ++m) ys[m] = make_float2(0.f, 0.f);
            ydup[0] = ydup[1] = make_float2(0.f, 0.f);
            for (int s = 0; s < P.m0; ++s) {
                float2 v[C::PPT];
                load_pts<T, PPT>(v, x, pos + (long long)s * L, n_samples, j);
                if (P.n_cfo > 0) rotate_pts<T, PPT>(P, v, h, pos + (long long)s * L, j);
                fft_group<C, false, GROUPS>(v, tws, b0, b1, j, gid, pp);
                sync_accumulate<C>(P, v, s, j, ys, ydup, pw);
            }
            sync_finish<C, GROUPS>(P, ys, pw, tws, b0, b1, redf, redu, j, gid, pp, p_est, peak, delay);   // consumes ys
            if (peak > best_peak) {                              // first maximum, like np.argmax (SynchEstAndFO.py:274-277)
                best_peak = peak;
                best_delay = delay;
                best_hyp = h;
            }
        }
        if (j == 0) {
            peak_out[t] = best_peak;
            delay_out[t] = best_delay;
            if (hyp_out != nullptr) hyp_out[t] = best_hyp;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3  channel estimate for accepted locks (+ optional time-domain channel and equalised preamble).
// ------------------------------------------------------------------------------------------------
template <class C, int GROUPS>
__global__ void __launch_bounds__(C::T *GROUPS, MinBlocks<C>::value)
rx_chanest_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples,
                  const long long *__restrict__ pos_in, const int *__restrict__ delay_in,
                  const int *__restrict__ hyp_in, long long n_locks, int iters, float2 *__restrict__ H_out,
                  float2 *__restrict__ htime_out, float2 *__restrict__ synch_eq_out,
                  const int *__restrict__ row_in, const int *__restrict__ count_dev) {
    // count_dev (block path, ofdm_rx_block_work): the number of listed locks is decided by a kernel earlier in the
    // stream, so the launch is sized for the most there can be and the actual count is read here; row_in: the row of
    // H_out each lock's estimate belongs to (the block's est_chan_freq_P state) instead of its list position
    if (count_dev != nullptr) {
        n_locks = *count_dev < n_locks ? *count_dev : n_locks;
        if (n_locks <= 0) return;
    }
    OFDM_KERNEL_PROLOGUE(GROUPS)
    constexpr int N = C::N;
    const int L = N + P.cp;
    for (int it = 0; it < iters; ++it) {
        long long i = gg + (long long)it * total_groups;
        i = i < n_locks - 1 ? i : n_locks - 1;
        const long long hrow = row_in != nullptr ? (long long)row_in[i] : i;
        const long long pos = pos_in[i];
        const int delay = delay_in[i];
        const int hyp = (hyp_in != nullptr && P.n_cfo > 0) ? hyp_in[i] : -1;
        float2 ys[C::PPT], ydup[2], H[C::PPT], G[C::PPT];
        float pw = 0.f, p_est, peak;
        int dfound;
#pragma unroll
        for (int m = 0; m < PPT; ++m) ys[m] = make_float2(0.f, 0.f);
        ydup[0] = ydup[1] = make_float2(0.f, 0.f);
        for (int s = 0; s < P.m0; ++s) {
            float2 v[C::PPT];
            load_pts<T, PPT>(v, x, pos + (long long)s * L, n_samples, j);
            if (hyp >= 0) rotate_pts<T, PPT>(P, v, hyp, pos + (long long)s * L, j);
            fft_group<C, false, GROUPS>(v, tws, b0, b1, j, gid, pp);
            sync_accumulate<C>(P, v, s, j, ys, ydup, pw);
        }
        {
            float2 c[C::PPT];
#pragma unroll
            for (int m = 0; m < PPT; ++m) c[m] = ys[m];
            sync_finish<C, GROUPS>(P, c, pw, tws, b0, b1, redf, redu, j, gid, pp, p_est, peak, dfound);
        }
        chan_core<C>(P, ys, ydup, p_est, delay, P.inv_reg_eq_synch, j, H, G);
#pragma unroll
        for (int m = 0; m < PPT; ++m) H_out[hrow * N + j + m * T] = H[m];
        if (htime_out != nullptr) {                              // ifft(chan_est1): SynchAndChanEst.py:202
            float2 c[C::PPT];
#pragma unroll
            for (int m = 0; m < PPT; ++m) c[m] = H[m];
            fft_group<C, true, GROUPS>(c, tws, b0, b1, j, gid, pp);
#pragma unroll
            for (int m = 0; m < PPT; ++m) htime_out[i * N + j + m * T] = cscale(c[m], 1.f / (float)N);
        }
        if (synch_eq_out != nullptr) {                           // eq_gain_ext * data_recov: :213-218
            for (int s = 0; s < P.m0; ++s) {
                float2 v[C::PPT];
                load_pts<T, PPT>(v, x, pos + (long long)s * L, n_samples, j);
                if (hyp >= 0) rotate_pts<T, PPT>(P, v, hyp, pos + (long long)s * L, j);
                fft_group<C, false, GROUPS>(v, tws, b0, b1, j, gid, pp);
#pragma unroll
                for (int m = 0; m < PPT; ++m) {
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

// ------------------------------------------------------------------------------------------------
// K4  data symbols from an explicit symbol list (used by the stream blocks' work()).
// ------------------------------------------------------------------------------------------------
template <class C, int BPS, int GROUPS, bool PIL>
__global__ void __launch_bounds__(C::T *GROUPS, MinBlocks<C>::value)
rx_demod_kernel(const DevParams P, const float2 *__restrict__ x, long long n_samples,
                const long long *__restrict__ start_in, const int *__restrict__ delay_in,
                const int *__restrict__ hrow_in, const int *__restrict__ hyp_in, const float2 *__restrict__ Hmat,
                long long n_sym, int iters,
                float2 *__restrict__ eq_out, float *__restrict__ llr_out, uint8_t *__restrict__ hard_out,
                const int *__restrict__ count_dev) {
    if (count_dev != nullptr) {                              // see rx_chanest_kernel
        n_sym = *count_dev < n_sym ? *count_dev : n_sym;
        if (n_sym <= 0) return;
    }
    OFDM_KERNEL_PROLOGUE(GROUPS)
    constexpr int N = C::N;
    const DataCtx<C, BPS> D(P, j);
    for (int it = 0; it < iters; ++it) {
        long long i = gg + (long long)it * total_groups;
        i = i < n_sym - 1 ? i : n_sym - 1;
        const int delay = delay_in[i];
        const long long hrow = hrow_in[i];
        float2 v[C::PPT];
        Gains<C, false> G;
        load_pts<T, PPT>(v, x, start_in[i], n_samples, j);
        if (hyp_in != nullptr && P.n_cfo > 0) rotate_pts<T, PPT>(P, v, hyp_in[i], start_in[i], j);
#pragma unroll
        for (int m = 0; m < PPT; ++m) {
            const int k = j + m * T;
            const float2 h = __ldg(Hmat + hrow * N + k);
            const float2 w = __ldg(P.tw + ((delay * k) & (N - 1)));
            const float den = cabs2(h) + P.inv_reg_eq;
            G.r[m] = PIL ? make_float2(w.x, -w.y)
                         : cscale(cmulc(make_float2(h.x, -h.y), w), den > 0.f ? __fdividef(1.f, den) : 0.f);
        }
        fft_group<C, false, GROUPS>(v, tws, b0, b1, j, gid, pp);
        if (PIL)
            data_epilogue_pilot<C, BPS, GROUPS, false>(P, D, v, G, i, redf, hb, hp, j, gid, eq_out, llr_out, hard_out);
        else
            data_epilogue<C, BPS, GROUPS, false>(P, D, v, G, i, redf, hb, ls, j, gid, eq_out, llr_out, hard_out);
    }
}

// ------------------------------------------------------------------------------------------------
// K5  transmitter: one OFDM symbol per group iteration.
//   bits -> QAM into data bins / ZC into synch bins (MultiAntennaSystem.py:135-183)
//   -> IFFT -> CP -> unit energy per sample -> divide by sqrt(var)   (:189-219)
// Both normalisations are scale factors derived from two sums over the L = N+CP output samples
// (CP samples count twice), so the symbol is written to HBM once, already normalised.
// ------------------------------------------------------------------------------------------------
// transmitter sizes from this one on exchange through split-phase (mbarrier) barriers; 13 = never (A/B in DESIGN.md 3.11)
#ifndef OFDM_SPLIT_TX_MIN_LOG2N
#define OFDM_SPLIT_TX_MIN_LOG2N 11    // B200: LTE TX 0.317 -> 0.314 ms, N=4096 (16 points per thread) 0.197 -> 0.193 ms; N=1024 0.148 -> 0.150 ms
#endif
#ifndef OFDM_TX_MIDZERO
#define OFDM_TX_MIDZERO 1                 // see tx_kernel: guard band compiled out for grids below 5/8 fill (LTE)
#endif
#ifndef OFDM_TX_LUT32
#define OFDM_TX_LUT32 1                   // B200: +2 % at N=64, +2.5 % at N=2048 (0 = float2 entries)
#endif
#ifndef OFDM_TX_TWREG_MAX_LOG2N
#define OFDM_TX_TWREG_MAX_LOG2N 7      // B200: +3 % at N=64, +1.5 % at N=128, nothing from N=256 on
#endif
template <class C, int BPS, int GROUPS, bool PIL, bool CPP = false>      // CPP: CP pattern, see rx_frames_kernel
__global__ void __launch_bounds__(C::T *GROUPS, MinBlocksTx<C>::value)
tx_kernel(const DevParams P, const uint8_t *__restrict__ bits, float2 *__restrict__ out, long long n_sym_total,
          int iters) {
    OFDM_KERNEL_PROLOGUE_(GROUPS, true, false)
    constexpr int N = C::N;
    // LUT entry that maps unused bins to 0: right behind the constellation, i.e. in a bank of its own below 64-QAM
    // (entry 64 shared bank 0 with point 0: a 2-way conflict on every lookup that mixes guard and data bins)
    constexpr int ZERO = 1 << BPS;
    const int L = N + P.cp, fs = P.m0 + P.d;
    // constellation table in shared memory: one LDS per bin instead of arithmetic on the bits
    // (bank-private copies of the table were tried: the 2-way conflicts of random lookups cost less than the
    // larger table's index arithmetic -- 2-3 % slower at every size, B200)
    float2 *lut = S::lut(smem_raw);
    // LUT32: entries are the integer levels (+-1, +-3, ..) as two halves -- 4 bytes per lookup instead of 8, so a warp's
    // lookup is one shared-memory wavefront less.  The levels are exact in fp16 and the common amplitude cancels in the
    // symbol's power normalisation; only the comb-pilot variant, whose pilots come at true scale, multiplies it back.
    constexpr bool LUT32 = OFDM_TX_LUT32 != 0;
    constexpr float AMP = BPS == 1 ? 1.f : (BPS == 2 ? 0.70710678118654752440f
                                                     : (BPS == 4 ? 0.31622776601683793320f : 0.15430334996209191026f));
    unsigned *lut32 = reinterpret_cast<unsigned *>(lut);
    for (int i = tid; i <= ZERO; i += S::THREADS) {
        const float2 c = i < (1 << BPS) ? map_bits<BPS>((unsigned)i) : make_float2(0.f, 0.f);
        if (LUT32) {
            const __half2 h = __floats2half2_rn(rintf(c.x / AMP), rintf(c.y / AMP));
            lut32[i] = *reinterpret_cast<const unsigned *>(&h);
        } else {
            lut[i] = c;
        }
    }
    __syncthreads();
    auto look = [&](unsigned idx) -> float2 {
        if (!LUT32) return lut[idx];
        const unsigned e = lut32[idx];
        const float2 c = __half22float2(*reinterpret_cast<const __half2 *>(&e));
        return PIL ? cscale(c, AMP) : c;
    };
    // the last pass' twiddles live in registers where the transmitter has them to spare (N <= OFDM_TX_TWREG_MAX_LOG2N)
    constexpr bool TWR = C::LOG2N <= OFDM_TX_TWREG_MAX_LOG2N && C::CH == 0;
    float2 twreg[LastPassTw<C>::COUNT];
    if constexpr (TWR) LastPassTw<C>::load(twreg, tws, j);
    const BinMap map = make_bin_map<T, PPT>(P.ndb, j);
    // bit position of this thread's first bin of each spectrum half inside a payload row; the other
    // bins of the half are T*BPS bits (a whole number of bytes) further on
    const int bop = map.base_p * BPS, bon = (map.base_n - (PPT / 2) * T) * BPS;
    const int byp = bop >> 3, shp = 16 - (bop & 7) - BPS, byn = bon >> 3, shn = 16 - (bon & 7) - BPS;
    // this thread's samples n = j + m*T that also go into the cyclic prefix (n >= N - CP)
    // (tested as j >= ncp0 - m*T where it is needed: one compare against a uniform value per sample; a per-thread bit
    // mask made the compiler keep sixteen flag registers and rebuild the mask every symbol)
    // symbol length entering the two normalisations: N + cp, or N + cp + cp_extra for the first symbol of a slot
    const float sqrtL0 = sqrtf((float)L), invL0 = 1.f / (float)L;
    const float sqrtL1 = sqrtf((float)(L + P.cp_extra)), invL1 = 1.f / (float)(L + P.cp_extra);
    (void)sqrtL1; (void)invL1;
    const float emin = 1e-30f * (float)N * (float)N;
    // The payload buffer, viewed from its 4-byte-aligned base.  A symbol's packed row is copied word by word into one
    // of the group's two row buffers with cp.async while the previous symbol is built: no registers and no scoreboard
    // are held across the iteration (the register version stalled 12 % of the kernel on a false dependency between
    // the staging store and the next row's loads), and the copy's src-size zero-fills the bytes past the caller's
    // buffer, so nothing outside it is touched (16-byte chunks from the buffer's 16-byte-aligned base, one per thread).
    const long long mis = (long long)(reinterpret_cast<unsigned long long>(bits) & 15ull);
    const uint8_t *abits = bits - mis;
    const long long limit = mis + (n_sym_total / fs) * P.d * (long long)P.hard_stride;
    const int nchunks = (P.hard_stride + 15 + 15) / 16;      // 16-byte chunks that can hold a row at any byte phase (<= T)
    constexpr int RB = S::TX_ROW_BYTES;
    const unsigned hb_s = (unsigned)__cvta_generic_to_shared(hb);
    // The group's work items are the symbols q = gg, gg + TG, ... (q counts symbols over all frames; items past the end
    // repeat the last symbol).  Symbol type s = q mod fs, payload row offset r = mis + (q - (f+1)*m0) * hard_stride and
    // the output offset are advanced incrementally: no 64-bit multiply or divide inside the loop.
    const long long TG = total_groups, last = n_sym_total - 1;
    const int step_s = (int)(TG % fs);
    const long long dr = (long long)P.hard_stride * (TG - (TG / fs) * P.m0), dr_wrap = (long long)P.hard_stride * P.m0;
    const long long dout = TG * L;
    auto locate = [&](long long qq, int &ss, long long &rr) {
        const long long ff = qq / fs;
        ss = (int)(qq - ff * fs);
        rr = mis + (qq - (ff + 1) * P.m0) * (long long)P.hard_stride;
    };
    int s_last;
    long long r_last;
    locate(last, s_last, r_last);
    long long q = gg < last ? gg : last, r;
    int s;
    locate(q, s, r);
    long long oo = q * L + P.cp + j;                          // first useful sample of the symbol in `out` (uniform prefix)
    // CP pattern: the frame index is needed as well (every slot before the symbol shifts it by cp_extra samples)
    const long long step_f = CPP ? TG / fs : 0, f_last = CPP ? last / fs : 0;
    long long f = CPP ? q / fs : 0;
    const int nslots = CPP ? (fs + P.slot - 1) / P.slot : 0;
    auto request = [&](long long r, int buf) {               // synch symbols (r < 0) carry no payload: nothing to copy
        if (r >= 0) {
            const long long base = r & ~15ll, left = limit - base;            // left > 0: the row starts inside the buffer
            const int room = left < (1 << 30) ? (int)left : (1 << 30);          // bytes from the row's first chunk to the end
            if (j < nchunks) {                                                  // one chunk per thread covers any row
                const int sz = min(max(room - 16 * j, 0), 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(hb_s + buf * RB + 16 * j),
                             "l"(abits + base + (sz ? 16 * j : 0)), "r"(sz) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cb = 0;
    request(s >= P.m0 ? r : -1, 0);
    for (int it = 0; it < iters; ++it) {
        // next work item; its payload row is requested as soon as the current one has landed
        long long qn = q + TG, rn = r + dr, on = oo + dout, fn = f + step_f;
        int sn = s + step_s;
        if (sn >= fs) { sn -= fs; rn -= dr_wrap; if (CPP) ++fn; }
        if (qn > last) { qn = last; sn = s_last; rn = r_last; on = last * L + P.cp + j; fn = f_last; }
        // this symbol's prefix length and its extra displacement under a CP pattern
        int ncp0 = N - P.cp;
        long long xo = 0;
        float sqrtL = sqrtL0, invL = invL0;
        if (CPP) {
            const int sl = s / P.slot;
            xo = (long long)P.cp_extra * (f * nslots + sl + 1);
            if (s == sl * P.slot) { ncp0 -= P.cp_extra; sqrtL = sqrtL1; invL = invL1; }
        }
        float2 v[C::PPT];
        {
            // Every thread picks the bits of its own bins from the staged row.  Groups that share a warp can be on
            // different symbol types, so the barrier is executed unconditionally and synch symbols simply overwrite
            // the result with their Zadoff-Chu bins.  The barrier also orders the previous symbol's reads of the
            // other row buffer before the copies that refill it.
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            group_sync<T, GROUPS>(gid);
            request(sn >= P.m0 ? rn : -1, cb ^ 1);
            const uint8_t *rp = hb + cb * RB + (int)(r & 15);
            cb ^= 1;
            if (PIL) {
#pragma unroll
                for (int m = 0; m < PPT; ++m) {
                    const bool used = (map.mask & (1u << m)) != 0;
                    const PilotPos pq = pilot_pos(P, used ? bin_out<T, PPT>(map, m) : 0);
                    const int bo = pq.dj * BPS;
                    const uint8_t *q = rp + (bo >> 3);
                    unsigned w = (unsigned)q[0] << 8;
                    if (BPS == 6) w |= (unsigned)q[1];
                    const unsigned val = (w >> (16 - (bo & 7) - BPS)) & ((1u << BPS) - 1u);
                    v[m] = !used ? make_float2(0.f, 0.f) : (pq.is_pilot ? __ldg(P.pilot_tx + j + m * T) : look(val));
                }
                if (s < P.m0) {
                    const float2 *z = P.zc_tx + s * N + j;
#pragma unroll
                    for (int m = 0; m < PPT; ++m) v[m] = __ldg(z + m * T);
                }
                fft_group<C, true, GROUPS, false, TWR>(v, tws, b0, b1, j, gid, pp, twreg);
            } else {
                // Bits -> points -> inverse transform.  MZ ("middle zero"): a data symbol whose grid fills less than
                // 5/8 of the transform (every LTE bandwidth) has bins 5T..11T-1 empty for every thread, i.e. inputs
                // m = 5..10 of the first radix-16 pass are zero: their fetch + lookup and 12 of that pass' 32 first-layer
                // additions are compiled out.  The choice is uniform per symbol group, and groups are whole warps here.
                auto front = [&](auto mz) {
                    constexpr bool MZ = decltype(mz)::value;
#pragma unroll
                    for (int m = 0; m < PPT; ++m) {
                        if (MZ && m >= 5 && m < 11) {
                            v[m] = make_float2(0.f, 0.f);
                            continue;
                        }
                        // (skipping the fetch + lookup for single guard bins was tried, per thread and per row with a
                        // branch on kernel parameters only: 4-19 % slower either way -- the straight-line form lets all
                        // loads issue together)
                        const uint8_t *q = rp + (m < PPT / 2 ? byp : byn) + m * (T * BPS / 8);
                        unsigned val;
                        if (BPS == 6) {                          // a 6-bit field can straddle two bytes
                            const unsigned w = ((unsigned)q[0] << 8) | (unsigned)q[1];
                            val = (w >> (m < PPT / 2 ? shp : shn)) & 63u;
                        } else {
                            val = ((unsigned)q[0] >> ((m < PPT / 2 ? shp : shn) - 8)) & ((1u << BPS) - 1u);
                        }
                        v[m] = look((map.mask & (1u << m)) ? val : ZERO);
                    }
                    if (!MZ && s < P.m0) {
                        const float2 *z = P.zc_tx + s * N + j;
#pragma unroll
                        for (int m = 0; m < PPT; ++m) v[m] = __ldg(z + m * T);
                    }
                    fft_group<C, true, GROUPS, false, TWR, MZ, (C::LOG2N >= OFDM_SPLIT_TX_MIN_LOG2N)>(v, tws, b0, b1, j, gid, pp, twreg, mbar);
                };
                constexpr bool MZ_OK = OFDM_TX_MIDZERO != 0 && PPT == 16 && T >= 32;
                if constexpr (MZ_OK) {
                    if (s >= P.m0 && P.ndb < 10 * T) front(std::true_type{});
                    else front(std::false_type{});
                } else {
                    front(std::false_type{});
                }
            }
        }
        // Power / mean sums with the CP samples counted twice, one scale, stores with the CP copy.  With CP <= N/4 (every
        // numerology in use) only the last quarter of a thread's samples can repeat in the prefix, so the common case
        // is compiled without the per-sample compare, weight select and predicated second store for the other three
        // quarters; the branch is on a kernel parameter, hence uniform.
        auto finish = [&](auto quarter) {
            constexpr int MCP = decltype(quarter)::value ? 3 * PPT / 4 : 0;      // first m whose sample can be in the CP
            float2 pw2 = make_float2(0.f, 0.f), mean2 = make_float2(0.f, 0.f);      // (re^2, im^2) partial sums; sum of the prefix copies
#pragma unroll
            for (int m = 0; m < PPT; ++m) {
                // sum of t: the N body samples add up to N * X[0], and the DC bin is never used (data, pilots and the
                // Zadoff-Chu bins all skip it), so only the prefix copies contribute to the symbol's mean
                if (m < MCP) {
                    pw2 = cpow_acc(pw2, v[m]);
                } else {
                    const float c = (j >= ncp0 - m * T) ? 1.f : 0.f;
                    pw2 = cpow_acc_w(pw2, v[m], c + 1.f);
                    mean2 = caxpy(mean2, c, v[m]);
                }
            }
            float3 acc = make_float3(pw2.x + pw2.y, mean2.x, mean2.y);
            acc = group_sum3<T, GROUPS>(acc, redf, j, gid);
            // v holds N*ifft(X) = N*t.  Reference: t -> t*sqrt(L/sum|t|^2) -> /sqrt(var); with the first
            // factor applied the mean power is 1, so var = 1 - |mean|^2 and both steps are one scale g.
            const float s1 = acc.x > emin ? sqrtL * rsqrtf(acc.x) : 1.f / (float)N;
            const float mx = s1 * acc.y * invL, my = s1 * acc.z * invL;
            const float var = s1 * s1 * acc.x * invL - (mx * mx + my * my);
            const float g = var > 0.f ? s1 * rsqrtf(var) : 0.f;
            float2 *o = out + oo + xo;
            float2 *ocp = o - N;                              // sample n >= N-CP is also output n - (N-CP)
#pragma unroll
            for (int m = 0; m < PPT; ++m) {
                const float2 z = cscale(v[m], g);
                st_out(o + m * T, z);
                if (m >= MCP && j >= ncp0 - m * T) st_out(ocp + m * T, z);
            }
        };
        if (4 * (P.cp + (CPP ? P.cp_extra : 0)) <= N) finish(std::true_type{}); else finish(std::false_type{});
        q = qn;
        s = sn;
        r = rn;
        oo = on;
        f = fn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace ofdm
