// fft_core.cuh -- batched power-of-two FFT for one "symbol group" of T = N/8 threads.
//
// Every thread keeps 8 complex points in registers.  The transform is a Stockham autosort
// sequence of radix-8 passes followed by at most one radix-4 or radix-2 pass; between passes the
// group exchanges data through shared memory (two ping-pong buffers, one group barrier per
// exchange).  Invariants used by every kernel in this library:
//
//   * before each pass and after the last one, thread j holds element  j + m*T  in v[m], m=0..7.
//     Inputs can therefore be loaded straight from HBM with warp-contiguous 8-byte accesses and
//     the output bins come out warp-contiguous too -- no staging pass on either side;
//   * pass twiddles live in a per-CTA shared-memory table laid out [pass][q][j mod Ns], so lanes
//     read consecutive (or broadcast) entries and no registers are pinned between symbols;
//   * the exchange layouts are chosen per pass so that both the scattered write and the
//     contiguous read are shared-memory bank-conflict free for 8-byte accesses.
//
// The arithmetic being replaced is np.fft.fft / np.fft.ifft at
// GRR/gr-utsa_ofdm/python/SynchAndChanEst.py:152,:202,:230 and
// GRR/LEGACY/gr-ofdm-rx/python/txrx_mod/MultiAntennaSystem.py:199.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ofdm {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 mul_mi(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// pass structure: N8 radix-8 passes, then one radix-4 (LOG2N%3==2) or radix-2 (LOG2N%3==1) pass
__host__ __device__ constexpr int fft_radix(int log2n, int p) { return p < log2n / 3 ? 8 : (1 << (log2n % 3)); }
__host__ __device__ constexpr int fft_ns(int log2n, int p) { return 1 << (3 * (p < log2n / 3 ? p : log2n / 3)); }
// shared twiddle table: pass p >= 1 holds (R-1) rows of Ns entries, row q-1 entry r = W_N^(q*r*N/(Ns*R))
__host__ __device__ constexpr int fft_tw_off(int log2n, int p) {
    int o = 0;
    for (int i = 1; i < p; ++i) o += (fft_radix(log2n, i) - 1) * fft_ns(log2n, i);
    return o;
}

template <int LOG2N> struct FftCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / 8;                       // threads per symbol group
    static constexpr int N8 = LOG2N / 3;                  // number of radix-8 passes
    static constexpr int REM = LOG2N % 3;                 // 0: none, 1: radix-2, 2: radix-4 last pass
    static constexpr int NPASS = N8 + (REM ? 1 : 0);
    static constexpr int NEXCH = NPASS - 1;               // shared-memory exchanges per transform
    static constexpr int TW_ENTRIES = fft_tw_off(LOG2N, NPASS);
    // padded buffer length in complex slots; N == 64 uses a pad-per-8 layout (see lay())
    static constexpr int BUF = (LOG2N == 6) ? 72 : N;
    __host__ __device__ static constexpr int radix(int p) { return fft_radix(LOG2N, p); }
    __host__ __device__ static constexpr int ns(int p) { return fft_ns(LOG2N, p); }
    __host__ __device__ static constexpr int tw_off(int p) { return fft_tw_off(LOG2N, p); }
};

// Exchange layout for the buffer written by pass P (Ns = 8^P):
//   N == 64            : i + (i >> 3)             (two 8-lane groups share a half-warp; base stride 72)
//   Ns == 1  (P == 0)  : i ^ ((i >> 4) & 7)       stride-8 scatter -> distinct bank pairs
//   Ns == 8  (P == 1)  : i ^ ((i >> 3) & 8)       bit 6 folded into bit 3
//   Ns >= 64           : identity (writes are already lane-contiguous)
template <int LOG2N, int P> __device__ __forceinline__ int lay(int i) {
    if (LOG2N == 6) return i + (i >> 3);
    if (P == 0) return i ^ ((i >> 4) & 7);
    if (P == 1) return i ^ ((i >> 3) & 8);
    return i;
}

template <bool INV> __device__ __forceinline__ void bfly2(float2 &a, float2 &b) {
    float2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

// 4-point DFT, natural-order output in (a0,a1,a2,a3)
template <bool INV> __device__ __forceinline__ void bfly4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
    float2 s13 = cadd(a1, a3), d13 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}

// 8-point DFT, natural-order output in a[0..7]
template <bool INV> __device__ __forceinline__ void bfly8(float2 (&a)[8]) {
    const float h = 0.70710678118654752440f;
    float2 b0 = cadd(a[0], a[4]), c0 = csub(a[0], a[4]);
    float2 b1 = cadd(a[1], a[5]), c1 = csub(a[1], a[5]);
    float2 b2 = cadd(a[2], a[6]), c2 = csub(a[2], a[6]);
    float2 b3 = cadd(a[3], a[7]), c3 = csub(a[3], a[7]);
    // c_n *= W8^n  (forward W8 = exp(-i*pi/4); inverse = conjugate)
    if (INV) {
        c1 = make_float2((c1.x - c1.y) * h, (c1.x + c1.y) * h);
        c2 = make_float2(-c2.y, c2.x);
        c3 = make_float2((-c3.x - c3.y) * h, (c3.x - c3.y) * h);
    } else {
        c1 = make_float2((c1.x + c1.y) * h, (c1.y - c1.x) * h);
        c2 = make_float2(c2.y, -c2.x);
        c3 = make_float2((c3.y - c3.x) * h, (-c3.x - c3.y) * h);
    }
    bfly4<INV>(b0, b1, b2, b3);   // X0 X2 X4 X6
    bfly4<INV>(c0, c1, c2, c3);   // X1 X3 X5 X7
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = c0; a[3] = c1; a[5] = c2; a[7] = c3;
}

// ------------------------------------------------------------------------------------------------
// group barrier: T <= 32 -> the group lives inside one warp; otherwise either the whole CTA is one
// group (bar 0) or several groups share a CTA and use named barriers 1..15.
// ------------------------------------------------------------------------------------------------
template <int T, int GROUPS> __device__ __forceinline__ void group_sync(int gid) {
    if (T <= 32) {
        __syncwarp();
    } else if (GROUPS == 1) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(T) : "memory");
    }
}

// Fill the CTA's shared twiddle table from the global table tw[k] = exp(-2*pi*i*k/N).
template <int LOG2N>
__device__ __forceinline__ void fill_twiddles(float2 *tws, const float2 *__restrict__ tw, int tid, int nthreads) {
    using C = FftCfg<LOG2N>;
#pragma unroll
    for (int p = 1; p < C::NPASS; ++p) {
        const int R = C::radix(p), Ns = C::ns(p), off = C::tw_off(p);
        const int step = C::N / (Ns * R);
        for (int i = tid; i < (R - 1) * Ns; i += nthreads) {
            const int q = i / Ns + 1, r = i & (Ns - 1);
            tws[off + i] = __ldg(&tw[(q * r * step) & (C::N - 1)]);
        }
    }
}

template <bool INV> __device__ __forceinline__ float2 twmul(float2 a, float2 w) {
    return INV ? cmulc(a, w) : cmul(a, w);
}

// One pass P of the transform, fully unrolled by the caller through template recursion.
template <int LOG2N, bool INV, int GROUPS, int P> struct Pass {
    using C = FftCfg<LOG2N>;
    static __device__ __forceinline__ void run(float2 (&v)[8], const float2 *tws, float2 *buf0, float2 *buf1,
                                               int j, int gid, int &pp) {
        constexpr int R = C::radix(P), Ns = C::ns(P);
        if (P > 0) {
            constexpr int off = C::tw_off(P);
            if (R == 8) {
                const float2 *t = tws + off + (j & (Ns - 1));
#pragma unroll
                for (int q = 1; q < 8; ++q) v[q] = twmul<INV>(v[q], t[(q - 1) * Ns]);
            } else if (R == 4) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const float2 *t = tws + off + ((j + b * C::T) & (Ns - 1));
#pragma unroll
                    for (int q = 1; q < 4; ++q) v[b + 2 * q] = twmul<INV>(v[b + 2 * q], t[(q - 1) * Ns]);
                }
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) v[b + 4] = twmul<INV>(v[b + 4], tws[off + ((j + b * C::T) & (Ns - 1))]);
            }
        }
        if (R == 8) {
            bfly8<INV>(v);
        } else if (R == 4) {
            bfly4<INV>(v[0], v[2], v[4], v[6]);
            bfly4<INV>(v[1], v[3], v[5], v[7]);
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) bfly2<INV>(v[b], v[b + 4]);
        }
        if (P < C::NPASS - 1) {
            // only radix-8 passes are followed by an exchange
            float2 *buf = (pp & 1) ? buf1 : buf0;
            pp ^= 1;
            const int e = (j / Ns) * (Ns * 8) + (j & (Ns - 1));
#pragma unroll
            for (int q = 0; q < 8; ++q) buf[lay<LOG2N, P>(e + q * Ns)] = v[q];
            group_sync<C::T, GROUPS>(gid);
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = buf[lay<LOG2N, P>(j + m * C::T)];
            Pass<LOG2N, INV, GROUPS, P + 1>::run(v, tws, buf0, buf1, j, gid, pp);
        }
    }
};
template <int LOG2N, bool INV, int GROUPS> struct Pass<LOG2N, INV, GROUPS, 4> {
    static __device__ __forceinline__ void run(float2 (&)[8], const float2 *, float2 *, float2 *, int, int, int &) {}
};

// In-place transform of the group's symbol.  pp is the running ping-pong selector (kept across
// calls so that consecutive transforms never need a second barrier per exchange).
template <int LOG2N, bool INV, int GROUPS>
__device__ __forceinline__ void fft_group(float2 (&v)[8], const float2 *tws, float2 *buf0, float2 *buf1, int j,
                                          int gid, int &pp) {
    Pass<LOG2N, INV, GROUPS, 0>::run(v, tws, buf0, buf1, j, gid, pp);
}

// ------------------------------------------------------------------------------------------------
// group reductions
// ------------------------------------------------------------------------------------------------
template <int T, int GROUPS> __device__ __forceinline__ float group_sum(float x, float *red, int j, int gid) {
    constexpr int W = T < 32 ? T : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (T > 32) {
        // red: T/32 floats per group, double-buffered by the caller through the pointer it passes
        if ((j & 31) == 0) red[j >> 5] = x;
        group_sync<T, GROUPS>(gid);
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) s += red[w];
        x = s;
    }
    return x;
}

template <int T, int GROUPS>
__device__ __forceinline__ unsigned long long group_max_u64(unsigned long long x, unsigned long long *red, int j,
                                                            int gid) {
    constexpr int W = T < 32 ? T : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
        x = y > x ? y : x;
    }
    if (T > 32) {
        if ((j & 31) == 0) red[j >> 5] = x;
        group_sync<T, GROUPS>(gid);
        unsigned long long s = 0ull;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) s = red[w] > s ? red[w] : s;
        x = s;
    }
    return x;
}

}  // namespace ofdm
