// fft_core.cuh -- batched power-of-two FFT for one "symbol group" of T = N/PPT threads.
//
// Every thread keeps PPT complex points in registers (PPT = 8 for N <= 512, 16 above: half the
// threads, two exchanges instead of three at N = 2048, twice the independent work per thread).
// The transform is a Stockham autosort sequence of radix-PPT passes followed by at most one smaller
// power-of-two pass; between passes the group exchanges data through shared memory (two ping-pong
// buffers, one group barrier per exchange).  Invariants used by every kernel in this library:
//
//   * before each pass and after the last one, thread j holds element  j + m*T  in v[m], m=0..PPT-1.
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

// Complex arithmetic on Blackwell's packed FP32 pipe.  A complex number is one 64-bit register pair (re, im) and
// add.rn.f32x2 / mul.rn.f32x2 / fma.rn.f32x2 (FADD2 / FMUL2 / FFMA2) work on both halves at once; their operands
// take a half-swap (.LO_HI), per-half signs and a 32-bit broadcast (.F32) for free, so
//   a +- b         = 1 FADD2                          (2 FADD before)
//   a * w          = FMUL2(a, w.x) ; FFMA2(swap(a), (-w.y, +w.y), .)     (2 FMUL + 2 FFMA before)
//   a * (+-i)      = nothing: the swap and the signs fold into the instruction that consumes the product
// which halves the floating-point instruction count of every butterfly (the N >= 1024 kernels are bound by
// instruction issue).  -DOFDM_F32X2=0 builds the scalar forms for A/B runs.
#ifndef OFDM_F32X2
#define OFDM_F32X2 1
#endif
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
// ptxas 12.9 copies the result of a packed instruction into a fresh register pair (two MOVs) when that result is the
// data operand of a 64-bit store; naming the two halves as scalars first (an empty asm, no instruction) makes it store
// from the pair the instruction wrote.  Used on every value that goes from packed arithmetic to shared or global memory.
__device__ __forceinline__ float2 pin(float2 r) {
#if OFDM_F32X2
    asm("" : "+f"(r.x), "+f"(r.y));
#endif
    return r;
}
#if OFDM_F32X2
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return __ffma2_rn(cswap(a), make_float2(-b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return __ffma2_rn(cswap(a), make_float2(b.y, -b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
// (OFDM_F32X2_ADD=0: additions and subtractions as two scalar instructions each, multiplies packed -- an A/B of how the
// two FP32 pipes share the work; B200: see DESIGN.md 3.11)
#ifndef OFDM_F32X2_ADD
#define OFDM_F32X2_ADD 1
#endif
#if OFDM_F32X2_ADD
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a + s * b with a real s
__device__ __forceinline__ float2 caxpy(float2 a, float s, float2 b) { return __ffma2_rn(b, make_float2(s, s), a); }
// acc + (a.x^2, a.y^2): a power sum kept as two partial sums, added up once at the end
__device__ __forceinline__ float2 cpow_acc(float2 acc, float2 a) { return __ffma2_rn(a, a, acc); }
// acc + a * b
__device__ __forceinline__ float2 cmul_acc(float2 acc, float2 a, float2 b) {
    return __ffma2_rn(cswap(a), make_float2(-b.y, b.y), __ffma2_rn(a, make_float2(b.x, b.x), acc));
}
// acc + w * (a.x^2, a.y^2)
__device__ __forceinline__ float2 cpow_acc_w(float2 acc, float2 a, float w) {
    return __ffma2_rn(__fmul2_rn(a, make_float2(w, w)), a, acc);
}
#else
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 caxpy(float2 a, float s, float2 b) { return make_float2(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y)); }
__device__ __forceinline__ float2 cpow_acc(float2 acc, float2 a) { return make_float2(fmaf(a.x, a.x, acc.x), fmaf(a.y, a.y, acc.y)); }
__device__ __forceinline__ float2 cmul_acc(float2 acc, float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y)));
}
__device__ __forceinline__ float2 cpow_acc_w(float2 acc, float2 a, float w) {
    return make_float2(fmaf(w * a.x, a.x, acc.x), fmaf(w * a.y, a.y, acc.y));
}
#endif
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 mul_mi(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// pass structure: NFULL radix-PPT passes, then one pass of radix 2^(LOG2N mod log2 PPT) if that is > 1
__host__ __device__ constexpr int fft_lp(int ppt) { return ppt == 16 ? 4 : 3; }
__host__ __device__ constexpr int fft_radix(int log2n, int ppt, int p) {
    return p < log2n / fft_lp(ppt) ? ppt : (1 << (log2n % fft_lp(ppt)));
}
__host__ __device__ constexpr int fft_ns(int log2n, int ppt, int p) {
    return 1 << (fft_lp(ppt) * (p < log2n / fft_lp(ppt) ? p : log2n / fft_lp(ppt)));
}
// shared twiddle table: pass p >= 1 holds rows of Ns entries, row for multiplier q entry r = W_N^(q*r*N/(Ns*R)).
// Full table (CH 0): rows q = 1..R-1.  Chained table (CH 1): rows q = 1, 2, 4, .. only; the other multipliers are
// products of those (3 = 1*2, 5 = 1*4, ...), which trades shared-memory reads for FMA-pipe work.  Power table (CH 2):
// row q = 1 only, W^q = W^(q/2) * W^(q - q/2) -- one shared-memory read per butterfly and R-2 complex multiplies, two
// packed instructions each.  With the butterflies on the packed pipe the shared-memory data path is the busiest unit
// of the N >= 1024 kernels (ncu: l1tex data-pipe wavefronts 72 % in the LTE transmitter, FMA pipe 24 %), so reads are
// what to save.  A product of k unit-modulus fp32 factors is off by <= k * 6e-8: W^15 to 4e-7, far inside the 1e-4 bar.
__host__ __device__ constexpr int fft_log2(int r) { return r <= 1 ? 0 : 1 + fft_log2(r / 2); }
__host__ __device__ constexpr int fft_tw_rows(int r, int ch) { return ch == 2 ? 1 : (ch == 1 ? fft_log2(r) : r - 1); }
__host__ __device__ constexpr int fft_tw_off(int log2n, int ppt, int p, int ch = 0) {
    int o = 0;
    for (int i = 1; i < p; ++i) o += fft_tw_rows(fft_radix(log2n, ppt, i), ch) * fft_ns(log2n, ppt, i);
    return o;
}

template <int LOG2N_, int PPT_, int CH_ = 0, bool PAD_ = false> struct FftCfg {
    static constexpr int LOG2N = LOG2N_, PPT = PPT_, CH = CH_;
    static constexpr bool PAD = PAD_;
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / PPT;                     // threads per symbol group
    static constexpr int LP = fft_lp(PPT);
    static constexpr int NFULL = LOG2N / LP;              // number of radix-PPT passes
    static constexpr int REM = LOG2N % LP;                // log2 of the last pass' radix (0: none)
    static constexpr int NPASS = NFULL + (REM ? 1 : 0);
    static constexpr int NEXCH = NPASS - 1;               // shared-memory exchanges per transform
    static constexpr int TW_ENTRIES = fft_tw_off(LOG2N, PPT, NPASS, CH);
    // padded buffer length in complex slots; N == 64 uses a pad-per-8 layout (see lay())
    // (PAD: one pad slot per 16, resp. 8 per 64, instead of XOR swizzles -- see lay())
    static constexpr int BUF = (LOG2N == 6) ? 72 : (PAD ? (PPT == 16 ? N + N / 16 : N + N / 8) : N);
    __host__ __device__ static constexpr int radix(int p) { return fft_radix(LOG2N, PPT, p); }
    __host__ __device__ static constexpr int ns(int p) { return fft_ns(LOG2N, PPT, p); }
    __host__ __device__ static constexpr int tw_off(int p) { return fft_tw_off(LOG2N, PPT, p, CH); }
};

// Points per thread for each transform size and kernel family (overridable for A/B runs, e.g.
// -DOFDM_PPT_RX_4096=8).  Measured on B200 (tools/build_variant.sh + tools/run_variants.sh): 16 points per thread speed the
// transmitter up by 11-28 % at N >= 256 and the receiver by 14 % at N = 4096; at N = 1024/2048 the
// receiver also carries 16 equaliser gains per thread and is faster with 8.
#ifndef OFDM_PPT_TX_LARGE
#define OFDM_PPT_TX_LARGE 16
#endif
#ifndef OFDM_PPT_RX_1024
#define OFDM_PPT_RX_1024 8
#endif
#ifndef OFDM_PPT_RX_2048
#define OFDM_PPT_RX_2048 8
#endif
#ifndef OFDM_PPT_RX_4096
#define OFDM_PPT_RX_4096 16
#endif
template <int LOG2N> struct PptRx {
    static constexpr int value = LOG2N <= 9 ? 8 : (LOG2N == 10 ? OFDM_PPT_RX_1024 : (LOG2N == 11 ? OFDM_PPT_RX_2048 : OFDM_PPT_RX_4096));
};
#ifndef OFDM_PPT_TX_MIN_LOG2N
#define OFDM_PPT_TX_MIN_LOG2N 8           // B200: N=256 +11 % with 16 points per thread (one exchange), N=128 -3 %
#endif
template <int LOG2N> struct PptTx {
    static constexpr int value = LOG2N >= OFDM_PPT_TX_MIN_LOG2N ? OFDM_PPT_TX_LARGE : 8;
};
// reduced twiddle tables (FftCfg CH) from this size on; 13 = never.  Round 1 (scalar butterflies): chained rows +2 % for
// the framed receiver at N >= 1024, +3 % for the transmitter at N = 4096.  With packed butterflies the FMA pipe has
// room and the shared-memory data path is the busiest unit, so the transmitter takes the reduced table from N = 256.
#ifndef OFDM_TW_CHAIN_TX_MIN_LOG2N
#define OFDM_TW_CHAIN_TX_MIN_LOG2N 8
#endif
#ifndef OFDM_TW_CHAIN_RXF_MIN_LOG2N
#define OFDM_TW_CHAIN_RXF_MIN_LOG2N 10
#endif
// which reduced table the sizes at or above those thresholds use: 1 = chained rows, 2 = first row + powers
// B200, packed butterflies (same box, ms per launch):  transmitter LTE 64-QAM 0.338 (full table) -> 0.316 (chained rows)
// / 0.317 (powers), N=2048 0.176 -> 0.163 / 0.162, N=4096 0.199 -> 0.199 / 0.197, N=256 0.113 -> 0.111, N=512 0.136 ->
// 0.133 (powers);  framed receiver, chained -> powers: LTE 64-QAM 0.702 -> 0.675, N=2048 16-QAM 0.338 -> 0.333, N=4096
// 0.410 -> 0.403, N=1024 0.298 -> 0.295; below N=1024 the receiver is HBM-bound and does not care.
#ifndef OFDM_TW_CHAIN_TX_MODE
#define OFDM_TW_CHAIN_TX_MODE 2
#endif
#ifndef OFDM_TW_CHAIN_RXF_MODE
#define OFDM_TW_CHAIN_RXF_MODE 2
#endif
#ifndef OFDM_TX_LAY_PAD
#define OFDM_TX_LAY_PAD 1
#endif
// N = 64: exchange through warp shuffles instead of shared memory (see Pass::run).  B200 A/B in DESIGN.md 3.11.
#ifndef OFDM_EXCH_SHFL_64
#define OFDM_EXCH_SHFL_64 0
#endif
// (round 2, with the arithmetic on the packed pipe the address arithmetic of the XOR swizzles weighs more: padded layouts in the
// receiver are now +1..3 % at almost every size -- sweep N=2048 0.746 / 0.830 / 0.770 -> 0.765 / 0.850 / 0.789, N=4096 QPSK 0.571 ->
// 0.586, LTE 0.645 -> 0.637 ms -- and within +-1 % elsewhere; round 1 had measured 0..-1 %)
#ifndef OFDM_RX_LAY_PAD
#define OFDM_RX_LAY_PAD 1
#endif
template <int LOG2N> using CfgRx = FftCfg<LOG2N, PptRx<LOG2N>::value, 0, OFDM_RX_LAY_PAD != 0>;
// the overlap-save search (api.cu) keeps the unpadded buffers: its CTA must stay under half an SM's shared memory at NB = 4096
template <int LOG2N>
using CfgOs = FftCfg<LOG2N, (LOG2N <= 9 ? 8 : (LOG2N == 10 ? OFDM_PPT_RX_1024 : (LOG2N == 11 ? OFDM_PPT_RX_2048 : OFDM_PPT_RX_4096))),
                     (LOG2N >= 10 ? 2 : 0), false>;

// The fused framed receiver (rx_frames_kernel) has its own knobs, because it is the one kernel that
// keeps per-frame state (the equaliser gains) alive across symbols:
//   OFDM_PPT_RXF_<N>        points per thread
//   OFDM_GSM_MIN_LOG2N      from this size on the gains live in thread-private shared-memory slots
//                           instead of registers (frees 2*PPT registers per thread)
//   OFDM_SB_MIN_LOG2N       from this size on the FFT exchanges use ONE buffer and two barriers
//                           instead of two ping-pong buffers and one barrier (halves that footprint)
#ifndef OFDM_PPT_RXF_1024
#define OFDM_PPT_RXF_1024 OFDM_PPT_RX_1024
#endif
#ifndef OFDM_PPT_RXF_2048
#define OFDM_PPT_RXF_2048 OFDM_PPT_RX_2048
#endif
#ifndef OFDM_PPT_RXF_4096
#define OFDM_PPT_RXF_4096 OFDM_PPT_RX_4096
#endif
#ifndef OFDM_GSM_MIN_LOG2N
#define OFDM_GSM_MIN_LOG2N 13
#endif
#ifndef OFDM_SB_MIN_LOG2N
#define OFDM_SB_MIN_LOG2N 13
#endif
// ... and up to this size as well (0 = none).  For groups inside one warp the second barrier is a __syncwarp.
#ifndef OFDM_SB_SMALL_MAX_LOG2N
#define OFDM_SB_SMALL_MAX_LOG2N 0
#endif
template <int LOG2N> struct PptRxF {
    static constexpr int value = LOG2N <= 9 ? 8 : (LOG2N == 10 ? OFDM_PPT_RXF_1024 : (LOG2N == 11 ? OFDM_PPT_RXF_2048 : OFDM_PPT_RXF_4096));
};
template <int LOG2N>
using CfgRxF = FftCfg<LOG2N, PptRxF<LOG2N>::value, (LOG2N >= OFDM_TW_CHAIN_RXF_MIN_LOG2N ? OFDM_TW_CHAIN_RXF_MODE : 0), OFDM_RX_LAY_PAD != 0>;
// The per-lane-epilogue variants of the framed receiver (sparse and dense) at N = 2048: 16 points per thread (two
// exchanges instead of three, half the barriers) are 4-8 % faster there (B200: 1920-bin grid 16-QAM 0.345 -> 0.326 ms,
// QPSK 0.311 -> 0.287 ms; LTE grid QPSK 0.539 -> 0.500 ms, 16-QAM 0.585 -> 0.561 ms), while the bulk-row 64-QAM kernel
// loses 10 % with it (0.711 -> 0.788 ms) and keeps 8.
#ifndef OFDM_PPT_RXFD_2048
#define OFDM_PPT_RXFD_2048 16
#endif
template <int LOG2N>
using CfgRxFD = FftCfg<LOG2N, (LOG2N == 11 ? OFDM_PPT_RXFD_2048 : PptRxF<LOG2N>::value),
                       (LOG2N >= OFDM_TW_CHAIN_RXF_MIN_LOG2N ? OFDM_TW_CHAIN_RXF_MODE : 0), OFDM_RX_LAY_PAD != 0>;
template <int LOG2N> struct RxFOpts {
    static constexpr bool GSM = LOG2N >= OFDM_GSM_MIN_LOG2N;
    static constexpr bool SB = LOG2N >= OFDM_SB_MIN_LOG2N || LOG2N <= OFDM_SB_SMALL_MAX_LOG2N;
};
template <int LOG2N>
using CfgTx = FftCfg<LOG2N, PptTx<LOG2N>::value, (LOG2N >= OFDM_TW_CHAIN_TX_MIN_LOG2N ? OFDM_TW_CHAIN_TX_MODE : 0), OFDM_TX_LAY_PAD != 0>;

// Exchange layout for the buffer written by pass P (Ns = PPT^P), chosen so that the scattered write
// out[(j/Ns)*Ns*R + j%Ns + q*Ns] and the contiguous read in[j + m*T] both hit 16 distinct 8-byte
// bank pairs per half-warp:
//   N == 64 (PPT 8)    : i + (i >> 3)             (two 8-lane groups share a half-warp; base stride 72)
//   PPT 8,  P == 0     : i ^ ((i >> 4) & 7)       stride-8 scatter
//   PPT 8,  P == 1     : i ^ ((i >> 3) & 8)       bit 6 folded into bit 3
//   PPT 16, P == 0     : i ^ ((i >> 4) & 15)      stride-16 scatter
//   otherwise          : identity (writes are already lane-contiguous)
// C::PAD replaces the XOR swizzles by padding -- P == 0: i + (i >> 4), PPT 8 P == 1: i + 8 * (i >> 6) -- which is
// conflict-free on both sides as well and makes every scattered write `thread base + constant` (77 fewer
// instructions per 2048-point transmit symbol), at the price of N/16 (N/8) more slots per buffer.  Measured on B200:
// transmitter +1..2 %, framed receiver 0..-1 % -> transmitter only (OFDM_TX_LAY_PAD / OFDM_RX_LAY_PAD).
template <class C, int P> __device__ __forceinline__ int lay(int i) {
    if (C::LOG2N == 6) return i + (i >> 3);
    if (C::PAD) {
        if (P == 0) return i + (i >> 4);
        if (P == 1 && C::PPT == 8) return i + ((i >> 6) << 3);
        return i;
    }
    if (C::PPT == 16) return P == 0 ? (i ^ ((i >> 4) & 15)) : i;
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
    // c_n *= W8^n  (forward W8 = exp(-i*pi/4); inverse = conjugate):  W8 c = h (c -+ i c),  W8^3 c = h (-c -+ i c)
    c1 = cscale(cadd(c1, mul_mi<INV>(c1)), h);
    c2 = mul_mi<INV>(c2);
    c3 = cscale(csub(mul_mi<INV>(c3), c3), h);
    bfly4<INV>(b0, b1, b2, b3);   // X0 X2 X4 X6
    bfly4<INV>(c0, c1, c2, c3);   // X1 X3 X5 X7
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = c0; a[3] = c1; a[5] = c2; a[7] = c3;
}

// 16-point DFT as 4 x 4 (n = 4*n1 + n2, k = k1 + 4*k2): DFT4 over n1, twiddle W16^(n2*k1), DFT4 over n2.
template <bool INV> __device__ __forceinline__ float2 mul_w16(float2 a, int m) {
    // W16^m = exp(-2*pi*i*m/16) (forward), conjugate for the inverse; m in {1,2,3,4,6,9}
    const float c = 0.92387953251128675613f, s = 0.38268343236508977173f, h = 0.70710678118654752440f;
    float wr, wi;
    switch (m) {
        case 1: wr = c; wi = -s; break;
        case 2: wr = h; wi = -h; break;
        case 3: wr = s; wi = -c; break;
        case 4: wr = 0.f; wi = -1.f; break;
        case 6: wr = -h; wi = -h; break;
        default: wr = -c; wi = s; break;   // 9
    }
    if (INV) wi = -wi;
    if (m == 4) return mul_mi<INV>(a);
    if (m == 2) return cscale(cadd(a, mul_mi<INV>(a)), h);                  // W8   = h (1 -+ i)
    if (m == 6) return cscale(csub(mul_mi<INV>(a), a), h);                  // W8^3 = h (-1 -+ i)
    return cmul(a, make_float2(wr, wi));
}

// 4-point DFT whose second and / or third input is known to be zero (Z1: a1 == 0, Z2: a2 == 0): x + 0 is not folded
// by the compiler without fast-math, so the shorter forms are written out
template <bool INV, bool Z1, bool Z2>
__device__ __forceinline__ void bfly4z(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 s02 = Z2 ? a0 : cadd(a0, a2), d02 = Z2 ? a0 : csub(a0, a2);
    const float2 s13 = Z1 ? a3 : cadd(a1, a3);
    const float2 d13 = mul_mi<INV>(Z1 ? make_float2(-a3.x, -a3.y) : csub(a1, a3));
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}

// MZ ("middle zero"): inputs 5..10 are known to be zero -- the guard band of a grid that fills less than 5/8 of the
// transform (every LTE bandwidth), as seen by the first pass of the transmitter's inverse transform, where input m of
// a thread is bin j + m*T.  Saves 12 of the 32 complex additions of the first radix-4 layer.
template <bool INV, bool MZ = false> __device__ __forceinline__ void bfly16(float2 (&a)[16]) {
    float2 y[16];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        float2 t0 = a[n2], t1 = a[4 + n2], t2 = a[8 + n2], t3 = a[12 + n2];
        if (MZ && n2 == 0) bfly4z<INV, false, true>(t0, t1, t2, t3);         // a8 = 0
        else if (MZ && n2 == 3) bfly4z<INV, true, false>(t0, t1, t2, t3);    // a7 = 0
        else if (MZ) bfly4z<INV, true, true>(t0, t1, t2, t3);                // a5 = a9 = 0, a6 = a10 = 0
        else bfly4<INV>(t0, t1, t2, t3);            // k1 = 0..3
        y[n2] = t0;
        y[4 + n2] = n2 ? mul_w16<INV>(t1, n2) : t1;         // y[4*k1 + n2] *= W16^(n2*k1)
        y[8 + n2] = n2 ? mul_w16<INV>(t2, 2 * n2) : t2;
        y[12 + n2] = n2 ? mul_w16<INV>(t3, 3 * n2) : t3;
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 t0 = y[4 * k1], t1 = y[4 * k1 + 1], t2 = y[4 * k1 + 2], t3 = y[4 * k1 + 3];
        bfly4<INV>(t0, t1, t2, t3);                 // k2 = 0..3
        a[k1] = t0;
        a[k1 + 4] = t1;
        a[k1 + 8] = t2;
        a[k1 + 12] = t3;
    }
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
template <class C>
__device__ __forceinline__ void fill_twiddles(float2 *tws, const float2 *__restrict__ tw, int tid, int nthreads) {
#pragma unroll
    for (int p = 1; p < C::NPASS; ++p) {
        const int R = C::radix(p), Ns = C::ns(p), off = C::tw_off(p);
        const int step = C::N / (Ns * R);
        for (int i = tid; i < fft_tw_rows(R, C::CH) * Ns; i += nthreads) {
            const int row = i / Ns, r = i & (Ns - 1);
            const int q = C::CH ? (1 << row) : row + 1;
            tws[off + i] = __ldg(&tw[(q * r * step) & (C::N - 1)]);
        }
    }
}

// number of last-pass twiddles a thread owns, and their load from the CTA's table (call after fill_twiddles + barrier)
template <class C> struct LastPassTw {
    static constexpr int P = C::NPASS - 1, R = C::radix(P), Ns = C::ns(P), B = C::PPT / R;
    static constexpr int COUNT = P > 0 ? B * (R - 1) : 1;
    static __device__ __forceinline__ void load(float2 (&w)[COUNT], const float2 *tws, int j) {
        static_assert(C::CH == 0, "register twiddles are taken from the full table");
        if (P > 0) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
#pragma unroll
                for (int q = 1; q < R; ++q) w[b * (R - 1) + q - 1] = tws[C::tw_off(P) + (q - 1) * Ns + ((j + b * C::T) & (Ns - 1))];
            }
        }
    }
};

template <bool INV> __device__ __forceinline__ float2 twmul(float2 a, float2 w) {
    return INV ? cmulc(a, w) : cmul(a, w);
}

// radix-R butterfly on registers v[b], v[b+B], ..., v[b+(R-1)B]
template <bool INV, int R, int B, int PPT, bool MZ = false>
__device__ __forceinline__ void bfly_strided(float2 (&v)[PPT], int b) {
    if constexpr (R == 16) {
        bfly16<INV, MZ>(v);
    } else if constexpr (R == 8) {
        float2 t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = v[b + q * B];
        bfly8<INV>(t);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[b + q * B] = t[q];
    } else if constexpr (R == 4) {
        bfly4<INV>(v[b], v[b + B], v[b + 2 * B], v[b + 3 * B]);
    } else {
        bfly2<INV>(v[b], v[b + B]);
    }
}

// Split-phase group barrier (mbarrier).  An exchange is "write my outputs -- everybody has written -- read my inputs".
// With bar.sync a warp that has written sits idle until the slowest warp of the group arrives; the LTE receiver spends a
// fifth of its warp cycles there (ncu: barrier 1.58 of 7.6 cycles per issued instruction, two CTAs of eight warps per
// SM).  Arriving on an mbarrier does not block: the thread arrives, fetches the NEXT pass' twiddle row and forms its
// powers (work that does not depend on the exchanged data), and only then waits for the phase to complete.
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ unsigned long long mbar_arrive(unsigned bar) {
    unsigned long long tok;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(tok) : "r"(bar) : "memory");
    return tok;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned long long tok) {
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "l"(tok) : "memory");
    } while (!done);
}

// Twiddles of pass P for one thread: B butterflies x (R-1) multipliers, from the full, chained or power table.
template <class C, int P> struct PassTw {
    static constexpr int R = C::radix(P), Ns = C::ns(P), B = C::PPT / (R > 0 ? R : 1);
    static constexpr int COUNT = (P > 0 && P < C::NPASS) ? B * (R - 1) : 1;
    static __device__ __forceinline__ void load(float2 (&wo)[COUNT], const float2 *tws, int j) {
        if constexpr (P > 0 && P < C::NPASS) {
            constexpr int off = C::tw_off(P);
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float2 *t = tws + off + ((j + b * C::T) & (Ns - 1));
                float2 w[R];
                if constexpr (C::CH == 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) w[q] = t[(q - 1) * Ns];
                } else if constexpr (C::CH == 2) {
                    // one row: W^r; every other multiplier is a product of two smaller ones (depth log2 R)
                    w[1] = t[0];
#pragma unroll
                    for (int q = 2; q < R; ++q) w[q] = cmul(w[q / 2], w[q - q / 2]);
                } else {
                    // rows hold W^(1*r), W^(2*r), W^(4*r), ..; multiplier q = 2^a + rest is built as w[2^a] * w[rest]
#pragma unroll
                    for (int a = 0; (1 << a) < R; ++a) w[1 << a] = t[a * Ns];
#pragma unroll
                    for (int q = 3; q < R; ++q) {
                        const int hi = 1 << (31 - __builtin_clz((unsigned)q));
                        if (q != hi) w[q] = cmul(w[hi], w[q - hi]);
                    }
                }
#pragma unroll
                for (int q = 1; q < R; ++q) wo[b * (R - 1) + q - 1] = w[q];
            }
        }
    }
    template <bool INV> static __device__ __forceinline__ void apply(float2 (&v)[C::PPT], const float2 *w) {
        if constexpr (P > 0 && P < C::NPASS) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[b + q * B] = twmul<INV>(v[b + q * B], w[b * (R - 1) + q - 1]);
            }
        }
    }
};

// One pass P of the transform, fully unrolled through template recursion.  Butterfly b of a thread
// works on registers b + q*B (B = PPT/R butterflies per thread), i.e. on elements j + b*T + q*N/R.
// TWR: the twiddles of the LAST pass depend only on the thread (its element j + b*T never leaves the first Ns of the
// pass), so a kernel with registers to spare keeps them in twreg[b*(R-1) + q-1] instead of reading the table.
// SPLIT: exchanges wait on the group's mbarrier (address `mbar`) after preparing the next pass' twiddles; TWDONE: this
// pass' twiddles were applied by the caller (the previous pass, in its split-phase exchange).
template <class C, bool INV, int GROUPS, int P, bool SB = false, bool TWR = false, bool MZ = false, bool SPLIT = false,
          bool TWDONE = false>
struct Pass {
    static __device__ __forceinline__ void run(float2 (&v)[C::PPT], const float2 *tws, float2 *buf0, float2 *buf1,
                                               int j, int gid, int &pp, const float2 *twreg = nullptr, unsigned mbar = 0) {
        constexpr int R = C::radix(P), Ns = C::ns(P), B = C::PPT / R;
        if constexpr (P > 0 && !TWDONE) {
            if constexpr (TWR && P == C::NPASS - 1) {
                PassTw<C, P>::template apply<INV>(v, twreg);
            } else {
                float2 w[PassTw<C, P>::COUNT];
                PassTw<C, P>::load(w, tws, j);
                PassTw<C, P>::template apply<INV>(v, w);
            }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) bfly_strided<INV, R, B, C::PPT, (MZ && P == 0)>(v, b);
        if constexpr (P < C::NPASS - 1 && C::LOG2N == 6 && OFDM_EXCH_SHFL_64 != 0) {
            // N = 64: the one exchange is an 8 x 8 transpose between the 8 lanes of the group and their 8 registers
            // (element 8j + q leaves thread j, arrives in register j of thread q).  Shuffle form, for the A/B against the
            // shared-memory form (OFDM_EXCH_SHFL_64, default 0): three stages, lane distance 4, 2, 1; per stage four
            // register pairs swap their off-diagonal halves with one 64-bit shuffle each.
#pragma unroll
            for (int D = 4; D >= 1; D >>= 1) {
                const bool up = (j & D) != 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (r & D) continue;
                    const float2 send = up ? v[r] : v[r | D];
                    float2 recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, D);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, D);
                    if (up) v[r] = recv; else v[r | D] = recv;
                }
            }
            Pass<C, INV, GROUPS, P + 1, SB, TWR, MZ, SPLIT>::run(v, tws, buf0, buf1, j, gid, pp, twreg, mbar);
        } else if constexpr (P < C::NPASS - 1) {
            // only full-radix passes are followed by an exchange (B == 1)
            // SB: one buffer, so a second barrier keeps the next exchange's writes behind these reads
            float2 *buf = SB ? buf0 : ((pp & 1) ? buf1 : buf0);
            if (!SB) pp ^= 1;
            const int e = (j / Ns) * (Ns * C::PPT) + (j & (Ns - 1));
#pragma unroll
            for (int q = 0; q < C::PPT; ++q) buf[lay<C, P>(e + q * Ns)] = pin(v[q]);
            if constexpr (SPLIT && !SB && C::T > 32 && !(TWR && P + 1 == C::NPASS - 1)) {
                const unsigned long long tok = mbar_arrive(mbar);
                float2 wn[PassTw<C, P + 1>::COUNT];
                PassTw<C, P + 1>::load(wn, tws, j);
                mbar_wait(mbar, tok);
#pragma unroll
                for (int m = 0; m < C::PPT; ++m) v[m] = buf[lay<C, P>(j + m * C::T)];
                PassTw<C, P + 1>::template apply<INV>(v, wn);
                Pass<C, INV, GROUPS, P + 1, SB, TWR, MZ, SPLIT, true>::run(v, tws, buf0, buf1, j, gid, pp, twreg, mbar);
            } else {
                group_sync<C::T, GROUPS>(gid);
#pragma unroll
                for (int m = 0; m < C::PPT; ++m) v[m] = buf[lay<C, P>(j + m * C::T)];
                if (SB) group_sync<C::T, GROUPS>(gid);
                Pass<C, INV, GROUPS, P + 1, SB, TWR, MZ, SPLIT>::run(v, tws, buf0, buf1, j, gid, pp, twreg, mbar);
            }
        }
    }
};
template <class C, bool INV, int GROUPS, bool SB, bool TWR, bool MZ, bool SPLIT, bool TWDONE>
struct Pass<C, INV, GROUPS, 4, SB, TWR, MZ, SPLIT, TWDONE> {
    static __device__ __forceinline__ void run(float2 (&)[C::PPT], const float2 *, float2 *, float2 *, int, int, int &,
                                               const float2 * = nullptr, unsigned = 0) {}
};

// In-place transform of the group's symbol.  pp is the running ping-pong selector (kept across
// calls so that consecutive transforms never need a second barrier per exchange).
template <class C, bool INV, int GROUPS, bool SB = false, bool TWR = false, bool MZ = false, bool SPLIT = false>
__device__ __forceinline__ void fft_group(float2 (&v)[C::PPT], const float2 *tws, float2 *buf0, float2 *buf1, int j,
                                          int gid, int &pp, const float2 *twreg = nullptr, unsigned mbar = 0) {
    static_assert(!MZ || (C::PPT == 16 && C::radix(0) == 16), "MZ prunes the first radix-16 pass");
    Pass<C, INV, GROUPS, 0, SB, TWR, MZ, SPLIT>::run(v, tws, buf0, buf1, j, gid, pp, twreg, mbar);
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
