// tx_wide.cuh -- the transmitter for N = 1024 / 2048 / 4096 with SIXTY-FOUR points per thread.
//
// The 16-point-per-thread transmitter (tx_kernel) needs two shared-memory exchanges and two CTA barriers per 2048-point
// symbol and runs at half the HBM roofline: no unit is saturated (ncu, round 2: issue slots 55 %, shared-memory data path
// 53 %, FMA pipe 52 %, DRAM 37 %), the warps wait on each other and on short dependencies.  Here a symbol belongs to
// T = N/64 threads -- ONE warp at N = 2048 -- and the transform is N = 64 x (N/64): a radix-64 pass in registers
// (8 x 8), one exchange, a radix-16/32/64 pass in registers.  Per symbol that is one exchange instead of two (half the
// shared-memory traffic), no CTA barrier at all for N <= 2048 (__syncwarp), a quarter of the per-thread bookkeeping, and
// 64 independent points of instruction-level parallelism per thread, which is what keeps the pipes busy at 8 warps per
// SM.  Same arithmetic as tx_kernel (MultiAntennaSystem.py:135-219: bits -> points / ZC -> IFFT -> CP -> the two
// normalisations folded into one scale); same data layout; selected by launch_tx for the non-pilot plans.
//
// Stockham bookkeeping (as fft_core.cuh): before each pass thread j holds elements j + m*T, m = 0..63.  Pass 0 (Ns = 1,
// radix 64) works on v[0..63] and writes element 64 j + q; the exchange buffer is padded by one slot per 64
// (index i + (i >> 6)), so the write is `thread base + constant` at 65 j + q and the read of element j + m T is
// `thread base + constant` as well, both free of bank conflicts.  Pass 1 (Ns = 64, radix R1 = N/64, B = 64/R1
// butterflies per thread on registers b + B q) multiplies by W_N^(q (j + T b)) -- successive powers of ONE table entry
// per butterfly, formed in two interleaved chains -- and leaves element j + m T in v[m] again: stores are
// warp-contiguous, 8 bytes per lane.
#pragma once
#include "kernels.cuh"

namespace ofdm {

__device__ constexpr float W64_COS[64] = {1.0f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f, 0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f, 0.0f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f, -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f, -1.0f, -0.995184727f, -0.98078528f, -0.956940336f, -0.923879533f, -0.881921264f, -0.831469612f, -0.773010453f, -0.707106781f, -0.634393284f, -0.555570233f, -0.471396737f, -0.382683432f, -0.290284677f, -0.195090322f, -0.0980171403f, 0.0f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f, 0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f};
__device__ constexpr float W64_SIN[64] = {0.0f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f, 0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f, 1.0f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f, 0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f, 0.0f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f, -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f, -1.0f, -0.995184727f, -0.98078528f, -0.956940336f, -0.923879533f, -0.881921264f, -0.831469612f, -0.773010453f, -0.707106781f, -0.634393284f, -0.555570233f, -0.471396737f, -0.382683432f, -0.290284677f, -0.195090322f, -0.0980171403f};

template <int I, int E, class F> __device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < E) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, E>(f);
    }
}

// a * W_64^E (forward: exp(-2 pi i E / 64); INV: the conjugate)
template <bool INV, int E> __device__ __forceinline__ float2 mul_w64(float2 a) {
    constexpr int e = E & 63;
    if constexpr (e == 0) return a;
    else if constexpr (e == 16) return mul_mi<INV>(a);
    else if constexpr (e == 32) return make_float2(-a.x, -a.y);
    else if constexpr (e == 48) return mul_mi<!INV>(a);
    else return cmul(a, make_float2(W64_COS[e], INV ? W64_SIN[e] : -W64_SIN[e]));
}

// 64-point DFT, natural order in and out: n = 8 n1 + n2, k = k1 + 8 k2
template <bool INV> __device__ __forceinline__ void bfly64(float2 (&v)[64]) {
    float2 y[64];
    static_for<0, 8>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        float2 t[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) t[n1] = v[8 * n1 + n2];
        bfly8<INV>(t);
        static_for<0, 8>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            y[8 * k1 + n2] = mul_w64<INV, k1 * n2>(t[k1]);
        });
    });
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        float2 t[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) t[n2] = y[8 * k1 + n2];
        bfly8<INV>(t);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) v[k1 + 8 * k2] = t[k2];
    }
}

// 32-point DFT, natural order in and out: n = 8 n1 + n2 (n1 < 4), k = k1 + 4 k2
template <bool INV> __device__ __forceinline__ void bfly32(float2 (&a)[32]) {
    float2 y[32];
    static_for<0, 8>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        float2 t0 = a[n2], t1 = a[8 + n2], t2 = a[16 + n2], t3 = a[24 + n2];
        bfly4<INV>(t0, t1, t2, t3);
        y[n2] = t0;
        y[8 + n2] = mul_w64<INV, 2 * n2>(t1);
        y[16 + n2] = mul_w64<INV, 4 * n2>(t2);
        y[24 + n2] = mul_w64<INV, 6 * n2>(t3);
    });
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 t[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) t[n2] = y[8 * k1 + n2];
        bfly8<INV>(t);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) a[k1 + 4 * k2] = t[k2];
    }
}

template <int LOG2N> struct TxWide {
    static constexpr int N = 1 << LOG2N, PPT = 64, T = N / 64, R1 = N / 64, B = 64 / R1;
    static constexpr int THREADS = 128, G = THREADS / T;          // symbol groups per CTA
    static constexpr int BUF = 65 * T;                            // padded exchange buffer, complex slots per group
    static constexpr int ROW = 3 * N / 4 + 32;                    // one packed payload row at any 16-byte phase
    static constexpr int OFF_ROW = (G * BUF * 8 + 15) & ~15;
    static constexpr int OFF_RED = OFF_ROW + G * 2 * ROW;
    static constexpr int OFF_LUT = OFF_RED + G * 2 * 8 * 4;       // two slots of 8 floats per group (N = 4096 only)
    static constexpr int BYTES = OFF_LUT + 80 * 4;
    static_assert(LOG2N >= 10 && LOG2N <= 12, "64 points per thread: N = 1024, 2048, 4096");
    static_assert(OFF_ROW % 16 == 0 && ROW % 16 == 0, "cp.async row buffers need 16-byte alignment");
};

template <int T> __device__ __forceinline__ void wide_sync(int gid) {
    if (T <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(T) : "memory");
}

template <int LOG2N, int BPS, bool CPP>
__global__ void __launch_bounds__(128, 2)
tx_wide_kernel(const DevParams P, const uint8_t *__restrict__ bits, float2 *__restrict__ out, long long n_sym_total,
               int iters) {
    using W = TxWide<LOG2N>;
    constexpr int N = W::N, T = W::T, R1 = W::R1, B = W::B, G = W::G, PPT = 64;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, gid = tid / T, j = tid % T;
    float2 *buf = reinterpret_cast<float2 *>(smem_raw) + gid * W::BUF;
    uint8_t *hb = smem_raw + W::OFF_ROW + gid * 2 * W::ROW;
    unsigned *lut32 = reinterpret_cast<unsigned *>(smem_raw + W::OFF_LUT);
    // constellation table: integer levels as two halves (see tx_kernel, LUT32); entry 1 << BPS maps unused bins to zero
    constexpr int ZERO = 1 << BPS;
    constexpr float AMP = BPS == 1 ? 1.f : (BPS == 2 ? 0.70710678118654752440f
                                                     : (BPS == 4 ? 0.31622776601683793320f : 0.15430334996209191026f));
    for (int i = tid; i <= ZERO; i += W::THREADS) {
        const float2 c = i < ZERO ? map_bits<BPS>((unsigned)i) : make_float2(0.f, 0.f);
        const __half2 h = __floats2half2_rn(rintf(c.x / AMP), rintf(c.y / AMP));
        lut32[i] = *reinterpret_cast<const unsigned *>(&h);
    }
    __syncthreads();
    auto look = [&](unsigned idx) -> float2 {
        const unsigned e = lut32[idx];
        return __half22float2(*reinterpret_cast<const __half2 *>(&e));
    };
    const int L = N + P.cp, fs = P.m0 + P.d;
    // this thread's bins k = j + m T: m < 32 the positive half (incl. DC), m >= 32 the negative half
    const int h = P.ndb >> 1;
    const int base_p = j - 1 + h, base_n = j + h - N / 2;                  // used-bin-list index of bin j, of bin j + N/2
    unsigned long long mask = 0ull;
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        const int k = j + m * T;
        const bool used = (m < 32) ? (k >= 1 && k <= h) : (k >= N - h);
        mask |= used ? (1ull << m) : 0ull;
    }
    const int bop = base_p * BPS, bon = (base_n - 32 * T) * BPS;
    const int byp = bop >> 3, shp = 16 - (bop & 7) - BPS, byn = bon >> 3, shn = 16 - (bon & 7) - BPS;
    // pass-1 twiddle bases W_N^(j + T b) (inverse transform: conjugate)
    float2 ub[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const float2 w = __ldg(P.tw + j + T * b);
        ub[b] = make_float2(w.x, -w.y);
    }
    const float sqrtL0 = sqrtf((float)L), invL0 = 1.f / (float)L;
    const float sqrtL1 = sqrtf((float)(L + P.cp_extra)), invL1 = 1.f / (float)(L + P.cp_extra);
    (void)sqrtL1; (void)invL1;
    const float emin = 1e-30f * (float)N * (float)N;
    // payload rows: staged with cp.async, one symbol ahead (see tx_kernel)
    const long long mis = (long long)(reinterpret_cast<unsigned long long>(bits) & 15ull);
    const uint8_t *abits = bits - mis;
    const long long limit = mis + (n_sym_total / fs) * P.d * (long long)P.hard_stride;
    const int nchunks = (P.hard_stride + 15 + 15) / 16;
    const unsigned hb_s = (unsigned)__cvta_generic_to_shared(hb);
    const long long TG = (long long)gridDim.x * G, gg = (long long)blockIdx.x * G + gid, last = n_sym_total - 1;
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
    long long oo = q * L + P.cp + j;
    const long long step_f = CPP ? TG / fs : 0, f_last = CPP ? last / fs : 0;
    long long f = CPP ? q / fs : 0;
    const int nslots = CPP ? (fs + P.slot - 1) / P.slot : 0;
    auto request = [&](long long rr, int b) {
        if (rr >= 0) {
            const long long base = rr & ~15ll, left = limit - base;
            const int room = left < (1 << 30) ? (int)left : (1 << 30);
            for (int c = j; c < nchunks; c += T) {
                const int sz = min(max(room - 16 * c, 0), 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(hb_s + b * W::ROW + 16 * c),
                             "l"(abits + base + (sz ? 16 * c : 0)), "r"(sz) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cb = 0;
    request(s >= P.m0 ? r : -1, 0);
    for (int it = 0; it < iters; ++it) {
        long long qn = q + TG, rn = r + dr, on = oo + dout, fn = f + step_f;
        int sn = s + step_s;
        if (sn >= fs) { sn -= fs; rn -= dr_wrap; if (CPP) ++fn; }
        if (qn > last) { qn = last; sn = s_last; rn = r_last; on = last * L + P.cp + j; fn = f_last; }
        int ncp0 = N - P.cp;
        long long xo = 0;
        float sqrtL = sqrtL0, invL = invL0;
        if (CPP) {
            const int sl = s / P.slot;
            xo = (long long)P.cp_extra * (f * nslots + sl + 1);
            if (s == sl * P.slot) { ncp0 -= P.cp_extra; sqrtL = sqrtL1; invL = invL1; }
        }
        float2 v[PPT];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        wide_sync<T>(gid);                       // the row is visible to the group; the previous symbol's reads are done
        request(sn >= P.m0 ? rn : -1, cb ^ 1);
        const uint8_t *rp = hb + cb * W::ROW + (int)(r & 15);
        cb ^= 1;
        // bits -> points.  MZ: a grid that fills at most 5/8 of the transform (every LTE bandwidth) leaves bins
        // 5N/16+1 .. 11N/16-1 empty, i.e. m = 21..43 for every thread: no fetch, no lookup
        auto front = [&](auto mz) {
            constexpr bool MZ = decltype(mz)::value;
#pragma unroll
            for (int m = 0; m < PPT; ++m) {
                if (MZ && m >= 21 && m < 44) {
                    v[m] = make_float2(0.f, 0.f);
                    continue;
                }
                const uint8_t *qp = rp + (m < 32 ? byp : byn) + m * (T * BPS / 8);
                unsigned val;
                if (BPS == 6) {
                    const unsigned w = ((unsigned)qp[0] << 8) | (unsigned)qp[1];
                    val = (w >> (m < 32 ? shp : shn)) & 63u;
                } else {
                    val = ((unsigned)qp[0] >> ((m < 32 ? shp : shn) - 8)) & ((1u << BPS) - 1u);
                }
                v[m] = look(((mask >> m) & 1ull) ? val : ZERO);
            }
        };
        if (s < P.m0) {
            const float2 *z = P.zc_tx + s * N + j;
#pragma unroll
            for (int m = 0; m < PPT; ++m) v[m] = __ldg(z + m * T);
        } else if (P.ndb * 8 <= 5 * N) {
            front(std::true_type{});
        } else {
            front(std::false_type{});
        }
        // pass 0: radix 64 in registers, then the one exchange
        bfly64<true>(v);
        {
            float2 *wbase = buf + 65 * j;
#pragma unroll
            for (int qq = 0; qq < PPT; ++qq) wbase[qq] = pin(v[qq]);
            wide_sync<T>(gid);
            const float2 *rbase = buf + j;
#pragma unroll
            for (int m = 0; m < PPT; ++m) v[m] = rbase[m * T + ((m * T) >> 6)];
        }
        // pass 1: twiddles W^(q p) as successive powers of W^p (two interleaved chains), radix-R1 butterflies
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const float2 u1 = ub[b], u2 = cmul(u1, u1);
            float2 wo = u1, we = u2;
#pragma unroll
            for (int qq = 1; qq < R1; qq += 2) {
                v[b + B * qq] = cmul(v[b + B * qq], wo);
                if (qq + 1 < R1) v[b + B * (qq + 1)] = cmul(v[b + B * (qq + 1)], we);
                if (qq + 2 < R1) {
                    wo = cmul(wo, u2);
                    we = cmul(we, u2);
                }
            }
            if constexpr (R1 == 64) {
                bfly64<true>(v);
            } else if constexpr (R1 == 32) {
                float2 t[32];
#pragma unroll
                for (int qq = 0; qq < 32; ++qq) t[qq] = v[b + B * qq];
                bfly32<true>(t);
#pragma unroll
                for (int qq = 0; qq < 32; ++qq) v[b + B * qq] = t[qq];
            } else {
                float2 t[16];
#pragma unroll
                for (int qq = 0; qq < 16; ++qq) t[qq] = v[b + B * qq];
                bfly16<true>(t);
#pragma unroll
                for (int qq = 0; qq < 16; ++qq) v[b + B * qq] = t[qq];
            }
        }
        // power / mean sums (prefix samples count twice; the N body samples sum to N X[0] = 0), one scale, stores
        constexpr int MCP = 48;                               // CP <= N/4: only m >= 48 can repeat in the prefix
        float2 pw2 = make_float2(0.f, 0.f), mean2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < PPT; ++m) {
            if (m < MCP) {
                pw2 = cpow_acc(pw2, v[m]);
            } else {
                const float c = (j >= ncp0 - m * T) ? 1.f : 0.f;
                pw2 = cpow_acc_w(pw2, v[m], c + 1.f);
                mean2 = caxpy(mean2, c, v[m]);
            }
        }
        float3 acc = make_float3(pw2.x + pw2.y, mean2.x, mean2.y);
        constexpr int WD = T < 32 ? T : 32;
#pragma unroll
        for (int o = WD / 2; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        }
        if constexpr (T > 32) {
            // N = 4096: a group is two warps; partial sums through shared memory (two slots, alternating per symbol,
            // so the next symbol's writes never meet these reads), one group barrier
            float *slot = reinterpret_cast<float *>(smem_raw + W::OFF_RED) + (gid * 2 + (it & 1)) * 8;
            if ((j & 31) == 0) {
                slot[3 * (j >> 5) + 0] = acc.x;
                slot[3 * (j >> 5) + 1] = acc.y;
                slot[3 * (j >> 5) + 2] = acc.z;
            }
            wide_sync<T>(gid);
            acc = make_float3(slot[0] + slot[3], slot[1] + slot[4], slot[2] + slot[5]);
        }
        const float s1 = acc.x > emin ? sqrtL * rsqrtf(acc.x) : 1.f / (float)N;
        const float mx = s1 * acc.y * invL, my = s1 * acc.z * invL;
        const float var = s1 * s1 * acc.x * invL - (mx * mx + my * my);
        const float g = var > 0.f ? s1 * rsqrtf(var) : 0.f;
        float2 *o = out + oo + xo;
        float2 *ocp = o - N;
#pragma unroll
        for (int m = 0; m < PPT; ++m) {
            const float2 z = cscale(v[m], g);
            st_out(o + m * T, z);
            if (m >= MCP && j >= ncp0 - m * T) st_out(ocp + m * T, z);
        }
        q = qn;
        s = sn;
        r = rn;
        oo = on;
        f = fn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace ofdm
