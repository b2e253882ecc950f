// launch.cuh -- host-side launchers, instantiated once per log2(NFFT) in inst_n*.cu so the
// translation units compile in parallel.
#pragma once
#include <cstdlib>
#include <mutex>

#include "kernels.cuh"
#include "tx_wide.cuh"

namespace ofdm {

template <class C> struct GroupsFor {
    // threads per CTA = max(T, 128): several small symbols share a CTA
    static constexpr int value = C::T >= 128 ? 1 : (128 / C::T);
};

struct LaunchDims {
    int grid, iters, ctas_per_sm;
};

void count_launch();

// grid = min(#CTAs needed, resident CTAs on the whole device): persistent groups stride over the work.
// The occupancy query is cached per device and dynamic shared-memory size (one slot per device: a process that drives
// several GPUs from several threads keeps every device's answer, and the attribute is only ever raised).
template <auto kernel>
static inline cudaError_t plan_launch(int threads, int &smem, int groups, long long work, int sms, LaunchDims &dims) {
    constexpr int MAXDEV = 64;
    struct Slot { int smem = -1, bps = 0, max_smem = -1; };
    static Slot slots[MAXDEV];
    static std::mutex mu;                      // blocks of one flowgraph call work() from their own threads
    std::lock_guard<std::mutex> lock(mu);
    // tuning aid: OFDM_SMEM_PAD=<bytes> inflates the dynamic shared-memory request to lower occupancy
    static const int pad = getenv("OFDM_SMEM_PAD") ? atoi(getenv("OFDM_SMEM_PAD")) : 0;
    smem += pad;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAXDEV) return cudaErrorInvalidDevice;
    Slot &c = slots[dev];
    if (smem != c.smem) {
        if (smem > c.max_smem) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            c.max_smem = smem;
        }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.bps, kernel, threads, smem);
        if (e != cudaSuccess) return e;
        if (c.bps < 1) c.bps = 1;
        c.smem = smem;
    }
    const long long need = (work + groups - 1) / groups;
    const long long cap = (long long)sms * c.bps;
    long long grid = need < cap ? need : cap;
    if (grid < 1) grid = 1;
    dims.grid = (int)grid;
    dims.iters = (int)((work + grid * groups - 1) / (grid * groups));
    dims.ctas_per_sm = c.bps;
    return cudaSuccess;
}

#define OFDM_BPS_SWITCH(bps, ...)         \
    switch (bps) {                        \
        case 1: { constexpr int BPS = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int BPS = 2; __VA_ARGS__; } break; \
        case 4: { constexpr int BPS = 4; __VA_ARGS__; } break; \
        case 6: { constexpr int BPS = 6; __VA_ARGS__; } break; \
        default: return cudaErrorInvalidValue; \
    }

// sizes (bit log2 N) whose dense receiver takes the EPI_DENSE_EQ variant when the equalised symbols are wanted
#ifndef OFDM_EQC_SIZES_DEFAULT
#define OFDM_EQC_SIZES_DEFAULT 0x0e00u
#endif
template <int LOG2N>
cudaError_t launch_rx_frames(const DevParams &P, const float2 *x, long long n_samples, long long first,
                             long long frame_stride, long long n_frames, int *delay, float *peak, uint8_t *locked,
                             int *hyp, float2 *H, float2 *eq, float *llr, uint8_t *hard, cudaStream_t st, int sms) {
    using C = CfgRxF<LOG2N>;
    constexpr int G = GroupsFor<C>::value, TH = C::T * G;
    constexpr bool GSM = RxFOpts<LOG2N>::GSM, SB = RxFOpts<LOG2N>::SB;
    if (n_frames <= 0) return cudaSuccess;
    if (P.pil_k && P.n_cfo) return cudaErrorInvalidValue;       // the hypothesis bank is not combined with comb pilots
    // the per-lane epilogues (sparse and dense) may want a different split of the transform over threads than the
    // bulk-row, pilot and hypothesis-bank variants (CfgRxFD, fft_core.cuh)
    using CD = CfgRxFD<LOG2N>;
    constexpr int GD = GroupsFor<CD>::value;
#define OFDM_LAUNCH_RXF(PIL, CFO, EPI, Q) OFDM_LAUNCH_RXF_(PIL, CFO, EPI, false, Q)
#define OFDM_LAUNCH_RXF_(PIL, CFO, EPI, CPP, Q) OFDM_LAUNCH_RXF_C(C, G, PIL, CFO, EPI, CPP, Q)
#define OFDM_LAUNCH_RXF_C(CC, GG, PIL, CFO, EPI, CPP, Q)                                                               \
    {                                                                                                                \
        constexpr auto k = rx_frames_kernel<CC, BPS, GG, PIL, GSM, SB, CFO, EPI, CPP>;                               \
        int SM = Smem<CC, GG, false, PIL, GSM, SB, (EPI == EPI_BULK ? OFDM_PF_DEPTH_BULK : OFDM_PF_DEPTH)>::BYTES +    \
                 (PIL ? 0 : GG * (Q.ls_stride * 4 + Q.es_stride * 8));                                               \
        cudaError_t e = plan_launch<k>(CC::T * GG, SM, GG, n_frames, sms, d);                                        \
        if (e != cudaSuccess) return e;                                                                              \
        k<<<d.grid, CC::T * GG, SM, st>>>(Q, x, n_samples, first, frame_stride, n_frames, d.iters, delay, peak,      \
                                          locked, hyp, H, eq, llr, hard);                                            \
    }
    // The per-lane variant only knows the 64-QAM LLR staging of small transforms (api.cu); the bulk-store variant
    // (64-QAM only, see data_epilogue_bulk) needs 16-byte aligned rows and must not cost a resident CTA.
    DevParams Qp = P;
    Qp.es_stride = 0;
    if (P.bulk_rows && !(P.bps == 6 && LOG2N <= 6)) Qp.ls_stride = 0;
    const bool bulk_ok = P.bulk_rows && !P.pil_k && !P.n_cfo && ((reinterpret_cast<unsigned long long>(eq) & 15ull) == 0) &&
                         ((reinterpret_cast<unsigned long long>(llr) & 15ull) == 0);
    const bool dense = (long long)P.ndb * 64 >= (long long)OFDM_DENSE_MIN_FILL64 * C::N;
    const bool cpp = P.cp_extra != 0;                           // CP pattern: never together with pilots / hypothesis bank
    if (cpp && (P.pil_k || P.n_cfo)) return cudaErrorInvalidValue;
    OFDM_BPS_SWITCH(P.bps, {
        LaunchDims d;
        if (P.pil_k) OFDM_LAUNCH_RXF(true, false, EPI_SPARSE, Qp)
        else if (P.n_cfo) OFDM_LAUNCH_RXF(false, true, EPI_SPARSE, Qp)
        else {
            bool bulk = false;
            // (bulk-row kernels exist from N = 1024 on; below that the plan never asks for them)
            if constexpr (LOG2N >= 10) {
                if (bulk_ok) {
                    LaunchDims db, dp;
                    int smb = Smem<C, G, false, false, GSM, SB, OFDM_PF_DEPTH_BULK>::BYTES + G * (P.ls_stride * 4 + P.es_stride * 8);
                    int smp = Smem<C, G, false, false, GSM, SB>::BYTES + G * Qp.ls_stride * 4;
                    const cudaError_t eb = plan_launch<rx_frames_kernel<C, BPS, G, false, GSM, SB, false, EPI_BULK>>(TH, smb, G, n_frames, sms, db);
                    const cudaError_t ep = plan_launch<rx_frames_kernel<C, BPS, G, false, GSM, SB, false, EPI_SPARSE>>(TH, smp, G, n_frames, sms, dp);
                    if (eb != cudaSuccess) cudaGetLastError();              // rows do not fit next to the transform: not an error
                    bulk = eb == cudaSuccess && ep == cudaSuccess && db.ctas_per_sm >= dp.ctas_per_sm;
                }
                if (bulk && cpp) OFDM_LAUNCH_RXF_(false, false, EPI_BULK, true, P)
                else if (bulk) OFDM_LAUNCH_RXF(false, false, EPI_BULK, P)
            }
            // densely filled grids take the branch-free epilogue (same shared-memory footprint as the sparse one)
            // (EPI_DENSE_EQ: the equalised symbols are wanted and this size runs faster with that switch compiled out --
            // OFDM_EQC_SIZES: bit log2 N set = use it at that size)
            static const unsigned eqc_sizes = getenv("OFDM_EQC_SIZES") ? (unsigned)strtoul(getenv("OFDM_EQC_SIZES"), nullptr, 0) : OFDM_EQC_SIZES_DEFAULT;
            const bool eqc = eq != nullptr && ((eqc_sizes >> LOG2N) & 1u);
            if (!bulk && dense && cpp && eqc) OFDM_LAUNCH_RXF_C(CD, GD, false, false, EPI_DENSE_EQ, true, Qp)
            else if (!bulk && dense && eqc) OFDM_LAUNCH_RXF_C(CD, GD, false, false, EPI_DENSE_EQ, false, Qp)
            else if (!bulk && dense && cpp) OFDM_LAUNCH_RXF_C(CD, GD, false, false, EPI_DENSE, true, Qp)
            else if (!bulk && dense) OFDM_LAUNCH_RXF_C(CD, GD, false, false, EPI_DENSE, false, Qp)
            else if (!bulk && cpp) OFDM_LAUNCH_RXF_C(CD, GD, false, false, EPI_SPARSE, true, Qp)
            else if (!bulk) OFDM_LAUNCH_RXF_C(CD, GD, false, false, EPI_SPARSE, false, Qp)
        }
        count_launch();
    })
#undef OFDM_LAUNCH_RXF
#undef OFDM_LAUNCH_RXF_
#undef OFDM_LAUNCH_RXF_C
    return cudaGetLastError();
}

template <int LOG2N>
cudaError_t launch_rx_demod(const DevParams &P, const float2 *x, long long n_samples, const long long *start,
                            const int *delay, const int *hrow, const int *hyp, const float2 *H, long long n_sym,
                            float2 *eq, float *llr, uint8_t *hard, cudaStream_t st, int sms, const int *count_dev) {
    using C = CfgRx<LOG2N>;
    constexpr int G = GroupsFor<C>::value, TH = C::T * G;
    int SM = Smem<C, G>::BYTES;
    if (n_sym <= 0) return cudaSuccess;
    OFDM_BPS_SWITCH(P.bps, {
        LaunchDims d;
        if (P.pil_k) {
            constexpr auto k = rx_demod_kernel<C, BPS, G, true>;
            cudaError_t e = plan_launch<k>(TH, SM, G, n_sym, sms, d);
            if (e != cudaSuccess) return e;
            k<<<d.grid, TH, SM, st>>>(P, x, n_samples, start, delay, hrow, hyp, H, n_sym, d.iters, eq, llr, hard, count_dev);
        } else {
            constexpr auto k = rx_demod_kernel<C, BPS, G, false>;
            DevParams Q = P;
            Q.es_stride = 0;
            if (P.bulk_rows && !(P.bps == 6 && LOG2N <= 6)) Q.ls_stride = 0;
            SM += G * Q.ls_stride * 4;
            cudaError_t e = plan_launch<k>(TH, SM, G, n_sym, sms, d);
            if (e != cudaSuccess) return e;
            k<<<d.grid, TH, SM, st>>>(Q, x, n_samples, start, delay, hrow, hyp, H, n_sym, d.iters, eq, llr, hard, count_dev);
        }
        count_launch();
    })
    return cudaGetLastError();
}

template <int LOG2N>
cudaError_t launch_tx(const DevParams &P, const uint8_t *bits, float2 *out, long long n_sym_total, cudaStream_t st,
                      int sms) {
    using C = CfgTx<LOG2N>;
    constexpr int G = GroupsFor<C>::value, TH = C::T * G;
    int SM = Smem<C, G, true, false>::BYTES;
    if (n_sym_total <= 0) return cudaSuccess;
    if (P.cp_extra && P.pil_k) return cudaErrorInvalidValue;     // CP pattern: not combined with comb pilots
    // The 64-point-per-thread transmitter (tx_wide.cuh), for plans without comb pilots and with a prefix of at most N/4.
    // B200, same box, ms per launch, 16 points per thread -> 64: N=4096 0.378 -> 0.331 (QPSK / 16-QAM), 0.385 -> 0.365
    // (64-QAM); N=2048 on a 1920-bin grid 0.309 -> 0.293; LTE grid (1200 of 2048 bins, "middle zero" front) 0.317 -> 0.322;
    // N=1024 0.147 -> 0.151.  Default: N=4096, and N=2048 on grids above 5/8 fill.  OFDM_TX_WIDE=0 never, =2 wherever
    // it applies (A/B runs and the parity tests, which cover it at all three sizes).
    if constexpr (LOG2N >= 10) {
        const char *we = getenv("OFDM_TX_WIDE");
        const int wmode = we ? atoi(we) : 1;
        const bool pick = wmode == 2 || (wmode == 1 && (LOG2N == 12 || (LOG2N == 11 && P.ndb * 8 > 5 * (1 << LOG2N))));
        if (pick && !P.pil_k && 4 * (P.cp + P.cp_extra) <= (1 << LOG2N)) {
            using W = TxWide<LOG2N>;
            int SMW = W::BYTES;
            OFDM_BPS_SWITCH(P.bps, {
                LaunchDims d;
                if (P.cp_extra) {
                    constexpr auto k = tx_wide_kernel<LOG2N, BPS, true>;
                    cudaError_t e = plan_launch<k>(W::THREADS, SMW, W::G, n_sym_total, sms, d);
                    if (e != cudaSuccess) return e;
                    k<<<d.grid, W::THREADS, SMW, st>>>(P, bits, out, n_sym_total, d.iters);
                } else {
                    constexpr auto k = tx_wide_kernel<LOG2N, BPS, false>;
                    cudaError_t e = plan_launch<k>(W::THREADS, SMW, W::G, n_sym_total, sms, d);
                    if (e != cudaSuccess) return e;
                    k<<<d.grid, W::THREADS, SMW, st>>>(P, bits, out, n_sym_total, d.iters);
                }
                count_launch();
            })
            return cudaGetLastError();
        }
    }
    OFDM_BPS_SWITCH(P.bps, {
        LaunchDims d;
        if (P.pil_k) {
            constexpr auto k = tx_kernel<C, BPS, G, true>;
            cudaError_t e = plan_launch<k>(TH, SM, G, n_sym_total, sms, d);
            if (e != cudaSuccess) return e;
            k<<<d.grid, TH, SM, st>>>(P, bits, out, n_sym_total, d.iters);
        } else if (P.cp_extra) {
            constexpr auto k = tx_kernel<C, BPS, G, false, true>;
            cudaError_t e = plan_launch<k>(TH, SM, G, n_sym_total, sms, d);
            if (e != cudaSuccess) return e;
            k<<<d.grid, TH, SM, st>>>(P, bits, out, n_sym_total, d.iters);
        } else {
            constexpr auto k = tx_kernel<C, BPS, G, false>;
            cudaError_t e = plan_launch<k>(TH, SM, G, n_sym_total, sms, d);
            if (e != cudaSuccess) return e;
            k<<<d.grid, TH, SM, st>>>(P, bits, out, n_sym_total, d.iters);
        }
        count_launch();
    })
    return cudaGetLastError();
}

template <int LOG2N>
cudaError_t launch_sync_search(const DevParams &P, const float2 *x, long long n_samples, long long first_start,
                               long long stride, long long n_trials, float *peak, int *delay, int *hyp,
                               cudaStream_t st, int sms) {
    using C = CfgRx<LOG2N>;
    constexpr int G = GroupsFor<C>::value, TH = C::T * G;
    int SM = Smem<C, G>::BYTES;
    if (n_trials <= 0) return cudaSuccess;
    constexpr auto k = rx_sync_search_kernel<C, G>;
    LaunchDims d;
    cudaError_t e = plan_launch<k>(TH, SM, G, n_trials, sms, d);
    if (e != cudaSuccess) return e;
    k<<<d.grid, TH, SM, st>>>(P, x, n_samples, first_start, stride, n_trials, d.iters, peak, delay, hyp);
    count_launch();
    return cudaGetLastError();
}

template <int LOG2N>
cudaError_t launch_chanest(const DevParams &P, const float2 *x, long long n_samples, const long long *pos,
                           const int *delay, const int *hyp, long long n_locks, float2 *H, float2 *htime,
                           float2 *synch_eq, cudaStream_t st, int sms, const int *row, const int *count_dev) {
    using C = CfgRx<LOG2N>;
    constexpr int G = GroupsFor<C>::value, TH = C::T * G;
    int SM = Smem<C, G>::BYTES;
    if (n_locks <= 0) return cudaSuccess;
    constexpr auto k = rx_chanest_kernel<C, G>;
    LaunchDims d;
    cudaError_t e = plan_launch<k>(TH, SM, G, n_locks, sms, d);
    if (e != cudaSuccess) return e;
    k<<<d.grid, TH, SM, st>>>(P, x, n_samples, pos, delay, hyp, n_locks, d.iters, H, htime, synch_eq, row, count_dev);
    count_launch();
    return cudaGetLastError();
}

// one table of launchers per transform size
struct KernelSet {
    cudaError_t (*rx_frames)(const DevParams &, const float2 *, long long, long long, long long, long long, int *,
                             float *, uint8_t *, int *, float2 *, float2 *, float *, uint8_t *, cudaStream_t, int);
    cudaError_t (*rx_demod)(const DevParams &, const float2 *, long long, const long long *, const int *, const int *,
                            const int *, const float2 *, long long, float2 *, float *, uint8_t *, cudaStream_t, int,
                            const int *);
    cudaError_t (*tx)(const DevParams &, const uint8_t *, float2 *, long long, cudaStream_t, int);
    cudaError_t (*sync_search)(const DevParams &, const float2 *, long long, long long, long long, long long, float *,
                               int *, int *, cudaStream_t, int);
    cudaError_t (*chanest)(const DevParams &, const float2 *, long long, const long long *, const int *, const int *,
                           long long, float2 *, float2 *, float2 *, cudaStream_t, int, const int *, const int *);
};

#define OFDM_DEFINE_KERNELSET(LOG2N)                                                               \
    extern const KernelSet kset_##LOG2N = {launch_rx_frames<LOG2N>, launch_rx_demod<LOG2N>,        \
                                           launch_tx<LOG2N>, launch_sync_search<LOG2N>,            \
                                           launch_chanest<LOG2N>};

extern const KernelSet kset_6, kset_7, kset_8, kset_9, kset_10, kset_11, kset_12;

}  // namespace ofdm
