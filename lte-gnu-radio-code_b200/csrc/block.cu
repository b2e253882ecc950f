// block.cu -- one C-ABI call per work() of the two live receiver blocks (the scheduler boundary).
//
// A GNU Radio work() call hands the block a few thousand samples.  Driving the modular entry points from Python costs
// three launches and six synchronous copies per call (0.3-0.6 ms, all latency).  ofdm_rx_block_work does the whole
// call on one stream with ONE synchronisation:
//
//   memcpy into a pinned staging buffer -> H2D -> sliding search (loop A's arithmetic, every trial)
//   -> block_control_kernel: the lock rule and the bookkeeping of loop A, and the symbol list of loop B, ON THE DEVICE
//      (gr-utsa_ofdm/python/SynchAndChanEst.py:166-175,:219,:221-224; gr-RXOFDM/python/synch_and_chan_est.py:170-179,:224-227)
//   -> channel estimate of the new locks (count read from device memory) -> data symbols of the list (likewise)
//   -> the small control block comes back with one D2H copy; equalised rows and the auxiliary outputs are written by the
//      kernels straight into mapped pinned host memory, so only rows that exist cross the link.
//
// The block's state that the reference keeps in NumPy arrays (corr_obs, time_synch_ref, est_chan_freq_P) lives in
// device memory between calls; the Python wrapper mirrors what it needs from the result.
#include <cstring>
#include <new>

#include "plan.cuh"

using namespace ofdm;

namespace {

struct BlockGeom {
    int mode;                  // 0: gr-utsa_ofdm (first acceptable lock, break; loop B over every frame with row 0)
                               // 1: gr-RXOFDM   (every acceptable lock; loop B = first data symbol after each lock)
    int N, CP, L, m0, d, fsyms, stride, rows, max_lock, max_sym;
    float thresh;              // gate * M0*nsb
};

// layout of the control block (device copy `ctl`, host copy in pinned memory): header, state, lists
struct CtlHeader {
    int obs;                   // corr_obs / cor_obs as the block leaves it (the reference resets it to 0 after every call)
    int n_new;                 // locks accepted in this call
    int n_demod;               // data symbols listed for loop B
    int obs_after_a;           // the counter after loop A (row of the last new lock), for the wrapper's diagnostics
};
struct CtlLayout {
    size_t tsr, new_pos, new_delay, new_row, dm_start, dm_delay, dm_hrow, dm_row, bytes;
};
CtlLayout ctl_layout(const BlockGeom &g) {
    CtlLayout l;
    size_t o = sizeof(CtlHeader);
    auto take = [&](size_t n) { size_t at = (o + 7) & ~(size_t)7; o = at + n; return at; };
    l.tsr = take((size_t)g.max_lock * 3 * sizeof(long long));          // (start, delay, int(peak)) per state row
    l.new_pos = take((size_t)g.max_lock * sizeof(long long));
    l.new_delay = take((size_t)g.max_lock * sizeof(int));
    l.new_row = take((size_t)g.max_lock * sizeof(int));
    l.dm_start = take((size_t)g.max_sym * sizeof(long long));
    l.dm_delay = take((size_t)g.max_sym * sizeof(int));
    l.dm_hrow = take((size_t)g.max_sym * sizeof(int));
    l.dm_row = take((size_t)g.max_sym * sizeof(int));
    l.bytes = (o + 15) & ~(size_t)15;
    return l;
}

// One CTA.  peak / delay: the sliding search's result for trials P = 0..nv-1 (window start P*stride + CP).
__global__ void __launch_bounds__(256)
block_control_kernel(const BlockGeom g, const CtlLayout lay, const float *__restrict__ peak, const int *__restrict__ delay,
                     int nv, long long n, unsigned char *ctl) {
    CtlHeader *hd = reinterpret_cast<CtlHeader *>(ctl);
    long long *tsr = reinterpret_cast<long long *>(ctl + lay.tsr);
    long long *new_pos = reinterpret_cast<long long *>(ctl + lay.new_pos);
    int *new_delay = reinterpret_cast<int *>(ctl + lay.new_delay), *new_row = reinterpret_cast<int *>(ctl + lay.new_row);
    long long *dm_start = reinterpret_cast<long long *>(ctl + lay.dm_start);
    int *dm_delay = reinterpret_cast<int *>(ctl + lay.dm_delay), *dm_hrow = reinterpret_cast<int *>(ctl + lay.dm_hrow);
    int *dm_row = reinterpret_cast<int *>(ctl + lay.dm_row);
    const int tid = threadIdx.x;
    const long long gap = 2ll * g.CP + g.N;
    __shared__ int first_hit;
    __shared__ int s_obs, s_new;
    if (tid == 0) {
        first_hit = 0x7fffffff;
        s_obs = hd->obs;
        s_new = 0;
    }
    __syncthreads();
    if (g.mode == 0) {
        // loop A, SynchAndChanEst.py:143-219: the FIRST trial over the gate that lies more than 2CP+N samples past the
        // previous lock (or any, before the first lock) is accepted, then the loop breaks
        const int obs = s_obs;
        const long long prev = tsr[0];
        int mine = 0x7fffffff;
        for (int P = tid; P < nv; P += blockDim.x) {
            const long long start = (long long)P * g.stride + g.CP;
            if (peak[P] > g.thresh && (start - prev > gap || obs == -1)) {
                mine = P;
                break;
            }
        }
        if (mine != 0x7fffffff) atomicMin(&first_hit, mine);
        __syncthreads();
        if (tid == 0 && first_hit != 0x7fffffff) {
            const int P = first_hit;
            s_obs = obs + 1;                                             // :168
            if (s_obs < g.rows) {
                const long long start = (long long)P * g.stride + g.CP;
                tsr[0] = start;                                          // :169-171
                tsr[1] = delay[P];
                tsr[2] = (long long)peak[P];                             // int(dmax_val)
                new_pos[0] = start;
                new_delay[0] = delay[P];
                new_row[0] = s_obs;
                s_new = 1;
            }
        }
        __syncthreads();
        // loop B, :221-248: data symbol k of the frame that starts P symbols into the buffer sits at
        // pos + M0*L*(P+1) + L*k; frames whose first window does not fit, or whose rows do not exist, are skipped
        const long long pos = tsr[0];
        const int dl = (int)tsr[1];
        const long long n_unique = n / g.L;
        int n_dm = 0;
        for (long long P = 0; P < n_unique; P += g.fsyms) {             // the valid frames form a prefix: count them
            const long long base = pos + (long long)g.m0 * g.L * (P + 1);
            if (base + g.N - 1 <= n && P + g.d <= g.rows) n_dm += g.d; else break;
        }
        for (int e = tid; e < n_dm; e += blockDim.x) {
            const int fi = e / g.d, k = e - fi * g.d;
            const long long P = (long long)fi * g.fsyms;
            dm_start[e] = pos + (long long)g.m0 * g.L * (P + 1) + (long long)g.L * k;
            dm_delay[e] = dl;
            dm_hrow[e] = 0;                                              // row 0's channel, whatever lock came last (:242)
            dm_row[e] = (int)P + k;
        }
        if (tid == 0) {
            hd->n_new = s_new;
            hd->n_demod = n_dm;
            hd->obs_after_a = s_obs;
            hd->obs = 0;                                                 // :261
        }
    } else {
        // loop A, synch_and_chan_est.py:146-221: every trial over the gate that lies more than 2CP+N samples past the
        // PREVIOUS ACCEPTED lock is recorded; the decision chain is sequential, so one warp walks the trials 32 at a
        // time (ballot) and lane 0 applies the rule to the set bits in order
        if (tid < 32) {
            int obs = s_obs, n_new = 0;
            bool full = false;
            for (int b0 = 0; b0 < nv && !full; b0 += 32) {
                const int P = b0 + tid;
                unsigned m = __ballot_sync(0xffffffffu, P < nv && peak[P] > g.thresh);
                if (tid == 0) {
                    while (m) {
                        const int q = b0 + __ffs(m) - 1;
                        m &= m - 1;
                        const long long start = (long long)q * g.stride + g.CP;
                        const long long prev = tsr[3 * (obs > 0 ? obs : 0)];
                        if (start - prev > gap || obs == -1) {
                            if (obs + 1 >= g.max_lock) { full = true; break; }    // the reference would raise IndexError
                            ++obs;
                            tsr[3 * obs] = start;
                            tsr[3 * obs + 1] = delay[q];
                            tsr[3 * obs + 2] = (long long)peak[q];
                            new_pos[n_new] = start;
                            new_delay[n_new] = delay[q];
                            new_row[n_new] = obs;
                            ++n_new;
                        }
                    }
                }
                full = __shfl_sync(0xffffffffu, full ? 1 : 0, 0) != 0;
            }
            if (tid == 0) {
                // loop B, :224-250: the first data symbol after each recorded lock, with that lock's channel row
                int n_dm = 0;
                for (int r = 0; r <= obs; ++r) {
                    const long long pos = tsr[3 * r];
                    if (pos + (long long)g.m0 * g.L + g.N - 1 <= n) {
                        dm_start[n_dm] = pos + (long long)g.m0 * g.L;
                        dm_delay[n_dm] = (int)tsr[3 * r + 1];
                        dm_hrow[n_dm] = r;
                        dm_row[n_dm] = r;
                        ++n_dm;
                    }
                }
                hd->n_new = n_new;
                hd->n_demod = n_dm;
                hd->obs_after_a = obs;
                hd->obs = 0;                                             // :265
            }
        }
    }
}

// ---- the block's output stream on device buffers -------------------------------------------------------------------
// est_data_freq (SynchAndChanEst.py:93 / synch_and_chan_est.py:97) is block state: rows a call does not list keep what
// an earlier call left there, and the output stream is packed from the whole array.  For device buffers the array
// lives in device memory: this call's rows (list order) are scattered to their rows ...
__global__ void __launch_bounds__(64)
block_scatter_kernel(const unsigned char *__restrict__ ctl, const CtlLayout lay, const float2 *__restrict__ rows,
                     float2 *__restrict__ edf, int ndb) {
    const int i = blockIdx.x;
    if (i >= reinterpret_cast<const CtlHeader *>(ctl)->n_demod) return;
    const int r = reinterpret_cast<const int *>(ctl + lay.dm_row)[i];
    for (int b = threadIdx.x; b < ndb; b += blockDim.x) edf[(size_t)r * ndb + b] = rows[(size_t)i * ndb + b];
}

// ... and the stream is the first n_elems values of the kept rows laid end to end: every row in mode 1
// (est_data_freq[0:corr_size], synch_and_chan_est.py:252-254), the rows np.delete(.., range(3, rows, M0+D)) leaves in
// mode 0 (SynchAndChanEst.py:249-250 -- the deleted rows start at 3 whatever M0 is; reproduced).
__device__ __forceinline__ long long kept_row(long long k, int mode, int fsyms) {
    if (mode == 1 || k < 3) return k;
    const long long q = (k - 3) / (fsyms - 1), rem = (k - 3) - q * (fsyms - 1);       // fsyms >= 2 here (checked on the host)
    return 3 + q * fsyms + 1 + rem;
}
__global__ void __launch_bounds__(256)
block_pack_kernel(const float2 *__restrict__ edf, float2 *__restrict__ out, long long n_elems, int ndb, int mode, int fsyms) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e / ndb;
        out[e] = edf[kept_row(k, mode, fsyms) * ndb + (e - k * ndb)];
    }
}

// rows the packed stream can have at most
long long kept_rows_total(const BlockGeom &g, int h_rows) {
    if (g.mode == 1) return h_rows;
    const long long deleted = g.rows > 3 ? (g.rows - 3 + g.fsyms - 1) / g.fsyms : 0;   // len(range(3, rows, M0+D))
    return g.rows - deleted;
}

}  // namespace

struct ofdm_rx_block {
    ofdm_plan *pl;
    BlockGeom g;
    CtlLayout lay;
    cudaStream_t st;
    cudaEvent_t ev;            // orders a caller's stream before the block's (device-buffer entry point)
    unsigned char *d_ctl;      // control block, device copy (persistent state + this call's lists)
    float2 *d_H;               // est_chan_freq_P: max_lock rows of nfft (utsa: `rows` rows)
    int h_rows;
    float2 *d_x;               // the call's samples
    float *d_peak;
    int *d_delay;
    float2 *d_eq;              // device-buffer calls: this call's equalised rows in list order (when the caller gives none)
    float2 *d_edf;             // device-buffer calls: est_data_freq, h_rows x num_data_bins, persistent
    long long cap_samples, cap_trial_samples;
    // pinned (mapped) host memory
    float2 *p_in;              // staging for the samples
    unsigned char *p_ctl;      // host copy of the control block
    float2 *p_aux;             // [max_lock][nfft + mm]: time-domain channel, equalised preamble of this call's new locks
    float2 *p_eq;              // [max_sym][ndb]: equalised data rows of this call's list
};

static void block_free(ofdm_rx_block *b) {
    if (!b) return;
    if (b->st) cudaStreamDestroy(b->st);
    if (b->ev) cudaEventDestroy(b->ev);
    cudaFree(b->d_ctl);
    cudaFree(b->d_H);
    cudaFree(b->d_x);
    cudaFree(b->d_peak);
    cudaFree(b->d_delay);
    cudaFree(b->d_eq);
    cudaFree(b->d_edf);
    cudaFreeHost(b->p_in);
    cudaFreeHost(b->p_ctl);
    cudaFreeHost(b->p_aux);
    cudaFreeHost(b->p_eq);
    delete b;
}

// Buffers sized by the call: the search's per-trial results (always) and, for host input only, the sample staging pair.
static int block_reserve(ofdm_rx_block *b, long long n, bool staging) {
    if (n > b->cap_trial_samples) {
        long long cap = b->cap_trial_samples ? b->cap_trial_samples : 8192;
        while (cap < n) cap *= 2;
        const long long trials = cap / b->g.stride + 2;
        cudaFree(b->d_peak);
        cudaFree(b->d_delay);
        b->d_peak = nullptr, b->d_delay = nullptr;
        b->cap_trial_samples = 0;
        CU(cudaMalloc(&b->d_peak, (size_t)trials * sizeof(float)), "ofdm_rx_block(cudaMalloc)");
        CU(cudaMalloc(&b->d_delay, (size_t)trials * sizeof(int)), "ofdm_rx_block(cudaMalloc)");
        b->cap_trial_samples = cap;
    }
    if (staging && n > b->cap_samples) {
        long long cap = b->cap_samples ? b->cap_samples : 8192;
        while (cap < n) cap *= 2;
        cudaFree(b->d_x);
        cudaFreeHost(b->p_in);
        b->d_x = nullptr, b->p_in = nullptr;
        b->cap_samples = 0;
        CU(cudaMalloc(&b->d_x, (size_t)cap * sizeof(float2)), "ofdm_rx_block(cudaMalloc)");
        CU(cudaHostAlloc(&b->p_in, (size_t)cap * sizeof(float2), cudaHostAllocDefault), "ofdm_rx_block(cudaHostAlloc)");
        b->cap_samples = cap;
    }
    return OFDM_OK;
}

extern "C" {

int ofdm_rx_block_create(ofdm_plan *pl, int32_t mode, int32_t rows, ofdm_rx_block **out) {
    if (!pl || !out || (mode != 0 && mode != 1) || rows < 1) return fail(OFDM_EINVAL, "ofdm_rx_block_create: bad argument%s");
    *out = nullptr;
    if (pl->dev.n_cfo || pl->prm.pilot_spacing || pl->dev.cp_extra)
        return fail(OFDM_EINVAL, "ofdm_rx_block_create: the live blocks have no hypothesis bank, comb pilots or CP pattern%s");
    DevGuard dev_guard(pl);
    ofdm_rx_block *b = new (std::nothrow) ofdm_rx_block();
    if (!b) return fail(OFDM_ENOMEM, "out of host memory%s");
    memset(b, 0, sizeof(*b));
    b->pl = pl;
    const ofdm_params &p = pl->prm;
    BlockGeom &g = b->g;
    g.mode = mode;
    g.N = p.nfft, g.CP = p.cp_len, g.L = p.nfft + p.cp_len, g.m0 = p.m0, g.d = p.d, g.fsyms = p.m0 + p.d;
    g.stride = mode == 0 ? 1 : (p.cp_len - 1 > 0 ? p.cp_len - 1 : 1);        // SynchAndChanEst.py:77 / synch_and_chan_est.py:81
    g.rows = rows;
    g.max_lock = mode == 0 ? 1 : 100;                                        // max_num_corr, synch_and_chan_est.py:85
    g.max_sym = mode == 0 ? rows : 100;
    g.thresh = pl->dev.gate_thresh;
    b->lay = ctl_layout(g);
    b->h_rows = mode == 0 ? rows : 100;
    const int mm = pl->dev.mm, ndb = pl->dev.ndd;
    cudaError_t e = cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_ctl, b->lay.bytes);
    if (e == cudaSuccess) e = cudaMemset(b->d_ctl, 0, b->lay.bytes);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_H, (size_t)b->h_rows * g.N * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemset(b->d_H, 0, (size_t)b->h_rows * g.N * sizeof(float2));
    if (e == cudaSuccess) e = cudaHostAlloc(&b->p_ctl, b->lay.bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&b->p_aux, (size_t)g.max_lock * (g.N + mm) * sizeof(float2), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc(&b->p_eq, (size_t)g.max_sym * ndb * sizeof(float2), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        CtlHeader h0 = {-1, 0, 0, -1};                                       // corr_obs = -1: SynchAndChanEst.py:72
        e = cudaMemcpy(b->d_ctl, &h0, sizeof(h0), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        block_free(b);
        return cuda_fail(e, "ofdm_rx_block_create");
    }
    memset(b->p_ctl, 0, b->lay.bytes);
    *out = b;
    return OFDM_OK;
}

int ofdm_rx_block_destroy(ofdm_rx_block *b) {
    if (!b) return OFDM_OK;
    DevGuard dev_guard(b->pl);
    block_free(b);
    return OFDM_OK;
}

// The body both entry points share.  h_in: host samples, staged and uploaded (d_in == nullptr; the equalised rows go
// to the block's mapped pinned rows), or d_in: samples that already live in device memory and are read in place; then
// the rows stay in device memory (d_rows, or a buffer of the block's), est_data_freq is kept there as well, and
// d_stream (nullable) receives the block's output stream.
static int block_work(ofdm_rx_block *b, const float *h_in, const float *d_in, int64_t n, float *d_rows, float *d_stream,
                      int64_t pack_rows, int64_t out_capacity, void *caller_stream, ofdm_rx_block_result *res,
                      int64_t *n_written) {
    ofdm_plan *pl = b->pl;
    DevGuard dev_guard(pl);
    const BlockGeom &g = b->g;
    const int ndb = pl->dev.ndd;
    int rc = block_reserve(b, n > 0 ? n : 1, d_in == nullptr);
    if (rc) return rc;
    const float2 *x = d_in ? reinterpret_cast<const float2 *>(d_in) : b->d_x;
    float2 *eq_rows = b->p_eq;
    if (d_in) {
        if (!b->d_edf) {
            const size_t bytes = (size_t)b->h_rows * ndb * sizeof(float2);
            CU(cudaMalloc(&b->d_edf, bytes), "ofdm_rx_block_work_device(cudaMalloc)");
            CU(cudaMemsetAsync(b->d_edf, 0, bytes, b->st), "ofdm_rx_block_work_device(memset)");   // np.zeros, :93 / :97
        }
        if (!d_rows && !b->d_eq)
            CU(cudaMalloc(&b->d_eq, (size_t)g.max_sym * ndb * sizeof(float2)), "ofdm_rx_block_work_device(cudaMalloc)");
        eq_rows = d_rows ? reinterpret_cast<float2 *>(d_rows) : b->d_eq;
        // whatever the caller queued on its stream (the upstream block's kernel writing d_in, the previous consumer of
        // d_out) comes first
        if (!b->ev) CU(cudaEventCreateWithFlags(&b->ev, cudaEventDisableTiming), "ofdm_rx_block_work_device(event)");
        CU(cudaEventRecord(b->ev, (cudaStream_t)caller_stream), "ofdm_rx_block_work_device(record)");
        CU(cudaStreamWaitEvent(b->st, b->ev, 0), "ofdm_rx_block_work_device(wait)");
    }
    // trials P = 0.. that pass the window guard M0*L + P*stride + N + CP < n  (SynchAndChanEst.py:139,:144)
    long long nv = 0;
    {
        const long long n_trials = (long long)nearbyint((double)n / (double)g.stride);     // np.around: half to even
        const long long room = n - (long long)g.m0 * g.L - g.N - g.CP;
        if (room > 0) {
            nv = (room - 1) / g.stride + 1;
            if (nv > n_trials) nv = n_trials;
        }
    }
    if (n > 0 && !d_in) {
        memcpy(b->p_in, h_in, (size_t)n * sizeof(float2));
        CU(cudaMemcpyAsync(b->d_x, b->p_in, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, b->st), "ofdm_rx_block_work(H2D)");
    }
    if (nv > 0) {
        rc = ofdm_rx_sync_search(pl, (const float *)x, n, g.CP, g.stride, nv, b->d_peak, b->d_delay, b->st);
        if (rc) return rc;
    }
    block_control_kernel<<<1, 256, 0, b->st>>>(g, b->lay, b->d_peak, b->d_delay, (int)nv, (long long)n, b->d_ctl);
    count_launch();
    CU(cudaGetLastError(), "ofdm_rx_block_work(control)");
    const int *d_n_new = &reinterpret_cast<const CtlHeader *>(b->d_ctl)->n_new;
    const int *d_n_dm = &reinterpret_cast<const CtlHeader *>(b->d_ctl)->n_demod;
    float2 *aux = b->p_aux;                                  // mapped: the device writes host memory directly (UVA)
    CU(pl->ks->chanest(pl->dev, x, n, reinterpret_cast<const long long *>(b->d_ctl + b->lay.new_pos),
                       reinterpret_cast<const int *>(b->d_ctl + b->lay.new_delay), nullptr, g.max_lock, b->d_H, aux,
                       aux + (size_t)g.max_lock * g.N, b->st, pl->sms,
                       reinterpret_cast<const int *>(b->d_ctl + b->lay.new_row), d_n_new),
       "ofdm_rx_block_work(chanest)");
    CU(pl->ks->rx_demod(pl->dev, x, n, reinterpret_cast<const long long *>(b->d_ctl + b->lay.dm_start),
                        reinterpret_cast<const int *>(b->d_ctl + b->lay.dm_delay),
                        reinterpret_cast<const int *>(b->d_ctl + b->lay.dm_hrow), nullptr, b->d_H, g.max_sym, eq_rows, nullptr,
                        nullptr, b->st, pl->sms, d_n_dm),
       "ofdm_rx_block_work(demod)");
    CU(cudaMemcpyAsync(b->p_ctl, b->d_ctl, b->lay.bytes, cudaMemcpyDeviceToHost, b->st), "ofdm_rx_block_work(D2H)");
    if (d_in) {
        block_scatter_kernel<<<g.max_sym, 64, 0, b->st>>>(b->d_ctl, b->lay, eq_rows, b->d_edf, ndb);
        count_launch();
        CU(cudaGetLastError(), "ofdm_rx_block_work_device(scatter)");
        long long n_elems = 0;
        if (d_stream) {
            // data_out[0:m], m = min(len(out), data_out.size): SynchAndChanEst.py:250-259 / synch_and_chan_est.py:253-263
            long long k = pack_rows < 0 ? 0 : pack_rows;
            const long long have = kept_rows_total(g, b->h_rows);
            if (k > have) k = have;
            n_elems = k * ndb;
            if (n_elems > out_capacity) n_elems = out_capacity;
            if (n_elems > 0) {
                const int grid = (int)((n_elems + 255) / 256 < 4 * pl->sms ? (n_elems + 255) / 256 : 4 * pl->sms);
                block_pack_kernel<<<grid, 256, 0, b->st>>>(b->d_edf, reinterpret_cast<float2 *>(d_stream), n_elems, ndb,
                                                           g.mode, g.fsyms);
                count_launch();
                CU(cudaGetLastError(), "ofdm_rx_block_work_device(pack)");
            }
        }
        if (n_written) *n_written = n_elems;
    }
    CU(cudaStreamSynchronize(b->st), "ofdm_rx_block_work(sync)");
    const CtlHeader *hd = reinterpret_cast<const CtlHeader *>(b->p_ctl);
    res->n_new_locks = hd->n_new;
    res->n_demod = hd->n_demod;
    res->obs_after_search = hd->obs_after_a;
    res->n_state_rows = g.max_lock;
    res->time_synch_ref = reinterpret_cast<const int64_t *>(b->p_ctl + b->lay.tsr);
    res->new_rows = reinterpret_cast<const int32_t *>(b->p_ctl + b->lay.new_row);
    res->chan_time = reinterpret_cast<const float *>(b->p_aux);
    res->synch_freq = reinterpret_cast<const float *>(b->p_aux + (size_t)g.max_lock * g.N);
    res->demod_rows = reinterpret_cast<const int32_t *>(b->p_ctl + b->lay.dm_row);
    res->data_freq = reinterpret_cast<const float *>(eq_rows);
    return OFDM_OK;
}

int ofdm_rx_block_work(ofdm_rx_block *b, const float *h_in, int64_t n, ofdm_rx_block_result *res) {
    if (!b || !res || n < 0 || (n > 0 && !h_in)) return fail(OFDM_EINVAL, "ofdm_rx_block_work: bad argument%s");
    return block_work(b, h_in, nullptr, n, nullptr, nullptr, 0, 0, nullptr, res, nullptr);
}

int ofdm_rx_block_work_device(ofdm_rx_block *b, const float *d_in, int64_t n, float *d_rows, float *d_stream_out,
                              int64_t pack_rows, int64_t out_capacity, void *stream, ofdm_rx_block_result *res,
                              int64_t *n_written) {
    if (!b || !res || n < 1 || !d_in || (d_stream_out && (pack_rows < 0 || out_capacity < 0)))
        return fail(OFDM_EINVAL, "ofdm_rx_block_work_device: bad argument%s");
    if (n_written) *n_written = 0;
    return block_work(b, nullptr, d_in, n, d_rows, d_stream_out, pack_rows, out_capacity, stream, res, n_written);
}

int ofdm_rx_block_max_rows(const ofdm_rx_block *b) { return b ? b->g.max_sym : OFDM_EINVAL; }

int ofdm_rx_block_chan_freq(ofdm_rx_block *b, float *h_H, int32_t rows) {
    if (!b || !h_H || rows < 0 || rows > b->h_rows) return fail(OFDM_EINVAL, "ofdm_rx_block_chan_freq: bad argument%s");
    DevGuard dev_guard(b->pl);
    CU(cudaMemcpyAsync(h_H, b->d_H, (size_t)rows * b->g.N * sizeof(float2), cudaMemcpyDeviceToHost, b->st), "ofdm_rx_block_chan_freq");
    CU(cudaStreamSynchronize(b->st), "ofdm_rx_block_chan_freq(sync)");
    return OFDM_OK;
}

}  // extern "C"
