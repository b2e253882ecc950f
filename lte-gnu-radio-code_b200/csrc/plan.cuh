// plan.cuh -- what the translation units behind the C ABI share: the plan object, error reporting, the device guard.
#pragma once
#include "../../include/ofdm_b200.h"
#include "launch.cuh"

namespace ofdm {
typedef float2 float2_t;
// error text returned by ofdm_last_error() (thread-local); both return the code they are given / OFDM_ECUDA
int fail(int code, const char *fmt, const char *a = "", long long b = 0);
int cuda_fail(cudaError_t e, const char *where);
}  // namespace ofdm

#define CU(call, where)                                         \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return ofdm::cuda_fail(e__, where); \
    } while (0)

struct ofdm_plan {
    ofdm_params prm;
    ofdm::DevParams dev;
    const ofdm::KernelSet *ks;
    int device, sms, log2n;
    void *d_tables;                // one allocation holding every table
    void *d_cfo;                   // frequency-offset hypothesis bank (optional)
    // staging for the host-buffer entry points (grown on demand, freed with the plan)
    static const int NSLOT = 3;
    cudaStream_t st[NSLOT];
    cudaEvent_t ev[NSLOT];
    bool streams_made;
    void *stage[NSLOT];
    size_t stage_bytes[NSLOT];
    void *small;                   // per-frame results of the host-buffer receiver (delay, peak, locked)
    size_t small_bytes;
    // overlap-save search (api.cu, rx_sync_search_os_kernel): block size 2^os_log2nb (0: not available for this plan),
    // positions per block, the block transform's twiddles and the taps' spectrum
    int os_log2nb, os_pb, os_smem_set;
    const ofdm::float2_t *d_tw_os, *d_gos;
    int tdg_smem_set;              // the same for the general (M0 = 2 / hypothesis bank) kernel
    int td_smem_set;               // dynamic shared memory the time-domain search kernel has been allowed on this device
};

// Makes the plan's device current for the duration of one entry point (and restores the caller's), so a plan created
// on cuda:1 launches there -- with that device's SM count and shared-memory attributes -- whatever device the calling
// thread happens to have current.
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(const ofdm_plan *pl) {
        if (pl && cudaGetDevice(&prev) == cudaSuccess && prev != pl->device) switched = cudaSetDevice(pl->device) == cudaSuccess;
    }
    ~DevGuard() {
        if (switched) cudaSetDevice(prev);
    }
};
