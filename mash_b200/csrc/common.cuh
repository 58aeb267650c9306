// common.cuh -- context, error handling and small device-buffer helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <utility>

#include "../../include/mashgpu.h"

namespace mashgpu {

struct EventPair { cudaEvent_t a, b; };

}  // namespace mashgpu

namespace mashgpu {
// grow-only device scratch owned by the context (no cudaMalloc/cudaFree on the hot path after warm-up)
struct Scratch {
    void *p = nullptr;
    size_t bytes = 0;
    ~Scratch() { if (p) cudaFree(p); }
    template <typename T>
    T *get(size_t count)
    {
        size_t need = count * sizeof(T);
        if (need == 0) need = 16;
        if (need > bytes) {
            if (p) cudaFree(p);
            p = nullptr; bytes = 0;
            size_t want = need + need / 8;
            if (cudaMalloc(&p, want) != cudaSuccess) {
                cudaGetLastError();
                if (cudaMalloc(&p, need) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
                want = need;
            }
            bytes = want;
        }
        return reinterpret_cast<T *>(p);
    }
};
}  // namespace mashgpu

struct mashgpu_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D staging
    std::string err;
    bool attr_dist = false, attr_merge = false, attr_select = false;
    // instrumentation
    bool timing = false;
    uint64_t kernel_launches = 0;
    uint64_t scan_launches = 0, dist_launches = 0, exact_reruns = 0;
    double scan_ms = 0, dist_ms = 0;
    std::vector<mashgpu::EventPair> scan_events, dist_events;
    // sketch_stream_core scratch
    mashgpu::Scratch sc_start, sc_t, sc_off, sc_log2, sc_flags, sc_maxhash, sc_keys, sc_cnt, sc_tmax, sc_first, sc_last, sc_qtarget, sc_qtstar;
    // mashgpu_sketch_batch: wave stream buffers and outputs
    mashgpu::Scratch sc_wave[2], sc_inval[2], sc_runs[2], sc_out_hashes, sc_out_n, sc_out_counts;
    mashgpu::Scratch sc_sep[2], sc_codes[2];
    mashgpu::Scratch sc_big_units, sc_big_n, sc_big_off, sc_big_bounds, sc_big_comp, sc_big_sorted, sc_big_tmp;   // select of tables > 2^14 slots
    void *pinned_sep[2] = {nullptr, nullptr};
    void *pinned_codes[2] = {nullptr, nullptr};     // packed feed path: host code buffers
    size_t pinned_codes_bytes[2] = {0, 0};
    cudaEvent_t pack_copied[2] = {nullptr, nullptr};
    void *flags_pinned = nullptr;                   // per-unit status flags of the sketch pass in flight (sketch.cu)
    size_t flags_pinned_n = 0;
    cudaStream_t pack_stream = nullptr;             // packed uploads (next to the ASCII copies on copy_stream)
    // screen: chunk staging buffers and the finish() outputs outlive a job (cudaMalloc of a few hundred MB inside a timed pass cost
    // 1-50 ms depending on the box); scr_stage_owner = the job using the staging pair, others allocate their own
    mashgpu::Scratch scr_stage[2], scr_out;
    const void *scr_stage_owner = nullptr;
    void *scr_pinned[2] = {nullptr, nullptr};       // packed host feed of screen: pinned codes + mask per slot (with scr_stage)
    size_t scr_pinned_bytes[2] = {0, 0};
    void *pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes[2] = {0, 0};
    cudaEvent_t wave_copied[2] = {nullptr, nullptr};
};

namespace mashgpu {

inline int fail(mashgpu_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define MG_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return mashgpu::fail((ctx), MASHGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call,           \
                                 cudaGetErrorString(e__), __FILE__, __LINE__);                      \
    } while (0)

#define MG_TRY(expr)                      \
    do {                                  \
        int rc__ = (expr);                \
        if (rc__ != MASHGPU_OK) return rc__; \
    } while (0)

// RAII device buffer (freed on scope exit)
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count)
    {
        release();
        const cudaError_t e = cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; cudaGetLastError(); return e; }    // n stays 0: a later `n < need` test re-allocates
        n = count;
        return e;
    }
    // grow-only: keeps the buffer when it already holds `count` entries
    cudaError_t reserve(size_t count) { return (p && n >= count) ? cudaSuccess : alloc(count + count / 4); }
};

template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { release(); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count)
    {
        release();
        const cudaError_t e = cudaMallocHost((void **)&p, (count ? count : 1) * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; cudaGetLastError(); return e; }
        n = count;
        return e;
    }
};

inline uint32_t ceil_log2(uint64_t x)
{
    uint32_t l = 0;
    while ((1ull << l) < x) l++;
    return l;
}

// Begin/end a timed region for one of the dominant kernels (no-ops unless ctx->timing).
inline void time_begin(mashgpu_ctx *ctx, std::vector<EventPair> &list, cudaStream_t s)
{
    if (!ctx->timing) return;
    EventPair ev;
    cudaEventCreate(&ev.a);
    cudaEventCreate(&ev.b);
    cudaEventRecord(ev.a, s);
    list.push_back(ev);
}
inline void time_end(mashgpu_ctx *ctx, std::vector<EventPair> &list, cudaStream_t s)
{
    if (!ctx->timing) return;
    cudaEventRecord(list.back().b, s);
}

}  // namespace mashgpu
