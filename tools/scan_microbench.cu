// scan_microbench.cu -- standalone timing harness for scan_kernel<21, canonical> variants (kernel tuning only).
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo [-DMG_*_FMA=0/1] -I mash_b200/csrc tools/scan_microbench.cu -o mb
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "scan.cuh"
using namespace mashgpu;

__global__ void fill_kernel(uint8_t *p, uint64_t n, uint64_t seed)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ULL; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
        p[i] = "ACGT"[x & 3];
    }
}

int main(int argc, char **argv)
{
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 4000000000ull;
    uint8_t *d; cudaMalloc(&d, n + 64);
    fill_kernel<<<148 * 8, 256>>>(d, n, 12345);
    const uint64_t unit_start_h[2] = {0, n};
    uint64_t *unit_start, *unit_t, *tab_off, *keys; uint32_t *log2c, *cnt, *flags, *maxh;
    cudaMalloc(&unit_start, 16); cudaMalloc(&unit_t, 8); cudaMalloc(&tab_off, 8); cudaMalloc(&log2c, 4); cudaMalloc(&flags, 4); cudaMalloc(&maxh, 4);
    const uint32_t lg = 23; const uint64_t cap = 1ull << lg;
    cudaMalloc(&keys, cap * 8); cudaMalloc(&cnt, cap * 4);
    cudaMemset(keys, 0xFF, cap * 8); cudaMemset(cnt, 0, cap * 4); cudaMemset(flags, 0, 4); cudaMemset(maxh, 0, 4);
    const uint64_t t = (uint64_t)(3000.0 / 5e6 * 18446744073709551616.0), zero = 0;
    cudaMemcpy(unit_start, unit_start_h, 16, cudaMemcpyHostToDevice); cudaMemcpy(unit_t, &t, 8, cudaMemcpyHostToDevice);
    cudaMemcpy(tab_off, &zero, 8, cudaMemcpyHostToDevice); cudaMemcpy(log2c, &lg, 4, cudaMemcpyHostToDevice);
    ScanArgs a; memset(&a, 0, sizeof a);
    a.stream = d; a.stream_len = n; a.tile_begin = 0; a.tile_end = (n + SCAN_TILE - 1) / SCAN_TILE; a.coarse_t = t;
    a.seed = 42; a.use64 = 1; a.mode = SCAN_SKETCH; a.unit_start = unit_start; a.n_units = 1; a.unit_t = unit_t; a.tab_off = tab_off;
    a.tab_log2 = log2c; a.tab_keys = keys; a.tab_cnt = cnt; a.unit_flags = flags; a.unit_maxhash = maxh; a.only_unit = -1;
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<21, true, false>, SCAN_THREADS, 0);
    const int grid = per_sm * 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) scan_kernel<21, true, false><<<grid, SCAN_THREADS>>>(a);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; i++) scan_kernel<21, true, false><<<grid, SCAN_THREADS>>>(a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    uint32_t c0 = 0; std::vector<uint32_t> hc(cap); cudaMemcpy(hc.data(), cnt, cap * 4, cudaMemcpyDeviceToHost);
    uint64_t distinct = 0, total = 0; for (auto v : hc) { distinct += v != 0; total += v; }
    (void)c0;
    printf("%s occupancy=%d/SM  %.3f ms  %.1f Gbp/s  survivors=%llu (distinct %llu) err=%s\n", VARIANT, per_sm, ms, n / ms / 1e6,
           (unsigned long long)total / (reps + 2), (unsigned long long)distinct, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
