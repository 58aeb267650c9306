// scan_inst.cu -- explicit instantiations of scan_kernel<K, CANON> for K in [SCAN_K_LO, SCAN_K_HI].
// Compiled several times with different -DSCAN_K_LO/-DSCAN_K_HI/-DSCAN_PART so the parts build in parallel.
#include "scan.cuh"

#ifndef SCAN_K_LO
#error "compile with -DSCAN_K_LO=.. -DSCAN_K_HI=.. -DSCAN_PART=.."
#endif

namespace mashgpu {

template <int K, bool CANON>
static void launch_scan(const ScanArgs &a, int grid, cudaStream_t stream)
{
    scan_kernel<K, CANON><<<grid, SCAN_THREADS, 0, stream>>>(a);
}

template <int K>
static scan_launch_fn pick(int k, bool canonical)
{
    if constexpr (K > SCAN_K_HI) {
        return nullptr;
    } else {
        if (k == K) return canonical ? &launch_scan<K, true> : &launch_scan<K, false>;
        return pick<K + 1>(k, canonical);
    }
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
scan_launch_fn CAT(get_scan_launcher_part, SCAN_PART)(int k, bool canonical)
{
    if (k < SCAN_K_LO || k > SCAN_K_HI) return nullptr;
    return pick<SCAN_K_LO>(k, canonical);
}

}  // namespace mashgpu
