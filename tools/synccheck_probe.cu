// Which construct ahead of the first block barrier does compute-sanitizer's synccheck report as "divergent thread(s)
// in warp"?  One kernel per variant, same prologue shapes as lin_umma_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I show-attend-and-tell_b200/csrc tools/synccheck_probe.cu -o tools/synccheck_probe
//   compute-sanitizer --tool synccheck tools/synccheck_probe <variant>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "sat_common.cuh"
using namespace sat;

template <int V>
__global__ void __launch_bounds__(320, 1) probe(int* out, int S) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (V == 1 || V == 2 || V == 4 || V == 5) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
            fence_mbar_init();
        }
    } else if (V == 3) {
        if (warp == 0) {
            if (lane < S) mbar_init(&bars[lane], 1);
            fence_mbar_init();
            __syncwarp();
        }
    }
    if (V >= 2 && V != 6) {
        if (warp == 1) { tmem_alloc(tptr, 32); tmem_relinquish(); }
        tc_fence_before();
    }
    if (V == 4) asm volatile("barrier.sync 0;" ::: "memory");
    else __syncthreads();
    if (V >= 2 && V != 6) {
        tc_fence_after();
        const uint32_t t = *tptr;
        __syncthreads();
        if (warp == 1) tmem_dealloc(t, 32);
    }
    if (threadIdx.x == 0) out[blockIdx.x] = S;
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 1;
    int* out = nullptr;
    cudaMalloc(&out, 4096);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(8);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = 4096;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = v == 5 ? 1 : 0;
    cudaError_t e = cudaSuccess;
    switch (v) {
        case 1: e = cudaLaunchKernelEx(&cfg, probe<1>, out, 4); break;
        case 2: e = cudaLaunchKernelEx(&cfg, probe<2>, out, 4); break;
        case 3: e = cudaLaunchKernelEx(&cfg, probe<3>, out, 4); break;
        case 4: e = cudaLaunchKernelEx(&cfg, probe<4>, out, 4); break;
        case 5: e = cudaLaunchKernelEx(&cfg, probe<5>, out, 4); break;
        default: e = cudaLaunchKernelEx(&cfg, probe<6>, out, 4); break;
    }
    cudaError_t s = cudaDeviceSynchronize();
    printf("variant %d: launch %s, sync %s\n", v, cudaGetErrorString(e), cudaGetErrorString(s));
    return 0;
}
