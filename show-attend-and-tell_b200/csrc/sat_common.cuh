// sat_common.cuh — sm_100a PTX wrappers shared by the sat_b200 kernels:
// mbarrier, TMA (cp.async.bulk / cp.async.bulk.tensor), tcgen05 (TMEM alloc,
// UMMA issue, commit, TMEM load), proxy fences.  Hand-written inline PTX; no
// CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#ifndef SAT_SPIN_LIMIT_CYCLES
// A barrier wait that lasts longer than this many SM cycles (~4 s) is a bug:
// trap instead of hanging the GPU box.
#define SAT_SPIN_LIMIT_CYCLES (8000000000ll)
#endif

namespace sat {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time; a polling loop that also
// watches something else wants test_wait)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SAT_SPIN_LIMIT_CYCLES) {
            printf("sat_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------- TMA
// 1-D bulk copy global -> shared, completion on an mbarrier (UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Bring a span of global memory into L2 without a destination (no shared memory, no barrier): used by kernels that
// were launched early (programmatic dependent launch) to pull their weight stream towards the SMs while the
// predecessor is still in its tail and the HBM channels are idle.
__device__ __forceinline__ void prefetch_l2_bulk(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// L2 eviction-priority policies for TMA loads: 0 = none, 1 = evict_first (streamed once per step),
// 2 = evict_last (weights: keep resident in the 126 MB L2 across decode steps), 3 = evict_normal.
__device__ __forceinline__ uint64_t l2_policy(int kind) {
    uint64_t pol = 0;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 3) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                  int kind, uint64_t pol) {
    if (kind == 0) {
        tma_bulk_g2s(dst_smem, src_gmem, bytes, bar);
        return;
    }
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_tensor2d_g2s_hint(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar,
                                                      int kind, uint64_t pol);

// 2-D tiled tensor copy global -> shared through a CUtensorMap (UTMALDG).
__device__ __forceinline__ void tma_tensor2d_g2s(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_tensor2d_g2s_hint(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar,
                                                      int kind, uint64_t pol) {
    if (kind == 0) {
        tma_tensor2d_g2s(dst_smem, tmap, c0, c1, bar);
        return;
    }
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / UMMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// generic-proxy global writes (possibly by other SMs, already acquired) -> visible to TMA reads
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// One lane of a CONVERGED warp (elect.sync).  The async-unit instructions (cp.async.bulk, tcgen05.mma / commit) take their
// operands from the uniform register file: issued from `if (lane == 0)` code their addresses live in per-thread registers
// and every instruction costs a handful of register->uniform moves plus an elect/branch loop (~70 cycles per MMA measured
// on the chained dense launch); issued as `if (elect_one()) ...` from a loop the whole warp runs, the operands are
// computed in uniform registers to begin with.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs / fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued UMMAs of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// TMEM -> registers: 32 lanes x 32 bit, 16 consecutive columns per thread.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100 format, version field = 1).
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
//   bits [61,64) layout: 0 = no swizzle (interleave), 2 = 128B swizzle
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   bit 15 a_major=0 (K)  bit 16 b_major=0 (K)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------ programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in
// the stream is still draining; it must not touch memory the predecessor (or anything before it) writes, nor
// write anything they read, before pdl_wait() returns (= all prerequisite grids complete and flushed).
// Every kernel here calls pdl_launch_dependents() only after its own pdl_wait(): a kernel therefore never starts
// before the predecessor of its predecessor has completed, and may read that older data without waiting.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------ thread-block clusters / distributed shared memory
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split form for a barrier that only orders "I am done reading your shared memory" (no data is published)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsmem_map(uint32_t local_smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cluster_addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(cluster_addr)
                 : "memory");
    return v;
}

// mbarrier signalling between the CTAs of a cluster (no barrier.cluster: only the threads that need it take part)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// arrive on an mbarrier in the shared memory of CTA `cta_rank` of this cluster; release at cluster scope: the writes of
// this thread (and, through a preceding bar.sync, of its CTA) are visible to whoever acquires the barrier's phase
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta_rank) {
    const uint32_t raddr = dsmem_map(smem_u32(local_bar), cta_rank);
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    long long t0 = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if ((++spins & 255) == 0) {
            if (spins == 256) t0 = clock64();
            else if (clock64() - t0 > SAT_SPIN_LIMIT_CYCLES) {
                printf("sat_b200: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
                __trap();
            }
        }
    }
}

// in-kernel timeline stamps (debug option "trace"): ns since an arbitrary origin, one row of 16 per CTA
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_stamp(unsigned long long* dbg, int slot) {
    if (dbg) dbg[(size_t)blockIdx.x * 16 + slot] = globaltimer_ns();
}

// loop timeline (debug option "trace" = 3): per launch four cells: min CTA start, max CTA end, min "go" (first
// CTA past its dependency wait), max "main loop done" (accumulator complete / streaming complete)
__device__ __forceinline__ void tl_begin(unsigned long long* tl) {
    if (tl) atomicMin(tl, globaltimer_ns());
}
__device__ __forceinline__ void tl_end(unsigned long long* tl) {
    if (tl) atomicMax(tl + 1, globaltimer_ns());
}
__device__ __forceinline__ void tl_go(unsigned long long* tl) {
    if (tl) atomicMin(tl + 2, globaltimer_ns());
}
__device__ __forceinline__ void tl_main_done(unsigned long long* tl) {
    if (tl) atomicMax(tl + 3, globaltimer_ns());
}

// --------------------------------------------------------------- misc math
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// Activations of the fused dense epilogues.  Those epilogues run once per launch from a cold instruction cache,
// so they are built from the short ex2/rcp forms: a handful of instructions each, absolute error ~2e-7 (the
// parity budget of the decode path is 1e-3 relative).
__device__ __forceinline__ float act_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float act_tanh(float x) {
    const float e = __expf(-2.0f * fabsf(x));                 // in (0, 1]: no overflow, no cancellation blow-up
    return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// spin until a monotonic arrival counter has reached `target` (wrap-safe); traps instead of hanging the GPU box
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p);
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target, const char* what) {
    if ((int)(ld_acquire_gpu(ctr) - target) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu(ctr) - target) < 0) {
        if (clock64() - t0 > SAT_SPIN_LIMIT_CYCLES) {
            printf("sat_b200: %s timed out (block %d)\n", what, (int)blockIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace sat
