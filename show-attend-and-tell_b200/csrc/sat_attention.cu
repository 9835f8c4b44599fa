// sat_attention.cu — fused soft attention of one decode step:
//   e[b,l]   = sum_a w2[a] * (T1[img(b),l,a] + q[b,a])        (attend, model.py:427-434)
//   alpha    = softmax_l(e)                                    (model.py:435)
//   z[b,:]   = sum_l alpha[b,l] * ctx[img(b),l,:]              (model.py:263-264)
// in ONE persistent kernel.  T1 = tanh(ctx*W1a + b1a) is the step-invariant feature
// branch (model.py:417-420) produced once per image batch by sat_prepare_contexts;
// q = tanh(h*W1b + b1b) is the state branch (model.py:421-424).  The 1-layer scorer
// (model.py:401-414) runs through the same kernel with T = ctx, vec = fc_a kernel,
// q = null and eadd = h*fc_b.
//
// The kernel is HBM-bound (it streams T1 and ctx once: 4*[B*L*(D+A)] bytes), so
// the work is cut so that every SM pulls the same number of bytes:
//   phase 1: the NI*L rows of T are split into equal contiguous ranges, one per CTA;
//            chunks of rows arrive by 1-D bulk TMA; one warp per row, warp-shuffle
//            reduction; e goes to global, a per-image row counter is released.
//   phase 2: the NI*(D/32) items (image, 32-wide feature slice) are split into
//            equal contiguous ranges; each item is a [L x 32] box fetched by 2-D
//            tensor TMA; the CTA waits (acquire) for the image's row counter,
//            recomputes the softmax of that image's G rows and forms z for the slice.
// Both phases share one ring of shared-memory slots fed by a single producer
// thread that runs ahead across the phase boundary (phase-2 boxes do not depend
// on phase-1 results), so ctx tiles are already resident when the softmax inputs
// arrive.  G rows (beams) of one image share the image's T1/ctx traffic.
// The grid never exceeds the SM count and is launched cooperatively, so all
// CTAs are co-resident and the counter wait cannot deadlock.
#include "sat_common.cuh"
#include "sat_attention.cuh"

namespace sat {

constexpr int kAttConsumerWarps = 8;
constexpr int kAttThreads = (kAttConsumerWarps + 1) * 32;

template <int G, int RV>
__global__ void __launch_bounds__(kAttThreads, 1)
att_fused_kernel(const __grid_constant__ CUtensorMap ctx_map, const __grid_constant__ AttParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [slots][barriers 2*nslots*8][vec RL][q G*RL][alpha G*Lp][red 2*8*G*32]
    uint8_t* slots = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.nslots * p.slot_bytes);
    uint64_t* empty = full + p.nslots;
    float* vec_s = reinterpret_cast<float*>(empty + p.nslots);
    float* q_s = vec_s + p.RL;
    const int Lp = (p.L + 3) & ~3;
    float* alpha_s = q_s + (size_t)G * p.RL;
    float* red = alpha_s + (size_t)G * Lp;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int P = gridDim.x, c = blockIdx.x;
    const int L = p.L, RL = p.RL;
    const int nds = p.D / 32;
    const long long NR = (long long)p.NI * L;
    const long long NIt = (long long)p.NI * nds;
    const int r_begin = (int)(NR * c / P), r_end = (int)(NR * (c + 1) / P);
    const int i_begin = (int)(NIt * c / P), i_end = (int)(NIt * (c + 1) / P);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nslots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kAttConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kAttConsumerWarps) {
        // ============================ producer ============================
        if (lane == 0) {
            tma_prefetch_desc(&ctx_map);
            const uint64_t pol_t = l2_policy(p.l2_t), pol_c = l2_policy(p.l2_ctx);
            int idx = 0;
            for (int r = r_begin; r < r_end;) {
                const int img = r / L;
                int n = min(p.rch, r_end - r);
                n = min(n, (img + 1) * L - r);
                const int s = idx % p.nslots;
                const uint32_t ph = (uint32_t)(idx / p.nslots) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                const uint32_t bytes = (uint32_t)n * RL * 4u;
                mbar_arrive_expect_tx(&full[s], bytes);
                tma_bulk_g2s_hint(slots + (size_t)s * p.slot_bytes, p.T + (size_t)r * RL, bytes, &full[s], p.l2_t, pol_t);
                r += n;
                ++idx;
            }
            for (int it = i_begin; it < i_end; ++it, ++idx) {
                const int img = it / nds, ds = it - img * nds;
                const int s = idx % p.nslots;
                const uint32_t ph = (uint32_t)(idx / p.nslots) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], (uint32_t)L * 128u);
                tma_tensor2d_g2s_hint(slots + (size_t)s * p.slot_bytes, &ctx_map, ds * 32, img * L, &full[s], p.l2_ctx,
                                      pol_c);
            }
        }
        return;
    }

    // ============================== consumers ==============================
    const int ct = threadIdx.x;  // 0..255
    constexpr int NT = kAttConsumerWarps * 32;
    for (int j = ct; j < RL; j += NT) vec_s[j] = p.vec[j];
    int cur_q_img = -1, cur_a_img = -1;
    int idx = 0;
    int img_lo = -1, img_hi = -1;   // images whose rows this CTA scored (for the release at the end of phase 1)
    int cnt_lo = 0, cnt_mid_first = 0;
    (void)cnt_lo; (void)cnt_mid_first;

    // ---------------- phase 1: attention logits ----------------
    // RV > 0: the row length is RV*128 floats and w2 / q live in registers (one float4 per lane per 128 floats)
    float4 wreg[RV > 0 ? RV : 1];
    float4 qreg[G][RV > 0 ? RV : 1];
    for (int r = r_begin; r < r_end;) {
        const int img = r / L;
        int n = min(p.rch, r_end - r);
        n = min(n, (img + 1) * L - r);
        const int s = idx % p.nslots;
        const uint32_t ph = (uint32_t)(idx / p.nslots) & 1u;
        if (img != cur_q_img) {
            if (cur_q_img < 0) {
                named_bar_sync(1, NT);  // vec_s visible
                if (RV > 0) {
#pragma unroll
                    for (int k = 0; k < (RV > 0 ? RV : 1); ++k) wreg[k] = reinterpret_cast<const float4*>(vec_s)[lane + 32 * k];
                }
                img_lo = img;
            }
            if (RV > 0) {
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int k = 0; k < (RV > 0 ? RV : 1); ++k)
                        qreg[g][k] = p.q ? __ldg(reinterpret_cast<const float4*>(p.q + ((size_t)img * G + g) * RL) + lane + 32 * k)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                if (cur_q_img >= 0) named_bar_sync(1, NT);  // everyone done with the previous image's q_s
                if (p.q) {
                    const float4* src = reinterpret_cast<const float4*>(p.q + (size_t)img * G * RL);
                    float4* dst = reinterpret_cast<float4*>(q_s);
                    for (int j = ct; j < G * RL / 4; j += NT) dst[j] = src[j];
                }
                named_bar_sync(1, NT);
            }
            cur_q_img = img;
            img_hi = img;
        }
        mbar_wait(&full[s], ph);
        const float* buf = reinterpret_cast<const float*>(slots + (size_t)s * p.slot_bytes);
        for (int row = warp; row < n; row += kAttConsumerWarps) {
            float acc[G];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] = 0.f;
            const float4* trow = reinterpret_cast<const float4*>(buf + (size_t)row * RL);
            if (RV > 0) {
#pragma unroll
                for (int k = 0; k < (RV > 0 ? RV : 1); ++k) {
                    const float4 t = trow[lane + 32 * k];
                    const float4 w = wreg[k];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float4 qq = qreg[g][k];
                        acc[g] = fmaf(w.x, t.x + qq.x, acc[g]);
                        acc[g] = fmaf(w.y, t.y + qq.y, acc[g]);
                        acc[g] = fmaf(w.z, t.z + qq.z, acc[g]);
                        acc[g] = fmaf(w.w, t.w + qq.w, acc[g]);
                    }
                }
            } else {
                for (int j = lane; j < RL / 4; j += 32) {
                    const float4 t = trow[j];
                    const float4 w = reinterpret_cast<const float4*>(vec_s)[j];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        float4 qq = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.q) qq = reinterpret_cast<const float4*>(q_s + (size_t)g * RL)[j];
                        acc[g] = fmaf(w.x, t.x + qq.x, acc[g]);
                        acc[g] = fmaf(w.y, t.y + qq.y, acc[g]);
                        acc[g] = fmaf(w.z, t.z + qq.z, acc[g]);
                        acc[g] = fmaf(w.w, t.w + qq.w, acc[g]);
                    }
                }
            }
            const int l = r + row - img * L;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float sum = warp_sum(acc[g]);
                if (lane == 0) {
                    const size_t o = ((size_t)img * G + g) * L + l;
                    if (p.eadd) sum += p.eadd[o];
                    p.e[o] = sum;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        r += n;
        ++idx;
    }
    // release the logits: ONE gpu-scope fence per CTA (not per chunk), then per-image row counts
    if (r_begin < r_end) {
        named_bar_sync(1, NT);
        if (ct == 0) {
            __threadfence();
            for (int img = img_lo; img <= img_hi; ++img) {
                const int a = max(r_begin, img * L), b = min(r_end, (img + 1) * L);
                if (b > a) atomicAdd(p.rowcnt + img, (unsigned)(b - a));
            }
        }
    }

    // ---------------- phase 2: softmax + context vector ----------------
    // item = [L x 32] floats; lanes 0-7 / 8-15 / 16-23 / 24-31 read four consecutive rows as float4
    const int sub = lane >> 3, l4 = lane & 7;
    int parity = 0;
    for (int it = i_begin; it < i_end; ++it, ++idx) {
        const int img = it / nds, ds = it - img * nds;
        const int s = idx % p.nslots;
        const uint32_t ph = (uint32_t)(idx / p.nslots) & 1u;
        if (img != cur_a_img) {
            if (ct == 0) {
                const long long t0 = clock64();
                while (ld_acquire_gpu(p.rowcnt + img) < p.target) {
                    if (clock64() - t0 > SAT_SPIN_LIMIT_CYCLES) {
                        printf("sat_b200: attention row-counter wait timed out (block %d img %d)\n", c, img);
                        __trap();
                    }
                }
            }
            named_bar_sync(1, NT);
            for (int g = warp; g < G; g += kAttConsumerWarps) {
                const float* er = p.e + ((size_t)img * G + g) * L;
                float m = -INFINITY;
                for (int l = lane; l < L; l += 32) m = fmaxf(m, __ldcg(er + l));
                m = warp_max(m);
                float sum = 0.f;
                for (int l = lane; l < L; l += 32) {
                    const float ex = expf(__ldcg(er + l) - m);
                    alpha_s[g * Lp + l] = ex;
                    sum += ex;
                }
                sum = warp_sum(sum);
                const float inv = 1.0f / sum;
                for (int l = lane; l < L; l += 32) {
                    const float a = alpha_s[g * Lp + l] * inv;
                    alpha_s[g * Lp + l] = a;
                    if (ds == 0) p.alpha[((size_t)img * G + g) * L + l] = a;
                }
            }
            cur_a_img = img;
            named_bar_sync(1, NT);
        }
        mbar_wait(&full[s], ph);
        const float4* buf4 = reinterpret_cast<const float4*>(slots + (size_t)s * p.slot_bytes);
        float4 acc[G];
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = warp * 4 + sub; l < L; l += kAttConsumerWarps * 4) {
            const float4 x = buf4[l * 8 + l4];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float a = alpha_s[g * Lp + l];
                acc[g].x = fmaf(a, x.x, acc[g].x);
                acc[g].y = fmaf(a, x.y, acc[g].y);
                acc[g].z = fmaf(a, x.z, acc[g].z);
                acc[g].w = fmaf(a, x.w, acc[g].w);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        float4* rp4 = reinterpret_cast<float4*>(red + (size_t)parity * kAttConsumerWarps * G * 32);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float4 v = acc[g];
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
                v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
            }
            if (sub == 0) rp4[(warp * G + g) * 8 + l4] = v;
        }
        named_bar_sync(1, NT);
        const float* rp = reinterpret_cast<const float*>(rp4);
        for (int g = warp; g < G; g += kAttConsumerWarps) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < kAttConsumerWarps; ++w) sum += rp[(w * G + g) * 32 + lane];
            p.z[((size_t)img * G + g) * p.D + ds * 32 + lane] = sum;
        }
        parity ^= 1;
    }
}

size_t att_smem_bytes(const AttParams& p) {
    const int Lp = (p.L + 3) & ~3;
    return (size_t)p.nslots * p.slot_bytes + 2 * (size_t)p.nslots * 8 +
           4 * ((size_t)p.RL + (size_t)p.G * p.RL + (size_t)p.G * Lp + 2 * (size_t)kAttConsumerWarps * p.G * 32);
}

// Fill in chunking / ring parameters from the device limits.  Returns false if the shape is unsupported.
bool att_plan(AttParams& p, int smem_optin) {
    if (p.G < 1 || p.G > 4 || p.L < 1 || p.L > 256 || (p.D % 32) || (p.RL % 4)) return false;
    const int box = p.L * 128;
    int rch = box / (p.RL * 4);
    if (rch < 1) rch = 1;
    int slot = rch * p.RL * 4;
    if (slot < box) slot = box;
    slot = (slot + 127) & ~127;
    p.rch = rch;
    p.slot_bytes = slot;
    AttParams t = p;
    t.nslots = 0;
    const size_t fixed = att_smem_bytes(t) + 64;
    if ((size_t)smem_optin < fixed + 2 * (size_t)slot) return false;
    int n = (int)(((size_t)smem_optin - fixed) / ((size_t)slot + 16));
    if (n > 16) n = 16;
    p.nslots = n;
    return n >= 2;
}

template <int G, int RV>
static cudaError_t att_launch_gr(const CUtensorMap& map, const AttParams& p, int grid, cudaStream_t st, bool coop) {
    const size_t smem = att_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(att_fused_kernel<G, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kAttThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = coop ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, att_fused_kernel<G, RV>, map, p);
}

template <int G>
static cudaError_t att_launch_g(const CUtensorMap& map, const AttParams& p, int grid, cudaStream_t st, bool coop) {
    // register-resident w2/q when a row is exactly 512 floats (dim_attend_layer = 512, the reference default)
    if (p.RL == 512) return att_launch_gr<G, 4>(map, p, grid, st, coop);
    return att_launch_gr<G, 0>(map, p, grid, st, coop);
}

cudaError_t att_launch(const CUtensorMap& map, const AttParams& p, int num_sms, cudaStream_t st, bool coop) {
    long long items = (long long)p.NI * (p.D / 32);
    int grid = (int)(items < num_sms ? items : num_sms);
    if (grid < 1) grid = 1;
    switch (p.G) {
        case 1: return att_launch_g<1>(map, p, grid, st, coop);
        case 2: return att_launch_g<2>(map, p, grid, st, coop);
        case 3: return att_launch_g<3>(map, p, grid, st, coop);
        case 4: return att_launch_g<4>(map, p, grid, st, coop);
    }
    return cudaErrorInvalidValue;
}

// mean over the L locations (model.py:240): out[i, d] = (1/L) sum_l ctx[i, l, d]
__global__ void ctx_mean_kernel(const float* __restrict__ ctx, float* __restrict__ out, int L, int D) {
    const int i = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const float* p = ctx + (size_t)i * L * D + d;
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += p[(size_t)l * D];
    out[(size_t)i * D + d] = s / (float)L;
}

cudaError_t ctx_mean_launch(const float* ctx, float* out, int NI, int L, int D, cudaStream_t st) {
    dim3 grid((D + 127) / 128, NI);
    ctx_mean_kernel<<<grid, 128, 0, st>>>(ctx, out, L, D);
    return cudaGetLastError();
}

}  // namespace sat
