// sat_rows.cu — per-row vocabulary kernels and the device-side beam bookkeeping.
//   softmax / argmax over V            model.py:288-289
//   top-(beam+1) words per row         base_model.py:215-219
//   TopN / CaptionData heap updates    base_model.py:222-232, utils/misc.py:38-87
#include "sat_common.cuh"
#include "sat_rows.cuh"

namespace sat {

constexpr int kRowThreads = 256;

struct ValIdx {
    float v;
    int i;
};
// ordering of the reference's stable descending sort (base_model.py:217-218) and of
// tf.argmax (first maximum): larger value first, then lower index.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ ValIdx block_best(ValIdx x, ValIdx* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, x.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, x.i, o);
        if (better(ov, oi, x.v, x.i)) { x.v = ov; x.i = oi; }
    }
    __syncthreads();
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    ValIdx r = sm[0];
    for (int w = 1; w < kRowThreads / 32; ++w)
        if (better(sm[w].v, sm[w].i, r.v, r.i)) r = sm[w];
    return r;
}

__device__ __forceinline__ float block_sum(float x, float* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < kRowThreads / 32; ++w) r += sm[w];
    return r;
}

// One block per row.  The row is read from global memory ONCE into shared memory (CACHED: V % 4 == 0 and 4*V bytes
// of dynamic shared memory available) and the three passes — arg-max, sum of exponentials, probabilities + top-k
// — run on that copy in small ROLLED loops (this kernel runs once per step from a cold instruction cache: a
// fully unrolled register-cached version executed 6 k instructions per warp and was instruction-fetch bound).
template <bool CACHED>
__global__ void __launch_bounds__(kRowThreads) rows_softmax_kernel(const RowsParams p) {
    extern __shared__ __align__(16) float row_s[];
    __shared__ ValIdx sm_vi[kRowThreads / 32];
    __shared__ float sm_f[kRowThreads / 32];
    const int row = blockIdx.x;
    const float* x = p.logits + (size_t)row * p.V;
    const int V = p.V, n4 = V >> 2;
    const float4* src4 = CACHED ? reinterpret_cast<const float4*>(row_s) : reinterpret_cast<const float4*>(x);

    ValIdx best = {-INFINITY, 0x7fffffff};
    if (CACHED) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        float4* d4 = reinterpret_cast<float4*>(row_s);
#pragma unroll 2
        for (int i4 = threadIdx.x; i4 < n4; i4 += kRowThreads) {
            const float4 v = x4[i4];
            d4[i4] = v;
            const int i = 4 * i4;
            if (better(v.x, i, best.v, best.i)) { best.v = v.x; best.i = i; }
            if (better(v.y, i + 1, best.v, best.i)) { best.v = v.y; best.i = i + 1; }
            if (better(v.z, i + 2, best.v, best.i)) { best.v = v.z; best.i = i + 2; }
            if (better(v.w, i + 3, best.v, best.i)) { best.v = v.w; best.i = i + 3; }
        }
    } else {
        for (int i = threadIdx.x; i < V; i += kRowThreads) {
            const float v = x[i];
            if (better(v, i, best.v, best.i)) { best.v = v; best.i = i; }
        }
    }
    best = block_best(best, sm_vi);       // (its barriers also publish row_s)
    const float m = best.v;
    if (threadIdx.x == 0) {
        if (p.argmax) p.argmax[row] = best.i;
        if (p.tokens) p.tokens[(size_t)row * p.tokens_ld + p.step] = best.i;
        if (p.next_word) p.next_word[row] = p.forced ? p.forced[(size_t)row * p.forced_ld + p.step] : best.i;
    }
    if (!p.probs && p.topk == 0) return;  // greedy / teacher-forced loops only need the argmax

    float s = 0.f;
    if (CACHED) {
#pragma unroll 1
        for (int i4 = threadIdx.x; i4 < n4; i4 += kRowThreads) {
            const float4 v = src4[i4];
            s += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
        }
    } else {
        for (int i = threadIdx.x; i < V; i += kRowThreads) s += expf(x[i] - m);
    }
    s = block_sum(s, sm_f);
    const float inv = 1.0f / s;

    // probabilities + thread-local top-kMaxTopK, kept sorted by (prob desc, index asc) with a compare-and-swap
    // chain on registers (static indices only)
    float tv[kMaxTopK];
    int ti[kMaxTopK];
#pragma unroll
    for (int k = 0; k < kMaxTopK; ++k) { tv[k] = -1.f; ti[k] = 0x7fffffff; }
    auto offer = [&](float pv, int i) {
        if (!better(pv, i, tv[kMaxTopK - 1], ti[kMaxTopK - 1])) return;
#pragma unroll
        for (int k = 0; k < kMaxTopK; ++k) {
            if (better(pv, i, tv[k], ti[k])) {
                const float fv = tv[k]; const int fi = ti[k];
                tv[k] = pv; ti[k] = i;
                pv = fv; i = fi;
            }
        }
    };
    float* pr = p.probs ? p.probs + (size_t)row * V : nullptr;
    const int n_el = CACHED ? V : 0;
#pragma unroll 1
    for (int i = threadIdx.x; i < n_el; i += kRowThreads) {      // element-wise: one copy of `offer` in the binary
        const float pv = expf(row_s[i] - m) * inv;
        if (pr) pr[i] = pv;
        if (p.topk > 0) offer(pv, i);
    }
    if (!CACHED) {
#pragma unroll 1
        for (int i = threadIdx.x; i < V; i += kRowThreads) {
            const float pv = expf(x[i] - m) * inv;
            if (pr) pr[i] = pv;
            if (p.topk > 0) offer(pv, i);
        }
    }
    if (p.topk > 0) {
#pragma unroll 1
        for (int k = 0; k < p.topk; ++k) {
            ValIdx cnd;
            cnd.v = tv[0];
            cnd.i = ti[0];
            const ValIdx w = block_best(cnd, sm_vi);
            if (w.i == cnd.i && w.v == cnd.v && cnd.i != 0x7fffffff) {   // this thread's head won: pop it
#pragma unroll
                for (int q = 0; q + 1 < kMaxTopK; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
                tv[kMaxTopK - 1] = -1.f;
                ti[kMaxTopK - 1] = 0x7fffffff;
            }
            if (threadIdx.x == 0) {
                p.topk_idx[(size_t)row * p.topk + k] = w.i;
                p.topk_p[(size_t)row * p.topk + k] = w.v;
            }
        }
    }
}

cudaError_t rows_softmax_launch(const RowsParams& p, int rows, cudaStream_t st) {
    if (p.topk > kMaxTopK) return cudaErrorInvalidValue;
    const size_t bytes = (size_t)p.V * sizeof(float);
    const bool cached = (p.V % 4) == 0 && bytes <= 96 * 1024 && (reinterpret_cast<uintptr_t>(p.logits) % 16) == 0;
    if (cached) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(rows_softmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        rows_softmax_kernel<true><<<rows, kRowThreads, bytes, st>>>(p);
    } else {
        rows_softmax_kernel<false><<<rows, kRowThreads, 0, st>>>(p);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ beam search
// Python heapq semantics (CPython Lib/heapq.py) on tiny arrays, so that ties are
// broken exactly as utils/misc.py:62-87 (TopN) does.
struct PItem { double score; int parent; int word; };
struct CItem { double score; int slot; int len; };

template <typename T>
__device__ void heap_siftdown(T* h, int startpos, int pos) {
    T item = h[pos];
    while (pos > startpos) {
        const int parent = (pos - 1) >> 1;
        if (item.score < h[parent].score) { h[pos] = h[parent]; pos = parent; continue; }
        break;
    }
    h[pos] = item;
}
template <typename T>
__device__ void heap_siftup(T* h, int n, int pos) {
    const int startpos = pos;
    T item = h[pos];
    int child = 2 * pos + 1;
    while (child < n) {
        const int right = child + 1;
        if (right < n && !(h[child].score < h[right].score)) child = right;
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    h[pos] = item;
    heap_siftdown(h, startpos, pos);
}

__global__ void __launch_bounds__(256) beam_update_kernel(const BeamParams p) {
    __shared__ PItem newp[kMaxBeam];
    __shared__ int newn;
    __shared__ int s_idx[kMaxBeam * (kMaxBeam + 1)];
    __shared__ float s_p[kMaxBeam * (kMaxBeam + 1)];
    __shared__ double s_ps[kMaxBeam];
    const int img = blockIdx.x;
    const int beam = p.beam, K = p.beam + 1, G = p.nlive, T = p.T, idx = p.step;
    const int* sent_cur = p.sent[idx & 1] + (size_t)img * beam * T;
    int* sent_next = p.sent[(idx + 1) & 1] + (size_t)img * beam * T;

    // the candidates of this image (G live beams x K words) and the beams' scores: one parallel round trip
    if ((int)threadIdx.x < G * K) {
        s_idx[threadIdx.x] = p.topk_idx[(size_t)img * G * K + threadIdx.x];
        s_p[threadIdx.x] = p.topk_p[(size_t)img * G * K + threadIdx.x];
    }
    if ((int)threadIdx.x < G) s_ps[threadIdx.x] = idx == 0 ? 1.0 : p.part_score[(size_t)img * beam + threadIdx.x];   // base_model.py:178
    __syncthreads();

    if (threadIdx.x == 0) {
        int np = 0;
        CItem* ch = p.comp_heap + (size_t)img * beam;
        int cn = p.comp_n[img];
        for (int b = 0; b < G; ++b) {
            const double ps = s_ps[b];
            for (int j = 0; j < K; ++j) {
                const int w = s_idx[b * K + j];
                const double sc = ps * (double)s_p[b * K + j];                           // base_model.py:224
                if (w == p.eos_id) {                                                   // base_model.py:229-230
                    int slot = -1;
                    if (cn < beam) {
                        slot = cn;
                        ch[cn].score = sc; ch[cn].slot = slot; ch[cn].len = idx + 1;
                        ++cn;
                        heap_siftdown(ch, 0, cn - 1);
                    } else if (ch[0].score < sc) {
                        slot = ch[0].slot;
                        ch[0].score = sc; ch[0].len = idx + 1;
                        heap_siftup(ch, cn, 0);
                    }
                    if (slot >= 0) {
                        int* dst = p.comp_sent + ((size_t)img * beam + slot) * T;
                        for (int t = 0; t < idx; ++t) dst[t] = sent_cur[(size_t)b * T + t];
                        dst[idx] = w;
                    }
                } else {                                                               // base_model.py:231-232
                    if (np < beam) {
                        newp[np].score = sc; newp[np].parent = b; newp[np].word = w;
                        ++np;
                        heap_siftdown(newp, 0, np - 1);
                    } else if (newp[0].score < sc) {
                        newp[0].score = sc; newp[0].parent = b; newp[0].word = w;
                        heap_siftup(newp, np, 0);
                    }
                }
            }
        }
        p.comp_n[img] = cn;
        newn = np;
        p.part_n[img] = np;
        for (int j = 0; j < np; ++j) p.part_score[(size_t)img * beam + j] = newp[j].score;
    }
    __syncthreads();
    // materialise the surviving beams: sentences, last word, LSTM state rows (float4 rows, all beams in one sweep)
    const int np = newn;
    for (int u = threadIdx.x; u < np * idx; u += blockDim.x) {
        const int j = u / idx, t = u - j * idx;
        sent_next[(size_t)j * T + t] = sent_cur[(size_t)newp[j].parent * T + t];
    }
    if ((int)threadIdx.x < np) {
        sent_next[(size_t)threadIdx.x * T + idx] = newp[threadIdx.x].word;
        p.next_word[(size_t)img * beam + threadIdx.x] = newp[threadIdx.x].word;
    }
    const int H4 = p.H >> 2;   // H % 32 == 0 (sat_create)
    for (int u = threadIdx.x; u < np * H4; u += blockDim.x) {
        const int j = u / H4, q = u - j * H4;
        const int b = newp[j].parent;
        reinterpret_cast<float4*>(p.c_next + ((size_t)img * beam + j) * p.H)[q] =
            reinterpret_cast<const float4*>(p.c_out + ((size_t)img * G + b) * p.H)[q];
        reinterpret_cast<float4*>(p.h_next + ((size_t)img * beam + j) * p.H)[q] =
            reinterpret_cast<const float4*>(p.h_out + ((size_t)img * G + b) * p.H)[q];
    }
}

// base_model.py:234-238: complete captions if any, else the partial ones, sorted by
// descending score (list.sort(reverse=True) is stable: equal scores keep heap order).
__global__ void beam_finalize_kernel(const BeamParams p) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= p.NI) return;
    const int beam = p.beam, T = p.T;
    const int cn = p.comp_n[img];
    const bool use_comp = cn > 0;
    const int n = use_comp ? cn : p.part_n[img];
    int order[kMaxBeam];
    double sc[kMaxBeam];
    for (int j = 0; j < n; ++j) {
        order[j] = j;
        sc[j] = use_comp ? p.comp_heap[(size_t)img * beam + j].score : p.part_score[(size_t)img * beam + j];
    }
    for (int a = 1; a < n; ++a) {  // stable insertion sort, descending
        const int o = order[a];
        const double s = sc[a];
        int k = a;
        while (k > 0 && sc[k - 1] < s) { sc[k] = sc[k - 1]; order[k] = order[k - 1]; --k; }
        sc[k] = s; order[k] = o;
    }
    const int* part_sent = p.sent[p.step & 1] + (size_t)img * beam * T;   // p.step = number of steps taken
    for (int j = 0; j < beam; ++j) {
        int* dst = p.res_sent + ((size_t)img * beam + j) * T;
        if (j < n) {
            const int o = order[j];
            int len;
            const int* src;
            if (use_comp) {
                const CItem& it = p.comp_heap[(size_t)img * beam + o];
                len = it.len;
                src = p.comp_sent + ((size_t)img * beam + it.slot) * T;
            } else {
                len = p.step;
                src = part_sent + (size_t)o * T;
            }
            for (int t = 0; t < T; ++t) dst[t] = t < len ? src[t] : -1;
            p.res_len[(size_t)img * beam + j] = len;
            p.res_score[(size_t)img * beam + j] = sc[j];
        } else {
            for (int t = 0; t < T; ++t) dst[t] = -1;
            p.res_len[(size_t)img * beam + j] = 0;
            p.res_score[(size_t)img * beam + j] = 0.0;
        }
    }
    p.res_n[img] = n;
    p.res_complete[img] = use_comp ? 1 : 0;
}

cudaError_t beam_update_launch(const BeamParams& p, cudaStream_t st) {
    if (p.beam > kMaxBeam) return cudaErrorInvalidValue;
    beam_update_kernel<<<p.NI, 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t beam_finalize_launch(const BeamParams& p, cudaStream_t st) {
    beam_finalize_kernel<<<(p.NI + 63) / 64, 64, 0, st>>>(p);
    return cudaGetLastError();
}

size_t beam_citem_bytes() { return sizeof(CItem); }

}  // namespace sat
