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

__global__ void __launch_bounds__(kRowThreads) rows_softmax_kernel(const RowsParams p) {
    __shared__ ValIdx sm_vi[kRowThreads / 32];
    __shared__ float sm_f[kRowThreads / 32];
    const int row = blockIdx.x;
    const float* x = p.logits + (size_t)row * p.V;

    ValIdx best = {-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < p.V; i += kRowThreads) {
        const float v = x[i];
        if (better(v, i, best.v, best.i)) { best.v = v; best.i = i; }
    }
    best = block_best(best, sm_vi);
    const float m = best.v;
    if (threadIdx.x == 0) {
        if (p.argmax) p.argmax[row] = best.i;
        if (p.tokens) p.tokens[(size_t)row * p.tokens_ld + p.step] = best.i;
        if (p.next_word) p.next_word[row] = p.forced ? p.forced[(size_t)row * p.forced_ld + p.step] : best.i;
    }
    if (!p.probs && p.topk == 0) return;  // greedy / teacher-forced loops only need the argmax

    float s = 0.f;
    for (int i = threadIdx.x; i < p.V; i += kRowThreads) s += expf(x[i] - m);
    s = block_sum(s, sm_f);
    const float inv = 1.0f / s;

    // probabilities + thread-local top-K (K <= kMaxTopK), sorted by (prob desc, index asc)
    float tv[kMaxTopK];
    int ti[kMaxTopK];
#pragma unroll
    for (int k = 0; k < kMaxTopK; ++k) { tv[k] = -1.f; ti[k] = 0x7fffffff; }
    float* pr = p.probs ? p.probs + (size_t)row * p.V : nullptr;
    for (int i = threadIdx.x; i < p.V; i += kRowThreads) {
        const float pv = expf(x[i] - m) * inv;
        if (pr) pr[i] = pv;
        if (p.topk > 0 && better(pv, i, tv[p.topk - 1], ti[p.topk - 1])) {
            int k = p.topk - 1;
            while (k > 0 && better(pv, i, tv[k - 1], ti[k - 1])) { tv[k] = tv[k - 1]; ti[k] = ti[k - 1]; --k; }
            tv[k] = pv; ti[k] = i;
        }
    }
    if (p.topk > 0) {
        int head = 0;
        for (int k = 0; k < p.topk; ++k) {
            ValIdx c;
            c.v = head < p.topk ? tv[head] : -1.f;
            c.i = head < p.topk ? ti[head] : 0x7fffffff;
            // local arrays are indexed dynamically only through `head`; keep it simple
            const ValIdx w = block_best(c, sm_vi);
            if (w.i == c.i && w.v == c.v && c.i != 0x7fffffff) ++head;
            if (threadIdx.x == 0) {
                p.topk_idx[(size_t)row * p.topk + k] = w.i;
                p.topk_p[(size_t)row * p.topk + k] = w.v;
            }
        }
    }
}

cudaError_t rows_softmax_launch(const RowsParams& p, int rows, cudaStream_t st) {
    if (p.topk > kMaxTopK) return cudaErrorInvalidValue;
    rows_softmax_kernel<<<rows, kRowThreads, 0, st>>>(p);
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

__global__ void __launch_bounds__(128) beam_update_kernel(const BeamParams p) {
    __shared__ PItem newp[kMaxBeam];
    __shared__ int newn;
    const int img = blockIdx.x;
    const int beam = p.beam, K = p.beam + 1, G = p.nlive, T = p.T, idx = p.step;
    const int* sent_cur = p.sent[idx & 1] + (size_t)img * beam * T;
    int* sent_next = p.sent[(idx + 1) & 1] + (size_t)img * beam * T;

    if (threadIdx.x == 0) {
        int np = 0;
        CItem* ch = p.comp_heap + (size_t)img * beam;
        int cn = p.comp_n[img];
        for (int b = 0; b < G; ++b) {
            const double ps = idx == 0 ? 1.0 : p.part_score[(size_t)img * beam + b];   // base_model.py:178
            const size_t row = (size_t)img * G + b;
            for (int j = 0; j < K; ++j) {
                const int w = p.topk_idx[row * K + j];
                const double sc = ps * (double)p.topk_p[row * K + j];                   // base_model.py:224
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
    // materialise the surviving beams: sentences, last word, LSTM state rows
    const int np = newn;
    for (int j = 0; j < np; ++j) {
        const int b = newp[j].parent, w = newp[j].word;
        for (int t = threadIdx.x; t < idx; t += blockDim.x) sent_next[(size_t)j * T + t] = sent_cur[(size_t)b * T + t];
        if (threadIdx.x == 0) {
            sent_next[(size_t)j * T + idx] = w;
            p.next_word[(size_t)img * beam + j] = w;
        }
        const float* cs = p.c_out + ((size_t)img * G + b) * p.H;
        const float* hs = p.h_out + ((size_t)img * G + b) * p.H;
        float* cd = p.c_next + ((size_t)img * beam + j) * p.H;
        float* hd = p.h_next + ((size_t)img * beam + j) * p.H;
        for (int u = threadIdx.x; u < p.H; u += blockDim.x) { cd[u] = cs[u]; hd[u] = hs[u]; }
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
    beam_update_kernel<<<p.NI, 128, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t beam_finalize_launch(const BeamParams& p, cudaStream_t st) {
    beam_finalize_kernel<<<(p.NI + 63) / 64, 64, 0, st>>>(p);
    return cudaGetLastError();
}

size_t beam_citem_bytes() { return sizeof(CItem); }

}  // namespace sat
