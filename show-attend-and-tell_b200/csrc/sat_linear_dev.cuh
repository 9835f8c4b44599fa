// sat_linear_dev.cuh — device helpers shared by the dense kernels (sat_linear.cu: one launch per layer;
// sat_chain.cu: the three dense layers of a decode step in one persistent launch).  Include after sat_common.cuh
// and sat_linear.cuh.
#pragma once
#include "sat_common.cuh"
#include "sat_linear.cuh"

namespace sat {

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ void split_bf16x8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
    float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]);
        __nv_bfloat16 h1 = __float2bfloat16_rn(x[2 * i + 1]);
        __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
        __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// TF LSTMCell (un-vendored TF 1.7 dependency; gate order i, j, f, o and forget_bias 1.0
// confirmed on the reference's recorded GraphDef, tests/golden/graph_fixture.json):
//   c = sigmoid(f + 1) * c_prev + sigmoid(i) * tanh(j);  h = sigmoid(o) * tanh(c)
__device__ __forceinline__ void lstm_gates(const LinProblem& P, float4 g, float cp, int b, int unit, int mode, bool dry) {
    const float c = act_sigmoid(g.z + 1.0f) * cp + act_sigmoid(g.x) * act_tanh(g.y);
    const float h = act_sigmoid(g.w) * act_tanh(c);
    if (dry) return;   // instruction-cache warm-up pass: no side effects
    P.c_out[(size_t)b * P.H + unit] = c;
    P.h_out[(size_t)b * P.H + unit] = h;
    if (P.out_pa) pa_store(P.out_pa, mode, P.row_tile, P.H >> 6, b, unit, h);   // h feeds the next dense layers
}

}  // namespace sat
