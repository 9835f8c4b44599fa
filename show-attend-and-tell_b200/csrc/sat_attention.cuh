// sat_attention.cuh — parameters of the fused attention kernel (see sat_attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sat {

struct AttParams {
    const float* T;      // phase-1 rows [NI*L, RL]: tanh-projected contexts (2-layer) or raw contexts (1-layer)
    const float* vec;    // [RL]: attend/fc_2 kernel (2-layer) or attend/fc_a kernel (1-layer)
    const float* q;      // [NI*G, RL] state branch tanh(h*W1b+b1b), or null (1-layer)
    const float* eadd;   // [NI*G, L] additive logits h*fc_b (1-layer), or null
    float* e;            // [NI*G, L] attention logits (scratch, L2 resident)
    unsigned* rowcnt;    // [NI] rows of T finished per image; zero at launch
    unsigned target;     // = L
    float* alpha;        // [NI*G, L]
    float* z;            // [NI*G, D]
    int NI, G, L, D, RL;
    int rch;             // phase-1 rows per TMA chunk
    int slot_bytes;
    int nslots;
    int l2_t, l2_ctx;    // L2 eviction policies of the two streams (see l2_policy)
};

bool att_plan(AttParams& p, int smem_optin);
size_t att_smem_bytes(const AttParams& p);
cudaError_t att_launch(const CUtensorMap& map, const AttParams& p, int num_sms, cudaStream_t st, bool coop);
cudaError_t ctx_mean_launch(const float* ctx, float* out, int NI, int L, int D, cudaStream_t st);

}  // namespace sat
