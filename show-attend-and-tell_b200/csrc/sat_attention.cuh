// sat_attention.cuh — parameters of the fused attention kernel (see sat_attention.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sat {

struct AttParams {
    const float* T;      // scored rows [NI*L, RL]: tanh-projected contexts (2-layer) or raw contexts (1-layer)
    const float* ctx;    // [NI*L, D] contexts (un-dropped, model.py:263)
    const float* vec;    // [RL]: attend/fc_2 kernel (2-layer) or attend/fc_a kernel (1-layer)
    const float* q;      // [NI*G, RL] state branch tanh(h*W1b+b1b), or null (1-layer)
    const float* eadd;   // [NI*G, L] additive logits h*fc_b (1-layer), or null
    float* e;            // [NI*G, L] attention logits (scratch, L2 resident)
    float* part;         // [grid, segmax, G, D+2] partial contexts + (max, sum) per CTA segment
    unsigned* rowcnt;    // [NI] CTAs finished per image; zero between launches (self-resetting)
    float* alpha;        // [NI*G, L]
    float* z;            // [NI*G, D]
    int NI, G, L, D, RL;
    int rch, cch;        // rows of T / ctx per TMA chunk
    int slot_bytes;
    int nslots;
    int grid, segmax;
    int occ;             // requested CTAs per SM (1 or 2; 2 only for G == 1)
    int warps;           // consumer warps per CTA: 8, or 16 (G == 1 only)
    int wpc;             // in: warp-per-chunk kernel wanted; out of att_plan: used (needs RL == D == 512)
    int l2_t, l2_ctx;    // L2 eviction policies of the two streams (see l2_policy)
    uint8_t* pa_z;       // optional packed copy of z for the dense layers that consume it
    int pa_row_tile, pa_mode;
    // side job while the first TMA chunks are in flight: embedding lookup of this step's words
    // (model.py:272-274), packed for the LSTM / decode layers that follow
    const float* emb;    // [V, E] or null
    const int32_t* emb_word;  // [NI*G]
    uint8_t* emb_pa;
    int emb_E;
    unsigned long long* dbg;  // optional [grid][16] timeline stamps
    unsigned long long* tl;   // optional {min start, max end} of this launch
    int pdl;             // launched with programmatic stream serialization
    int nowait;          // (with pdl, warp-per-chunk kernel) nothing of the immediate predecessor is read: run beside it
                         // and wait for it only before exiting
    const unsigned* qflag;   // (with nowait) q is produced by a phase of the STILL RUNNING predecessor (the chained dense
    unsigned qtarget;        // launch, sat_chain.cu): spin until *qflag has reached qtarget before reading q
};

bool att_plan(AttParams& p, int smem_optin, int num_sms);
size_t att_smem_bytes(const AttParams& p);
size_t att_part_floats(const AttParams& p);
cudaError_t att_launch(const AttParams& p, cudaStream_t st);
cudaError_t ctx_mean_launch(const float* ctx, float* out, int NI, int L, int D, cudaStream_t st);
cudaError_t ctx_mean_pack_launch(const float* ctx, float* out, uint8_t* pa, int row_tile, int layout_mode, int NI, int L, int D,
                                 cudaStream_t st);

}  // namespace sat
