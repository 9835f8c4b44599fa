// sat_rows.cuh — per-row vocabulary kernels + device beam bookkeeping (see sat_rows.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sat {

constexpr int kMaxTopK = 8;
constexpr int kMaxBeam = 7;

struct RowsParams {
    const float* logits;  // [rows, V]
    int V;
    float* probs;         // [rows, V] or null
    int32_t* argmax;      // [rows] or null
    int32_t* tokens;      // [rows, tokens_ld] or null: tokens[row, step] = argmax
    int tokens_ld;
    int step;
    int32_t* next_word;   // [rows] or null: word fed to the next step
    const int32_t* forced;  // [rows, forced_ld] teacher-forced words or null (greedy)
    int forced_ld;
    int topk;             // 0 or beam+1
    int32_t* topk_idx;    // [rows, topk]
    float* topk_p;        // [rows, topk]
};

struct PItem;
struct CItem;

struct BeamParams {
    int NI, beam, nlive, T, step, eos_id, H;
    const int32_t* topk_idx;  // [NI*nlive, beam+1]
    const float* topk_p;
    double* part_score;       // [NI, beam]   partial heap (array order == row order of the next step)
    int32_t* part_n;          // [NI]
    int32_t* sent[2];         // [NI, beam, T] ping-pong by step parity
    CItem* comp_heap;         // [NI, beam]
    int32_t* comp_n;          // [NI]
    int32_t* comp_sent;       // [NI, beam, T]
    const float* c_out;       // [NI*nlive, H] states computed this step
    const float* h_out;
    float* c_next;            // [NI*beam, H] states fed to the next step
    float* h_next;
    int32_t* next_word;       // [NI*beam]
    // results (finalize)
    int32_t* res_sent;        // [NI, beam, T], -1 padded
    int32_t* res_len;         // [NI, beam]
    double* res_score;        // [NI, beam]
    int32_t* res_n;           // [NI]
    int32_t* res_complete;    // [NI]
};

cudaError_t rows_softmax_launch(const RowsParams& p, int rows, cudaStream_t st);
cudaError_t beam_update_launch(const BeamParams& p, cudaStream_t st);
cudaError_t beam_finalize_launch(const BeamParams& p, cudaStream_t st);
size_t beam_citem_bytes();

}  // namespace sat
