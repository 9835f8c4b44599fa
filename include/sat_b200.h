/* sat_b200.h — C ABI of libsat_b200.so: the soft-attention LSTM decode path of
 * Cheng-Lin-Li/show-attend-and-tell on NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI: its boundary is the Python attribute surface of
 * CaptionGenerator consumed through tf.Session.run feeds/fetches.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference
 * repository root).
 *
 * Conventions
 *   - return 0 (SAT_OK) on success, a negative SAT_ERR_* code otherwise; nothing is
 *     thrown or aborted across the boundary; sat_last_error() describes the failure
 *     (thread local).
 *   - every tensor argument is a DEVICE pointer unless the name ends in _host;
 *     the caller owns every tensor buffer; the library owns its workspace and its
 *     repacked copies of the weights.
 *   - all tensors are dense row-major fp32 (model.py:205-213) except word ids, which
 *     are int32 (model.py:214-216).
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation except in *_host entry points, which return after the results
 *     are in host memory.
 *   - a handle is bound to the CUDA device current at sat_create and is not thread
 *     safe (the reference's driver loop is single threaded: base_model.py:184-212).
 */
#ifndef SAT_B200_H_
#define SAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAT_OK 0
#define SAT_ERR_INVALID (-1)     /* bad argument / shape mismatch (TF InvalidArgumentError) */
#define SAT_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define SAT_ERR_STATE (-3)       /* weights missing, contexts not prepared, ... */
#define SAT_ERR_UNSUPPORTED (-4) /* shape outside what the kernels implement */
#define SAT_ERR_NOMEM (-5)

typedef struct sat_handle sat_handle;

/* Static shape of the decoder graph.  Field names follow config.py:9-17,67 (including the
 * `initalize` spelling); L and D are num_ctx / dim_ctx of model.py:54-59,103-108. */
typedef struct sat_dims {
    int32_t max_batch;            /* max rows per step = images x beam            */
    int32_t num_ctx;              /* L                                            */
    int32_t dim_ctx;              /* D                                            */
    int32_t num_lstm_units;       /* H                                            */
    int32_t dim_embedding;        /* E                                            */
    int32_t dim_attend_layer;     /* A                                            */
    int32_t dim_decode_layer;     /* Dd                                           */
    int32_t dim_initalize_layer;  /* I                                            */
    int32_t vocabulary_size;      /* V                                            */
    int32_t num_attend_layers;    /* 1 or 2                                       */
    int32_t num_decode_layers;    /* 1 or 2                                       */
    int32_t num_initalize_layers; /* 1 or 2                                       */
    int32_t max_caption_length;   /* max steps of a loop / beam search            */
    int32_t max_beam;             /* max beam_size (<= 4); 0 or 1 = no beam search */
} sat_dims;

/* replaces CaptionGenerator(config) (main.py:48,61,69; model.py:7-13, build_rnn :190-356) */
int sat_create(const sat_dims* dims, sat_handle** out);
void sat_destroy(sat_handle* h);
const char* sat_last_error(void);
int sat_version(void);

/* integer knobs (defaults in brackets; all of them are for experiments and tests, none changes results beyond
 * the summation order noted):
 *   "gemm"        1 = tcgen05 [1], 0 = CUDA-core bring-up kernels
 *   "umma_layout" 0 = interleaved [0], 1 = 128B swizzle; must be set before sat_set_weight
 *   "graphs"      1 = replay CUDA graphs in loops [1]
 *   "hoist"       1 = project the contexts once per image batch [1]; 0 = recompute every step like model.py:259-262
 *   "pa"          1 = activations travel between dense layers as packed UMMA operands [1]
 *   "xpack"       1 = cooperative activation pre-pass when operands are not packed [1]; 0 = producer warps
 *   "pdl"         1 = launches carry the programmatic-dependent-launch attribute [1]
 *   "overlap"     launch layout of the greedy loop: 2 = one stream, the attention kernel of step t+1 runs beside the
 *                 vocabulary layer of step t without waiting for it [2]; 1 = two streams, fork/join; 0 = in order.
 *                 (the attention grid, hence the split-L merge order, differs between 0 and 1/2)
 *   "xbatch"      1 = sat_decode_loop runs its prologue (context projection, initialize) on a stream of the library's
 *                 own into one of two buffer sets, so that it overlaps the decode steps of the previous call [0].
 *                 Contract: the contexts passed to sat_decode_loop are COMPLETE when the call is made (not produced
 *                 by earlier work queued on the same stream).  The pipelined host API does this by itself.
 *   "prologue1"   1 = the pass that packs the contexts for the hoisted projection also takes their mean over the
 *                 locations (one pass over the conv features for initialize and attend/fc_1a) [1]
 *   "chain"       1 = greedy loops at 64-row batches run the three dense layers of a step as phases of ONE persistent
 *                 launch (sat_chain.cu; an experiment, slower than the default chained launches) [0]
 *   "warm"        1 = idle epilogue warps pre-run the epilogue code to warm the instruction caches [1]
 *   "att_wpc"     1 = warp-per-chunk attention kernel for 512-float rows [1]
 *   "att_sms", "att_occ", "att_warps", "l2_w", "l2_vocab", "l2_t", "l2_ctx", "l2_prefetch": grid / cache-policy knobs
 *   "profile"     1 = record CUDA events around every eager kernel launch; read back with sat_get_info
 *                 "prof_ns_<family>" / "prof_n_<family>"
 *   "trace"       1 / 2 / 3 = in-kernel globaltimer stamps of a dense launch ("trace_at") / of the attention kernel /
 *                 of every launch of a loop (tools/trace*.py, tools/timeline.py)
 * Environment: SAT_PDL=0 creates handles with "pdl" off (for tools that expect one kernel of a stream at a time).
 * A handle expects the GPU to itself while a loop runs: the fused arg-max of the vocabulary layer ends in a grid-wide
 * rendezvous of its one-wave launch (a stuck rendezvous traps with a message after a few seconds). */
int sat_set_option(sat_handle* h, const char* key, int64_t value);
int sat_get_info(sat_handle* h, const char* key, int64_t* value);

/* replaces BaseModel.load's per-variable assign (base_model.py:257-278).  `tf_var_name` is
 * the TF variable name with or without ":0" (e.g. "lstm/lstm_cell/kernel"); `dev` holds the
 * variable in the reference's layout: dense kernels [in, units] (rows=in, cols=units), biases
 * [units] (rows=1), the LSTM kernel [D+E+H, 4H] with column blocks i,j,f,o, the embedding [V,E].
 * The repack into the library's own storage is queued on `stream` (no host synchronisation: a training loop that
 * refreshes the decode weights pays one sync, not twenty): `dev` must stay valid until `stream` has passed this
 * call, and work that uses the weights must be ordered after it (the same stream, or an event). */
int sat_set_weight(sat_handle* h, const char* tf_var_name, const float* dev, int64_t rows, int64_t cols,
                   void* stream);
/* number of variables still missing (0 = ready) */
int sat_weights_missing(sat_handle* h);

/* replaces sess.run([conv_feats, initial_memory, initial_output], {images})
 * (base_model.py:168-170) minus the CNN: projects the contexts (attend fc_1a, model.py:417-420,
 * hoisted out of the step loop) and runs initialize (model.py:239-242, 358-393).
 * contexts [n_img, L, D]; initial_memory / initial_output [n_img, H] may be NULL. */
int sat_prepare_contexts(sat_handle* h, const float* contexts, int32_t n_img, float* initial_memory,
                         float* initial_output, void* stream);

/* replaces sess.run([memory, output, probs], {contexts, last_word, last_memory, last_output})
 * (base_model.py:207-212; graph model.py:258-290).  All of memory/output [B,H] are required;
 * logits / probs [B,V] and alpha [B,L] may be NULL.  `contexts` must be the buffer last given
 * to sat_prepare_contexts with n_img == B (otherwise the projection is redone here). */
int sat_decode_step(sat_handle* h, const float* contexts, const int32_t* last_word, const float* last_memory,
                    const float* last_output, float* memory, float* output, float* logits, float* probs,
                    float* alpha, int32_t B, void* stream);

/* T steps on the device without host round trips: prepare + initialize + T x step, the word fed
 * to step t+1 being argmax of step t (model.py:289) or forced_words[b, t] (teacher forcing,
 * model.py:310).  tokens [B,T] receives the argmax words; logits_all [T,B,V] may be NULL. */
int sat_decode_loop(sat_handle* h, const float* contexts, int32_t B, int32_t T, const int32_t* forced_words,
                    int32_t* tokens, float* logits_all, void* stream);

/* replaces BaseModel.beam_search (base_model.py:163-240) with TopN/CaptionData semantics of
 * utils/misc.py:38-87; `eos_id` stands for vocabulary.words[w] == '.' (base_model.py:229).
 * Outputs per image, sorted by descending score: sentences [n_img, beam, T] (-1 padded),
 * lengths [n_img, beam], scores [n_img, beam] (fp64 products of probabilities,
 * base_model.py:224), n_results [n_img], is_complete [n_img]. */
int sat_beam_search(sat_handle* h, const float* contexts, int32_t n_img, int32_t beam_size, int32_t T,
                    int32_t eos_id, int32_t* sentences, int32_t* lengths, double* scores, int32_t* n_results,
                    int32_t* is_complete, void* stream);

/* host-buffer forms: copy in, run, copy out, synchronise (what a sess.run caller sees) */
int sat_decode_step_host(sat_handle* h, const float* contexts_host, int32_t contexts_changed,
                         const int32_t* last_word_host, const float* last_memory_host,
                         const float* last_output_host, float* memory_host, float* output_host,
                         float* probs_host, int32_t B, void* stream);
int sat_decode_loop_host(sat_handle* h, const float* contexts_host, int32_t B, int32_t T,
                         const int32_t* forced_words_host, int32_t* tokens_host, void* stream);
/* Pipelined form of sat_decode_loop_host for a stream of batches (what a caption server / the eval loop of
 * base_model.py:94-140 does batch after batch).  submit() enqueues upload (own copy stream) -> loop -> download for
 * staging slot 0 or 1 and returns; wait() blocks until that slot's tokens are in tokens_host.  Submit batch i+1 on
 * the other slot before waiting for batch i and its upload overlaps batch i's decode.  The host buffers must stay
 * valid (and, for true overlap, be pinned) until wait(). */
int sat_decode_loop_host_submit(sat_handle* h, const float* contexts_host, int32_t B, int32_t T,
                                const int32_t* forced_words_host, int32_t* tokens_host, int32_t slot, void* stream);
int sat_decode_loop_host_wait(sat_handle* h, int32_t slot);
int sat_beam_search_host(sat_handle* h, const float* contexts_host, int32_t n_img, int32_t beam_size, int32_t T,
                         int32_t eos_id, int32_t* sentences_host, int32_t* lengths_host, double* scores_host,
                         int32_t* n_results_host, int32_t* is_complete_host, void* stream);

/* individually callable kernels of one step (profiling / unit tests).  Rows = n_img * group;
 * `group` rows of one image share its contexts (beams).
 *   sat_attention_fwd : attend + context vector (model.py:262-264) given the state h [rows,H]
 *   sat_lstm_fwd      : embedding lookup + LSTMCell (model.py:272-279)
 *   sat_vocab_gemm    : decode (model.py:282-287) -> logits [rows,V]                         */
int sat_attention_fwd(sat_handle* h, const float* contexts, const float* output, float* alpha, float* context,
                      int32_t n_img, int32_t group, void* stream);
int sat_lstm_fwd(sat_handle* h, const float* context, const int32_t* last_word, const float* last_memory,
                 const float* last_output, float* memory, float* output, int32_t rows, void* stream);
int sat_vocab_gemm(sat_handle* h, const float* output, const float* context, const int32_t* last_word,
                   float* logits, int32_t rows, void* stream);
/* generic dense layer y = act(x W + b) through the same tensor-core kernel (tests):
 * x [rows,K], w_tf [K,n_out] (TF layout), b [n_out] or NULL, act 0 none / 1 tanh. */
int sat_dense_fwd(sat_handle* h, const float* x, const float* w_tf, const float* b, float* y, int32_t rows,
                  int32_t K, int32_t n_out, int32_t act, int32_t splits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step (model.py:250-334 losses, :461-511 optimizer; driver base_model.py:39-68), for the 1- and 2-layer
 * variants of initialize / attend / decode (config.py:15-19; sat_train_var enumerates the variables of the chosen graph).
 * The trainable variables live in ONE flat fp32 device buffer owned by the caller (parameters), with
 * parallel buffers for the gradient and the two Adam slots; sat_train_var describes the layout
 * (TF variable name, offset in floats, TF shape).  A data-parallel step is
 *     sat_train_forward_backward  ->  all-reduce(sum) of `grads` across ranks (NCCL)  ->  sat_train_apply
 * Dropout masks come from a counter-based generator keyed by `seed` (0 = dropout off), so a step is
 * reproducible; ranks must use different seeds.
 * Knobs: sat_set_option "train_tc" 1 = the large products of the step run on the tcgen05 dense kernel [1], 0 = fp32
 * CUDA-core SGEMM everywhere.  Environment switches, read once per process, for A/B timing (results agree to round-off):
 * SAT_TRAIN_PDL=0 launches the step's kernels without the programmatic-serialization attribute; SAT_TRAIN_DEC_ALL=0 keeps
 * the decode layers inside the time loop (default: one stacked product per layer for all T steps); SAT_TRAIN_SIDE=0 keeps
 * the attend/fc_1a products on the caller's stream (default: a second, low-priority stream of the library's own; 2 / 3:
 * only the forward / only the backward ones);
 * SAT_TRAIN_FUSE_SOFTMAX=0 / 2 un-fuses the softmax kernels (both directions / the backward one only); SAT_TRAIN_FUSE_PACK=0
 * packs the operands of the batch-row products in launches of their own instead of in their producer kernels.
 * Word ids outside [0, vocabulary_size) read as zero rows, contribute no gradient and are counted
 * (sat_get_info "train_bad_ids"). */
int sat_train_init(sat_handle* h, int32_t B, int32_t T, float fc_drop_rate, float lstm_drop_rate,
                   float attention_loss_factor, float fc_kernel_regularizer_scale);
int sat_train_num_vars(sat_handle* h);
/* the counter-based dropout generator, U[0,1) with 24 bits: mask = floor(keep + u) (host function, no GPU) */
float sat_train_rng_uniform(uint64_t seed, uint64_t stream, uint64_t index);
/* i in [0, num_vars): name / offset / rows / cols / regularised of variable i; total = floats in the flat buffer */
int sat_train_var(sat_handle* h, int32_t i, const char** name, int64_t* offset, int64_t* rows, int64_t* cols,
                  int32_t* regularised, int64_t* total);
/* replaces the forward+backward half of sess.run(opt_op) (base_model.py:57-60).  contexts [B,L,D], sentences
 * int32 [B,T], masks [B,T]; global_mask_sum / global_batch are the normalisers of the WHOLE (all ranks) batch
 * (model.py:316-318, 324-326).  losses (device, 4 floats): cross_entropy, accuracy, attention, reg.
 * grads is overwritten with this shard's gradient WITHOUT the regulariser term. */
int sat_train_forward_backward(sat_handle* h, const float* params, float* grads, const float* contexts,
                               const int32_t* sentences, const float* masks, int32_t B, int32_t T, uint64_t seed,
                               double global_mask_sum, int32_t global_batch, float* losses, void* stream);
/* the same with the whole-batch mask sum read from DEVICE memory (one double, e.g. the output of an all-reduce
 * queued on `stream` just before): nothing of a data-parallel step then waits for the host. */
int sat_train_forward_backward_dsum(sat_handle* h, const float* params, float* grads, const float* contexts,
                                    const int32_t* sentences, const float* masks, int32_t B, int32_t T, uint64_t seed,
                                    const double* global_mask_sum_dev, int32_t global_batch, float* losses, void* stream);
/* adds the L2-regulariser gradient, clips by the global norm (clip_gradients, config.py:36) and applies TF Adam
 * (config.py:32-43).  step counts from 1.  grad_norm (device, 1 float, may be NULL) receives the squared norm. */
int sat_train_apply(sat_handle* h, float* params, float* grads, float* adam_m, float* adam_v, int64_t step, float lr,
                    float beta1, float beta2, float epsilon, float clip, float* grad_norm, void* stream);

/* The optimizer of model.py:479-503 as plain data (field names of config.py:30-43).  kind: SAT_OPT_*.
 *   Adam      tf.train.AdamOptimizer(learning_rate, beta1, beta2, epsilon)              slots: m, v
 *   RMSProp   tf.train.RMSPropOptimizer(learning_rate, decay, momentum, epsilon, centered)
 *                                                   slots: rms (STARTS AT ONE: sat_train_fill), mg (centered only), momentum
 *   Momentum  tf.train.MomentumOptimizer(learning_rate, momentum, use_nesterov)        slots: accumulator
 *   SGD       tf.train.GradientDescentOptimizer(learning_rate)                          slots: none
 * clip_gradients: optimize_loss(clip_gradients=...) = clip_by_global_norm, applied for every optimizer.
 * (The reference passes an Optimizer INSTANCE to tf.contrib.layers.optimize_loss, so its learning_rate_decay_fn only feeds
 * the "learning_rate" summary: the optimizer keeps initial_learning_rate.  The facade reproduces that and offers the
 * decayed rate as an option; this ABI simply takes the rate to use.) */
#define SAT_OPT_ADAM 0
#define SAT_OPT_RMSPROP 1
#define SAT_OPT_MOMENTUM 2
#define SAT_OPT_SGD 3
typedef struct sat_optimizer {
    int32_t kind;
    float learning_rate, beta1, beta2, epsilon, decay, momentum;
    int32_t centered, use_nesterov;
    float clip_gradients;
} sat_optimizer;
/* sat_train_apply for any of the four optimizers; slot0..2 in the order listed above (unused ones may be NULL). */
int sat_train_apply_opt(sat_handle* h, float* params, float* grads, float* slot0, float* slot1, float* slot2, int64_t step,
                        const sat_optimizer* opt, float* grad_norm, void* stream);
/* buf[0..n) = value on the device (initial value of an optimizer slot) */
int sat_train_fill(sat_handle* h, float* buf, float value, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAT_B200_H_ */
