/*
 * vqa_b200.h -- C ABI of libvqa_b200.so, the B200-native (sm_100a) VQAScore engine.
 *
 * This is the drop-in boundary for the scoring hot path of linzhiqiu/t2v_metrics: one call computes, for a batch of
 * (image, question) pairs, the score the reference's VQAScoreModel.forward() returns
 * (t2v_metrics/models/vqascore_models/vqa_model.py:10-18, called from t2v_metrics/score.py:104-106).
 * For CLIP-FlanT5 that is exp(-mean CE(logits, labels)) of T5ForConditionalGeneration on the spliced
 * [text | 576 CLIP patch features | text] sequence (v3.0 clip_t5_model.py; arithmetic in
 * transformers/models/t5/modeling_t5.py:992-1133 and transformers/models/clip/modeling_clip.py:647-696).
 *
 * Conventions
 *   - plain C types only; every pointer documented as HOST or DEVICE memory;
 *   - the caller owns all buffers (weights, inputs, outputs, workspace); the library never frees caller memory and
 *     keeps only borrowed device pointers to the bound weights;
 *   - all work is enqueued on the given CUDA stream (passed as void* = cudaStream_t); no internal device sync;
 *   - every function returns 0 on success or a negative vqa_status; vqa_last_error() gives the message;
 *   - a handle belongs to one device and one in-flight call; distinct handles are independent.
 */
#ifndef VQA_B200_H
#define VQA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQA_B200_ABI_VERSION 1

typedef enum {
    VQA_OK = 0,
    VQA_ERR_INVALID_ARG = -1,
    VQA_ERR_MISSING_WEIGHT = -2,
    VQA_ERR_CUDA = -3,
    VQA_ERR_WORKSPACE = -4,
    VQA_ERR_UNSUPPORTED = -5
} vqa_status;

typedef enum { VQA_DTYPE_BF16 = 0, VQA_DTYPE_F32 = 1, VQA_DTYPE_I32 = 2 } vqa_dtype;

/* CLIP-FlanT5 architecture description (replaces the HF config objects the reference loads in
 * t2v_metrics/models/vqascore_models/mm_utils.py:182-241). */
typedef struct {
    /* CLIP ViT vision tower (transformers/models/clip/modeling_clip.py) */
    int32_t image_size;       /* 336 */
    int32_t patch_size;       /* 14 */
    int32_t vit_hidden;       /* 1024 */
    int32_t vit_heads;        /* 16 (head_dim must be 64) */
    int32_t vit_mlp;          /* 4096 */
    int32_t vit_layers_run;   /* encoder layers executed = select_layer -2 -> num_hidden_layers - 1 = 23 */
    float   vit_ln_eps;       /* 1e-5 */
    /* T5 encoder-decoder (transformers/models/t5/modeling_t5.py) */
    int32_t d_model;          /* 4096 (xxl) / 2048 (xl) */
    int32_t n_heads;          /* 64 / 32 (d_kv must be 64) */
    int32_t d_ff;             /* 10240 / 5120 */
    int32_t enc_layers;       /* 24 */
    int32_t dec_layers;       /* 24 */
    int32_t vocab;            /* 32128 */
    int32_t rel_buckets;      /* 32 */
    int32_t rel_max_distance; /* 128 */
    float   t5_ln_eps;        /* 1e-6 */
    int32_t image_token_id;   /* -200, t2v_metrics/constants.py:7 */
    int32_t pad_token_id;     /* 0 */
    int32_t decoder_start_id; /* 0 */
    int32_t emulate_bf16_rounding; /* 1: round scores/bias adds to bf16 where the reference's eager path does */
    int32_t cross_attention_mode;  /* 0: absorbed (q.(Wk x) = (Wk^T q).x, needs t5.dec.{i}.ckT); 1: project K/V of all
                                      encoder rows in every decoder layer, as modeling_t5.py:297-299 does */
} vqa_clipt5_config;

/* A named device tensor handed to vqa_bind_weights (borrowed pointer, bf16, row-major contiguous). */
typedef struct {
    const char* name;  /* HOST string, canonical engine name (see INTEGRATION.md for the HF -> engine mapping) */
    const void* data;  /* DEVICE pointer */
    int64_t shape[4];
    int32_t ndim;
    int32_t dtype;     /* vqa_dtype; weights must be VQA_DTYPE_BF16 */
} vqa_tensor;

typedef struct vqa_handle vqa_handle;

/* ABI / build info ("vqa_b200 abi=1 sm_100a ..."). Never fails. */
const char* vqa_version(void);

/* Create an engine for a CLIP-FlanT5 model on CUDA device `device`. */
int vqa_create_clipt5(const vqa_clipt5_config* cfg, int device, vqa_handle** out);

/* Borrow device pointers of the model weights. May be called several times; names are matched exactly.
 * Replaces model_cls.from_pretrained(...).to(device, bf16) (mm_utils.py:201,228) -- loading stays in the host. */
int vqa_bind_weights(vqa_handle* h, const vqa_tensor* tensors, int32_t n);

/* Check that every weight the forward needs is bound with the right shape. */
int vqa_finalize_weights(vqa_handle* h);

/* Bytes of device workspace vqa_clipt5_score needs for up to `batch` pairs, `n_images` distinct images,
 * text length `text_len` (ids per row incl. the image slot) and `label_len` target tokens. */
size_t vqa_clipt5_workspace_bytes(vqa_handle* h, int32_t batch, int32_t n_images, int32_t text_len,
                                  int32_t label_len);

/* Score `batch` (image, question) pairs.
 *   pixels      DEVICE [n_images, 3, image_size, image_size], dtype pixel_dtype (F32 or BF16): CLIPImageProcessor output
 *   image_index DEVICE int32 [batch] -> which image each pair uses, or NULL for identity (n_images == batch)
 *   input_ids   DEVICE int32 [batch, text_len], image slot = cfg.image_token_id, right-padded with pad_token_id
 *   text_lens   DEVICE int32 [batch] valid ids per row
 *   labels      DEVICE int32 [batch, label_len], -100 = ignored (padding)
 *   out_scores  DEVICE float [batch] : exp(-mean CE)  (v3.0 CLIPT5Model.forward)
 *   out_logprobs DEVICE float [batch, label_len] per-token log-probabilities, or NULL
 *   workspace   DEVICE, >= vqa_clipt5_workspace_bytes(...)
 *   stream      cudaStream_t
 */
int vqa_clipt5_score(vqa_handle* h, const void* pixels, int32_t pixel_dtype, int32_t n_images,
                     const int32_t* image_index, const int32_t* input_ids, const int32_t* text_lens,
                     const int32_t* labels, int32_t batch, int32_t text_len, int32_t label_len, float* out_scores,
                     float* out_logprobs, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels the last vqa_clipt5_score call launched (for bench.py's gpu_launches). */
int64_t vqa_last_launch_count(vqa_handle* h);

/* Optional device-side timing of the forward: with profiling on, every launch of vqa_clipt5_score is bracketed by CUDA
 * events on the caller's stream. After the caller synchronised the stream, vqa_profile_read returns, per category
 * {0 gemm, 1 attention, 2 norm, 3 other} (arrays of 4): device ms, algorithmic FLOPs (2MNK / 4*S*S*d) and scope counts
 * of the LAST call. Used by bench.py for the roofline object. */
int vqa_set_profile(vqa_handle* h, int32_t enable);
int vqa_profile_read(vqa_handle* h, float* ms, double* flops, int64_t* scopes);

const char* vqa_last_error(vqa_handle* h);
void vqa_destroy(vqa_handle* h);

/* ---- kernel-level entry points (used by tests/ and bench.py to exercise single kernels through the same ABI) ---- */

/* C[M,N] = epilogue(A[M,K] . W[N,K]^T); all DEVICE bf16 row-major. epilogue: 0 store, 1 quick_gelu, 2 gelu(erf),
 * 3 gated gelu_new (W = [gate; up] with `gate_up_offset` rows between them, C is [M, N/2]), 5 relu.
 * bias [N] / residual [M, ldr] may be NULL. variant: 0 auto, else (BLOCK_N * 10 + cta_group), e.g. 2562, 2561, 641. */
int vqa_op_gemm_bf16(const void* A, int32_t lda, const void* W, int32_t ldw, int32_t w_rows, void* C, int32_t ldc,
                     int32_t M, int32_t N, int32_t K, const void* bias, const void* residual, int32_t ldr,
                     int32_t epilogue, int32_t gate_up_offset, int32_t variant, void* stream);

/* Fused lm_head + log-softmax gather: logprob[m] = (h[m].W[label[m]]) - logsumexp_n(h[m].W[n]); logits never stored.
 * scratch: DEVICE float, >= 4*M*ceil(N/128) + M floats. */
int vqa_op_lmhead_logprob(const void* H, int32_t ldh, const void* W, int32_t ldw, int32_t M, int32_t N, int32_t K,
                          const int32_t* labels, float* logprob, float* scratch, void* stream);

/* Bidirectional attention, head_dim 64, packed qkv [B*S, 3*H*64] -> out [B*S, H*64].
 * bias_table: DEVICE float [H, 2S-1] or NULL; seq_lens DEVICE int32 [B] or NULL. */
int vqa_op_attention_d64(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, const int32_t* seq_lens,
                         const float* bias_table, float scale, int32_t round_scores, void* stream);

/* T5LayerNorm / nn.LayerNorm on [rows, D] bf16. beta == NULL selects T5 RMS norm. */
int vqa_op_norm(const void* x, const void* gamma, const void* beta, void* y, int32_t rows, int32_t D, float eps,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_B200_H */
