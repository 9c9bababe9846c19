// Glue kernels of the Qwen2.5-VL scoring path (rotary embeddings, window re-ordering, feature splice, last-token gather).
// All index arrays (positions, window order, feature rows) come from the host, which mirrors the reference's Python index
// logic bit for bit (t2v_metrics_b200/qwen_host.py <-> transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py:382-451,
// 1024-1133).
#pragma once
#include "ptx.cuh"
#include "elementwise.cuh"

namespace vqa {

// cos/sin table for rotary embeddings: for i < half: angle = pos[axis_of_dim[i]][row] * inv_freq[i];
// table[row, i] = table[row, i + half] = cos/sin(angle)   (emb = cat(freqs, freqs), modeling_qwen2_5_vl.py:487, :611)
// pos: [n_axes, rows] int32. round_bf16: the text model casts cos/sin to the activation dtype (:614).
__global__ void rope_table_kernel(const int* __restrict__ pos, int rows, const int* __restrict__ axis_of_dim,
                                  const float* __restrict__ inv_freq, int half, float* __restrict__ cos_t,
                                  float* __restrict__ sin_t, int round_bf16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * half) return;
    const int row = idx / half, i = idx % half;
    const float ang = (float)pos[(size_t)axis_of_dim[i] * rows + row] * inv_freq[i];
    float c = cosf(ang), s = sinf(ang);
    if (round_bf16) { c = bf16_round(c); s = bf16_round(s); }
    const size_t o = (size_t)row * (2 * half);
    cos_t[o + i] = c; cos_t[o + i + half] = c;
    sin_t[o + i] = s; sin_t[o + i + half] = s;
}

// In-place rotary embedding on `n_heads` consecutive heads of a packed row-major buffer (head h at column col0 + h*head_stride,
// `head_dim` <= head_stride real dims): x' = x*cos + rotate_half(x)*sin  (modeling_qwen2_5_vl.py:156-167, :650-669).
// bf16_products: the text path multiplies in bf16 (each product and the sum are rounded); the vision path works in fp32.
// One warp per (row, head); lane handles dims lane, lane+32, ... of the first half.
__global__ void rope_inplace_kernel(__nv_bfloat16* __restrict__ x, int ld, int col0, int n_heads, int head_stride,
                                    int head_dim, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                    long long rows, int bf16_products) {
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= rows * n_heads) return;
    const int lane = threadIdx.x & 31;
    const long long row = w / n_heads;
    const int h = (int)(w % n_heads);
    __nv_bfloat16* px = x + (size_t)row * ld + col0 + h * head_stride;
    const float* pc = cos_t + (size_t)row * head_dim;
    const float* ps = sin_t + (size_t)row * head_dim;
    const int half = head_dim >> 1;
    for (int i = lane; i < half; i += 32) {
        const float x1 = __bfloat162float(px[i]), x2 = __bfloat162float(px[i + half]);
        const float c1 = pc[i], s1 = ps[i], c2 = pc[i + half], s2 = ps[i + half];
        float o1, o2;
        if (bf16_products) {
            o1 = bf16_round(x1 * c1) + bf16_round(-x2 * s1);
            o2 = bf16_round(x2 * c2) + bf16_round(x1 * s2);
        } else {
            o1 = x1 * c1 - x2 * s1;
            o2 = x2 * c2 + x1 * s2;
        }
        px[i] = __float2bfloat16_rn(o1);
        px[i + half] = __float2bfloat16_rn(o2);
    }
}

// dst row r = src row index[r / group] * group + r % group (window re-ordering of 2x2 patch groups, and its inverse on the
// merged tokens with group = 1). D % 8 == 0. One block per destination row.
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                   const int* __restrict__ index, int group, int D) {
    const size_t r = blockIdx.x;
    const size_t s = (size_t)index[r / group] * group + r % group;
    const uint4* sp = reinterpret_cast<const uint4*>(src + s * D);
    uint4* dp = reinterpret_cast<uint4*>(dst + r * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dp[i] = sp[i];
}

// fp32 (or bf16) pixel patches -> bf16 (the Conv3d input cast, modeling_qwen2_5_vl.py:112-113)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    } else {
        for (size_t k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
    }
}

// inputs_embeds = embed_tokens(ids); image-token positions are overwritten with the merged vision features
// (masked_scatter, modeling_qwen2_5_vl.py:1298-1307). feat_index[b*S+s] = row of `feats` or -1. Rows >= seq_len stay zero.
__global__ void qwen_embed_kernel(const int* __restrict__ ids, const int* __restrict__ feat_index, const int* __restrict__ seq_lens,
                                  const __nv_bfloat16* __restrict__ embed, const __nv_bfloat16* __restrict__ feats,
                                  __nv_bfloat16* __restrict__ out, int S, int D) {
    const int b = blockIdx.x / S, s = blockIdx.x % S;
    const uint4* src = nullptr;
    if (s < seq_lens[b]) {
        const int f = feat_index[blockIdx.x];
        src = (f >= 0) ? reinterpret_cast<const uint4*>(feats + (size_t)f * D)
                       : reinterpret_cast<const uint4*>(embed + (size_t)ids[blockIdx.x] * D);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)blockIdx.x * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = src ? src[i] : make_uint4(0, 0, 0, 0);
}

// out[b] = x[b*S + seq_len[b] - 1]: the last prompt position, the only one lm_head is applied to (logits_to_keep = 1).
__global__ void gather_last_rows_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ seq_lens,
                                        __nv_bfloat16* __restrict__ out, int S, int D) {
    const int b = blockIdx.x;
    const int last = max(seq_lens[b], 1) - 1;
    const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)b * S + last) * D);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)b * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = src[i];
}

// prob[b] = exp(logprob[b])
__global__ void exp_kernel(const float* __restrict__ lp, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = expf(lp[i]);
}

}  // namespace vqa
