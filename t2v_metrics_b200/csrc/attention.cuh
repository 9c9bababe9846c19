// Small attention kernels of the T5 decoder rows (head_dim = 64). The encoder / vision-tower attention is the tcgen05 kernel in
// attention_sm100.cuh.
//
//  * t5_decoder_self_attn_kernel / t5_cross_attn_kernel: the decoder rows (causal self-attention with
//    unidirectional buckets; cross-attention with zero bias + encoder padding mask, modeling_t5.py:312-325).
#pragma once
#include "ptx.cuh"
#include "elementwise.cuh"

namespace vqa {

// Decoder self-attention over the T target positions of each pair (T = 2 for the "Yes" answer, tens of tokens in VisualGPTScore
// mode where the caption itself is the target): causal, unidirectional relative buckets, no scale (modeling_t5.py:253-344).
// qkv: [B*T, 3*H*64] packed. One warp per (b, h, query t); lanes cover d in pairs; the <= T scores of the row live in shared
// memory (dynamic: warps per block * T floats).
__global__ void t5_decoder_self_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                            const __nv_bfloat16* __restrict__ rel_emb,  // [num_buckets, H]
                                            const int* __restrict__ bucket_lut,         // [2*max_dist+1] unidirectional
                                            int max_dist, int B, int T, int H, int round_scores) {
    pdl_launch_dependents();
    extern __shared__ float dec_sc[];
    const int widx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (widx >= B * H * T) return;
    const int lane = threadIdx.x & 31;
    float* sc = dec_sc + (size_t)(threadIdx.x >> 5) * T;
    const int tq = widx % T, h = (widx / T) % H, b = widx / (T * H);
    const int ld = 3 * H * 64;
    const __nv_bfloat16* qp = qkv + ((size_t)b * T + tq) * ld + h * 64;
    const float2 qv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(qp + 2 * lane));
    float mx = -INFINITY;
    for (int tk = 0; tk <= tq; ++tk) {
        const __nv_bfloat16* kp = qkv + ((size_t)b * T + tk) * ld + H * 64 + h * 64;
        const float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kp + 2 * lane));
        float v = warp_sum(qv.x * kv.x + qv.y * kv.y);
        if (round_scores) v = bf16_round(v);
        int rel = tk - tq;
        rel = min(max(rel, -max_dist), max_dist);
        v += __bfloat162float(rel_emb[bucket_lut[rel + max_dist] * H + h]);
        if (round_scores) v = bf16_round(v);
        if (lane == 0) sc[tk] = v;
        mx = fmaxf(mx, v);
    }
    __syncwarp();
    float denom = 0.f;
    for (int tk = 0; tk <= tq; ++tk) denom += __expf(sc[tk] - mx);
    float ox = 0.f, oy = 0.f;
    for (int tk = 0; tk <= tq; ++tk) {
        const __nv_bfloat16* vp = qkv + ((size_t)b * T + tk) * ld + 2 * H * 64 + h * 64;
        const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vp + 2 * lane));
        const float pw = bf16_round(__expf(sc[tk] - mx) / denom);   // attn_weights are cast back to bf16 before the PV matmul
        ox += pw * vv.x;
        oy += pw * vv.y;
    }
    *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + tq) * (H * 64) + h * 64 + 2 * lane) = pack_bf16x2(ox, oy);
}

// Cross-attention of T decoder rows against the encoder keys/values of the same sample: zero position bias,
// encoder padding mask, no scale. kv: [B*S, 2*H*64] (k | v). One block (4 warps) per (b, h); each warp owns a
// strided subset of keys for ALL T queries, partial softmax states are merged through shared memory.
template <int TMAX>
__global__ void __launch_bounds__(128) t5_cross_attn_kernel(const __nv_bfloat16* __restrict__ q,   // [B*T, H*64]
                                                           const __nv_bfloat16* __restrict__ kv,  // [B*S, ldkv]
                                                           __nv_bfloat16* __restrict__ out,       // [B*T, H*64]
                                                           const int* __restrict__ seq_lens, int ldkv, int B, int T,
                                                           int S, int H, int round_scores) {
    pdl_launch_dependents();
    const int h = blockIdx.x % H, b = blockIdx.x / H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int len = seq_lens ? seq_lens[b] : S;
    __shared__ float sm_m[4][TMAX], sm_l[4][TMAX];
    __shared__ float sm_o[4][TMAX][64];

    float2 qv[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
        qv[t] = (t < T) ? unpack_bf16x2(*reinterpret_cast<const uint32_t*>(
                              q + ((size_t)b * T + t) * (H * 64) + h * 64 + 2 * lane))
                        : make_float2(0.f, 0.f);
    float m[TMAX], l[TMAX], ox[TMAX], oy[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) { m[t] = -INFINITY; l[t] = 0.f; ox[t] = 0.f; oy[t] = 0.f; }

    for (int s = warp; s < len; s += 4) {
        const __nv_bfloat16* kp = kv + ((size_t)b * S + s) * ldkv + h * 64;
        const float2 kk = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kp + 2 * lane));
        const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kp + H * 64 + 2 * lane));
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
            if (t < T) {
                float sc = warp_sum(qv[t].x * kk.x + qv[t].y * kk.y);
                if (round_scores) sc = bf16_round(sc);
                const float m_new = fmaxf(m[t], sc);
                const float c = __expf(m[t] - m_new);
                const float pw = __expf(sc - m_new);
                l[t] = l[t] * c + pw;
                ox[t] = ox[t] * c + pw * vv.x;
                oy[t] = oy[t] * c + pw * vv.y;
                m[t] = m_new;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
        if (lane == 0) { sm_m[warp][t] = m[t]; sm_l[warp][t] = l[t]; }
        sm_o[warp][t][2 * lane] = ox[t];
        sm_o[warp][t][2 * lane + 1] = oy[t];
    }
    __syncthreads();
    // merge: thread (t, d) pairs
    for (int i = threadIdx.x; i < T * 64; i += blockDim.x) {
        const int t = i / 64, d = i % 64;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, sm_m[w][t]);
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (sm_m[w][t] > -INFINITY) {
                const float c = __expf(sm_m[w][t] - M);
                L += sm_l[w][t] * c;
                O += sm_o[w][t][d] * c;
            }
        }
        out[((size_t)b * T + t) * (H * 64) + h * 64 + d] = __float2bfloat16_rn(L > 0.f ? O / L : 0.f);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Absorbed cross-attention helpers (see vqa_b200.cu: q.(Wk h) == (Wk^T q).h and sum p (Wv h) == Wv (sum p h)).

// [B, S, D] -> [B, D, Sp] (Sp >= S, multiple of 8; columns >= S are zero). 32x32 tiles through shared memory.
__global__ void transpose_bsd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ xt, int S, int D,
                                     int Sp) {
    pdl_launch_dependents();
    __shared__ __nv_bfloat16 tile[32][33];
    const int b = blockIdx.z;
    const int s0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int s = s0 + i, d = d0 + tx;
        tile[i][tx] = (s < S && d < D) ? x[((size_t)b * S + s) * D + d] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int d = d0 + i, s = s0 + tx;
        if (d < D && s < Sp) xt[((size_t)b * D + d) * Sp + s] = tile[tx][i];
    }
}

// In-place masked softmax over the key axis of scores [B*rows_per_b, Sp] (bf16): keys >= seq_len[b] get probability 0.
// fp32 softmax, probabilities cast back to bf16 (modeling_t5.py:331). One warp per row.
__global__ void cross_softmax_kernel(__nv_bfloat16* __restrict__ sc, const int* __restrict__ seq_lens, int rows_per_b,
                                     int total_rows, int S, int Sp) {
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int lane = threadIdx.x & 31;
    const int b = row / rows_per_b;
    const int len = seq_lens ? min(seq_lens[b], S) : S;
    __nv_bfloat16* r = sc + (size_t)row * Sp;
    float mx = -INFINITY;
    for (int i = lane; i < len; i += 32) mx = fmaxf(mx, __bfloat162float(r[i]));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < len; i += 32) sum += __expf(__bfloat162float(r[i]) - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int i = lane; i < Sp; i += 32)
        r[i] = __float2bfloat16_rn(i < len ? __expf(__bfloat162float(r[i]) - mx) * inv : 0.f);
}

}  // namespace vqa
