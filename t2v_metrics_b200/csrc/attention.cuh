// Fused attention kernels for the CLIP-FlanT5 scoring path (head_dim = 64 everywhere on this path).
//
//  * flash_attn_d64_kernel: bidirectional self-attention with (a) T5's learned relative-position bias read from a
//    per-head [2S-1] table, no 1/sqrt(d) scale, key-padding mask (transformers/models/t5/modeling_t5.py:308-334),
//    or (b) CLIP's plain scaled attention (transformers/models/clip/modeling_clip.py:261-336). Scores never
//    touch HBM (the reference materialises [B,H,S,S] fp32). Online softmax in fp32.
//  * t5_decoder_self_attn_kernel / t5_cross_attn_kernel: the decoder rows (causal self-attention with
//    unidirectional buckets; cross-attention with zero bias + encoder padding mask, modeling_t5.py:312-325).
#pragma once
#include "ptx.cuh"
#include "elementwise.cuh"

namespace vqa {

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct FlashParams {
    const __nv_bfloat16* q;  // row (b*S + s), column h*64 + d, row stride ldq
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;        // [B*S, ldo], column h*64 + d
    int ldq, ldk, ldv, ldo;
    const int* seq_lens;     // [B] valid keys/queries per sample, or nullptr (all S valid)
    const float* bias_table; // [H, 2S-1] or nullptr
    int S, H;
    float scale;             // multiplies q.k (1 for T5)
    int round_scores;        // emulate the reference's bf16 matmul output / bias add rounding
};

constexpr int FA_BQ = 64, FA_BK = 64, FA_D = 64, FA_LD = 72;  // padded smem row stride (elements)

// grid (ceil(S/64), H, B), 128 threads. Dynamic smem: (1 + 2 + 2) tiles * 64*72*2 B + (2S-1)*4 B.
__global__ void __launch_bounds__(128) flash_attn_d64_kernel(const FlashParams p) {
    extern __shared__ __align__(16) uint8_t fa_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(fa_smem);
    __nv_bfloat16* sK = sQ + FA_BQ * FA_LD;           // [2][64][72]
    __nv_bfloat16* sV = sK + 2 * FA_BK * FA_LD;       // [2][64][72]
    float* sBias = reinterpret_cast<float*>(sV + 2 * FA_BK * FA_LD);

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int len = p.seq_lens ? p.seq_lens[b] : p.S;
    const int q0 = qt * FA_BQ;
    const size_t row_base = (size_t)b * p.S;

    if (q0 >= len) {  // padded query tile: deterministic zeros
        for (int i = tid; i < FA_BQ * 8; i += 128) {
            const int r = i >> 3, c = i & 7;
            if (q0 + r < p.S)
                *reinterpret_cast<uint4*>(p.o + (row_base + q0 + r) * p.ldo + h * FA_D + c * 8) = make_uint4(0, 0, 0, 0);
        }
        return;
    }

    // ---- async loads: Q tile, then K/V tile 0
    auto load_tile = [&](__nv_bfloat16* dst, const __nv_bfloat16* src, int ld, int r0) {
        for (int i = tid; i < 64 * 8; i += 128) {
            const int r = i >> 3, c = i & 7;
            const bool ok = (r0 + r) < len;
            const __nv_bfloat16* gp = src + (row_base + (ok ? r0 + r : 0)) * ld + h * FA_D + c * 8;
            cp_async_16(dst + r * FA_LD + c * 8, gp, ok);
        }
    };
    load_tile(sQ, p.q, p.ldq, q0);
    cp_async_commit();
    const int num_kt = (len + FA_BK - 1) / FA_BK;
    load_tile(sK, p.k, p.ldk, 0);
    load_tile(sV, p.v, p.ldv, 0);
    cp_async_commit();
    const int bias_w = 2 * p.S - 1;
    if (p.bias_table)
        for (int i = tid; i < bias_w; i += 128) sBias[i] = p.bias_table[(size_t)h * bias_w + i];

    cp_async_wait<1>();
    __syncthreads();
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4(qf[ks], sQ + r * FA_LD + c);
    }

    float o_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) o_acc[i][e] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int qrow0 = q0 + warp * 16 + g;  // rows qrow0 and qrow0 + 8

    for (int kt = 0; kt < num_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < num_kt) {
            load_tile(sK + (buf ^ 1) * FA_BK * FA_LD, p.k, p.ldk, (kt + 1) * FA_BK);
            load_tile(sV + (buf ^ 1) * FA_BK * FA_LD, p.v, p.ldv, (kt + 1) * FA_BK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __nv_bfloat16* cK = sK + buf * FA_BK * FA_LD;
        const __nv_bfloat16* cV = sV + buf * FA_BK * FA_LD;

        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t kf[4];
                const int r = np * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(kf, cK + r * FA_LD + c);
                mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }

        // ---- scale / bias / mask, online softmax
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kcol = kt * FA_BK + nt * 8 + 2 * t4 + (e & 1);
                const int qrow = qrow0 + (e >> 1) * 8;
                float v = s[nt][e] * p.scale;
                if (p.round_scores) v = bf16_round(v);
                if (p.bias_table) {
                    int bi = kcol - qrow + p.S - 1;
                    bi = min(max(bi, 0), bias_w - 1);
                    v += sBias[bi];
                    if (p.round_scores) v = bf16_round(v);
                }
                if (kcol >= len) v = -INFINITY;
                s[nt][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);   // finite: key 0 of tile 0 is always valid
            corr[r] = __expf(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float e0 = __expf(s[nt][0] - m_run[0]);
            float e1 = __expf(s[nt][1] - m_run[0]);
            float e2 = __expf(s[nt][2] - m_run[1]);
            float e3 = __expf(s[nt][3] - m_run[1]);
            rs[0] += e0 + e1;
            rs[1] += e2 + e3;
            const int kk = nt >> 1;
            if ((nt & 1) == 0) { pf[kk][0] = pack_bf16x2(e0, e1); pf[kk][1] = pack_bf16x2(e2, e3); }
            else               { pf[kk][2] = pack_bf16x2(e0, e1); pf[kk][3] = pack_bf16x2(e2, e3); }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
            l_run[r] = l_run[r] * corr[r] + rs[r];
        }
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            o_acc[dt][0] *= corr[0]; o_acc[dt][1] *= corr[0];
            o_acc[dt][2] *= corr[1]; o_acc[dt][3] *= corr[1];
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vf[4];
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = dp * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(vf, cV + r * FA_LD + c);
                mma_bf16_16816(o_acc[2 * dp], pf[kk], vf[0], vf[1]);
                mma_bf16_16816(o_acc[2 * dp + 1], pf[kk], vf[2], vf[3]);
            }
        }
        __syncthreads();  // all warps done with buf before it is refilled
    }

    // ---- normalise and store
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        const int col = h * FA_D + dt * 8 + 2 * t4;
        if (qrow0 < p.S) {
            const bool ok = qrow0 < len;
            *reinterpret_cast<uint32_t*>(p.o + (row_base + qrow0) * p.ldo + col) =
                ok ? pack_bf16x2(o_acc[dt][0] * inv0, o_acc[dt][1] * inv0) : 0u;
        }
        if (qrow0 + 8 < p.S) {
            const bool ok = qrow0 + 8 < len;
            *reinterpret_cast<uint32_t*>(p.o + (row_base + qrow0 + 8) * p.ldo + col) =
                ok ? pack_bf16x2(o_acc[dt][2] * inv1, o_acc[dt][3] * inv1) : 0u;
        }
    }
}

inline size_t flash_smem_bytes(int S, bool has_bias) {
    return (size_t)5 * FA_BQ * FA_LD * 2 + (has_bias ? (size_t)(2 * S - 1) * 4 : 0) + 16;
}

// Decoder self-attention over the T target positions of each pair (T = 2 for the "Yes" answer, tens of tokens in VisualGPTScore
// mode where the caption itself is the target): causal, unidirectional relative buckets, no scale (modeling_t5.py:253-344).
// qkv: [B*T, 3*H*64] packed. One warp per (b, h, query t); lanes cover d in pairs; the <= T scores of the row live in shared
// memory (dynamic: warps per block * T floats).
__global__ void t5_decoder_self_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                            const __nv_bfloat16* __restrict__ rel_emb,  // [num_buckets, H]
                                            const int* __restrict__ bucket_lut,         // [2*max_dist+1] unidirectional
                                            int max_dist, int B, int T, int H, int round_scores) {
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
