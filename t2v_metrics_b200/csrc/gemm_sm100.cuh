// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] . W[N,K]^T )
//
//   * operands are both K-major (activations [M,K] row-major, nn.Linear weights [N,K] row-major), fetched
//     tile-by-tile from HBM/L2 by TMA into a 128B-swizzled shared-memory ring;
//   * one elected thread issues tcgen05.mma (UMMA 128xBLOCK_Nx16, or 256xBLOCK_Nx16 across a CTA pair with
//     cta_group::2); accumulators live in TMEM, double-buffered so the epilogue of tile i overlaps the
//     main loop of tile i+1;
//   * four epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue (bias / activation /
//     gated-GELU / residual add / online log-sum-exp for the lm_head) before the only write to HBM.
//
// Replaces the cuBLAS bf16 GEMMs that the reference reaches through nn.Linear
// (transformers/models/t5/modeling_t5.py:109-111,178-181; transformers/models/clip/modeling_clip.py:296-299,344-345).
#pragma once
#include "ptx.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <unordered_map>

namespace vqa {

#ifndef VQA_GEMM_WAIT_HINT
#define VQA_GEMM_WAIT_HINT 1
#endif
// mbarrier wait used by the GEMM's producer / MMA / epilogue warps: with the suspend-time hint (default) or the plain polling loop
__device__ __forceinline__ void gemm_wait(uint64_t* bar, uint32_t parity) {
#if VQA_GEMM_WAIT_HINT
    mbar_wait(bar, parity);
#else
    mbar_wait_spin(bar, parity);
#endif
}


enum GemmEpilogue : int {
    EPI_STORE = 0,       // C = [residual +] bf16(acc [+ bias])
    EPI_QUICK_GELU = 1,  // C = quick_gelu(bf16(acc + bias))                 (CLIP MLP fc1)
    EPI_GELU_ERF = 2,    // C = gelu_erf(bf16(acc + bias))                   (mlp2x_gelu projector)
    EPI_GATED_GELU = 3,  // C = gelu_new(bf16(acc_gate)) * bf16(acc_up)      (T5 DenseGatedActDense wi_0/wi_1)
    EPI_LSE = 4,         // no C; per-row (max, sum exp) partials + label-logit gather (lm_head + CE)
    EPI_RELU = 5,        // C = relu(bf16(acc + bias))
    EPI_GATED_SILU = 6,  // C = silu(bf16(acc_gate + b_gate)) * bf16(acc_up + b_up)   (Qwen2.5-VL SwiGLU MLPs)
};
__host__ __device__ constexpr bool epi_is_gated(int epi) { return epi == EPI_GATED_GELU || epi == EPI_GATED_SILU; }

struct GemmParams {
    int M, N, K;              // N = number of OUTPUT columns of the logical GEMM (for GATED: 2 * out columns)
    __nv_bfloat16* C;         // [M, ldc]
    int ldc;
    const __nv_bfloat16* bias;      // [N] or nullptr
    const __nv_bfloat16* residual;  // [M, ldr] or nullptr
    int ldr;
    int c_f32;                // EPI_STORE only. 1: C and residual are fp32 buffers (ldc / ldr in floats): C = residual + bf16(acc + bias), unrounded
                              // 2: C is an fp32 buffer that receives the RAW accumulator (split-K partial sums: batches = K slices, reduced by
                              //    splitk_reduce_kernel, which also applies bias / rounding / residual)
                              // (the CLIP tower's residual stream stays fp32 under the reference's autocast: LayerNorm is on autocast's fp32 list)
    int gate_up_offset;       // GATED: row offset of the "up" weight block inside W (= d_ff)
    int c_group_in, c_group_out;   // EPI_STORE without residual, c_group_in > 0: logical output column c is stored at column
                              // (c / c_group_in) * c_group_out + c % c_group_in -- heads narrower than the attention kernel's 128-wide slots
                              // (Qwen2.5-VL vision tower, 80-wide) are produced at their native width and written into the slots; both multiples of 8
    // EPI_LSE
    float* lse_max;           // [M, num_n_tiles]
    float* lse_sum;           // [M, num_n_tiles]
    const int* labels;        // [M]
    float* label_logit;       // [M]
    float lse_scale;          // logits are multiplied by this (1 / temperature) after the bf16 rounding
    // optional repetition penalty (HF RepetitionPenaltyLogitsProcessor, generation/logits_process.py): bit n of row m's bitmap
    // set -> logit = logit < 0 ? logit * penalty : logit / penalty, applied to the bf16-rounded fp32 logit BEFORE the 1/T scale
    const uint32_t* penalty_bitmap;   // [M, penalty_words] or nullptr
    int penalty_words;
    float penalty;
    // RMSNorm folded into the GEMMs around it (T5LayerNorm modeling_t5.py:55-68 / Qwen2RMSNorm): the GEMM that WRITES the residual stream
    // (EPI_STORE with a residual) also emits, per row, the sum of squares of the bf16 values it stores for its slice of columns
    // (ssq_out[m * ssq_out_parts + n_tile * parts_per_tile + column_half]); the GEMM that CONSUMES the stream takes the un-normalised rows
    // as its A operand with gamma folded into W, sums the partials and multiplies its fp32 accumulator row by rsqrt(sum / dim + eps)
    // before the bf16 rounding of the Linear's output (EPI_STORE, EPI_GATED_*). No normalised copy of the stream is ever written.
    float* ssq_out;
    int ssq_out_parts;
    const float* ssq_in;
    int ssq_in_parts;
    float ssq_inv_dim, ssq_eps;
    // batching: `num_batches` independent GEMMs of the same M,N,K share one launch; batch b reads A at
    // (row + b*a_row_off, k + b*a_k_off), W at (row + b*w_row_off, k + b*w_k_off) and writes C + b*c_batch_stride.
    int num_batches;
    int a_row_off, a_k_off, w_row_off, w_k_off;
    long long c_batch_stride;
    // scheduling: W is cut into chunks of `chunk_n` N tiles; inside a chunk, groups of `group_m` M tiles sweep the chunk's N tiles
    int num_m_tiles, num_n_tiles, group_m, chunk_n;
    // L2 eviction priority of the two operand streams (ptx.cuh L2_EVICT_*), chosen per launch by launch_gemm_t
    unsigned long long a_policy, w_policy;
};

template <int BLOCK_N, int CG>
struct GemmConfig {
    static constexpr int BLOCK_M = 128;           // rows per CTA (UMMA M = 128 * CG)
    static constexpr int BLOCK_K = 64;            // 64 bf16 = 128 B = one swizzle atom
    static constexpr int UMMA_K = 16;
    static constexpr int B_ROWS_PER_CTA = BLOCK_N / CG;
    static constexpr int B_HALF_ROWS = BLOCK_N / 2;
    static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_STAGE_BYTES = B_ROWS_PER_CTA * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int SMEM_BUDGET = 200 * 1024;
    static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int NUM_THREADS = 320;       // warp0 TMA, warp1 MMA(+TMEM alloc), warps 2..9 epilogue
    // the second epilogue warp group takes half of the 32-column chunks when a tile has at least two of them
    static constexpr bool EPI_SPLIT = BLOCK_N >= 128;   // >= 2 chunks also for the gated epilogue (BLOCK_N/2 >= 64)
    static constexpr int EPI_ARRIVALS = EPI_SPLIT ? 8 : 4;
    static constexpr int LSE_PARTS = EPI_SPLIT ? 2 : 1;
    static constexpr int TMEM_COLS = 2 * BLOCK_N; // two accumulator stages
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(BLOCK_N >= 32 && BLOCK_N <= 256 && (BLOCK_N & (BLOCK_N - 1)) == 0, "BLOCK_N must be 32..256, pow2");
    static_assert(B_HALF_ROWS % 8 == 0, "B half tile must hold whole 8-row swizzle atoms");
    static_assert(STAGES >= 3, "pipeline too shallow");
};

__device__ __forceinline__ float act_quick_gelu(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float act_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float act_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float act_gelu_new(float x) {
    // 0.5 * x * (1 + tanh(sqrt(2/pi) * (x + 0.044715 x^3)))  -- transformers/activations.py NewGELUActivation, evaluated in fp32 like the
    // reference does under CUDA autocast (torch.pow is on autocast's fp32 list, so everything downstream of it is fp32).
    // tanh(u) = 1 - 2 / (exp(2u) + 1) with ex2.approx + rcp.approx: relative error ~1e-6. (tanh.approx.f32 is 2^-11 absolute, i.e. a
    // SYSTEMATIC 5e-4 error in 48 FFNs -- too coarse for parity with the fp32-evaluated reference.)
    const float k = 0.7978845608028654f;
    const float inner = k * (x + 0.044715f * x * x * x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(inner * 2.8853900817779268f));   // exp(2u) = 2^(2u log2 e)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    const float th = 1.0f - 2.0f * r;
    return 0.5f * x * (1.0f + th);
}
// CLIP's quick_gelu as eager bf16 tensor ops run it (`input * torch.sigmoid(1.702 * input)`, transformers/activations.py:117-123; none of
// the three ops is on an autocast list, so each produces a bf16 tensor): y bf16 -> bf16(1.702 y) -> bf16(sigmoid) -> product (rounded by caller)
__device__ __forceinline__ float act_quick_gelu_bf16_ops(float y) {
    const float a = bf16_round(1.702f * y);
    const float sg = bf16_round(__fdividef(1.0f, 1.0f + __expf(-a)));
    return y * sg;
}

template <int BLOCK_N, int CG, int EPI>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const GemmParams p) {
    using Cfg = GemmConfig<BLOCK_N, CG>;
    static_assert(!epi_is_gated(EPI) || BLOCK_N >= 64, "gated epilogue pairs two >=32-column half tiles");
    pdl_launch_dependents();
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BLOCK_M = Cfg::BLOCK_M;
    constexpr int BLOCK_K = Cfg::BLOCK_K;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                      // [STAGES][128 x 64 bf16]
    uint8_t* smem_b = smem + STAGES * Cfg::A_STAGE_BYTES;        // [STAGES][B_ROWS_PER_CTA x 64 bf16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tmem_full_bar = bars + 2 * STAGES;      // [2]
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2; // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool is_leader = (cta_rank == 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], CG);   // CG==2: one arrive per CTA's producer, on the leader's barrier
            mbar_init(&empty_bar[s], 1);   // one tcgen05.commit
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);        // one tcgen05.commit
            mbar_init(&tmem_empty_bar[s], CG * Cfg::EPI_ARRIVALS);  // one arrive per active epilogue warp (both CTAs)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc<CG>(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tcgen05_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();      // everything above touched only this CTA's shared memory / TMEM; from here on operands of earlier kernels are read

    const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int tiles_per_batch = p.num_m_tiles * p.num_n_tiles;
    const int num_tiles = tiles_per_batch * p.num_batches;
    const int num_workers = gridDim.x / CG;
    const int worker = blockIdx.x / CG;
    const int tiles_per_chunk = p.num_m_tiles * p.chunk_n;

    // Tile order (L2 blocking): W chunk (chunk_n N tiles, sized to stay resident in L2) -> group of group_m M tiles -> the chunk's
    // N tiles -> the group's M tiles (fastest). One wave of CTAs therefore shares group_m A row panels and a few W panels, the W chunk
    // is re-read from L2 by every M group, and A streams through once per chunk.
    auto tile_coords = [&](int t, int& batch, int& m_blk, int& n_blk) {
        batch = t / tiles_per_batch;
        t -= batch * tiles_per_batch;
        const int c = t / tiles_per_chunk;
        t -= c * tiles_per_chunk;
        const int n_first = c * p.chunk_n;
        const int cn = min(p.chunk_n, p.num_n_tiles - n_first);
        const int tiles_per_group = p.group_m * cn;
        const int g = t / tiles_per_group;
        const int r = t - g * tiles_per_group;
        const int m_first = g * p.group_m;
        const int gm = min(p.group_m, p.num_m_tiles - m_first);
        m_blk = m_first + r % gm;
        n_blk = n_first + r / gm;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = worker; t < num_tiles; t += num_workers) {
                int batch, m_blk, n_blk;
                tile_coords(t, batch, m_blk, n_blk);
                const int m_row = (m_blk * CG + (int)cta_rank) * BLOCK_M + batch * p.a_row_off;
                const int w_row_base = batch * p.w_row_off;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    gemm_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sa = smem_a + stage * Cfg::A_STAGE_BYTES;
                    uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
                    const int k0 = kb * BLOCK_K + batch * p.a_k_off;
                    const int k0w = kb * BLOCK_K + batch * p.w_k_off;
                    if constexpr (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d_hint(sa, &tmap_a, &full_bar[stage], k0, m_row, p.a_policy);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            int n_row = epi_is_gated(EPI) ? h * p.gate_up_offset + n_blk * Cfg::B_HALF_ROWS
                                                                : n_blk * BLOCK_N + h * Cfg::B_HALF_ROWS;
                            tma_load_2d_hint(sb + h * Cfg::B_HALF_ROWS * BLOCK_K * 2, &tmap_b, &full_bar[stage], k0w,
                                             n_row + w_row_base, p.w_policy);
                        }
                    } else {
                        const int h = (int)cta_rank;
                        int n_row = epi_is_gated(EPI) ? h * p.gate_up_offset + n_blk * Cfg::B_HALF_ROWS
                                                            : n_blk * BLOCK_N + h * Cfg::B_HALF_ROWS;
                        tma_load_2d_2sm_hint(sa, &tmap_a, &full_bar[stage], k0, m_row, p.a_policy);
                        tma_load_2d_2sm_hint(sb, &tmap_b, &full_bar[stage], k0w, n_row + w_row_base, p.w_policy);
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        else           mbar_arrive_cluster(&full_bar[stage], 0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only, one thread) =====================
        if (lane == 0 && is_leader) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(BLOCK_M * CG, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = worker; t < num_tiles; t += num_workers, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1u;
                gemm_wait(&tmem_empty_bar[as], aphase ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    gemm_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem_a + stage * Cfg::A_STAGE_BYTES));
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / Cfg::UMMA_K; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-byte units
                        umma_f16<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit<CG>(&empty_bar[stage]);                     // frees the smem slot when MMAs retire
                    if (kb == num_k_blocks - 1) umma_commit<CG>(&tmem_full_bar[as]);  // accumulator ready
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue warps (TMEM -> registers -> HBM) =====================
        // Eight warps: warp w may only touch TMEM lanes 32*(w%4).., so warps {2..5} and {6..9} each cover all four lane
        // quadrants; the first group drains the first half of the tile's column chunks, the second group the rest.
        const uint32_t q = warp & 3u;
        const uint32_t half = (warp - 2u) >> 2;
        const int row_in_tile = q * 32 + lane;
        constexpr int OUT_TILE_COLS = epi_is_gated(EPI) ? BLOCK_N / 2 : BLOCK_N;
        constexpr int NCH = OUT_TILE_COLS / 32;                    // 32-column chunks per tile
        constexpr bool SPLIT = Cfg::EPI_SPLIT;
        constexpr int NCH_PER = SPLIT ? NCH / 2 : NCH;
        const int c_begin = SPLIT ? (int)half * NCH_PER : 0;
        if (SPLIT || half == 0) {
        int it = 0;
        for (int t = worker; t < num_tiles; t += num_workers, ++it) {
            int batch, m_blk, n_blk;
            tile_coords(t, batch, m_blk, n_blk);
            const long long c_off = (long long)batch * p.c_batch_stride;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1u;
            const int m = (m_blk * CG + (int)cta_rank) * BLOCK_M + row_in_tile;
            const bool row_ok = m < p.M;
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * BLOCK_N;
            float row_scale = 1.0f;     // fused RMSNorm of the A rows (consumer side)
            if (p.ssq_in && row_ok) {
                const float* sp = p.ssq_in + (size_t)m * p.ssq_in_parts;
                float ssum = 0.f;
                for (int i = 0; i < p.ssq_in_parts; i += 4) {
                    const float4 v4 = __ldg(reinterpret_cast<const float4*>(sp + i));
                    ssum += (v4.x + v4.y) + (v4.z + v4.w);
                }
                row_scale = rsqrtf(ssum * p.ssq_inv_dim + p.ssq_eps);
            }

            if constexpr (epi_is_gated(EPI)) {
                const int n_out0 = n_blk * OUT_TILE_COLS;
                gemm_wait(&tmem_full_bar[as], aphase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_begin + NCH_PER; ++c) {
                    uint32_t g[32], u[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, g);
                    tmem_ld_32x32b_x32(taddr + OUT_TILE_COLS + c * 32, u);
                    tmem_ld_wait();
                    if (row_ok) {
                        __nv_bfloat16* crow = p.C + c_off + (size_t)m * p.ldc + n_out0 + c * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t w[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float g0 = __uint_as_float(g[j * 8 + 2 * e]) * row_scale, g1 = __uint_as_float(g[j * 8 + 2 * e + 1]) * row_scale;
                                float u0 = __uint_as_float(u[j * 8 + 2 * e]) * row_scale, u1 = __uint_as_float(u[j * 8 + 2 * e + 1]) * row_scale;
                                if (p.bias) {   // [gate bias | up bias], `gate_up_offset` apart like the weight rows
                                    const int n = n_out0 + c * 32 + j * 8 + 2 * e;
                                    if (n < p.N / 2) {
                                        const float2 bg = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.bias + n));
                                        const float2 bu = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.bias + p.gate_up_offset + n));
                                        g0 += bg.x; g1 += bg.y; u0 += bu.x; u1 += bu.y;
                                    }
                                }
                                g0 = bf16_round(g0); g1 = bf16_round(g1); u0 = bf16_round(u0); u1 = bf16_round(u1);
                                float h0, h1;
                                if constexpr (EPI == EPI_GATED_GELU) {
                                    // T5DenseGatedActDense under CUDA autocast: act(wi_0 x) comes out in fp32 (pow -> fp32), the product
                                    // with the bf16 wi_1 x is fp32, and the single rounding is the cast in front of `wo` (modeling_t5.py:115-131)
                                    h0 = act_gelu_new(g0) * u0;
                                    h1 = act_gelu_new(g1) * u1;
                                } else {
                                    h0 = bf16_round(act_silu(g0)) * u0;
                                    h1 = bf16_round(act_silu(g1)) * u1;
                                }
                                w[e] = pack_bf16x2(h0, h1);
                            }
                            if (n_out0 + c * 32 + j * 8 < p.N / 2)
                                *reinterpret_cast<uint4*>(crow + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
            } else if constexpr (EPI == EPI_LSE) {
                const int n0 = n_blk * BLOCK_N;
                const int label = row_ok ? p.labels[m] : -1;
                float run_max = -INFINITY, run_sum = 0.f;
                gemm_wait(&tmem_full_bar[as], aphase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_begin + NCH_PER; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    float x[32];
                    float cmax = -INFINITY;
                    const uint32_t pen_bits = (p.penalty_bitmap && row_ok && n0 + c * 32 < p.N)
                                                  ? p.penalty_bitmap[(size_t)m * p.penalty_words + ((n0 + c * 32) >> 5)] : 0u;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + c * 32 + j;
                        float raw = bf16_round(__uint_as_float(v[j]));                  // reference logits are bf16, then fp32
                        if (pen_bits && ((pen_bits >> j) & 1u)) raw = raw < 0.f ? raw * p.penalty : __fdiv_rn(raw, p.penalty);
                        float val = raw * p.lse_scale;                                   // ... / T in fp32
                        x[j] = (n < p.N) ? val : -INFINITY;
                        cmax = fmaxf(cmax, x[j]);
                        if (n == label) p.label_logit[m] = val;
                    }
                    if (cmax > -INFINITY) {
                        const float new_max = fmaxf(run_max, cmax);
                        float sacc = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) sacc += __expf(x[j] - new_max);
                        run_sum = run_sum * __expf(run_max - new_max) + sacc;
                        run_max = new_max;
                    }
                }
                if (row_ok) {
                    // one partial per (row, n tile, column half): Cfg::LSE_PARTS partials per tile
                    const size_t slot = ((size_t)m * p.num_n_tiles + n_blk) * Cfg::LSE_PARTS + (SPLIT ? half : 0);
                    p.lse_max[slot] = run_max;
                    p.lse_sum[slot] = run_sum;
                }
            } else if (EPI == EPI_STORE && p.c_f32) {
                const int n0 = n_blk * BLOCK_N;
                float* Cf = reinterpret_cast<float*>(p.C);
                const float* Rf = reinterpret_cast<const float*>(p.residual);
                gemm_wait(&tmem_full_bar[as], aphase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_begin + NCH_PER; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    const int nc = n0 + c * 32;
                    if (row_ok && nc < p.N) {
                        float* crow = Cf + c_off + (size_t)m * p.ldc + nc;
                        const float* rrow = Rf ? Rf + c_off + (size_t)m * p.ldr + nc : nullptr;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (nc + j * 4 < p.N) {
                                if (p.c_f32 == 2) {
                                    *reinterpret_cast<float4*>(crow + j * 4) = make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]),
                                                                                          __uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3]));
                                    continue;
                                }
                                float4 r = rrow ? *reinterpret_cast<const float4*>(rrow + j * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                                float b[4] = {0.f, 0.f, 0.f, 0.f};
                                if (p.bias) {
                                    const uint2 bb = __ldg(reinterpret_cast<const uint2*>(p.bias + nc + j * 4));
                                    const float2 b0 = unpack_bf16x2(bb.x), b1 = unpack_bf16x2(bb.y);
                                    b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y;
                                }
                                r.x += bf16_round(__uint_as_float(v[j * 4]) + b[0]);
                                r.y += bf16_round(__uint_as_float(v[j * 4 + 1]) + b[1]);
                                r.z += bf16_round(__uint_as_float(v[j * 4 + 2]) + b[2]);
                                r.w += bf16_round(__uint_as_float(v[j * 4 + 3]) + b[3]);
                                *reinterpret_cast<float4*>(crow + j * 4) = r;
                            }
                        }
                    }
                }
            } else {
                const int n0 = n_blk * BLOCK_N;
                // Bias and residual for the FIRST chunk are requested before waiting for the accumulator, and the next
                // chunk's while the current one is being converted, so their L2 latency overlaps the TMEM traffic.
                uint4 bq[4], rq[4];
                auto prefetch = [&](int c) {
                    const int nc = n0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = (nc + j * 8) < p.N;
                        bq[j] = (p.bias && ok) ? __ldg(reinterpret_cast<const uint4*>(p.bias + nc + j * 8)) : make_uint4(0, 0, 0, 0);
                        rq[j] = (p.residual && ok && row_ok)
                                    ? *reinterpret_cast<const uint4*>(p.residual + c_off + (size_t)m * p.ldr + nc + j * 8)
                                    : make_uint4(0, 0, 0, 0);
                    }
                };
                prefetch(c_begin);
                float ssq_acc = 0.f;        // producer side of the fused RMSNorm: sum of squares of the bf16 values this thread stores
                gemm_wait(&tmem_full_bar[as], aphase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_begin + NCH_PER; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    uint4 bcur[4], rcur[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { bcur[j] = bq[j]; rcur[j] = rq[j]; }
                    if (c + 1 < c_begin + NCH_PER) prefetch(c + 1);
                    tmem_ld_wait();
                    const int nc = n0 + c * 32;
                    if (row_ok && nc < p.N) {
                        __nv_bfloat16* crow = p.C + c_off + (size_t)m * p.ldc + nc;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (nc + j * 8 < p.N) {
                                float f[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]) * row_scale;
                                {
                                    float2 b0 = unpack_bf16x2(bcur[j].x), b1 = unpack_bf16x2(bcur[j].y);
                                    float2 b2 = unpack_bf16x2(bcur[j].z), b3 = unpack_bf16x2(bcur[j].w);
                                    f[0] += b0.x; f[1] += b0.y; f[2] += b1.x; f[3] += b1.y;
                                    f[4] += b2.x; f[5] += b2.y; f[6] += b3.x; f[7] += b3.y;
                                }
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    float y = bf16_round(f[e]);   // the Linear's bf16 output in the reference
                                    if constexpr (EPI == EPI_QUICK_GELU) y = act_quick_gelu_bf16_ops(y);
                                    if constexpr (EPI == EPI_GELU_ERF) y = act_gelu_erf(y);
                                    if constexpr (EPI == EPI_RELU) y = fmaxf(y, 0.f);
                                    f[e] = y;
                                }
                                {
                                    float2 r0 = unpack_bf16x2(rcur[j].x), r1 = unpack_bf16x2(rcur[j].y);
                                    float2 r2 = unpack_bf16x2(rcur[j].z), r3 = unpack_bf16x2(rcur[j].w);
                                    f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
                                    f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
                                }
                                uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                     pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                                __nv_bfloat16* dst = crow + j * 8;
                                if (p.c_group_in > 0) {
                                    const int cc = nc + j * 8;
                                    dst = p.C + c_off + (size_t)m * p.ldc + (cc / p.c_group_in) * p.c_group_out + cc % p.c_group_in;
                                }
                                *reinterpret_cast<uint4*>(dst) = o;
                                if (p.ssq_out) {
                                    const float2 s0 = unpack_bf16x2(o.x), s1 = unpack_bf16x2(o.y), s2 = unpack_bf16x2(o.z), s3 = unpack_bf16x2(o.w);
                                    ssq_acc += (s0.x * s0.x + s0.y * s0.y) + (s1.x * s1.x + s1.y * s1.y) + (s2.x * s2.x + s2.y * s2.y) +
                                               (s3.x * s3.x + s3.y * s3.y);
                                }
                            }
                        }
                    }
                }
                if (p.ssq_out && row_ok)
                    p.ssq_out[(size_t)m * p.ssq_out_parts + n_blk * Cfg::LSE_PARTS + (SPLIT ? half : 0)] = ssq_acc;
            }
            // release this accumulator stage back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(&tmem_empty_bar[as]);
                else                   mbar_arrive_cluster(&tmem_empty_bar[as], 0);
            }
        }
        }
    }

    // ---- teardown
    tcgen05_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// bf16 row-major [rows, cols] (row stride ld elements) -> 2-D map, box {64 cols, box_rows}, 128B swizzle.
inline bool make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                              uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// process-wide tuning overrides (vqa_set_gemm_schedule); 0 = automatic
static int g_gemm_group_rows_override = 0;
static int g_gemm_chunk_rows_override = 0;

// The same (base, shape) pairs come back on every forward (weights and workspace slices are stable), and encoding a CUtensorMap costs
// a driver call: keep the encoded maps per thread. The map only describes address + extents, so a hit is valid whatever lives there.
struct TmapKey {
    const void* base; uint64_t rows, cols, ld; uint32_t box_rows;
    bool operator==(const TmapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
        h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
        h ^= (k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2));
        h ^= (k.ld * 0x165667B19E3779F9ull + k.box_rows + (h << 6) + (h >> 2));
        return (size_t)h;
    }
};
inline bool tmap_bf16_2d_cached(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{base, rows, cols, ld, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) { memcpy(map, &it->second, sizeof(CUtensorMap)); return true; }
    if (!make_tmap_bf16_2d(map, base, rows, cols, ld, box_rows)) return false;
    if (cache.size() > 16384) cache.clear();
    cache.emplace(key, *map);
    return true;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (function, device): remember which devices have it for this kernel.
struct PerDeviceOnce {
    std::atomic<uint64_t> done{0};
    template <typename F>
    cudaError_t ensure(F&& set) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const uint64_t bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
        e = set();
        if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
        return e;
    }
};

// number of (row) partial sums a residual-writing GEMM of N columns emits with this tile shape (GemmParams::ssq_out_parts)
template <int BLOCK_N, int CG>
inline int gemm_ssq_parts(int N) { return ((N + BLOCK_N - 1) / BLOCK_N) * GemmConfig<BLOCK_N, CG>::LSE_PARTS; }

struct GemmLaunch {
    const __nv_bfloat16* A; int lda;   // [a_rows, a_cols] visible to TMA (defaults: M x K)
    const __nv_bfloat16* W; int ldw;   // [w_rows, w_cols] visible to TMA (defaults: N x K; 2*d_ff rows for GATED)
    int w_rows;
    long long a_rows = 0, a_cols = 0, w_cols = 0;   // 0 = default
    GemmParams p;
};

template <int BLOCK_N, int CG, int EPI>
inline cudaError_t launch_gemm_t(const GemmLaunch& g, int num_sms, cudaStream_t stream) {
    using Cfg = GemmConfig<BLOCK_N, CG>;
    auto kernel = gemm_bf16_sm100_kernel<BLOCK_N, CG, EPI>;
    static PerDeviceOnce attr;
    {
        cudaError_t e = attr.ensure([&] { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); });
        if (e != cudaSuccess) return e;
    }
    CUtensorMap ta, tb;
    const uint64_t a_rows = g.a_rows ? (uint64_t)g.a_rows : (uint64_t)g.p.M;
    const uint64_t a_cols = g.a_cols ? (uint64_t)g.a_cols : (uint64_t)g.p.K;
    const uint64_t w_cols = g.w_cols ? (uint64_t)g.w_cols : (uint64_t)g.p.K;
    if (!tmap_bf16_2d_cached(&ta, g.A, a_rows, a_cols, (uint64_t)g.lda, Cfg::BLOCK_M)) return cudaErrorInvalidValue;
    if (!tmap_bf16_2d_cached(&tb, g.W, (uint64_t)g.w_rows, w_cols, (uint64_t)g.ldw, Cfg::B_HALF_ROWS))
        return cudaErrorInvalidValue;
    GemmParams p = g.p;
    if (p.num_batches < 1) p.num_batches = 1;
    const int rows_per_tile = Cfg::BLOCK_M * CG;
    p.num_m_tiles = (p.M + rows_per_tile - 1) / rows_per_tile;
    p.num_n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
    // Tile order = L2 blocking (see tile_coords). Either the W chunk or the A row group is the L2-resident operand:
    //   * W-stationary: W cut into chunks of <= ~36 MB, small A groups stream past each chunk;
    //   * A-stationary: one chunk (all of W streams), A groups of ~32 MB.
    // The choice per shape follows the sweep in profiles/r02_gemm_raster.md (DRAM bytes from ncu, isolated CUDA-event times).
    // vqa_set_gemm_schedule() / VQA_GEMM_GROUP_ROWS / VQA_GEMM_CHUNK_ROWS override (tuning); VQA_GEMM_SCHEDULE=r1 selects round 1's order.
    static const int env_rows = [] { const char* v = getenv("VQA_GEMM_GROUP_ROWS"); return (v && v[0]) ? atoi(v) : 0; }();
    static const int env_chunk = [] { const char* v = getenv("VQA_GEMM_CHUNK_ROWS"); return (v && v[0]) ? atoi(v) : 0; }();
    const int ov_rows = g_gemm_group_rows_override > 0 ? g_gemm_group_rows_override : env_rows;
    const int ov_chunk = g_gemm_chunk_rows_override != 0 ? g_gemm_chunk_rows_override : env_chunk;
    const long long w_bytes = (long long)g.w_rows * p.K * 2;
    const long long resident = 36ll << 20;
    static const int policy = [] { const char* v = getenv("VQA_GEMM_SCHEDULE"); return (v && v[0] == 'r' && v[1] == '1') ? 1 : 2; }();
    int group_rows, chunk_rows;   // chunk_rows < 0: no chunking (all N tiles in one chunk)
    if (policy == 1) {
        // round-1 order (A/B reference): no W chunks; A groups of ~32 MB when W is larger than ~L2, else 1024 rows
        chunk_rows = -1;
        if (w_bytes <= (96ll << 20)) group_rows = 1024;
        else group_rows = max(512, min(8192, (int)((32ll << 20) / ((long long)p.K * 2)) / 256 * 256));
    } else if (w_bytes <= resident) {
        group_rows = 1024; chunk_rows = -1;                       // W stays in L2 as a whole, A streams once
    } else {
        const long long n_chunks = (w_bytes + resident - 1) / resident;
        if (n_chunks <= 3) {
            // W-stationary: up to three L2-resident chunks, A streams once per chunk. Measured on B200 (profiles/r02_gemm_raster.md):
            // encoder qkv 7.7 -> 1.8 GB of DRAM reads, wo 7.6 -> 6.1 GB and 2.74 -> 2.40 ms in isolation
            const long long rows = ((long long)g.w_rows + n_chunks - 1) / n_chunks;
            chunk_rows = (int)((rows + BLOCK_N - 1) / BLOCK_N * BLOCK_N);
            group_rows = 1024;
        } else {
            // W much larger than L2 (FFN wi, 168 MB): A-stationary groups of ~32 MB are faster than many small W chunks (5.22 vs 5.38 ms)
            chunk_rows = -1;
            group_rows = max(512, min(8192, (int)((32ll << 20) / ((long long)p.K * 2)) / 256 * 256));
        }
    }
    if (ov_rows > 0) group_rows = ov_rows;
    if (ov_chunk != 0) chunk_rows = ov_chunk;
    p.group_m = max(1, group_rows / rows_per_tile);
    // the gated epilogue's W tile holds BLOCK_N/2 gate rows + BLOCK_N/2 up rows: a chunk of `chunk_rows` W rows = chunk_rows / BLOCK_N tiles either way
    p.chunk_n = chunk_rows > 0 ? max(1, chunk_rows / BLOCK_N) : p.num_n_tiles;
    if (p.chunk_n > p.num_n_tiles) p.chunk_n = p.num_n_tiles;
    // L2 eviction hints on the operand streams. Measured on B200 (profiles/r01_summary.md section 5): marking the re-read operand
    // evict_last and the streamed one evict_first RAISED the step's GEMM DRAM reads from 445 GB to 694 GB and cost 5 % of throughput
    // (an evict_first tile is dropped before the sibling CTAs that share it have fetched it), so the default is evict_normal on both.
    // VQA_GEMM_L2_POLICY=<a><w> with digits 0 normal / 1 first / 2 last overrides (tuning).
    static const int env_pol = [] { const char* v = getenv("VQA_GEMM_L2_POLICY"); return (v && v[0] && v[1]) ? (v[0] - '0') * 10 + (v[1] - '0') : -1; }();
    const unsigned long long pol[3] = {L2_EVICT_NORMAL, L2_EVICT_FIRST, L2_EVICT_LAST};
    p.a_policy = p.w_policy = L2_EVICT_NORMAL;
    if (env_pol >= 0 && env_pol / 10 < 3 && env_pol % 10 < 3) {
        p.a_policy = pol[env_pol / 10]; p.w_policy = pol[env_pol % 10];
    }
    const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.num_batches;
    int workers = num_sms / CG;
    if (workers > num_tiles) workers = num_tiles;
    if (workers < 1) workers = 1;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * CG);
    cfg.blockDim = dim3(Cfg::NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = CG;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, ta, tb, p);
}

}  // namespace vqa
