// tcgen05 flash attention for head_dim 64 (T5 encoder self-attention with relative-position bias + key padding mask,
// CLIP ViT self-attention with 1/sqrt(d) scale). One CTA = one (sample, head), looping over its 128-query tiles (TMEM allocation,
// barrier set-up and the bias table are paid once, the TMA producer runs ahead across query tiles); the score tile never leaves the SM:
//
//   warp 0   TMA producer : Q tile once, then K/V tiles (128 keys x 64) through a 2-stage smem ring
//   warp 1   MMA issuer   : S = Q K^T   (UMMA 128x128x16, SS: both operands in 128B-swizzled smem, D in TMEM)
//                           O += P V    (UMMA 128x64x16,  TS: P read from TMEM as bf16, V from smem MN-major)
//   warps 2.. softmax     : tcgen05.ld the scores in 32-key chunks, + bias (log2 domain), speculative row reference with lazy rescale
//                           of O in TMEM, exp2, P -> bf16 -> tcgen05.st; final O / l -> HBM.
//                           attn_tc_d64_split_kernel (shipped): 8 warps, two threads per query row (64 keys of a tile each);
//                           attn_tc_d64_stream_kernel (A/B, VQA_ATTN_VARIANT=30/31): 4 warps, one thread per row
//
// Two CTAs are resident per SM (<= 110 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
// Measurements and the history of the stage (round-1 kernel, two-pass stage with clock64 phase counters): profiles/r02_attention.md.
// Replaces the eager attention of transformers/models/t5/modeling_t5.py:308-334 (which materialises [B,H,S,S] scores in
// HBM) and transformers/models/clip/modeling_clip.py:261-336.
#pragma once
#include "ptx.cuh"
#include "gemm_sm100.cuh"   // make_tmap_bf16_2d

namespace vqa {

// launch with the programmatic-stream-serialisation attribute (the kernel calls pdl_wait() before its first global access)
template <typename Kernel, typename... Args>
inline cudaError_t launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}


constexpr int AT_BQ = 128, AT_BK = 128, AT_D = 64;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: Q, K or V tile
constexpr int AT_TMEM_COLS = 256;            // S: [0,128)  O: [128,192)  P (bf16 pairs): [192,256)
constexpr int AT_BIAS_PAD = 128;
constexpr int ATTN_STAGE_DEFAULT = 2;   // production softmax stage: 1 streaming (4 softmax warps), 2 split-row streaming (8 softmax warps)
constexpr int ATTN128_POLY_DEFAULT = 2;   // d128 (1 CTA / SM): score pairs of every 8 whose exp2 runs on the FMA pipe; 2 measured 1-3 % faster (profiles/r02_attention.md)


__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem, bf16 pairs] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MN-major operand tile: rows = K index (keys), 64 contiguous bf16 (128 B) of the MN index (head dim) per row, 128B
// swizzle as written by TMA. SBO = 8 rows * 128 B between 8-row groups along K; LBO = distance between 64-wide MN atoms.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------------------------------------------
// Softmax stage (round 2; round 1's kernel kept all 128 scores of a row in registers with scalar fp32 math, profiles/r01_ncu_kernels.md).
// The work per score bounds this kernel (head_dim 64: 512 tensor-core cycles per 128x128 tile against >= 1024 MUFU cycles for its
// 16384 exponentials and 64 KB of S through the TMEM read port):
//   * the row maximum is taken over the RAW scores (FMNMX3, two scores per instruction); the reference point of the exponentials is
//     the upper bound  scale * max_j s_ij + max(bias window)  -- softmax is invariant to it, P just carries a common factor <= 1;
//   * scale, bias and reference are applied with packed fp32x2 FFMA2 / FADD2 (two scores per instruction);
//   * the T5 bias of a "near" tile (|key tile - query tile| <= near_tiles) comes from a sliding-window table in shared memory,
//     Q[x] = (b[x], b[x+1], b[x+2], b[x+3]), so a row reads its 128 biases with 32 conflict-free LDS.128 (lane l's window is lane 0's
//     shifted by -l entries = -16 B); tiles further out see one constant per side (T5 buckets saturate at max_distance) folded into
//     the FFMA2 addend, as do bias-free heads (CLIP);
//   * POLY of every 8 score pairs can take their exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, relative error
//     7.5e-5, well under the bf16 rounding of P) instead of the MUFU (used by the d128 kernel only: the d64 stage is not MUFU-bound).
struct AttnTc2Params {
    __nv_bfloat16* o; int ldo;
    const int* seq_lens; const float* bias_table;   // [H, 2S-1] fp32 (natural-log domain) or nullptr
    int S, H, q_col0, k_col0, v_col0;
    float scale_log2e;
    int q_per_cta;      // split-row kernel: query tiles per CTA (grid.x CTAs per (sample, head)); the streaming kernel always takes all of them
    int near_tiles;     // key tiles with |kt - qt| <= near_tiles read the bias table; beyond, the table's end values (constant there)
    float scale;        // ROUND kernels: softmax scale applied after the bf16 rounding of the scores (1 for T5, 1/8 for CLIP)
};

inline size_t attn_tc2_smem_bytes(int near_tiles, bool has_bias, bool round_scores) {
    const size_t nq = has_bias ? (size_t)(2 * (128 * near_tiles + 127) + 1) : 0;
    return 1024 + 6 * AT_TILE_BYTES + ((nq * (round_scores ? 8 : 16) + 15) & ~size_t(15)) + 32 /*reduction scratch*/ + 14 * 8 + 16;
}

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// exp2 of two values on the FMA pipe: x = n + f, n = rint(x) via the 1.5*2^23 trick, 2^f by a degree-3 minimax polynomial on
// [-0.5, 0.5] (max relative error 7.5e-5), 2^n by adding n to the exponent field. x is clamped at -125 (result ~ 2^-125 ~ 0).
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& e0, float& e1) {
    float x0, x1;
    unpack2(x, x0, x1);
    x0 = fmaxf(x0, -125.f); x1 = fmaxf(x1, -125.f);
    const uint64_t xc = pack2(x0, x1);
    const uint64_t y = fadd2(xc, pack2(12582912.f, 12582912.f));
    const uint64_t yf = fadd2(y, pack2(-12582912.f, -12582912.f));
    const uint64_t f = ffma2(yf, pack2(-1.f, -1.f), xc);
    uint64_t pl = ffma2(f, pack2(0.0551716648f, 0.0551716648f), pack2(0.2426111251f, 0.2426111251f));
    pl = ffma2(pl, f, pack2(0.6932609677f, 0.6932609677f));
    pl = ffma2(pl, f, pack2(0.9999280572f, 0.9999280572f));
    float y0, y1, p0, p1;
    unpack2(y, y0, y1);
    unpack2(pl, p0, p1);
    e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(y0) << 23));
    e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(y1) << 23));
}

__device__ __forceinline__ uint32_t hadd2_bf16(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// One 32-key chunk of the exponential pass: sv = raw scores (fp32 bits) of this thread's row, out = 16 packed bf16 pairs of P.
//   !ROUND:  NEAR   t = s * c + bias_log2[j] - m           (fp32 log2-domain bias from the sliding-window table, LDS.128 per 4 keys)
//            !NEAR  t = s * c + addc                       (addc = constant bias (log2 domain) - m)
//   ROUND :  the reference's eager attention produces bf16 tensors before its fp32 softmax (modeling_t5.py:308-331: `scores = matmul(q, k^T)`
//            is a bf16 tensor, `scores += position_bias` a bf16 add; modeling_clip.py: bf16 matmul * 2^-3). With |score| ~ 10 that rounding
//            moves the exponent's argument by up to 0.04, far more than any other rounding on the path, so parity needs the same values:
//            x = bf16(s);  NEAR: x = bf16(x + bias[j]) (HADD2.BF16, bias in natural units from a bf16 sliding-window table, LDS.64 per 4 keys);
//            !NEAR with bias: x = bf16(x + bconst);  t = x * c - m with c = scale * log2(e).
template <bool NEAR, int POLY, bool ROUND>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&sv)[32], uint32_t (&pk)[16], uint32_t bq /*shared-space byte address*/,
                                              uint64_t cc, uint64_t addc, uint32_t bconst2 /*ROUND: far bias as bf16x2*/, bool far_bias,
                                              uint64_t& acc0, uint64_t& acc1) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint64_t t0, t1;
        if constexpr (ROUND) {
            uint32_t x01 = pack_bf16x2(__uint_as_float(sv[4 * q]), __uint_as_float(sv[4 * q + 1]));
            uint32_t x23 = pack_bf16x2(__uint_as_float(sv[4 * q + 2]), __uint_as_float(sv[4 * q + 3]));
            if (NEAR) {
                uint32_t b01, b23;
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b01), "=r"(b23) : "r"(bq + q * 32));
                x01 = hadd2_bf16(x01, b01);
                x23 = hadd2_bf16(x23, b23);
            } else if (far_bias) {
                x01 = hadd2_bf16(x01, bconst2);
                x23 = hadd2_bf16(x23, bconst2);
            }
            t0 = ffma2(pack2(__uint_as_float(x01 << 16), __uint_as_float(x01 & 0xffff0000u)), cc, addc);
            t1 = ffma2(pack2(__uint_as_float(x23 << 16), __uint_as_float(x23 & 0xffff0000u)), cc, addc);
        } else {
            t0 = pack2(__uint_as_float(sv[4 * q]), __uint_as_float(sv[4 * q + 1]));
            t1 = pack2(__uint_as_float(sv[4 * q + 2]), __uint_as_float(sv[4 * q + 3]));
            if (NEAR) {
                uint64_t b01, b23;   // four consecutive biases of this row: one conflict-free LDS.128 (see the kernel's header comment)
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b01), "=l"(b23) : "r"(bq + q * 64));
                t0 = fadd2(ffma2(t0, cc, b01), addc);
                t1 = fadd2(ffma2(t1, cc, b23), addc);
            } else {
                t0 = ffma2(t0, cc, addc);
                t1 = ffma2(t1, cc, addc);
            }
        }
        float e0, e1, e2, e3;
        // pairs 2q and 2q+1 of this chunk's 16: POLY of every 8 pairs go to the FMA pipe, spread so MUFU and FMA work interleave
        if (((2 * q) & 7) < POLY) exp2_poly2(t0, e0, e1);
        else { float a, b_; unpack2(t0, a, b_); e0 = fast_exp2(a); e1 = fast_exp2(b_); }
        if (((2 * q + 1) & 7) < POLY) exp2_poly2(t1, e2, e3);
        else { float a, b_; unpack2(t1, a, b_); e2 = fast_exp2(a); e3 = fast_exp2(b_); }
        acc0 = fadd2(acc0, pack2(e0, e1));
        acc1 = fadd2(acc1, pack2(e2, e3));
        pk[2 * q] = pack_bf16x2(e0, e1);
        pk[2 * q + 1] = pack_bf16x2(e2, e3);
    }
}

// Streaming variant of the softmax stage: the same CTA skeleton, barriers and MMA schedule, but a softmax thread never holds more than
// two 32-key chunks of its row. The 128 scores of a tile are exponentiated against a reference fixed BEFORE the tile starts (the running
// bound of the previous tiles; for the first key tile of a query tile a max-only pre-pass over S), so each chunk goes TMEM -> registers ->
// exp2 -> P in one pass while the next chunk's tcgen05.ld is in flight; the tile's own maximum is collected on the way and, if it
// exceeds the reference by more than 2^8, O and l are rescaled before the NEXT tile (P may exceed 1 inside a tile -- harmless in
// bf16 / fp32, the normaliser l carries the same factor). Why: profiles/r02_attention.md -- the 128-register score array of the
// two-pass stage costs spills at the 168-register cap of two CTAs per SM and unrolls to 67 KB of SASS (instruction-fetch stalls 11 %,
// branch-resolve stalls 16 % of the samples); this loop is ~10 KB and needs ~110 registers.
template <bool HAS_BIAS, bool ROUND>
__global__ void __launch_bounds__(192, 2)
attn_tc_d64_stream_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTc2Params p) {
    pdl_launch_dependents();
    pdl_wait();      // seq_lens / the bias table / qkv are outputs of earlier kernels
    constexpr int POLY = 0;
    const int h = blockIdx.y, b = blockIdx.z;
    const int len = p.seq_lens ? min(p.seq_lens[b], p.S) : p.S;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const size_t row_base = (size_t)b * p.S;
    const int nq = (len + AT_BQ - 1) / AT_BQ;     // query tiles that contain at least one valid row
    const int nkt = (len + AT_BK - 1) / AT_BK;    // key tiles that contain at least one valid key

    // rows past the last valid query tile: deterministic zeros
    for (int i = threadIdx.x; i < (p.S - nq * AT_BQ) * 8; i += blockDim.x) {
        const int r = nq * AT_BQ + (i >> 3), c = i & 7;
        *reinterpret_cast<uint4*>(p.o + (row_base + r) * p.ldo + h * AT_D + c * 8) = make_uint4(0, 0, 0, 0);
    }
    if (nq == 0) return;

    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                          // [2]
    uint8_t* sK = smem + 2 * AT_TILE_BYTES;      // [2]
    uint8_t* sV = smem + 4 * AT_TILE_BYTES;      // [2]
    // sliding-window bias table, entry x <-> rel = x - W: !ROUND: float4 log2e * (b[rel], .., b[rel+3]); ROUND: the same four as bf16 (8 bytes)
    uint8_t* sBiasQ = smem + 6 * AT_TILE_BYTES;
    constexpr uint32_t BQ_ENTRY = ROUND ? 8u : 16u;
    const int Wn = 128 * p.near_tiles + 127;
    const int nQ = HAS_BIAS ? 2 * Wn + 1 : 0;
    float* sRed = reinterpret_cast<float*>(smem + 6 * AT_TILE_BYTES + (((size_t)nQ * BQ_ENTRY + 15) & ~size_t(15)));   // [8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 8);
    uint64_t* q_full = bars;          // [2]
    uint64_t* q_empty = bars + 2;     // [2]
    uint64_t* kv_full = bars + 4;     // [2]
    uint64_t* kv_empty = bars + 6;    // [2]
    uint64_t* s_full = bars + 8;
    uint64_t* s_empty = bars + 9;
    uint64_t* p_full = bars + 10;
    uint64_t* o_done = bars + 11;
    uint64_t* o_free = bars + 12;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 4);
        mbar_init(p_full, 4);
        mbar_init(o_done, 1);
        mbar_init(o_free, 4);
        fence_barrier_init();
        // the first loads only need the barriers: get them in flight before the rest of the CTA finishes its set-up
        mbar_arrive_expect_tx(&q_full[0], AT_TILE_BYTES);
        tma_load_2d(sQ, &tmap_qkv, &q_full[0], p.q_col0 + h * AT_D, (int)row_base);
        mbar_arrive_expect_tx(&kv_full[0], 2 * AT_TILE_BYTES);
        tma_load_2d(sK, &tmap_qkv, &kv_full[0], p.k_col0 + h * AT_D, (int)row_base);
        tma_load_2d(sV, &tmap_qkv, &kv_full[0], p.v_col0 + h * AT_D, (int)row_base);
    }
    if (warp == 1) {
        tmem_alloc<1>(tmem_ptr_smem, AT_TMEM_COLS);
        tmem_relinquish<1>();
    }
    const float LOG2E = 1.4426950408889634f;
    const float BSC = ROUND ? 1.0f : LOG2E;   // domain the table / constants are kept in
    float b_left = 0.f, b_right = 0.f;      // constant bias of far tiles to the left / right of the diagonal
    if (HAS_BIAS) {
        const int width = 2 * p.S - 1;
        const float* src = p.bias_table + (size_t)h * width;
        b_left = __ldg(src) * BSC;
        b_right = __ldg(src + width - 1) * BSC;
        float lmax = -INFINITY;
        for (int x = threadIdx.x; x < nQ; x += blockDim.x) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = min(max(x + k - Wn + (p.S - 1), 0), width - 1);   // rel = x + k - W, clamped to the table
                v[k] = __ldg(src + idx) * BSC;
            }
            if constexpr (ROUND) reinterpret_cast<uint2*>(sBiasQ)[x] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            else                 reinterpret_cast<float4*>(sBiasQ)[x] = make_float4(v[0], v[1], v[2], v[3]);
            lmax = fmaxf(lmax, v[0]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if (lane == 0) sRed[warp] = lmax;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128, tmem_P = tmem_base + 192;
    const int total_tiles = nq * nkt;   // global tile index g = qi * nkt + j

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int qi = 0; qi < nq; ++qi) {
                const int qb = qi & 1;
                if (qi > 0) {   // (tile 0 was issued during set-up)
                    mbar_wait(&q_empty[qb], (((uint32_t)qi >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&q_full[qb], AT_TILE_BYTES);
                    tma_load_2d(sQ + qb * AT_TILE_BYTES, &tmap_qkv, &q_full[qb], p.q_col0 + h * AT_D, (int)(row_base + qi * AT_BQ));
                }
                for (int j = 0; j < nkt; ++j) {
                    const int g = qi * nkt + j;
                    if (g == 0) continue;
                    const int st = g & 1;
                    mbar_wait(&kv_empty[st], (((uint32_t)g >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE_BYTES);
                    tma_load_2d(sK + st * AT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.k_col0 + h * AT_D, (int)(row_base + j * AT_BK));
                    tma_load_2d(sV + st * AT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.v_col0 + h * AT_D, (int)(row_base + j * AT_BK));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128);
            constexpr uint32_t idesc_o = make_idesc_bf16_f32(128, 64) | (1u << 16);   // B operand MN-major
            auto issue_pv = [&](int g) {
                const int j = g % nkt, qi = g / nkt;
                mbar_wait(p_full, (uint32_t)g & 1u);
                if (j == 0 && qi > 0) mbar_wait(o_free, (uint32_t)(qi - 1) & 1u);   // previous query tile's O has been read out
                tcgen05_fence_after();
                const uint64_t vdesc = make_mnmajor_sw128_desc(smem_u32(sV + (g & 1) * AT_TILE_BYTES), AT_TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < AT_BK / 16; ++ks)
                    umma_f16_ts(tmem_O, tmem_P + ks * 8, vdesc + (uint64_t)(ks * (16 * 128 / 16)), idesc_o,
                                (j > 0 || ks > 0) ? 1u : 0u);
                umma_commit<1>(&kv_empty[g & 1]);
                umma_commit<1>(o_done);
            };
            for (int qi = 0; qi < nq; ++qi) {
                const int qb = qi & 1;
                mbar_wait(&q_full[qb], ((uint32_t)qi >> 1) & 1u);
                const uint64_t qdesc = make_kmajor_sw128_desc(smem_u32(sQ + qb * AT_TILE_BYTES));
                for (int j = 0; j < nkt; ++j) {
                    const int g = qi * nkt + j;
                    const int st = g & 1;
                    mbar_wait(&kv_full[st], ((uint32_t)g >> 1) & 1u);
                    mbar_wait(s_empty, ((uint32_t)g & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint64_t kdesc = make_kmajor_sw128_desc(smem_u32(sK + st * AT_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k)
                        umma_f16<1>(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                    umma_commit<1>(s_full);
                    if (j == nkt - 1) umma_commit<1>(&q_empty[qb]);   // Q buffer reusable once this tile's QK^T retired
                    if (g > 0) issue_pv(g - 1);
                }
            }
            issue_pv(total_tiles - 1);
        }
    } else {
        // ===================== softmax / correction / epilogue: one thread per query row, one 32-key chunk at a time =====================
        const uint32_t quad = warp & 3u;
        const int row = quad * 32 + lane;
        const uint32_t lane_off = (quad * 32u) << 16;
        float bmax_near = 0.f;
        if (HAS_BIAS) {
            bmax_near = sRed[0];
#pragma unroll
            for (int i = 1; i < 6; ++i) bmax_near = fmaxf(bmax_near, sRed[i]);
        }
        const uint64_t cc = pack2(p.scale_log2e, p.scale_log2e);
        const uint32_t bl2 = pack_bf16x2(b_left, b_left), br2 = pack_bf16x2(b_right, b_right);   // ROUND: far-tile bias as bf16 pairs
        auto chunk_max = [](const uint32_t (&v)[32], float& m0, float& m1, float& m2, float& m3) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            }
        };
        int g = 0;
        for (int qi = 0; qi < nq; ++qi) {
            const int q0 = qi * AT_BQ;
            const int qrow = q0 + row;
            __nv_bfloat16* orow = p.o + (row_base + qrow) * p.ldo + h * AT_D;
            if (q0 + (int)quad * 32 >= len) {
                // every row of this warp is padding in this query tile: keep the barrier protocol in lock-step, no math
                for (int j = 0; j < nkt; ++j, ++g) {
                    mbar_wait(s_full, (uint32_t)g & 1u);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_empty);
                    if (g > 0) mbar_wait(p_full, (uint32_t)(g - 1) & 1u);   // previous phase must be closed before arriving again
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_full);
                }
                mbar_wait(o_done, (uint32_t)(g - 1) & 1u);
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                if (qrow < p.S) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            float m_run = 0.f, l_run = 0.f, corr_pend = 1.f;
            bool pend = false;                 // O and l still have to be multiplied by corr_pend (decided at the end of the previous tile)
            for (int j = 0; j < nkt; ++j, ++g) {
                const int k0 = j * AT_BK;
                const int nch = min(4, (len - k0 + 31) >> 5);   // 32-key chunks that contain at least one valid key
                const int dt = j - qi;
                const bool near = HAS_BIAS && (dt <= p.near_tiles) && (dt >= -p.near_tiles);
                const float bias_ub = near ? bmax_near : (dt < 0 ? b_left : b_right);   // 0 without bias
                const uint32_t s_addr = tmem_S + lane_off;
                mbar_wait(s_full, (uint32_t)g & 1u);
                tcgen05_fence_after();
                if (j == 0) {
                    // first key tile of this query tile: no reference yet -> max-only pre-pass over S (S stays in TMEM for the real pass)
                    float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < nch; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(s_addr + c * 32, v);
                        tmem_ld_wait();
                        if (k0 + c * 32 + 32 > len) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (k0 + c * 32 + i >= len) v[i] = 0xff800000u;
                        }
                        chunk_max(v, a0, a1, a2, a3);
                    }
                    const float raw = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
                    m_run = ROUND ? (raw + bias_ub) * p.scale_log2e : fmaf(raw, p.scale_log2e, bias_ub);
                } else {
                    mbar_wait(o_done, (uint32_t)(g - 1) & 1u);   // P_{g-1} consumed, O complete up to tile g-1
                    tcgen05_fence_after();
                    if (pend) {
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            uint32_t ov[16];
                            tmem_ld_32x32b_x16(tmem_O + lane_off + c * 16, ov);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr_pend);
                            tmem_st_32x32b_x16(tmem_O + lane_off + c * 16, ov);
                        }
                        l_run *= corr_pend;
                        pend = false;
                    }
                }
                // ---- the tile: chunk c is exponentiated while chunk c + 1 is on its way from TMEM (ping-pong register buffers)
                const float a_add = near ? -m_run : (ROUND ? -m_run : bias_ub - m_run);
                const uint64_t addc = pack2(a_add, a_add);
                const uint32_t bq = near ? smem_u32(sBiasQ) + (uint32_t)(dt * 128 - row + Wn) * BQ_ENTRY : 0u;
                const uint32_t bfar = dt < 0 ? bl2 : br2;
                uint64_t acc0 = 0ull, acc1 = 0ull;
                float t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY, t3 = -INFINITY;
                uint32_t bufA[32], bufB[32];
                auto process = [&](uint32_t (&v)[32], int c) {
                    if (c == nch - 1) {      // every chunk of S has left TMEM: the next QK^T may overwrite it
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_empty);
                    }
                    if (k0 + c * 32 + 32 > len) {   // the one chunk that straddles the sample's length
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (k0 + c * 32 + i >= len) v[i] = 0xff800000u;
                    }
                    chunk_max(v, t0, t1, t2, t3);
                    uint32_t pk[16];
                    if (near) softmax_chunk<true, POLY, ROUND>(v, pk, bq + c * 32 * BQ_ENTRY, cc, addc, 0u, false, acc0, acc1);
                    else      softmax_chunk<false, POLY, ROUND>(v, pk, 0u, cc, addc, bfar, HAS_BIAS, acc0, acc1);
                    tmem_st_32x32b_x16(tmem_P + lane_off + c * 16, pk);
                };
                tmem_ld_32x32b_x32(s_addr, bufA);
#pragma unroll 1
                for (int c = 0; c < nch; c += 2) {
                    tmem_ld_wait();                                              // bufA = chunk c
                    if (c + 1 < nch) tmem_ld_32x32b_x32(s_addr + (c + 1) * 32, bufB);
                    process(bufA, c);
                    if (c + 1 < nch) {
                        tmem_ld_wait();                                          // bufB = chunk c + 1
                        if (c + 2 < nch) tmem_ld_32x32b_x32(s_addr + (c + 2) * 32, bufA);
                        process(bufB, c + 1);
                    }
                }
                if (nch < 4) {
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll 1
                    for (int c = nch; c < 4; ++c) tmem_st_32x32b_x16(tmem_P + lane_off + c * 16, z);
                }
                {
                    float a0, a1, a2, a3;
                    unpack2(acc0, a0, a1);
                    unpack2(acc1, a2, a3);
                    l_run += (a0 + a1) + (a2 + a3);
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                // ---- did this tile outgrow the reference? then the NEXT tile starts by rescaling O and l (after P.V of this one retired)
                const float raw = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                const float bound = ROUND ? (raw + bias_ub) * p.scale_log2e : fmaf(raw, p.scale_log2e, bias_ub);
                if (__any_sync(0xffffffffu, bound > m_run + 8.f)) {
                    const float m_new = fmaxf(m_run, bound);
                    corr_pend = fast_exp2(m_run - m_new);
                    m_run = m_new;
                    pend = true;
                }
            }
            // ---- epilogue of this query tile: O / l (a pending rescale would multiply both by the same factor: skipped)
            mbar_wait(o_done, (uint32_t)(g - 1) & 1u);
            tcgen05_fence_after();
            const float inv = (qrow < len) ? 1.f / l_run : 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, ov);
                tmem_ld_wait();
                if (c == 1) {   // O fully copied out: the next query tile's first P.V may overwrite it now
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(o_free);
                }
                if (qrow < p.S) {
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w[e] = pack_bf16x2(__uint_as_float(ov[gq * 8 + 2 * e]) * inv, __uint_as_float(ov[gq * 8 + 2 * e + 1]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + gq * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        tcgen05_fence_before();
    }

    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<1>(tmem_base, AT_TMEM_COLS);
}

// Split-row variant of the streaming stage: EIGHT softmax warps per CTA. Warps w and w + 4 own the same 32 query rows (same TMEM lane
// quadrant) and half of each tile's key columns (chunks {0,1} / {2,3}), so every SM sub-partition hosts four softmax warps (two per CTA, two
// CTAs per SM) instead of two -- the stage is latency-bound (profiles/r02_attention.md: 0.19 IPC per softmax warp, MUFU pipe 38 % busy), and
// the second pair of warps fills the issue slots the first pair leaves empty. The two halves of a row agree on the exponent reference through
// a split-phase exchange in shared memory (value slots + mbarriers, see post()/collect() below): once per query tile (pre-pass maximum), once per
// key tile (the tile's bound, published at its end and read at the start of the next tile, where the lazy rescale is decided) and once in the
// epilogue (row sums). Each half rescales / normalises / stores 32 of O's 64 columns. The wait for the previous tile's P.V (the P slot in TMEM is
// single-buffered) sits between the first chunk's exponentials and its store, not at the top of the tile.
template <bool HAS_BIAS, bool ROUND>
__global__ void __launch_bounds__(320, 2)
attn_tc_d64_split_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTc2Params p) {
    pdl_launch_dependents();
    pdl_wait();      // seq_lens / the bias table / qkv are outputs of earlier kernels
    constexpr int POLY = 0;
    const int h = blockIdx.y, b = blockIdx.z;
    const int len = p.seq_lens ? min(p.seq_lens[b], p.S) : p.S;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const size_t row_base = (size_t)b * p.S;
    const int nq_all = (len + AT_BQ - 1) / AT_BQ; // query tiles that contain at least one valid row
    const int nkt = (len + AT_BK - 1) / AT_BK;    // key tiles that contain at least one valid key
    // blockIdx.x > 0 only for small batches (launch_attn_tc: B * H CTAs would leave most SMs idle): this CTA owns q_per_cta consecutive query tiles
    const int q_first = blockIdx.x * p.q_per_cta;
    const int nq = max(0, min(p.q_per_cta, nq_all - q_first));

    // rows past the last valid query tile: deterministic zeros
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < (p.S - nq_all * AT_BQ) * 8; i += blockDim.x) {
            const int r = nq_all * AT_BQ + (i >> 3), c = i & 7;
            *reinterpret_cast<uint4*>(p.o + (row_base + r) * p.ldo + h * AT_D + c * 8) = make_uint4(0, 0, 0, 0);
        }
    }
    if (nq == 0) return;

    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                          // [2]
    uint8_t* sK = smem + 2 * AT_TILE_BYTES;      // [2]
    uint8_t* sV = smem + 4 * AT_TILE_BYTES;      // [2]
    // sliding-window bias table, entry x <-> rel = x - W: !ROUND: float4 log2e * (b[rel], .., b[rel+3]); ROUND: the same four as bf16 (8 bytes)
    uint8_t* sBiasQ = smem + 6 * AT_TILE_BYTES;
    constexpr uint32_t BQ_ENTRY = ROUND ? 8u : 16u;
    const int Wn = 128 * p.near_tiles + 127;
    const int nQ = HAS_BIAS ? 2 * Wn + 1 : 0;
    float* sRed = reinterpret_cast<float*>(smem + 6 * AT_TILE_BYTES + (((size_t)nQ * BQ_ENTRY + 15) & ~size_t(15)));   // [8]
    float* sXch = sRed + 16;                                     // [2 buffers][2 halves][128 rows]: row maxima / row sums exchanged between the halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(sXch + 512);
    uint64_t* q_full = bars;          // [2]
    uint64_t* q_empty = bars + 2;     // [2]
    uint64_t* kv_full = bars + 4;     // [2]
    uint64_t* kv_empty = bars + 6;    // [2]
    uint64_t* s_full = bars + 8;
    uint64_t* s_empty = bars + 9;
    uint64_t* p_full = bars + 10;
    uint64_t* o_done = bars + 11;
    uint64_t* o_free = bars + 12;
    uint64_t* xbar = bars + 13;       // [4 row quadrants][2 halves][2 phases]: "my value for exchange k is in shared memory"
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 29);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 8);
        mbar_init(p_full, 8);
        mbar_init(o_done, 1);
        mbar_init(o_free, 8);
#pragma unroll
        for (int i = 0; i < 16; ++i) mbar_init(&xbar[i], 1);
        fence_barrier_init();
        // the first loads only need the barriers: get them in flight before the rest of the CTA finishes its set-up
        mbar_arrive_expect_tx(&q_full[0], AT_TILE_BYTES);
        tma_load_2d(sQ, &tmap_qkv, &q_full[0], p.q_col0 + h * AT_D, (int)(row_base + q_first * AT_BQ));
        mbar_arrive_expect_tx(&kv_full[0], 2 * AT_TILE_BYTES);
        tma_load_2d(sK, &tmap_qkv, &kv_full[0], p.k_col0 + h * AT_D, (int)row_base);
        tma_load_2d(sV, &tmap_qkv, &kv_full[0], p.v_col0 + h * AT_D, (int)row_base);
    }
    if (warp == 1) {
        tmem_alloc<1>(tmem_ptr_smem, AT_TMEM_COLS);
        tmem_relinquish<1>();
    }
    const float LOG2E = 1.4426950408889634f;
    const float BSC = ROUND ? 1.0f : LOG2E;   // domain the table / constants are kept in
    float b_left = 0.f, b_right = 0.f;      // constant bias of far tiles to the left / right of the diagonal
    if (HAS_BIAS) {
        const int width = 2 * p.S - 1;
        const float* src = p.bias_table + (size_t)h * width;
        b_left = __ldg(src) * BSC;
        b_right = __ldg(src + width - 1) * BSC;
        float lmax = -INFINITY;
        for (int x = threadIdx.x; x < nQ; x += blockDim.x) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = min(max(x + k - Wn + (p.S - 1), 0), width - 1);   // rel = x + k - W, clamped to the table
                v[k] = __ldg(src + idx) * BSC;
            }
            if constexpr (ROUND) reinterpret_cast<uint2*>(sBiasQ)[x] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            else                 reinterpret_cast<float4*>(sBiasQ)[x] = make_float4(v[0], v[1], v[2], v[3]);
            lmax = fmaxf(lmax, v[0]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if (lane == 0) sRed[warp] = lmax;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128, tmem_P = tmem_base + 192;
    const int total_tiles = nq * nkt;   // global tile index g = qi * nkt + j

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int qi = 0; qi < nq; ++qi) {
                const int qb = qi & 1;
                if (qi > 0) {   // (tile 0 was issued during set-up)
                    mbar_wait(&q_empty[qb], (((uint32_t)qi >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&q_full[qb], AT_TILE_BYTES);
                    tma_load_2d(sQ + qb * AT_TILE_BYTES, &tmap_qkv, &q_full[qb], p.q_col0 + h * AT_D, (int)(row_base + (q_first + qi) * AT_BQ));
                }
                for (int j = 0; j < nkt; ++j) {
                    const int g = qi * nkt + j;
                    if (g == 0) continue;
                    const int st = g & 1;
                    mbar_wait(&kv_empty[st], (((uint32_t)g >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE_BYTES);
                    tma_load_2d(sK + st * AT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.k_col0 + h * AT_D, (int)(row_base + j * AT_BK));
                    tma_load_2d(sV + st * AT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.v_col0 + h * AT_D, (int)(row_base + j * AT_BK));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128);
            constexpr uint32_t idesc_o = make_idesc_bf16_f32(128, 64) | (1u << 16);   // B operand MN-major
            auto issue_pv = [&](int g) {
                const int j = g % nkt, qi = g / nkt;
                mbar_wait(p_full, (uint32_t)g & 1u);
                if (j == 0 && qi > 0) mbar_wait(o_free, (uint32_t)(qi - 1) & 1u);   // previous query tile's O has been read out
                tcgen05_fence_after();
                const uint64_t vdesc = make_mnmajor_sw128_desc(smem_u32(sV + (g & 1) * AT_TILE_BYTES), AT_TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < AT_BK / 16; ++ks)
                    umma_f16_ts(tmem_O, tmem_P + ks * 8, vdesc + (uint64_t)(ks * (16 * 128 / 16)), idesc_o,
                                (j > 0 || ks > 0) ? 1u : 0u);
                umma_commit<1>(&kv_empty[g & 1]);
                umma_commit<1>(o_done);
            };
            for (int qi = 0; qi < nq; ++qi) {
                const int qb = qi & 1;
                mbar_wait(&q_full[qb], ((uint32_t)qi >> 1) & 1u);
                const uint64_t qdesc = make_kmajor_sw128_desc(smem_u32(sQ + qb * AT_TILE_BYTES));
                for (int j = 0; j < nkt; ++j) {
                    const int g = qi * nkt + j;
                    const int st = g & 1;
                    mbar_wait(&kv_full[st], ((uint32_t)g >> 1) & 1u);
                    mbar_wait(s_empty, ((uint32_t)g & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint64_t kdesc = make_kmajor_sw128_desc(smem_u32(sK + st * AT_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k)
                        umma_f16<1>(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                    umma_commit<1>(s_full);
                    if (j == nkt - 1) umma_commit<1>(&q_empty[qb]);   // Q buffer reusable once this tile's QK^T retired
                    if (g > 0) issue_pv(g - 1);
                }
            }
            issue_pv(total_tiles - 1);
        }
    } else {
        // ===================== softmax / correction / epilogue: two threads per query row (64 keys of a tile each) =====================
        const uint32_t quad = warp & 3u;
        const uint32_t half = (warp - 2u) >> 2;           // 0: key chunks {0,1} of every tile and O columns [0,32); 1: chunks {2,3}, O columns [32,64)
        const int row = quad * 32 + lane;
        const uint32_t lane_off = (quad * 32u) << 16;
        // Split-phase exchange between the two halves of a row: post(v) publishes this half's value for exchange number xk, collect() waits
        // for the partner's value of the same exchange. Two value buffers and two mbarriers per (quadrant, half) alternate with xk, so a warp
        // that runs one exchange ahead of its partner (it cannot run two ahead: the S / P barriers need both) never reuses a live slot.
        uint32_t xk = 0;
        auto post = [&](float v) {
            sXch[(xk & 1u) * 256 + half * 128 + row] = v;
            __syncwarp();
            if (lane == 0) mbar_arrive(&xbar[(quad * 2 + half) * 2 + (xk & 1u)]);
        };
        auto collect = [&]() -> float {
            mbar_wait(&xbar[(quad * 2 + (half ^ 1u)) * 2 + (xk & 1u)], (xk >> 1) & 1u);
            const float other = sXch[(xk & 1u) * 256 + (half ^ 1u) * 128 + row];
            ++xk;
            return other;
        };
        float bmax_near = 0.f;
        if (HAS_BIAS) {
            bmax_near = sRed[0];
#pragma unroll
            for (int i = 1; i < 10; ++i) bmax_near = fmaxf(bmax_near, sRed[i]);
        }
        const uint64_t cc = pack2(p.scale_log2e, p.scale_log2e);
        const uint32_t bl2 = pack_bf16x2(b_left, b_left), br2 = pack_bf16x2(b_right, b_right);   // ROUND: far-tile bias as bf16 pairs
        auto chunk_max = [](const uint32_t (&v)[32], float& m0, float& m1, float& m2, float& m3) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            }
        };
        int g = 0;
        for (int qi = 0; qi < nq; ++qi) {
            const int q0 = (q_first + qi) * AT_BQ;
            const int qrow = q0 + row;
            __nv_bfloat16* orow = p.o + (row_base + qrow) * p.ldo + h * AT_D + half * 32;
            if (q0 + (int)quad * 32 >= len) {
                // every row of this warp is padding in this query tile: keep the barrier protocol in lock-step, no math
                for (int j = 0; j < nkt; ++j, ++g) {
                    mbar_wait(s_full, (uint32_t)g & 1u);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_empty);
                    if (g > 0) mbar_wait(p_full, (uint32_t)(g - 1) & 1u);   // previous phase must be closed before arriving again
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_full);
                }
                mbar_wait(o_done, (uint32_t)(g - 1) & 1u);
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                if (qrow < p.S) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            float m_run = 0.f, l_run = 0.f;       // l_run: this half's share of the row sum
            float bound_prev = 0.f;               // this half's upper bound of the previous tile's exponents (posted to the partner at its end)
            for (int j = 0; j < nkt; ++j, ++g) {
                const int k0 = j * AT_BK;
                const int nch = min(4, (len - k0 + 31) >> 5);   // 32-key chunks of the tile that contain at least one valid key
                const int c_lo = 2 * (int)half;
                const int n_mine = max(0, min(nch - c_lo, 2));  // this half's chunks: c_lo .. c_lo + n_mine - 1
                const int dt = j - (q_first + qi);
                const bool near = HAS_BIAS && (dt <= p.near_tiles) && (dt >= -p.near_tiles);
                const float bias_ub = near ? bmax_near : (dt < 0 ? b_left : b_right);   // 0 without bias
                const uint32_t s_addr = tmem_S + lane_off + c_lo * 32;
                mbar_wait(s_full, (uint32_t)g & 1u);
                tcgen05_fence_after();
                uint32_t buf[32];
                auto mask = [&](uint32_t (&v)[32], int c) {
                    if (k0 + c * 32 + 32 > len) {   // the one chunk that straddles the sample's length
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (k0 + c * 32 + i >= len) v[i] = 0xff800000u;
                    }
                };
                bool p_free = (j == 0);           // P of the previous tile consumed (o_done observed)? -- the first tile's P slot is free by construction
                if (j == 0) {
                    // first key tile of this query tile: no reference yet -> max-only pre-pass (S stays in TMEM for the real pass)
                    float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < n_mine; ++c) {
                        tmem_ld_32x32b_x32(s_addr + c * 32, buf);
                        tmem_ld_wait();
                        mask(buf, c_lo + c);
                        chunk_max(buf, a0, a1, a2, a3);
                    }
                    const float raw_half = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
                    const float bound_half = ROUND ? (raw_half + bias_ub) * p.scale_log2e : fmaf(raw_half, p.scale_log2e, bias_ub);
                    post(bound_half);
                    m_run = fmaxf(bound_half, collect());
                } else {
                    // did the previous tile outgrow the reference? (its row maximum = max of the two halves' bounds; both halves see the same
                    // value and take the same decision) then O and l are rescaled before this tile's P is added
                    const float bound = fmaxf(bound_prev, collect());
                    if (__any_sync(0xffffffffu, bound > m_run + 8.f)) {
                        const float m_new = fmaxf(m_run, bound);
                        const float corr = fast_exp2(m_run - m_new);
                        m_run = m_new;
                        l_run *= corr;
                        mbar_wait(o_done, (uint32_t)(g - 1) & 1u);   // P_{g-1} consumed, O complete up to tile g-1
                        tcgen05_fence_after();
                        p_free = true;
#pragma unroll 1
                        for (int c = 0; c < 2; ++c) {            // this half's 32 columns of O
                            uint32_t ov[16];
                            tmem_ld_32x32b_x16(tmem_O + lane_off + half * 32 + c * 16, ov);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
                            tmem_st_32x32b_x16(tmem_O + lane_off + half * 32 + c * 16, ov);
                        }
                    }
                }
                // ---- this half's (up to) two chunks: exponentials against the running reference, tile maximum tracked on the side
                const float a_add = near ? -m_run : (ROUND ? -m_run : bias_ub - m_run);
                const uint64_t addc = pack2(a_add, a_add);
                const uint32_t bq = near ? smem_u32(sBiasQ) + (uint32_t)(dt * 128 - row + Wn) * BQ_ENTRY : 0u;
                const uint32_t bfar = dt < 0 ? bl2 : br2;
                uint64_t acc0 = 0ull, acc1 = 0ull;
                float t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY, t3 = -INFINITY;
                if (n_mine == 0) {           // nothing of S to read for this half: release it right away
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_empty);
                }
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t pk[16];
                    if (c < n_mine) {
                        tmem_ld_32x32b_x32(s_addr + c * 32, buf);
                        tmem_ld_wait();
                        if (c == n_mine - 1) {   // this half's part of S has left TMEM (the next QK^T starts when all eight warps say so)
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(s_empty);
                        }
                        mask(buf, c_lo + c);
                        chunk_max(buf, t0, t1, t2, t3);
                        if (near) softmax_chunk<true, POLY, ROUND>(buf, pk, bq + (c_lo + c) * 32 * BQ_ENTRY, cc, addc, 0u, false, acc0, acc1);
                        else      softmax_chunk<false, POLY, ROUND>(buf, pk, 0u, cc, addc, bfar, HAS_BIAS, acc0, acc1);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    }
                    if (!p_free) {
                        // the P slot is single-buffered: P.V of the previous tile must have read it. Waiting HERE (after the first chunk's
                        // exponentials) instead of at the top of the tile hides the P.V latency behind them.
                        mbar_wait(o_done, (uint32_t)(g - 1) & 1u);
                        tcgen05_fence_after();
                        p_free = true;
                    }
                    tmem_st_32x32b_x16(tmem_P + lane_off + (c_lo + c) * 16, pk);
                }
                {
                    float a0, a1, a2, a3;
                    unpack2(acc0, a0, a1);
                    unpack2(acc1, a2, a3);
                    l_run += (a0 + a1) + (a2 + a3);
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                if (j + 1 < nkt) {           // publish this half's bound for the next tile's rescale decision
                    const float raw_half = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                    bound_prev = ROUND ? (raw_half + bias_ub) * p.scale_log2e : fmaf(raw_half, p.scale_log2e, bias_ub);
                    post(bound_prev);
                }
            }
            // ---- epilogue of this query tile: O / l, 32 columns per half
            post(l_run);
            const float l_row = l_run + collect();
            mbar_wait(o_done, (uint32_t)(g - 1) & 1u);
            tcgen05_fence_after();
            const float inv = (qrow < len) ? 1.f / l_row : 0.f;
            {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem_O + lane_off + half * 32, ov);
                tmem_ld_wait();
                tcgen05_fence_before();      // O copied out: the next query tile's first P.V may overwrite it
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                if (qrow < p.S) {
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w[e] = pack_bf16x2(__uint_as_float(ov[gq * 8 + 2 * e]) * inv, __uint_as_float(ov[gq * 8 + 2 * e + 1]) * inv);
                        *reinterpret_cast<uint4*>(orow + gq * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        tcgen05_fence_before();
    }

    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<1>(tmem_base, AT_TMEM_COLS);
}

// qkv: packed [B*S, ld] buffer; q/k/v head 0 start at columns q_col0/k_col0/v_col0.
// bias_const_from: the bias table is constant (per head and side) for |key - query| >= bias_const_from (T5: relative_attention_max_distance);
// <= 0 or >= S: no such guarantee, every tile reads the table.
template <bool HAS_BIAS, bool ROUND>
inline cudaError_t launch_attn_stream_t(const CUtensorMap& tm, const AttnTc2Params& p, int B, size_t smem, cudaStream_t stream) {
    auto kernel = attn_tc_d64_stream_kernel<HAS_BIAS, ROUND>;
    static std::atomic<size_t> max_set[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (smem > max_set[dev & 63].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        max_set[dev & 63].store(smem, std::memory_order_release);
    }
    return launch_pdl(kernel, dim3(1, p.H, B), dim3(192), smem, stream, tm, p);
}

template <bool HAS_BIAS, bool ROUND>
inline cudaError_t launch_attn_split_t(const CUtensorMap& tm, const AttnTc2Params& p, int B, size_t smem, cudaStream_t stream, int q_ctas) {
    auto kernel = attn_tc_d64_split_kernel<HAS_BIAS, ROUND>;
    static std::atomic<size_t> max_set[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (smem > max_set[dev & 63].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        max_set[dev & 63].store(smem, std::memory_order_release);
    }
    return launch_pdl(kernel, dim3((unsigned)q_ctas, p.H, B), dim3(320), smem, stream, tm, p);
}

// round_scores: reproduce the bf16 tensors of the reference's eager attention (scores, scores + bias) before the fp32 softmax.
inline cudaError_t launch_attn_tc(const __nv_bfloat16* qkv, int ld, int q_col0, int k_col0, int v_col0, __nv_bfloat16* o, int ldo,
                                  int B, int S, int H, const int* seq_lens, const float* bias_table, float scale, int bias_const_from,
                                  bool round_scores, cudaStream_t stream_) {
    CUtensorMap tm;
    if (!tmap_bf16_2d_cached(&tm, qkv, (uint64_t)B * S, (uint64_t)ld, (uint64_t)ld, 128)) return cudaErrorInvalidValue;
    // A/B switch (tools/bench_kernels.py): 30 / 31 = streaming stage without / with the score rounding, 40 / 41 = split-row stage
    static const int variant = [] { const char* v = getenv("VQA_ATTN_VARIANT"); return (v && v[0]) ? atoi(v) : -1; }();
    if (variant == 30 || variant == 40) round_scores = false;
    if (variant == 31 || variant == 41) round_scores = true;
    const int stage = variant >= 40 ? 2 : (variant >= 30 ? 1 : ATTN_STAGE_DEFAULT);
    AttnTc2Params p;
    p.o = o; p.ldo = ldo; p.seq_lens = seq_lens; p.bias_table = bias_table; p.S = S; p.H = H;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.scale = scale;
    const int n_tiles = (S + 127) / 128;
    int near = n_tiles;                                  // every tile reads the table
    if (bias_table && bias_const_from > 0 && bias_const_from < S) near = min(n_tiles, (bias_const_from - 1 + 127) / 128);
    p.near_tiles = bias_table ? near : 0;
    const size_t smem = attn_tc2_smem_bytes(p.near_tiles, bias_table != nullptr, round_scores);
    p.q_per_cta = n_tiles;
    if (stage == 2) {
        const size_t smem4 = smem + 2048 + 64 + 128;  // + the row-pair exchange buffers / barriers and the wider per-warp reduction scratch
        // small batches: B * H CTAs do not fill 148 SMs x 2 CTAs -> give each (sample, head) several CTAs, each with a share of the query tiles
        // (the K/V tiles are then read once per CTA instead of once per (sample, head): irrelevant at this size)
        int q_ctas = 1;
        if (B * H < 296 && n_tiles > 1) {
            q_ctas = min(n_tiles, (296 + B * H - 1) / (B * H));
            p.q_per_cta = (n_tiles + q_ctas - 1) / q_ctas;
            q_ctas = (n_tiles + p.q_per_cta - 1) / p.q_per_cta;
        }
        if (bias_table) return round_scores ? launch_attn_split_t<true, true>(tm, p, B, smem4, stream_, q_ctas) : launch_attn_split_t<true, false>(tm, p, B, smem4, stream_, q_ctas);
        return round_scores ? launch_attn_split_t<false, true>(tm, p, B, smem4, stream_, q_ctas) : launch_attn_split_t<false, false>(tm, p, B, smem4, stream_, q_ctas);
    }
    if (bias_table) return round_scores ? launch_attn_stream_t<true, true>(tm, p, B, smem, stream_) : launch_attn_stream_t<true, false>(tm, p, B, smem, stream_);
    return round_scores ? launch_attn_stream_t<false, true>(tm, p, B, smem, stream_) : launch_attn_stream_t<false, false>(tm, p, B, smem, stream_);
}

// ------------------------------------------------------------------------------------------------------------------
// head_dim 128 variant (Qwen2.5-VL language model: causal GQA; vision tower: windows / whole frames with the 80-wide heads
// zero-padded to 128). Variable-length sequences (cu_seqlens) or fixed stride S; one CTA per (query tile, head, sequence).
struct AttnTc128Params {
    __nv_bfloat16* o;          // [rows, ldo], column h*128 + d
    int ldo;
    const int* cu_seqlens;     // [n_seq + 1] row offsets (varlen) or nullptr
    const int* seq_lens;       // [n_seq] valid rows (fixed-stride mode) or nullptr
    int S;                     // fixed-stride mode: rows per sequence
    int q_col0, k_col0, v_col0;
    int kv_group;              // query heads per key/value head (GQA); 1 for MHA
    float scale_log2e;
    int o_head_stride, d_out;  // output column of head h = h * o_head_stride; only the first d_out (multiple of 8) of the 128 head dims are written
                               // (heads stored zero-padded to 128 in qkv produce d_out real output columns: the projection then contracts over them only)
    int pair, n_seq;           // varlen mode, pair != 0: CTA z covers the two consecutive sequences 2z and 2z+1 (each <= 64 rows) as ONE 128-row
                               // tile with a block-diagonal mask -- the 64-token windows of the Qwen2.5-VL vision tower (modeling_qwen2_5_vl.py
                               // get_window_index: 112-px windows = 8x8 patches) otherwise fill half of a tile and double the CTA count
    const int* kv_prefix;      // varlen mode, [n_seq] or nullptr: sequence b additionally attends to ALL rows of sequence kv_prefix[b] (>= 0),
                               // placed in front of its own keys -- the shared [system + vision] prefix of several prompts over one image
                               // (SURVEY 8(f)1): its K/V rows are computed once and read by every prompt's suffix
};

constexpr int A8_TILE = 128 * 128 * 2;       // 32 KB: 128 rows x 128 bf16 as two 64-column swizzle blocks
constexpr int A8_BLOCK = 128 * 64 * 2;       // 16 KB
constexpr int A8_TMEM_COLS = 512;            // S [0,128)  O [128,256)  P [256,320)
inline size_t attn_tc128_smem_bytes() { return 1024 + 5 * A8_TILE + 128; }

// Softmax stage as in the d64 kernel above (raw-score FMNMX3 maximum, packed FFMA2 scale/reference, one TMEM wait per tile; POLY of every
// 8 score pairs exponentiated on the FMA pipe).
template <bool CAUSAL, int POLY>
__global__ void __launch_bounds__(192, 1)
attn_tc_d128_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTc128Params p) {
    pdl_launch_dependents();
    pdl_wait();      // cu_seqlens / qkv are outputs of earlier kernels
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / p.kv_group;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    int row_base, len, write_rows, mid = 0;
    if (p.cu_seqlens && p.pair) {
        row_base = p.cu_seqlens[2 * b];
        mid = p.cu_seqlens[2 * b + 1] - row_base;                       // rows [0, mid): first sequence, [mid, len): second (absent for an odd tail)
        len = p.cu_seqlens[min(2 * b + 2, p.n_seq)] - row_base;
        write_rows = len;
    } else if (p.cu_seqlens) {
        row_base = p.cu_seqlens[b];
        len = p.cu_seqlens[b + 1] - row_base;
        write_rows = len;                       // rows past the sequence belong to the next one
    } else {
        row_base = b * p.S;
        len = p.seq_lens ? min(p.seq_lens[b], p.S) : p.S;
        write_rows = p.S;                       // padded rows of this sequence get zeros
    }
    const int q0 = qt * 128;
    if (q0 >= write_rows) return;
    if (q0 >= len) {
        for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
            const int r = i >> 4, c = i & 15;
            if (q0 + r < write_rows)
                if (c * 8 < p.d_out) *reinterpret_cast<uint4*>(p.o + (size_t)(row_base + q0 + r) * p.ldo + h * p.o_head_stride + c * 8) = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    extern __shared__ uint8_t a8_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a8_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + A8_TILE;        // [2]
    uint8_t* sV = smem + 3 * A8_TILE;    // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * A8_TILE);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 3;
    uint64_t* s_full = bars + 5;
    uint64_t* s_empty = bars + 6;
    uint64_t* p_full = bars + 7;
    uint64_t* o_done = bars + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

    // key tiles: first the nA tiles of the shared prefix sequence (all of it visible), then this sequence's own tiles (causal)
    int pre_base = 0, pre_len = 0;
    if (p.kv_prefix && p.cu_seqlens) {
        const int ps = p.kv_prefix[b];
        if (ps >= 0) { pre_base = p.cu_seqlens[ps]; pre_len = p.cu_seqlens[ps + 1] - pre_base; }
    }
    const int nA = (pre_len + 127) / 128;
    int nkt_own = (len + 127) / 128;
    if (CAUSAL) nkt_own = min(nkt_own, qt + 1);
    const int nkt = nA + nkt_own;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(q_full, 1);
        mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
        mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
        mbar_init(s_full, 1);
        mbar_init(s_empty, 4);
        mbar_init(p_full, 4);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc<1>(tmem_ptr_smem, A8_TMEM_COLS);
        tmem_relinquish<1>();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128, tmem_P = tmem_base + 256;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, A8_TILE);
            tma_load_2d(sQ, &tmap_qkv, q_full, p.q_col0 + h * 128, row_base + q0);
            tma_load_2d(sQ + A8_BLOCK, &tmap_qkv, q_full, p.q_col0 + h * 128 + 64, row_base + q0);
            for (int j = 0; j < nkt; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_empty[st], (((uint32_t)j >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&kv_full[st], 2 * A8_TILE);
                const int r = j < nA ? pre_base + j * 128 : row_base + (j - nA) * 128;
                tma_load_2d(sK + st * A8_TILE, &tmap_qkv, &kv_full[st], p.k_col0 + kvh * 128, r);
                tma_load_2d(sK + st * A8_TILE + A8_BLOCK, &tmap_qkv, &kv_full[st], p.k_col0 + kvh * 128 + 64, r);
                tma_load_2d(sV + st * A8_TILE, &tmap_qkv, &kv_full[st], p.v_col0 + kvh * 128, r);
                tma_load_2d(sV + st * A8_TILE + A8_BLOCK, &tmap_qkv, &kv_full[st], p.v_col0 + kvh * 128 + 64, r);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128);
            constexpr uint32_t idesc_o = make_idesc_bf16_f32(128, 128) | (1u << 16);   // B (= V) MN-major
            auto issue_pv = [&](int i) {
                mbar_wait(p_full, (uint32_t)i & 1u);
                tcgen05_fence_after();
                // V tile: two 64-wide head-dim atoms 16 KB apart (LBO), 8-key groups 1 KB apart (SBO)
                const uint64_t vdesc = make_mnmajor_sw128_desc(smem_u32(sV + (i & 1) * A8_TILE), A8_BLOCK);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_f16_ts(tmem_O, tmem_P + ks * 8, vdesc + (uint64_t)(ks * 128), idesc_o, (i > 0 || ks > 0) ? 1u : 0u);
                umma_commit<1>(&kv_empty[i & 1]);
                umma_commit<1>(o_done);
            };
            mbar_wait(q_full, 0);
            for (int j = 0; j < nkt; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_full[st], ((uint32_t)j >> 1) & 1u);
                mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
                tcgen05_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t qd = make_kmajor_sw128_desc(smem_u32(sQ + (k >> 2) * A8_BLOCK)) + 2 * (k & 3);
                    const uint64_t kd = make_kmajor_sw128_desc(smem_u32(sK + st * A8_TILE + (k >> 2) * A8_BLOCK)) + 2 * (k & 3);
                    umma_f16<1>(tmem_S, qd, kd, idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit<1>(s_full);
                if (j > 0) issue_pv(j - 1);
            }
            issue_pv(nkt - 1);
        }
    } else {
        const uint32_t quad = warp & 3u;
        const int row = quad * 32 + lane;
        const int qrow = q0 + row;
        const uint32_t lane_off = (quad * 32u) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        {
            const uint64_t cc = pack2(p.scale_log2e, p.scale_log2e);
            for (int j = 0; j < nkt; ++j) {
                const bool in_prefix = j < nA;
                const int k0 = in_prefix ? j * 128 : (j - nA) * 128;
                const int klen = in_prefix ? pre_len : len;
                mbar_wait(s_full, (uint32_t)j & 1u);
                tcgen05_fence_after();
                uint32_t sv[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tmem_S + lane_off + c * 32, sv[c]);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty);

                const bool diag = CAUSAL && !in_prefix && (j - nA) == qt;
                const bool edge = (k0 + 128 > klen) || diag || p.pair;
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (edge) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kcol = k0 + c * 32 + i;
                            if (kcol >= klen || (diag && kcol > qrow) || (p.pair && ((kcol < mid) != (qrow < mid)))) sv[c][i] = 0xff800000u;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        mx0 = fmax3(mx0, __uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1]));
                        mx1 = fmax3(mx1, __uint_as_float(sv[c][i + 2]), __uint_as_float(sv[c][i + 3]));
                        mx2 = fmax3(mx2, __uint_as_float(sv[c][i + 4]), __uint_as_float(sv[c][i + 5]));
                        mx3 = fmax3(mx3, __uint_as_float(sv[c][i + 6]), __uint_as_float(sv[c][i + 7]));
                    }
                }
                float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2e;
                if (!(tile_max > -1e30f)) tile_max = -1e30f;   // a padded query row may see no key at all in this tile (-inf * scale)
                float corr = 1.f;
                bool rescale = false;
                if (j == 0) {
                    m_run = tile_max;
                } else {
                    const bool need = tile_max > m_run + 8.f;
                    rescale = __any_sync(0xffffffffu, need);
                    if (rescale) {
                        const float m_new = fmaxf(m_run, tile_max);
                        corr = fast_exp2(m_run - m_new);
                        m_run = m_new;
                    }
                }
                if (j > 0) {
                    mbar_wait(o_done, (uint32_t)(j - 1) & 1u);
                    tcgen05_fence_after();
                    if (rescale) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint32_t ov[16];
                            tmem_ld_32x32b_x16(tmem_O + lane_off + c * 16, ov);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
                            tmem_st_32x32b_x16(tmem_O + lane_off + c * 16, ov);
                        }
                    }
                }
                uint64_t acc0 = 0ull, acc1 = 0ull;
                const uint64_t addc = pack2(-m_run, -m_run);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
                    softmax_chunk<false, POLY, false>(sv[c], pk, 0u, cc, addc, 0u, false, acc0, acc1);
                    tmem_st_32x32b_x16(tmem_P + lane_off + c * 16, pk);
                }
                {
                    float a0, a1, a2, a3;
                    unpack2(acc0, a0, a1);
                    unpack2(acc1, a2, a3);
                    l_run = l_run * corr + ((a0 + a1) + (a2 + a3));
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
            }
        }
        mbar_wait(o_done, (uint32_t)(nkt - 1) & 1u);
        tcgen05_fence_after();
        const float inv = (qrow < len && l_run > 0.f) ? 1.f / l_run : 0.f;
        __nv_bfloat16* orow = p.o + (size_t)(row_base + qrow) * p.ldo + h * p.o_head_stride;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 32 >= p.d_out) break;
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, ov);
            tmem_ld_wait();
            if (qrow < write_rows) {
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    if (c * 32 + gq * 8 >= p.d_out) break;
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = pack_bf16x2(__uint_as_float(ov[gq * 8 + 2 * e]) * inv, __uint_as_float(ov[gq * 8 + 2 * e + 1]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + gq * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<1>(tmem_base, A8_TMEM_COLS);
}

// rows: total rows of the packed buffer; max_len: longest sequence (grid sizing); kv_prefix: see AttnTc128Params (varlen mode only)
inline cudaError_t launch_attn_tc128(const __nv_bfloat16* qkv, int ld, long long rows, int q_col0, int k_col0, int v_col0,
                                     __nv_bfloat16* o, int ldo, int n_seq, int max_len, int S, int Hq, int kv_group,
                                     const int* cu_seqlens, const int* seq_lens, float scale, bool causal, cudaStream_t stream,
                                     const int* kv_prefix = nullptr, bool pair_sequences = false, int o_head_stride = 128, int d_out = 128) {
    CUtensorMap tm;
    if (!tmap_bf16_2d_cached(&tm, qkv, (uint64_t)rows, (uint64_t)ld, (uint64_t)ld, 128)) return cudaErrorInvalidValue;
    if (kv_prefix && !cu_seqlens) return cudaErrorInvalidValue;
    if (pair_sequences && (!cu_seqlens || causal || kv_prefix || max_len > 64)) return cudaErrorInvalidValue;
    AttnTc128Params p;
    if (d_out < 8 || d_out > 128 || (d_out & 7) || (o_head_stride & 7)) return cudaErrorInvalidValue;
    p.pair = pair_sequences ? 1 : 0; p.n_seq = n_seq;
    p.o_head_stride = o_head_stride; p.d_out = d_out;
    p.o = o; p.ldo = ldo; p.cu_seqlens = cu_seqlens; p.seq_lens = seq_lens; p.S = S;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.kv_group = kv_group;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.kv_prefix = kv_prefix;
    const size_t smem = attn_tc128_smem_bytes();
    static const int variant = [] { const char* v = getenv("VQA_ATTN128_VARIANT"); return (v && v[0]) ? atoi(v) : -1; }();   // 10 / 12: POLY 0 / 2
    const int poly = variant >= 10 ? variant - 10 : ATTN128_POLY_DEFAULT;
    dim3 grid(pair_sequences ? 1 : (max_len + 127) / 128, Hq, pair_sequences ? (n_seq + 1) / 2 : n_seq);
    auto go = [&](auto kernel, PerDeviceOnce& once) -> cudaError_t {
        cudaError_t e = once.ensure([&] { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
        if (e != cudaSuccess) return e;
        return launch_pdl(kernel, grid, dim3(192), smem, stream, tm, p);
    };
    static PerDeviceOnce once[4];
    if (causal) return poly == 0 ? go(attn_tc_d128_kernel<true, 0>, once[0]) : go(attn_tc_d128_kernel<true, 2>, once[1]);
    return poly == 0 ? go(attn_tc_d128_kernel<false, 0>, once[2]) : go(attn_tc_d128_kernel<false, 2>, once[3]);
}

}  // namespace vqa
