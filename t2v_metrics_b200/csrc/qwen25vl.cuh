// Host-side orchestration of the Qwen2.5-VL VQAScore forward: P(answer token | image, question) for a batch of prompts
// in ONE prefill (the reference runs generate(max_new_tokens=1) per sample: qwen2vl_model.py:190-230).
// Included by vqa_b200.cu (it uses that file's handle, GEMM/norm wrappers and ProfScope).
#pragma once

struct QwenVisLayerW {
    const bf16 *norm1, *norm2, *qkv_w, *qkv_b, *proj_w, *proj_b, *gu_w, *gu_b, *down_w, *down_b;
};
struct QwenLlmLayerW {
    const bf16 *ln1, *ln2, *qkv_w, *qkv_b, *o_w, *gu_w, *down_w;
};
struct QwenState {
    vqa_qwen25vl_config cfg;
    const bf16 *patch_w = nullptr, *merger_ln = nullptr, *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;
    const bf16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
    std::vector<QwenVisLayerW> vis;
    std::vector<QwenLlmLayerW> llm;
    // rotary metadata (device): axis of each frequency index and its inverse frequency
    int *text_axis = nullptr, *vis_axis = nullptr;
    float *text_inv_freq = nullptr, *vis_inv_freq = nullptr;
    bool rope_set = false;
};

static inline int qwen_mlp_pad(int mlp) { return (mlp + 127) / 128 * 128; }

struct QwenWorkspace {
    size_t patches, vx, vxn, vqkv, vattn, vff, vmerge_in, vfc1, vfeat, vfeat_orig, vcos, vsin;
    size_t x, xn, qkv, attn, ff, cos, sin, last, lastn, lse_max, lse_sum, label_logit, logprob, pen_bitmap, trace_logits;
    size_t total;
};
// B = prompts scored, M = token rows of the language model (B * S in the padded layout; fewer when prompts share a vision prefix)
static QwenWorkspace qwen_plan_rows(const vqa_qwen25vl_config& c, int B, size_t M, int n_patches) {
    Plan pl;
    QwenWorkspace w;
    const size_t L = n_patches;
    const size_t unit = (size_t)c.spatial_merge * c.spatial_merge;
    const size_t Dv = c.vit_hidden, Hv = c.vit_heads;
    const size_t mlp_pad = qwen_mlp_pad(c.vit_mlp);
    w.patches = pl.take(L * c.patch_dim * 2);
    w.vx = pl.take(L * Dv * 2);
    w.vxn = pl.take(L * Dv * 2);
    w.vqkv = pl.take(L * 3 * Hv * 128 * 2);
    w.vattn = pl.take(L * Hv * 128 * 2);      // used as [L, Hv * head_dim] (compact heads)
    w.vff = pl.take(L * mlp_pad * 2);
    w.vmerge_in = pl.take(L * Dv * 2);
    w.vfc1 = pl.take(L / unit * (Dv * unit) * 2);
    w.vfeat = pl.take(L / unit * c.out_hidden * 2);
    w.vfeat_orig = pl.take(L / unit * c.out_hidden * 2);
    w.vcos = pl.take(L * c.vit_head_dim * 4);
    w.vsin = pl.take(L * c.vit_head_dim * 4);
    const size_t qkv_cols = (size_t)(c.heads + 2 * c.kv_heads) * 128;
    w.x = pl.take(M * c.hidden * 2);
    w.xn = pl.take(M * c.hidden * 2);
    w.qkv = pl.take(M * qkv_cols * 2);
    w.attn = pl.take(M * c.heads * 128 * 2);
    w.ff = pl.take(M * c.mlp * 2);
    w.cos = pl.take(M * 128 * 4);
    w.sin = pl.take(M * 128 * 4);
    w.last = pl.take((size_t)B * c.hidden * 2);
    w.lastn = pl.take((size_t)B * c.hidden * 2);
    const size_t ntiles = (size_t)LMHEAD_PARTS * ((c.vocab + LMHEAD_BN - 1) / LMHEAD_BN);
    w.lse_max = pl.take((size_t)B * ntiles * 4);
    w.lse_sum = pl.take((size_t)B * ntiles * 4);
    w.label_logit = pl.take((size_t)B * 4);
    w.logprob = pl.take((size_t)B * 4);
    w.pen_bitmap = pl.take((size_t)B * ((c.vocab + 31) / 32) * 4);
    w.trace_logits = pl.take((size_t)B * c.vocab * 2);      // only written by vqa_qwen25vl_topk (trace output); the scoring path never stores logits
    w.total = pl.off;
    return w;
}

static QwenWorkspace qwen_plan(const vqa_qwen25vl_config& c, int B, int S, int n_patches) { return qwen_plan_rows(c, B, (size_t)B * S, n_patches); }

// Packed-rows layout of the language-model tokens (KV-prefix sharing): sequences are stored back to back (cu_seqlens);
// a prompt's [system + vision] prefix is ONE sequence shared by every prompt over that image, each prompt's remaining tokens are their own
// sequence whose kv_prefix points at it. Causal attention makes this exact: a prefix row never sees a suffix.
struct QwenPacked {
    int total_rows, n_seq, max_seq_len, max_prompt_len;
    const int* cu_seqlens;   // [n_seq + 1]
    const int* kv_prefix;    // [n_seq], -1 = none
    const int* pair_row;     // [B] packed row of each prompt's last token
    const int* pair_seq;     // [B] sequence holding that row
};

static int qwen_finalize(vqa_handle* h, QwenState& q) {
    const vqa_qwen25vl_config& c = q.cfg;
    bool ok = true;
    const int Dv = c.vit_hidden, Hv = c.vit_heads, unit = c.spatial_merge * c.spatial_merge, mlp_pad = qwen_mlp_pad(c.vit_mlp);
    q.patch_w = need(h, "vis.patch_embed", Dv, c.patch_dim, ok);
    q.vis.resize(c.vit_depth);
    for (int i = 0; i < c.vit_depth; ++i) {
        const std::string p = "vis." + std::to_string(i) + ".";
        QwenVisLayerW& L = q.vis[i];
        L.norm1 = need(h, p + "norm1", Dv, 1, ok); L.norm2 = need(h, p + "norm2", Dv, 1, ok);
        // native head width (80 for Qwen2.5-VL): the HF qkv weight is already ordered (q|k|v, head, dim); the GEMM epilogue scatters each head's
        // columns into the 128-wide slots the attention kernel reads, the attention writes compact heads, proj contracts over Hv * head_dim
        L.qkv_w = need(h, p + "qkv.weight", 3 * Hv * c.vit_head_dim, Dv, ok); L.qkv_b = need(h, p + "qkv.bias", 3 * Hv * c.vit_head_dim, 1, ok);
        L.proj_w = need(h, p + "proj.weight", Dv, Hv * c.vit_head_dim, ok); L.proj_b = need(h, p + "proj.bias", Dv, 1, ok);
        L.gu_w = need(h, p + "gate_up.weight", 2 * mlp_pad, Dv, ok); L.gu_b = need(h, p + "gate_up.bias", 2 * mlp_pad, 1, ok);
        L.down_w = need(h, p + "down.weight", Dv, mlp_pad, ok); L.down_b = need(h, p + "down.bias", Dv, 1, ok);
    }
    q.merger_ln = need(h, "vis.merger.ln_q", Dv, 1, ok);
    q.fc1_w = need(h, "vis.merger.fc1.weight", Dv * unit, Dv * unit, ok); q.fc1_b = need(h, "vis.merger.fc1.bias", Dv * unit, 1, ok);
    q.fc2_w = need(h, "vis.merger.fc2.weight", c.out_hidden, Dv * unit, ok); q.fc2_b = need(h, "vis.merger.fc2.bias", c.out_hidden, 1, ok);
    const int D = c.hidden, qkv_rows = (c.heads + 2 * c.kv_heads) * 128;
    q.embed = need(h, "llm.embed", c.vocab, D, ok);
    q.final_norm = need(h, "llm.norm", D, 1, ok);
    q.lm_head = need(h, "llm.lm_head", c.vocab, D, ok);
    q.llm.resize(c.layers);
    for (int i = 0; i < c.layers; ++i) {
        const std::string p = "llm." + std::to_string(i) + ".";
        QwenLlmLayerW& L = q.llm[i];
        L.ln1 = need(h, p + "ln1", D, 1, ok); L.ln2 = need(h, p + "ln2", D, 1, ok);
        L.qkv_w = need(h, p + "qkv.weight", qkv_rows, D, ok); L.qkv_b = need(h, p + "qkv.bias", qkv_rows, 1, ok);
        L.o_w = need(h, p + "o.weight", D, c.heads * 128, ok);
        L.gu_w = need(h, p + "gate_up.weight", 2 * c.mlp, D, ok);
        L.down_w = need(h, p + "down.weight", D, c.mlp, ok);
    }
    if (!ok) return VQA_ERR_MISSING_WEIGHT;
    if (!q.rope_set) return fail(h, VQA_ERR_INVALID_ARG, "vqa_qwen25vl_set_rope has not been called");
    return VQA_OK;
}

static int qwen_score(vqa_handle* h, QwenState& q, const void* pixel_patches, int pixel_dtype, int n_patches,
                      const int* vis_pos_hw, const int* window_index, const int* reverse_index, const int* cu_window,
                      int n_windows, int max_window_len, const int* cu_frames, int n_frames, int max_frame_len,
                      const int* input_ids, const int* seq_lens, const int* feat_index, const int* position_ids,
                      const int* answer_ids, int B, int S, float temperature, float repetition_penalty, float* out_probs,
                      float* out_logprobs,
                      void* workspace, size_t workspace_bytes, cudaStream_t st, const QwenPacked* pk = nullptr) {
    const vqa_qwen25vl_config& c = q.cfg;
    const QwenWorkspace w = pk ? qwen_plan_rows(c, B, (size_t)pk->total_rows, n_patches) : qwen_plan(c, B, S, n_patches);
    if (workspace_bytes < w.total) return fail(h, VQA_ERR_WORKSPACE, "workspace too small");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    auto P_ = [&](size_t off) { return reinterpret_cast<bf16*>(ws + off); };
    auto F_ = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    auto P8_ = [&](size_t off) { return ws + off; };
    h->launches = 0;
    h->prof.clear();
    h->ev_used = 0;
    int64_t* lc = &h->launches;
    const int nsm = h->num_sms;
#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)
    auto cuda_ok = [&](cudaError_t e, const char* what) -> int {
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return fail(h, VQA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return 0;
    };
    auto gemm = [&](const bf16* A, int lda, const bf16* W, int ldw, int w_rows, bf16* C, int ldc, int M_, int N_, int K_,
                    const bf16* bias, const bf16* res, int ldr, int epi, int gate_off) -> int {
        const double n_out = epi_is_gated(epi) ? N_ / 2 : N_;
        const double bytes = 2.0 * ((double)M_ * K_ + (double)N_ * K_ + (double)M_ * n_out * (res ? 2 : 1));
        ProfScope ps(h, CAT_GEMM, 2.0 * M_ * (double)N_ * K_, st, bytes);
        return cuda_ok(run_gemm(A, lda, W, ldw, w_rows, C, ldc, M_, N_, K_, bias, res, ldr, epi, gate_off, 0, nsm, st, lc), "gemm");
    };
    auto rms = [&](const bf16* x, const bf16* wgt, bf16* y, int rows, int D_) -> int {
        ProfScope ps(h, CAT_NORM, 0, st);
        return cuda_ok(run_rmsnorm(x, wgt, y, rows, D_, c.rms_eps, st, lc), "rmsnorm");
    };
    const int L = n_patches, Dv = c.vit_hidden, Hv = c.vit_heads, hdv = c.vit_head_dim;
    const int unit = c.spatial_merge * c.spatial_merge, mlp_pad = qwen_mlp_pad(c.vit_mlp);
    const int D = c.hidden, M = pk ? pk->total_rows : B * S;

    // ---------------- vision tower ----------------
    {
        ProfScope ps(h, CAT_OTHER, 0, st);
        *lc += 3;
        const size_t n = (size_t)L * c.patch_dim;
        bf16* patches = P_(w.patches);
        if (pixel_dtype == VQA_DTYPE_F32)
            cast_f32_bf16_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(reinterpret_cast<const float*>(pixel_patches), patches, n);
        else
            CUDA_TRY(h, cudaMemcpyAsync(patches, pixel_patches, n * 2, cudaMemcpyDeviceToDevice, st));
        // rotary table of the (already window-ordered) patch positions: [h freqs | w freqs] duplicated (:404-409, :487)
        rope_table_kernel<<<(L * (hdv / 2) + 255) / 256, 256, 0, st>>>(vis_pos_hw, L, q.vis_axis, q.vis_inv_freq, hdv / 2, F_(w.vcos), F_(w.vsin), 0);
        TRY(cuda_ok(cudaSuccess, "vision prologue"));
    }
    // patch embedding (Conv3d == GEMM), then re-order 2x2 groups into window order (:470-472)
    TRY(gemm(P_(w.patches), c.patch_dim, q.patch_w, c.patch_dim, Dv, P_(w.vxn), Dv, L, Dv, c.patch_dim, nullptr, nullptr, 0, EPI_STORE, 0));
    {
        ProfScope ps(h, CAT_OTHER, 0, st);
        ++*lc;
        gather_rows_kernel<<<L, 128, 0, st>>>(P_(w.vxn), P_(w.vx), window_index, unit, Dv);
        TRY(cuda_ok(cudaSuccess, "window reorder"));
        if (hdv < 128) {   // the qkv GEMMs write head_dim of every 128-wide head slot; the attention kernel contracts q.k over all 128: zero the rest once
            ++*lc;
            CUDA_TRY(h, cudaMemsetAsync(P_(w.vqkv), 0, (size_t)L * 3 * Hv * 128 * 2, st));
        }
    }
    const float vscale = 1.0f / sqrtf((float)hdv);
    for (int l = 0; l < c.vit_depth; ++l) {
        const QwenVisLayerW& Lw = q.vis[l];
        const bool full = (c.fullatt_mask >> l) & 1ull;
        TRY(rms(P_(w.vx), Lw.norm1, P_(w.vxn), L, Dv));
        {
            const double nq = 3.0 * Hv * hdv;
            ProfScope ps(h, CAT_GEMM, 2.0 * L * nq * Dv, st, 2.0 * ((double)L * Dv + nq * Dv + (double)L * nq));
            TRY(cuda_ok(run_gemm(P_(w.vxn), Dv, Lw.qkv_w, Dv, 3 * Hv * hdv, P_(w.vqkv), 3 * Hv * 128, L, 3 * Hv * hdv, Dv, Lw.qkv_b, nullptr, 0, EPI_STORE,
                                 0, 0, nsm, st, lc, false, nullptr, hdv < 128 ? hdv : 0, 128), "vision qkv gemm"));
        }
        {
            ProfScope ps(h, CAT_OTHER, 0, st);
            ++*lc;
            const long long nthr = (long long)L * 2 * Hv * (hdv / 16);
            rope_inplace_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(P_(w.vqkv), 3 * Hv * 128, 0, 2 * Hv, 128, hdv, F_(w.vcos),
                                                                            F_(w.vsin), L, 0);
            TRY(cuda_ok(cudaSuccess, "vision rope"));
        }
        {
            const int nseq = full ? n_frames : n_windows, mlen = full ? max_frame_len : max_window_len;
            ProfScope ps(h, CAT_ATTENTION, 4.0 * (double)L * mlen * Hv * hdv, st);
            ++*lc;
            static const bool no_pairs = getenv("VQA_ATTN128_NO_PAIRS") != nullptr;     // A/B switch
            TRY(cuda_ok(launch_attn_tc128(P_(w.vqkv), 3 * Hv * 128, L, 0, Hv * 128, 2 * Hv * 128, P_(w.vattn), Hv * hdv, nseq, mlen, 0, Hv, 1,
                                          full ? cu_frames : cu_window, nullptr, vscale, false, st, nullptr,
                                          /*two windows per tile*/ !full && max_window_len <= 64 && !no_pairs, /*compact heads*/ hdv, hdv),
                        "vision attention"));
        }
        TRY(gemm(P_(w.vattn), Hv * hdv, Lw.proj_w, Hv * hdv, Dv, P_(w.vx), Dv, L, Dv, Hv * hdv, Lw.proj_b, P_(w.vx), Dv, EPI_STORE, 0));
        TRY(rms(P_(w.vx), Lw.norm2, P_(w.vxn), L, Dv));
        TRY(gemm(P_(w.vxn), Dv, Lw.gu_w, Dv, 2 * mlp_pad, P_(w.vff), mlp_pad, L, 2 * mlp_pad, Dv, Lw.gu_b, nullptr, 0, EPI_GATED_SILU, mlp_pad));
        TRY(gemm(P_(w.vff), mlp_pad, Lw.down_w, mlp_pad, Dv, P_(w.vx), Dv, L, Dv, mlp_pad, Lw.down_b, P_(w.vx), Dv, EPI_STORE, 0));
    }
    // merger: RMSNorm -> [L/4, 4*Dv] -> Linear + GELU -> Linear; back to the original token order (:144-146, :512-513)
    TRY(rms(P_(w.vx), q.merger_ln, P_(w.vmerge_in), L, Dv));
    TRY(gemm(P_(w.vmerge_in), Dv * unit, q.fc1_w, Dv * unit, Dv * unit, P_(w.vfc1), Dv * unit, L / unit, Dv * unit, Dv * unit, q.fc1_b, nullptr, 0,
             EPI_GELU_ERF, 0));
    TRY(gemm(P_(w.vfc1), Dv * unit, q.fc2_w, Dv * unit, c.out_hidden, P_(w.vfeat), c.out_hidden, L / unit, c.out_hidden, Dv * unit, q.fc2_b, nullptr,
             0, EPI_STORE, 0));
    {
        ProfScope ps(h, CAT_OTHER, 0, st);
        *lc += 3;
        gather_rows_kernel<<<L / unit, 128, 0, st>>>(P_(w.vfeat), P_(w.vfeat_orig), reverse_index, 1, c.out_hidden);
        // ---------------- language model input ----------------
        // packed rows: one "sample" of M rows, all valid (its length is the last entry of cu_seqlens)
        qwen_embed_kernel<<<M, 128, 0, st>>>(input_ids, feat_index, pk ? pk->cu_seqlens + pk->n_seq : seq_lens, q.embed, P_(w.vfeat_orig), P_(w.x),
                                             pk ? M : S, D);
        rope_table_kernel<<<(M * 64 + 255) / 256, 256, 0, st>>>(position_ids, M, q.text_axis, q.text_inv_freq, 64, F_(w.cos), F_(w.sin), 1);
        TRY(cuda_ok(cudaSuccess, "llm prologue"));
    }
    const int Hq = c.heads, Hkv = c.kv_heads, qkv_cols = (Hq + 2 * Hkv) * 128;
    const float tscale = 1.0f / sqrtf(128.0f);
    for (int l = 0; l < c.layers; ++l) {
        const QwenLlmLayerW& Lw = q.llm[l];
        TRY(rms(P_(w.x), Lw.ln1, P_(w.xn), M, D));
        TRY(gemm(P_(w.xn), D, Lw.qkv_w, D, qkv_cols, P_(w.qkv), qkv_cols, M, qkv_cols, D, Lw.qkv_b, nullptr, 0, EPI_STORE, 0));
        {
            ProfScope ps(h, CAT_OTHER, 0, st);
            ++*lc;
            // q heads and k heads are contiguous in the packed buffer -> one launch rotates Hq + Hkv heads
            const long long nthr = (long long)M * (Hq + Hkv) * 8;
            rope_inplace_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(P_(w.qkv), qkv_cols, 0, Hq + Hkv, 128, 128, F_(w.cos), F_(w.sin), M,
                                                                            c.emulate_bf16_rounding);
            TRY(cuda_ok(cudaSuccess, "llm rope"));
        }
        {
            ProfScope ps(h, CAT_ATTENTION, 2.0 * B * (double)Hq * S * S * 128, st);
            ++*lc;
            if (pk)
                TRY(cuda_ok(launch_attn_tc128(P_(w.qkv), qkv_cols, M, 0, Hq * 128, (Hq + Hkv) * 128, P_(w.attn), Hq * 128, pk->n_seq, pk->max_seq_len, 0,
                                              Hq, Hq / Hkv, pk->cu_seqlens, nullptr, tscale, true, st, pk->kv_prefix), "llm attention (packed)"));
            else
                TRY(cuda_ok(launch_attn_tc128(P_(w.qkv), qkv_cols, M, 0, Hq * 128, (Hq + Hkv) * 128, P_(w.attn), Hq * 128, B, S, S, Hq, Hq / Hkv, nullptr,
                                              seq_lens, tscale, true, st), "llm attention"));
        }
        TRY(gemm(P_(w.attn), Hq * 128, Lw.o_w, Hq * 128, D, P_(w.x), D, M, D, Hq * 128, nullptr, P_(w.x), D, EPI_STORE, 0));
        TRY(rms(P_(w.x), Lw.ln2, P_(w.xn), M, D));
        TRY(gemm(P_(w.xn), D, Lw.gu_w, D, 2 * c.mlp, P_(w.ff), c.mlp, M, 2 * c.mlp, D, nullptr, nullptr, 0, EPI_GATED_SILU, c.mlp));
        TRY(gemm(P_(w.ff), c.mlp, Lw.down_w, c.mlp, D, P_(w.x), D, M, D, c.mlp, nullptr, P_(w.x), D, EPI_STORE, 0));
    }
    // ---------------- last position -> final norm -> lm_head with fused log-softmax gather ----------------
    {
        ProfScope ps(h, CAT_OTHER, 0, st);
        ++*lc;
        if (pk) gather_rows_by_index_kernel<<<B, 128, 0, st>>>(P_(w.x), pk->pair_row, P_(w.last), D);
        else    gather_last_rows_kernel<<<B, 128, 0, st>>>(P_(w.x), seq_lens, P_(w.last), S, D);
        TRY(cuda_ok(cudaSuccess, "last-row gather"));
    }
    TRY(rms(P_(w.last), q.final_norm, P_(w.lastn), B, D));
    const int ntiles = LMHEAD_PARTS * ((c.vocab + LMHEAD_BN - 1) / LMHEAD_BN);
    const int pen_words = (c.vocab + 31) / 32;
    const bool penalise = repetition_penalty != 1.0f;
    if (penalise) {
        // HF applies RepetitionPenaltyLogitsProcessor over the prompt ids before the scores are returned (generation/utils.py:2762-2770,
        // logits_process.py): one bit per (sample, vocabulary id) that occurs in the sample's prompt.
        ProfScope ps(h, CAT_OTHER, 0, st);
        *lc += 2;
        TRY(cuda_ok(cudaMemsetAsync(P8_(w.pen_bitmap), 0, (size_t)B * pen_words * 4, st), "penalty bitmap clear"));
        if (pk)
            token_bitmap_packed_kernel<<<dim3((pk->max_prompt_len + 255) / 256, B), 256, 0, st>>>(input_ids, pk->cu_seqlens, pk->kv_prefix, pk->pair_seq,
                                                                                                   c.vocab, reinterpret_cast<uint32_t*>(P8_(w.pen_bitmap)), pen_words);
        else
            token_bitmap_kernel<<<(B * S + 255) / 256, 256, 0, st>>>(input_ids, seq_lens, B, S, c.vocab,
                                                                     reinterpret_cast<uint32_t*>(P8_(w.pen_bitmap)), pen_words);
        TRY(cuda_ok(cudaSuccess, "penalty bitmap"));
    }
    {
        ProfScope ps(h, CAT_GEMM, 2.0 * B * (double)c.vocab * D, st, 2.0 * ((double)B * D + (double)c.vocab * D));
        TRY(cuda_ok(run_lmhead(P_(w.lastn), D, q.lm_head, D, B, c.vocab, D, answer_ids, F_(w.lse_max), F_(w.lse_sum), F_(w.label_logit), nsm, st,
                               lc, 1.0f / temperature, penalise ? reinterpret_cast<const uint32_t*>(P8_(w.pen_bitmap)) : nullptr, pen_words,
                               repetition_penalty), "lm_head"));
    }
    {
        ProfScope ps(h, CAT_OTHER, 0, st);
        ++*lc;
        // T = 1 target per row: score = exp(logprob) = softmax(logits / T)[answer] (qwen2vl_model.py:160-167)
        lse_finalize_kernel<<<(B + 3) / 4, 128, 0, st>>>(F_(w.lse_max), F_(w.lse_sum), F_(w.label_logit), answer_ids, out_probs,
                                                        out_logprobs ? out_logprobs : F_(w.logprob), B, 1, ntiles);
        TRY(cuda_ok(cudaSuccess, "lse finalize"));
    }
#undef TRY
    return VQA_OK;
}

// forward_with_trace support: top-k of the last-position distribution of the LAST scoring call (its final hidden states -- and, when a
// repetition penalty is in use, its prompt-token bitmap -- are still in the workspace). Materialises the [B, vocab] bf16 logits of that
// one position -- trace mode only.
static int qwen_topk(vqa_handle* h, QwenState& q, int B, size_t rows, int n_patches, int K, float temperature, float repetition_penalty, int* out_ids,
                     float* out_probs, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const vqa_qwen25vl_config& c = q.cfg;
    const QwenWorkspace w = qwen_plan_rows(c, B, rows, n_patches);
    if (workspace_bytes < w.total) return fail(h, VQA_ERR_WORKSPACE, "workspace too small");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    bf16* lastn = reinterpret_cast<bf16*>(ws + w.lastn);
    bf16* logits = reinterpret_cast<bf16*>(ws + w.trace_logits);
    const int pen_words = (c.vocab + 31) / 32;
    const bool penalise = repetition_penalty != 1.0f;
    const uint32_t* bitmap = reinterpret_cast<const uint32_t*>(ws + w.pen_bitmap);      // built by the scoring call with the same penalty
    CUDA_TRY(h, run_gemm(lastn, c.hidden, q.lm_head, c.hidden, c.vocab, logits, c.vocab, B, c.vocab, c.hidden, nullptr, nullptr, 0, EPI_STORE, 0, 0,
                         h->num_sms, st, nullptr));
    topk_softmax_kernel<<<B, 256, 0, st>>>(logits, (long long)c.vocab, c.vocab, 1.0f / temperature, penalise ? bitmap : nullptr, pen_words,
                                           repetition_penalty, K, out_ids, out_probs);
    CUDA_TRY(h, cudaGetLastError());
    return VQA_OK;
}
