#!/usr/bin/env python
"""Kernel-level measurements through the C ABI (run on the GPU box).

  python tools/bench_kernels.py attn            # head_dim-64 attention: T5 encoder shape and CLIP shape, CUDA-event timing
  python tools/bench_kernels.py gemm-time       # the four T5-encoder GEMM shapes under a list of tile schedules, CUDA-event timing
  python tools/bench_kernels.py gemm-ncu        # same launches once each, in a fixed order, for `ncu --metrics dram__bytes_*`
  python tools/bench_kernels.py gemm-list       # prints that order (shape, group_rows, chunk_rows) as JSON
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def time_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def attn():
    from t2v_metrics_b200.engine import ops
    out = {}
    torch.manual_seed(0)
    for name, (B, S, H, bias, scale) in dict(t5_enc=(64, 672, 64, True, 1.0), t5_enc_ragged=(64, 672, 64, True, 1.0),
                                             clip=(64, 577, 16, False, 0.125)).items():
        qkv = (torch.randn(B * S, 3 * H * 64, device="cuda") * (0.5 if bias else 1.0)).bfloat16()
        table = None
        if bias:
            rel = torch.arange(-(S - 1), S, device="cuda").clamp(-128, 128) + 128
            table = (torch.randn(H, 257, device="cuda")).bfloat16().float()[:, rel].contiguous()
        lens = torch.randint(639, 673, (B,), device="cuda", dtype=torch.int32) if "ragged" in name else None
        rnd = os.environ.get("ATTN_ROUND", "1") == "1"
        fn = lambda: ops.attention(qkv, B, S, H, seq_lens=lens, bias_table=table, scale=scale, bias_const_from=128 if bias else 0, round_scores=rnd)
        ms = time_ms(fn)
        o = fn()
        torch.cuda.synchronize()
        flops = 4.0 * B * H * S * S * 64
        out[name] = dict(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), checksum=float(o.float().abs().sum()))
    print(json.dumps(dict(variant=os.environ.get("VQA_ATTN_VARIANT", "default"), **out)))


def _t5_attn_inputs(B=64, S=672, H=64):
    torch.manual_seed(0)
    qkv = (torch.randn(B * S, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    rel = torch.arange(-(S - 1), S, device="cuda").clamp(-128, 128) + 128
    table = (torch.randn(H, 257, device="cuda")).bfloat16().float()[:, rel].contiguous()
    return qkv, table


def attn_one():
    """one launch of the production attention kernel at the T5-encoder shape (target of `ncu --set full`)."""
    from t2v_metrics_b200.engine import ops
    qkv, table = _t5_attn_inputs()
    for _ in range(3):
        ops.attention(qkv, 64, 672, 64, bias_table=table, scale=1.0, bias_const_from=128, round_scores=True)
    torch.cuda.synchronize()


def attn128():
    """head_dim-128 attention at the Qwen2.5-VL-7B shapes: causal GQA prefill (B=32, S=320, 28/4 heads), vision windows (64 tokens) and whole frames."""
    import ctypes as C
    from t2v_metrics_b200 import _lib
    from t2v_metrics_b200.engine import _ptr, _stream_ptr, _check
    lib = _lib.load()
    torch.manual_seed(0)
    res = {}

    def run(qkv, out_cols, n_seq, max_len, S, Hq, group, cu, scale, causal, q0, k0, v0):
        out = torch.empty(qkv.shape[0], out_cols, dtype=torch.bfloat16, device="cuda")
        fn = lambda: _check(lib.vqa_op_attention_d128(_ptr(qkv), qkv.shape[1], qkv.shape[0], q0, k0, v0, _ptr(out), out_cols, n_seq, max_len, S, Hq,
                                                       group, _ptr(cu), None, float(scale), 1 if causal else 0, _stream_ptr(qkv.device)), None, "d128")
        return time_ms(fn), float(out.float().abs().sum())
    B, S, Hq, Hkv = 32, 320, 28, 4
    qkv = (torch.randn(B * S, (Hq + 2 * Hkv) * 128, device="cuda") * 0.5).bfloat16()
    ms, cs = run(qkv, Hq * 128, B, S, S, Hq, Hq // Hkv, None, 128 ** -0.5, True, 0, Hq * 128, (Hq + Hkv) * 128)
    res["llm_causal"] = dict(ms=round(ms, 4), tflops=round(2.0 * B * Hq * S * S * 128 / ms / 1e9, 1), checksum=cs)
    L, Hv = 32 * 1024, 16
    x = torch.zeros(L, 3, Hv, 128, device="cuda")
    x[..., :80] = torch.randn(L, 3, Hv, 80, device="cuda") * 0.5
    vq = x.reshape(L, 3 * Hv * 128).bfloat16()
    for name, wlen in (("vis_window64", 64), ("vis_frame1024", 1024)):
        cu = torch.arange(0, L + 1, wlen, dtype=torch.int32, device="cuda")
        ms, cs = run(vq, Hv * 128, L // wlen, wlen, 0, Hv, 1, cu, 80 ** -0.5, False, 0, Hv * 128, 2 * Hv * 128)
        res[name] = dict(ms=round(ms, 4), tflops_padded=round(4.0 * L * wlen * Hv * 128 / ms / 1e9, 1), checksum=cs)
    print(json.dumps(dict(variant=os.environ.get("VQA_ATTN128_VARIANT", "default"), **res)))


SHAPES = dict(   # name: (M, N, K, epilogue, residual)   clip-flant5-xxl encoder layer at B=64, S=672
    qkv=(43008, 12288, 4096, "store", False),
    o=(43008, 4096, 4096, "store", True),
    wi=(43008, 20480, 4096, "gated_gelu", False),
    wo=(43008, 4096, 10240, "store", True),
)
SCHEDULES = [(0, 0), (4096, -1), (1024, -1), (512, 2048), (512, 4096), (1024, 3072), (1024, 4096), (1024, 6144), (2048, 4096), (1024, 8192),
             (2048, 2048), (512, 6144)]


def gemm(mode):
    from t2v_metrics_b200 import _lib
    from t2v_metrics_b200.engine import ops
    lib = _lib.load()
    torch.manual_seed(0)
    rows = []
    only = os.environ.get("SHAPES", "").split(",") if os.environ.get("SHAPES") else list(SHAPES)
    scheds = [tuple(int(x) for x in s.split(":")) for s in os.environ["SCHEDS"].split(",")] if os.environ.get("SCHEDS") else SCHEDULES
    ballast = None
    if os.environ.get("BALLAST_GB"):      # mimic the footprint of the real step (weights + workspace) around the operands
        ballast = torch.empty(int(os.environ["BALLAST_GB"]) << 30, dtype=torch.uint8, device="cuda").fill_(1)
    for name in only:
        M, N, K, epi, res = SHAPES[name]
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
        if os.environ.get("ZERO_OPERANDS"):
            a.zero_(); w.zero_()
        n_out = N // 2 if epi == "gated_gelu" else N
        c = torch.empty(M, n_out, dtype=torch.bfloat16, device="cuda")
        r = torch.randn(M, n_out, device="cuda").bfloat16() if res else None
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for (g, ch) in scheds:
            lib.vqa_set_gemm_schedule(g, ch)
            fn = lambda: ops.gemm(a, w, residual=r, epilogue=epi, out=c, gate_up_offset=N // 2 if epi == "gated_gelu" else 0)
            if mode == "ncu":
                flush.zero_()      # start every measured launch from a cold L2
                fn()
                torch.cuda.synchronize()
            else:
                ms = time_ms(fn, iters=10, warm=2)
                rows.append(dict(shape=name, group_rows=g, chunk_rows=ch, ms=round(ms, 4), tflops=round(2.0 * M * N * K / ms / 1e9, 1)))
        lib.vqa_set_gemm_schedule(0, 0)
        del a, w, c, r
    if mode != "ncu":
        print(json.dumps(rows))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "attn":
        attn()
    elif cmd == "attn-one":
        attn_one()
    elif cmd == "attn128":
        attn128()
    elif cmd == "gemm-time":
        gemm("time")
    elif cmd == "gemm-ncu":
        gemm("ncu")
    elif cmd == "gemm-list":
        only = os.environ.get("SHAPES", "").split(",") if os.environ.get("SHAPES") else list(SHAPES)
        print(json.dumps([dict(shape=n, group_rows=g, chunk_rows=c) for n in only for (g, c) in SCHEDULES]))
