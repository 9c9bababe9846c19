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
        fn = lambda: ops.attention(qkv, B, S, H, seq_lens=lens, bias_table=table, scale=scale, bias_const_from=128 if bias else 0)
        ms = time_ms(fn)
        o = fn()
        torch.cuda.synchronize()
        flops = 4.0 * B * H * S * S * 64
        out[name] = dict(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), checksum=float(o.float().abs().sum()))
    print(json.dumps(dict(variant=os.environ.get("VQA_ATTN_VARIANT", "default"), **out)))


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
    for name in only:
        M, N, K, epi, res = SHAPES[name]
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
        n_out = N // 2 if epi == "gated_gelu" else N
        c = torch.empty(M, n_out, dtype=torch.bfloat16, device="cuda")
        r = torch.randn(M, n_out, device="cuda").bfloat16() if res else None
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for (g, ch) in SCHEDULES:
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
    elif cmd == "gemm-time":
        gemm("time")
    elif cmd == "gemm-ncu":
        gemm("ncu")
    elif cmd == "gemm-list":
        only = os.environ.get("SHAPES", "").split(",") if os.environ.get("SHAPES") else list(SHAPES)
        print(json.dumps([dict(shape=n, group_rows=g, chunk_rows=c) for n in only for (g, c) in SCHEDULES]))
