#!/usr/bin/env python
"""bench.py -- VQAScore (image,text) pairs/s for clip-flant5-xxl on B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            # engine arm (this repo's sm_100a kernels)
  python bench.py --impl reference --gpus N ...            # reference arm: the reference algorithm on the host CPU cores
  torchrun --nproc-per-node N bench.py --gpus N ...        # weak scaling: one replica + one batch of 64 pairs per GPU

A "step" = one pass of the scoring hot path over one batch of 64 synthetic (image,text) pairs (BASELINE config 2:
512x512 uint8 images -> 336x336 CLIP input, 97 ids incl. the image slot (S_enc = 672), labels ['Yes', </s>]).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_PAIR = {"clip-flant5-xxl": 7.896e12, "clip-flant5-xl": 2.294e12,    # SURVEY 8(d), reference algorithm
                  "qwen2.5-vl-7b": 5.545e12}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="clip-flant5-xxl")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--text-len", type=int, default=97)
    ap.add_argument("--ragged", action="store_true", help="clip-flant5: text lens ~U[64, text_len] instead of all = text_len (SURVEY 8d)")
    ap.add_argument("--video", action="store_true", help="qwen: SURVEY 8(d) config 5 shape (grid 8x16x16, S=576, batch 8)")
    ap.add_argument("--pairs", type=int, default=0, help="clip-flant5: SURVEY 8(d) config 4 -- a JOB of this many pairs sharded "
                    "contiguously over the ranks in batches of --batch (+ tail), one all-gather; a step = the whole job; strong scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu", action="store_true", help="profiling pass: 2 device steps only, no JSON (run under ncu)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), burst=float(p["bf16_tflops"]), hbm=float(p["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (sustained cuBLAS bf16)")
    except Exception:
        return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def measured_traffic(model):
    """DRAM bytes per GEMM launch from the committed ncu capture of the same step (profiles/r01_gemm_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            t = json.load(f)
        return t.get(model)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU (reference) leg
def cpu_reference_pairs_per_s(model: str, text_len: int):
    """The reference algorithm on the host cores: oracle/clipt5_oracle.py (CPU restatement of the transformers T5/CLIP
    forward the reference delegates to; the reference package itself cannot be imported offline), fp32,
    batch = 1 as in the reference's own CPU-runnable config. Bounded sample: the full-WIDTH model at depth 1 and depth 2
    (ViT/encoder/decoder layers), one pair each; per-layer cost = t2 - t1, extrapolated to 23/24/24 layers."""
    import torch
    from oracle import clipt5_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    base = orc.ClipT5Config.xxl() if model.endswith("xxl") else orc.ClipT5Config.xl()
    import dataclasses
    times = {}
    for depth in (1, 2):
        cfg = dataclasses.replace(base, vit_layers=depth + 1, enc_layers=depth, dec_layers=depth)
        sd = orc.make_synthetic_state_dict(cfg, seed=0)
        inp = orc.make_synthetic_inputs(cfg, 1, text_len, seed=1)
        orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], mode="fp32")  # warm-up
        best = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], mode="fp32")
            best = min(best, time.perf_counter() - t0)
        times[depth] = best
        del sd
    per_layer = max(times[2] - times[1], 1e-9)
    full = times[1] + (base.enc_layers - 1) * per_layer
    return dict(value=1.0 / full, unit="pairs/s", cores=os.cpu_count(), kind="port",
                sample=(f"1 pair, fp32, full-width {model}: depth-1 pass {times[1]:.2f}s, depth-2 pass {times[2]:.2f}s; "
                        f"per-layer-triple {per_layer:.2f}s extrapolated to {base.enc_layers} layers = {full:.1f}s/pair"),
                seconds_per_pair=full)


def run_reference(args, rank, world):
    if rank != 0:
        return
    if not args.model.startswith("clip-flant5"):
        # the headline metric (BASELINE.json) is the CLIP-FlanT5 one; the Qwen line is a secondary bench without a CPU arm
        emit(dict(impl="reference", unavailable=f"the CPU reference arm is implemented for clip-flant5-* only, not {args.model}"))
        return
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(cpu_reference_pairs_per_s(args.model, args.text_len))
    best = max(vals, key=lambda v: v["value"])
    line = dict(impl="reference", metric="VQAScore (image,text) pairs/sec @ clip-flant5-xxl, 512px", value=best["value"],
                unit="pairs/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * best["seconds_per_pair"] * args.batch, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"{args.model} VQAScore, batch=1 on host CPU, 336px CLIP input, {args.text_len} ids (S_enc=672), T=2",
                            model=args.model),
                cpu_baseline=dict(value=best["value"], unit="pairs/s", cores=best["cores"], kind=best["kind"], sample=best["sample"]),
                e2e=dict(value=best["value"], unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=round(time.perf_counter() - t0, 1))
    emit(line)


# ------------------------------------------------------------------------------------------------ engine arm
def run_job(args, rank, world, dev, cfg, eng):
    """N pairs, contiguous shard per rank (parallel.shard_bounds), full batches of --batch plus one
    tail batch, ONE all-gather of the N fp32 scores at the end. A step is the whole job (strong scaling: the total is fixed)."""
    import torch
    import torch.distributed as dist
    from t2v_metrics_b200.synthetic import synthetic_batch
    from t2v_metrics_b200.parallel import gather_scores, shard_bounds
    B, L, N = args.batch, args.text_len, args.pairs
    start, end, per = shard_bounds(N, world, rank)
    n_local = end - start
    full = {k: v.to(dev) for k, v in synthetic_batch(cfg, B, L, seed=1 + rank, ragged=args.ragged).items()}
    tail_n = n_local % B
    tail = {k: v[:tail_n].contiguous() for k, v in full.items()} if tail_n else None
    local = torch.empty(n_local, dtype=torch.float32, device=dev)

    def job():
        o = 0
        for _ in range(n_local // B):
            eng.score_tensors(full["pixels"], full["input_ids"], full["text_lens"], full["labels"], out=local[o:o + B])
            o += B
        if tail is not None:
            eng.score_tensors(tail["pixels"], tail["input_ids"], tail["text_lens"], tail["labels"], out=local[o:o + tail_n])
        return gather_scores(local, N) if world > 1 else local

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(3):        # warm-up: three full batches (+ the tail shape once)
        eng.score_tensors(full["pixels"], full["input_ids"], full["text_lens"], full["labels"])
    if tail is not None:
        eng.score_tensors(tail["pixels"], tail["input_ids"], tail["text_lens"], tail["labels"])
    sync_all()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 2))
    sync_all()
    ev0.record()
    for _ in range(steps):
        out = job()
    ev1.record()
    sync_all()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_job = float(t) / steps
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        assert out.numel() == N and bool(torch.isfinite(out).all()) and float(out.min()) >= 0 and float(out.max()) <= 1
        emit(dict(
            metric="VQAScore (image,text) pairs/sec @ clip-flant5-xxl, 512px", value=N / (ms_job * 1e-3), unit="pairs/s", n_gpus=world,
            steps=steps, warmup=3, ms_per_step=ms_job, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
            data="synthetic",
            config=dict(workload=f"{args.model} VQAScore JOB: {N} pairs sharded contiguously over {world} GPU(s) = {per} per rank in batches of "
                                 f"{B} + tail {per % B}, one all-gather of {N} fp32 scores; S_enc={L - 1 + cfg.num_patches}, T=2",
                        model=args.model, global_batch=B * world, seq_len=L - 1 + cfg.num_patches, parallelism=f"dp{world}",
                        l2_policy="inputs larger than L2"),
            gpu_launches=int(eng.last_launch_count()) * (-(-n_local // B)) * steps, clocks=clocks))
    if world > 1:
        dist.destroy_process_group()


def run_engine(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from t2v_metrics_b200.config import CLIPT5_MODELS
    from t2v_metrics_b200.engine import ClipT5Engine
    from t2v_metrics_b200.synthetic import synthetic_engine_weights, synthetic_batch
    from t2v_metrics_b200.parallel import gather_scores

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CLIPT5_MODELS[args.model]["config"]()
    eng = ClipT5Engine(cfg, dev)
    eng.bind_engine_tensors(synthetic_engine_weights(cfg, dev, seed=0))
    B, L = args.batch, args.text_len
    if args.pairs:
        return run_job(args, rank, world, dev, cfg, eng)
    host = synthetic_batch(cfg, B, L, seed=1 + rank, ragged=args.ragged, raw_u8=True)
    raw_u8 = host.pop("raw_u8")                      # [B, 512, 512, 3] uint8 pinned: the decoded images of config 2
    d = {k: v.to(dev) for k, v in host.items()}
    total_pairs = B * world

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_device():
        s = eng.score_tensors(d["pixels"], d["input_ids"], d["text_lens"], d["labels"])
        return gather_scores(s, total_pairs) if world > 1 else s

    def step_host():
        # the call a user makes: decoded uint8 images + token ids on the HOST -> H2D -> device pre-processing (expand2square, PIL-exact
        # bicubic 512 -> 336, normalise) -> forward -> scores back on the host
        s = eng.score_images_u8(raw_u8, host["input_ids"], host["text_lens"], host["labels"])
        if world > 1:
            s = gather_scores(s.to(dev), total_pairs).cpu()
        return s

    if args.ncu:
        for _ in range(2):
            step_device()
        sync_all()
        if rank == 0:
            print(f"ncu pass: {eng.last_launch_count()} launches per step", flush=True)
        return
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    sync_all()

    # ---- timed region 1: inputs resident in HBM (kernel-side throughput) + per-category device timing
    eng.set_profile(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_acc = {}
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    sync_all()
    ms_local = ev0.elapsed_time(ev1)
    prof = eng.read_profile()          # categories of the last step
    launches = eng.last_launch_count()
    eng.set_profile(False)
    t = torch.tensor([ms_local], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D + D2H inside)
    for _ in range(2):
        step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_host()
    torch.cuda.synchronize(dev)
    e2e_local = (time.perf_counter() - t0) * 1000.0
    t = torch.tensor([e2e_local], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t) / args.steps
    h2d = raw_u8.numel() + sum(v.numel() * v.element_size() for k, v in host.items() if k != "pixels")
    d2h = B * 4

    if rank == 0:
        peaks = measured_peaks()
        gemm_ms, gemm_flops, gemm_n, gemm_bytes = prof["gemm"]
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        value = total_pairs / (ms_step * 1e-3)
        traffic = measured_traffic(args.model)
        fpp = FLOPS_PER_PAIR.get(args.model)
        line = dict(
            metric="VQAScore (image,text) pairs/sec @ clip-flant5-xxl, 512px", value=value, unit="pairs/s", n_gpus=world,
            steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload=f"{args.model} VQAScore: batch {B}/GPU, synthetic 512x512 uint8 images -> 336px CLIP input, "
                                 f"{'64..' if args.ragged else ''}{L} ids incl. image slot (S_enc={L - 1 + cfg.num_patches}), labels [Yes,</s>] (T=2)",
                        model=args.model, global_batch=total_pairs, seq_len=L - 1 + cfg.num_patches, parallelism=f"dp{world}",
                        l2_policy="inputs larger than L2: 22.6 GB of weights + 4.5 GB of activations stream per step"),
            roofline=dict(bound="tensor", achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s",
                          frac=(achieved / peaks["tflops"]) if achieved else None,
                          traffic=traffic.get("dram_bytes_per_launch") if traffic else None, traffic_source=traffic.get("source") if traffic else None,
                          algorithmic_bytes_per_launch=gemm_bytes / max(gemm_n, 1),
                          kernel="gemm_bf16_sm100_kernel (all tcgen05 GEMM launches of the step)",
                          flops_per_launch=gemm_flops / max(gemm_n, 1), launches=gemm_n, device_ms=gemm_ms, peak_source=peaks["source"],
                          whole_step_tflops=(value / world) * fpp / 1e12 if fpp else None),
            breakdown_ms={k: round(v[0], 3) for k, v in prof.items()},
            e2e=dict(value=total_pairs / (e2e_ms_step * 1e-3), unit="pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     ms_per_step=e2e_ms_step,
                     path="ClipT5Engine.score_images_u8: pinned uint8 512x512 images + ids -> H2D -> vqa_clip_preprocess -> "
                          "vqa_clipt5_score -> D2H scores"),
            gpu_launches=int(launches) * args.steps, clocks=clocks,
            sample_scores=[round(float(x), 6) for x in out[:4].float().cpu()])
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = {k: v for k, v in cpu_reference_pairs_per_s(args.model, L).items() if k != "seconds_per_pair"}
            except Exception as e:  # noqa
                line["cpu_baseline"] = dict(value=None, unit="pairs/s", cores=os.cpu_count(), kind="port", sample=f"failed: {e!r}")
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_engine_qwen(args, rank, local_rank, world):
    """BASELINE config 3: qwen2.5-vl-7b VQAScore, batch 32 per GPU, synthetic 448x448 images (1024 patches -> 256 vision tokens) +
    64 text ids (S = 320). Secondary bench line (`--model qwen2.5-vl-7b`); the default/headline line is clip-flant5-xxl."""
    import torch
    import torch.distributed as dist
    from t2v_metrics_b200 import qwen_host
    from t2v_metrics_b200.config import QWEN25VL_MODELS
    from t2v_metrics_b200.engine import QwenVLEngine
    from t2v_metrics_b200.synthetic import synthetic_qwen_engine_weights, synthetic_qwen_batch
    from t2v_metrics_b200.parallel import gather_scores

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = QWEN25VL_MODELS[args.model]["config"]()
    eng = QwenVLEngine(cfg, dev)
    eng.bind_engine_tensors(synthetic_qwen_engine_weights(cfg, dev, seed=0))
    # --video = SURVEY 8(d) config 5: 16 frames of 224x224 -> grid (8, 16, 16) = 2048 patches / 512 video tokens, S = 576, B = 8
    video = bool(args.video)
    B = args.batch if args.batch != 64 else (8 if video else 32)
    hw, frames = ((224, 224), 8) if video else ((448, 448), 1)
    host = synthetic_qwen_batch(cfg, B, hw, 64, seed=1 + rank, frames=frames)
    seq_len = 64 + frames * (hw[0] // 28) * (hw[1] // 28)
    flops_pair = 10.25e12 if video else FLOPS_PER_PAIR[args.model]
    idx = qwen_host.build_batch_indices(host["prompts"], host["grid_thw"], list(range(B)), cfg.image_token_id, cfg.spatial_merge_size,
                                        cfg.tokens_per_second, video_token_id=cfg.video_token_id,
                                        second_per_grid_ts=[1.0] * B if video else None)
    idx = {k: v.pin_memory() for k, v in idx.items()}
    ans_h = torch.tensor(host["answer_ids"], dtype=torch.int32).pin_memory()
    d_idx = {k: v.to(dev) for k, v in idx.items()}
    d_pix, d_ans = host["pixel_patches"].to(dev), ans_h.to(dev)
    total = B * world

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_device():
        s = eng.score_tensors(d_pix, host["grid_thw"], d_idx["input_ids"], d_idx["seq_lens"], d_idx["feat_index"], d_idx["position_ids"], d_ans)
        return gather_scores(s, total) if world > 1 else s

    def step_host():
        t = {k: v.to(dev, non_blocking=True) for k, v in idx.items()}
        s = eng.score_tensors(host["pixel_patches"].to(dev, non_blocking=True), host["grid_thw"], t["input_ids"], t["seq_lens"],
                              t["feat_index"], t["position_ids"], ans_h.to(dev, non_blocking=True))
        if world > 1:
            s = gather_scores(s, total)
        return s.cpu()

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    sync_all()
    eng.set_profile(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    sync_all()
    prof = eng.read_profile()
    launches = eng.last_launch_count()
    eng.set_profile(False)
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize(dev)
    t = torch.tensor([(time.perf_counter() - t0) * 1000.0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t) / args.steps
    h2d = host["pixel_patches"].numel() * 4 + sum(v.numel() * 4 for v in idx.values()) + B * 4
    if rank == 0:
        peaks = measured_peaks()
        gemm_ms, gemm_flops, gemm_n, gemm_bytes = prof["gemm"]
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        value = total / (ms_step * 1e-3)
        shape = ("synthetic 16-frame 224x224 videos (grid 8x16x16: 2048 patches, 512 video tokens) + 64 text ids (S=576)" if video else
                 "synthetic 448x448 images (1024 patches, 256 vision tokens) + 64 text ids (S=320)")
        line = dict(metric="VQAScore (video,text) pairs/sec @ qwen2.5-vl-7b, 16x224px" if video else
                    "VQAScore (image,text) pairs/sec @ qwen2.5-vl-7b, 448px", value=value, unit="pairs/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic",
                    config=dict(workload=f"{args.model} VQAScore: batch {B}/GPU, {shape}, one answer token", model=args.model,
                                global_batch=total, seq_len=seq_len,
                                parallelism=f"dp{world}", l2_policy="inputs larger than L2: 16.6 GB of weights stream per step"),
                    roofline=dict(bound="tensor", achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s",
                                  frac=(achieved / peaks["tflops"]) if achieved else None, traffic=None,
                                  kernel="gemm_bf16_sm100_kernel (all tcgen05 GEMM launches of the step)", launches=gemm_n, device_ms=gemm_ms,
                                  flops_per_launch=gemm_flops / max(gemm_n, 1), peak_source=peaks["source"],
                                  whole_step_tflops=(value / world) * flops_pair / 1e12),
                    breakdown_ms={k: round(v[0], 3) for k, v in prof.items()},
                    e2e=dict(value=total / (e2e_ms * 1e-3), unit="pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=B * 4, ms_per_step=e2e_ms),
                    gpu_launches=int(launches) * args.steps, clocks=clocks, sample_scores=[float(x) for x in out[:4].float().cpu()])
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def reserve_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner on stdout under torchrun), so
    keep a private handle on the real stdout for the result line and point file descriptor 1 at stderr for everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    reserve_stdout()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.model.startswith("qwen"):
        run_engine_qwen(args, rank, local_rank, world)
    else:
        run_engine(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
