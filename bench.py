#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" is one pass of the hot path over one batch of synthetic input: B uint8 512x512 RGB images ->
4 crops (centre / padded / subcrop1 / subcrop2) -> PIL-exact resize + normalise -> ViT-L/14 (random-init
weights of the openai architecture, bf16 tensor-core GEMMs, fp32 residual stream) -> f32 [B,4,768] unit-norm
embeddings.  `value` is timed with the batch already in HBM; `e2e` goes through the public API
(CLIP_Encoder.encode_images_u8) from pinned host memory with the H2D copy of the images and the D2H read of
the embeddings inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (4-crop ViT-L/14 embed) at 1/2/4/8 B200; dedup sim-pairs/sec"
MODEL = "ViT-L-14/openai"
IMG_HW = 512


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], src="measured (MEASURED_PEAKS.json)")
    return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_batch(B, seed, device="cpu"):
    """Synthetic 512x512 uint8 images: low-frequency colour field + noise (cheap torch version of SURVEY §8d).  The
    B200 arm generates its pool on the device (eight ranks sharing 16 host cores would spend a minute here otherwise);
    the CPU arms generate theirs on the host.  Same distribution either way."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    ar = torch.arange(IMG_HW, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(ar, ar, indexing="ij")
    f = torch.rand(B, 3, 2, generator=g, device=device) * 3.5 + 0.5
    ph = torch.rand(B, 3, 2, generator=g, device=device) * 6.2832
    img = 128 + 90 * torch.sin(6.2832 * f[..., 0, None, None] * xx / IMG_HW + ph[..., 0, None, None]) * \
        torch.cos(6.2832 * f[..., 1, None, None] * yy / IMG_HW + ph[..., 1, None, None])
    img = img + 20 * torch.randn(B, 3, IMG_HW, IMG_HW, generator=g, device=device)
    return img.clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()  # [B,H,W,3]


# --------------------------------------------------------------------------------------------- reference arm / cpu baseline
_CPU_CACHE = {}


def cpu_reference_images_per_s(n_images, seed=0, threads=None):
    """The reference's CPU path for the same workload, on the oracle port: PIL crops + torchvision transform
    (what utils/embedder.py:164-175 runs per image) and the fp32 tower + L2 normalise (utils/embedder.py:94-100)."""
    import torch
    from PIL import Image
    from clip_assisted_data_labeling_b200.embedder import CustomImageDataset, _open_clip_val_transform
    from oracle import vit_oracle
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    if "m" not in _CPU_CACHE:
        _CPU_CACHE["m"] = vit_oracle.build_visual("ViT-L-14", "openai", seed=0, perturb=False)
    m = _CPU_CACHE["m"]
    tf = _open_clip_val_transform(224)
    ds = CustomImageDataset([], ["centre_crop", "square_padded_crop", "subcrop1", "subcrop2"], tf)
    imgs = synth_batch(n_images, seed).numpy()
    t0 = time.perf_counter()
    crops = []
    for im in imgs:
        raw, _ = ds.extract_crops(Image.fromarray(im))
        crops.append(torch.stack([tf(c) for c in raw]))
    x = torch.cat(crops)
    t1 = time.perf_counter()
    with torch.no_grad():
        out = vit_oracle.encode_image_oracle(m, x)
    t2 = time.perf_counter()
    assert out.shape == (4 * n_images, 768)
    return n_images / (t2 - t0), {"preprocess_s": t1 - t0, "forward_s": t2 - t1}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    # calibrate the per-step sample so that the whole run stays within a few minutes
    ips, _ = cpu_reference_images_per_s(1)
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    n = int(max(1, min(32, ips * min(budget_s, 8.0))))
    for _ in range(args.warmup):
        cpu_reference_images_per_s(n)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_reference_images_per_s(n, seed=s)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} synthetic 512x512 images x 4 crops per step (PIL crops + torchvision transform + fp32 ViT-L/14 tower, torch CPU, {cores} threads)"
    _emit(args.out_fd, {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: ViT-L/14 4-crop embedding, synthetic 512x512 images, random-init openai architecture",
                   "images_per_step": n, "note": "open_clip is not installable offline: the tower is the oracle's restatement of it"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# --------------------------------------------------------------------------------------------- B200 arm
GEMM_STAGES = ("in_proj", "out_proj", "c_fc", "c_proj")


def stage_profile(enc, batches, cfg, n_crops, peaks, steps=2):
    """Per-stage device time of whole steps, measured live with CUDA events on the launching stream by the
    library's stage timer (include/b2c.h: b2c_prof_enable / b2c_prof_read): every stage of the step is bracketed by
    an event pair.  The dominant kernel is umma2_tile_kernel<GemmPolicy<mode>> (the four GEMMs of each block);
    roofline.achieved = their algorithmic FLOPs per step / their summed duration per step."""
    import torch
    from clip_assisted_data_labeling_b200 import _lib
    d, mlp, L = cfg["width"], cfg["mlp"], cfg["layers"]
    T = (cfg["image"] // cfg["patch"]) ** 2 + 1
    M = n_crops * T
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for i in range(steps):
        enc.encode_images_u8(batches[i % len(batches)])
    torch.cuda.synchronize()
    rec = _lib.prof_read()
    _lib.prof_enable(False)
    shapes = {"in_proj": (3 * d, d), "out_proj": (d, d), "c_fc": (mlp, d), "c_proj": (d, mlp)}
    per, tot_f, tot_ms = {}, 0.0, 0.0
    total_ms = sum(ms for ms, _ in rec.values())
    for name in GEMM_STAGES:
        ms, n = rec[name]
        N, K = shapes[name]
        fl = 2.0 * M * N * K * L * steps  # all launches of this shape in the profiled steps
        per[name] = {"M_per_step": M, "N": N, "K": K, "launches": n, "ms_per_launch": ms / n, "tflops": fl / ms / 1e9}
        tot_f += fl
        tot_ms += ms
    ach = tot_f / tot_ms / 1e9
    shares = {k: {"ms_per_step": ms / steps, "share": ms / total_ms, "stages_per_step": n // steps} for k, (ms, n) in rec.items()}
    return {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
            "traffic": None,
            "kernel": "umma2_tile_kernel<GemmPolicy<mode>>: tcgen05.mma cta_group::2 (256x256x16 per CTA pair), TMA-fed 4-6-stage ring, "
                      "TMEM double-buffered accumulators, fused LayerNorm/bias/QuickGELU (in_proj, c_fc) and residual-update/"
                      "bf16-copy/row-statistics (out_proj, c_proj) epilogues",
            "peak_source": peaks["src"] + ", sustained (kernel timed inside the step); burst is %.1f" % peaks["tf_burst"],
            "frac_of_burst": ach / peaks["tf_burst"],
            "gemm_share_of_step": tot_ms / total_ms,
            "per_shape": per, "stage_shares": shares,
            "how": "library stage timer: CUDA event pairs on the launching stream around every stage of %d whole steps (the timer "
                   "runs the pass as one lane so that stages do not overlap); achieved = algorithmic GEMM FLOPs (2*M*N*K, no "
                   "padding) / summed GEMM stage time" % steps}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from clip_assisted_data_labeling_b200 import _lib
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    from clip_assisted_data_labeling_b200.vit_arch import ARCHS, flops_per_crop

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    B = args.batch
    cfg = ARCHS["ViT-L-14"]

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # stdout carries exactly one JSON line
        enc = CLIP_Encoder(MODEL, device="cuda", seed=0, allow_random_init=True)
    enc.model.set_lanes(args.lanes)
    enc.model.set_fused_ln(not args.standalone_layernorm)
    pool_dev = [synth_batch(B, 100 * rank + i, device="cuda") for i in range(args.pool)]
    pool_host = [b.cpu().pin_memory() for b in pool_dev]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM
    for i in range(args.warmup):
        enc.encode_images_u8(pool_dev[i % args.pool])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = enc.encode_images_u8(pool_dev[i % args.pool])
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    value = world * B * args.steps / (ms_total / 1e3)
    norms_ok = bool(torch.allclose(out.norm(dim=-1), torch.ones_like(out[..., 0]), atol=1e-4))

    # ---------------- informational variant, NOT the headline: last block evaluated on the class-token row only
    variant = None
    if rank == 0 and not args.no_variants:
        enc.model.set_cls_only_last_block(True)
        for i in range(3):
            enc.encode_images_u8(pool_dev[i % args.pool])
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for i in range(args.steps):
            vout = enc.encode_images_u8(pool_dev[i % args.pool])
        v1.record()
        torch.cuda.synchronize()
        enc.model.set_cls_only_last_block(False)
        variant = {"cls_only_last_block": {
            "value": B * args.steps / (v0.elapsed_time(v1) / 1e3), "unit": "images/s (one GPU, device-resident)",
            "max_abs_vs_full": float((vout - out).abs().max()),  # same input batch as the last timed step of `value`
            "note": "opt-in b2c_vit_set_cls_only_last_block(1): block 24 computes K/V for all tokens but attention, out_proj, ln_2, "
                    "c_fc, c_proj only for the class-token row that ln_post/proj read (3.3 % fewer FLOPs). Off in `value` and `e2e`."}}

    # ---------------- stage shares + roofline of the dominant kernel (separate, event-instrumented steps)
    roof = stage_profile(enc, pool_dev, cfg, 4 * B, peaks) if rank == 0 else None

    # ---------------- e2e: host buffers, H2D + D2H inside the timed region, through the public bulk API
    # (CLIP_Encoder.encode_host_batches: pinned uint8 batches in, pinned f32 embeddings out, copies double-buffered)
    for _ in enc.encode_host_batches(pool_host[i % args.pool] for i in range(min(args.warmup, 3))):
        pass
    barrier()
    e0.record()
    n_out = 0
    for res in enc.encode_host_batches(pool_host[i % args.pool] for i in range(args.steps)):
        n_out += res.shape[0]
    e1.record()
    barrier()
    assert n_out == B * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (ms2.item() / 1e3)

    # ---------------- dedup secondary metric (sim-pairs/s), config 4 shape
    dedup = None
    if args.dedup_n > 0:
        from clip_assisted_data_labeling_b200.dedup import duplicate_pairs, duplicate_pairs_distributed
        n_local = args.dedup_n // world
        g = torch.Generator(device="cuda").manual_seed(7 + rank)
        e = torch.nn.functional.normalize(torch.randn(n_local, 768, device="cuda", generator=g), dim=1)
        k = n_local // 100  # plant 1% near-duplicates inside the shard
        src = torch.randint(0, n_local, (k,), device="cuda", generator=g)
        dst = torch.randperm(n_local, device="cuda", generator=g)[:k]
        c = torch.empty(k, device="cuda").uniform_(0.90, 0.999, generator=g)
        e[dst] = torch.nn.functional.normalize(e[src] + (1 / c ** 2 - 1).sqrt()[:, None] * torch.randn(k, 768, device="cuda", generator=g) / 768 ** 0.5, dim=1)
        e16 = e.to(torch.float16)
        fn = (lambda: duplicate_pairs_distributed(e16, 0.96)) if world > 1 else (lambda: duplicate_pairs(e16, 0.96))
        fn()  # warm-up (also sizes the pair buffer)
        barrier()
        t0 = time.perf_counter()
        pairs, _ = fn()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        n_tot = n_local * world
        npairs = n_tot * (n_tot - 1) / 2
        dedup = {"metric": "dedup sim-pairs/sec", "value": npairs / dt.item(), "unit": "pairs/s", "n_embeddings": n_tot, "dim": 768,
                 "threshold": 0.96, "seconds": dt.item(), "pairs_found": int(len(pairs)),
                 "tensor_frac_of_burst": npairs * 2 * 768 / dt.item() / 1e12 / (peaks["tf_burst"] * world),
                 "timing": "host wall clock around the whole call (normalise + all-gather + kernel + D2H + sort), max over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- rank 0: roofline of the dominant kernel, CPU baseline, JSON
    prof = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(prof):
        tr = json.load(open(prof))
        roof["traffic"] = tr.get("dram_bytes_per_launch")
        roof["traffic_detail"] = tr
    F = flops_per_crop(cfg) * 4  # FLOPs per image
    step_tf = value / world * F / 1e12
    cpu_v, cpu_parts = (None, None)
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_images
        cpu_v, cpu_parts = cpu_reference_images_per_s(n_cpu)
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: ViT-L/14 (openai arch, random-init) 4-crop embedding of synthetic 512x512 uint8 images, "
                               "bf16 GEMMs / fp32 residual, per-GPU batch %d images = %d crops per step" % (B, 4 * B),
                   "global_batch": B * world, "per_gpu_batch": B, "image": "512x512x3 uint8", "crops_per_image": 4,
                   "l2": "inputs larger than L2 (%.0f MB of uint8 per step, %d distinct batches cycled)" % (B * IMG_HW * IMG_HW * 3 / 1e6, args.pool),
                   "parallelism": "dp%d (images sharded, no collective on the embedding path)" % world,
                   "lanes": args.lanes,
                   "layernorm": "stand-alone kernels" if args.standalone_layernorm else
                                "fused into the GEMM epilogues (ln_1/ln_2 folded into in_proj/c_fc, residual update + bf16 copy + row statistics in out_proj/c_proj)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * IMG_HW * IMG_HW * 3,
                "d2h_bytes_per_step": B * 4 * cfg["embed"] * 4, "api": "CLIP_Encoder.encode_host_batches (pinned host uint8 batches -> pinned host f32 embeddings; the H2D copy of the next batch and the D2H copy of the results overlap the compute)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "step_roofline": {"bound": "tensor", "achieved": step_tf, "unit": "TFLOP/s", "peak": peaks["tf_sustained"],
                          "frac": step_tf / peaks["tf_sustained"], "flops_per_image": F,
                          "peak_source": peaks["src"] + ", sustained"},
        "cpu_baseline": None if cpu_v is None else {
            "value": cpu_v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d synthetic 512x512 images x 4 crops, PIL crops + torchvision transform + fp32 ViT-L/14 oracle tower (%s)" % (
                args.cpu_images, ", ".join("%s=%.1fs" % kv for kv in cpu_parts.items()))},
        "dedup": dedup,
        "variants": variant,
        "checks": {"unit_norm": norms_ok, "weights": enc.weights_source},
    }
    _emit(args.out_fd, line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to fd 1 when
    NCCL_DEBUG is set in the environment): point fd 1 at stderr for the whole run and keep the real stdout for the line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_fd, line: dict):
    sys.stdout.flush()
    os.write(real_fd, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches cycled through")
    ap.add_argument("--dedup-n", type=int, default=1_000_000, help="embeddings in the dedup measurement (0 = skip)")
    ap.add_argument("--cpu-images", type=int, default=32, help="bounded CPU-baseline sample (~15 s of CPU work on 16 cores)")
    ap.add_argument("--lanes", type=int, default=2, help="independent sub-batches (own stream each) per pass, b2c_vit_set_lanes")
    ap.add_argument("--standalone-layernorm", action="store_true", help="A/B: stand-alone LayerNorm kernels instead of the fused epilogues")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the informational opt-in variants")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.out_fd = _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
