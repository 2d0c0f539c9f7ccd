#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code path on the host cores

Headline (configs[1]): a "step" is one pass of the hot path over one batch of synthetic input: B uint8 512x512 RGB
images -> 4 crops (centre / padded / subcrop1 / subcrop2) -> PIL-exact resize + normalise -> ViT-L/14 (random-init
weights of the openai architecture, bf16 tensor-core GEMMs, fp32 residual stream) -> f32 [B,4,768] unit-norm embeddings.
`value` is timed with the batch already in HBM; `e2e` goes through the public API (CLIP_Encoder.encode_host_batches) from
pinned host memory with the H2D copy of the images and the D2H read of the embeddings inside the timed region.

Beside the headline, the same JSON line carries every other BASELINE config at this N, each with its parity:
  parity     8 crops of the LAST timed 1024-crop pass against the fp32 CPU oracle tower (cos >= 0.9995, max-abs <= 2e-3);
  config5    ViT-H/14 (LAION arch) 4-crop + SimpleFC(4096,[264,128,64],1) scoring, images/s, roofline, e2e, parity;
  l14_336    ViT-L/14-336 (the reference's CLI default model), images/s, roofline, parity;
  dedup      1 M x 768 duplicate search (configs[3]): the SAME embedding set for every N (global seed, sliced per rank),
             planted pairs straddling the threshold incl. cross-shard ones, pairs_expected vs pairs_found, tensor
             roofline of the kernel, and e2e from a packed store on disk;
  cpu_baseline / gpu_library_bar (N = 1): the reference's own embed loop (verbatim copy under baseline/_ref, open_clip
             supplied by oracle/reference_shim.py) on the host cores, and the same code in its native CUDA fp16 mode.
One JSON line on stdout (rank 0).
"""
import argparse
import contextlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (4-crop ViT-L/14 embed) at 1/2/4/8 B200; dedup sim-pairs/sec"
MODEL = "ViT-L-14/openai"
MODEL_H = "ViT-H-14/laion2b_s32b_b79k"
MODEL_336 = "ViT-L-14-336/openai"
IMG_HW = 512
COS_MIN, MAX_ABS = 0.9995, 2e-3  # north_star's tolerance for the bf16 tower
WORKLOAD = ("configs[1]: ViT-L/14 (openai arch, random-init) 4-crop embedding of synthetic 512x512 uint8 images, "
            "bf16 GEMMs / fp32 residual")


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


def synth_batch(B, seed, device="cpu", hw=IMG_HW):
    """Synthetic hw x hw uint8 images: low-frequency colour field + noise (cheap torch version of SURVEY §8d).  The
    B200 arm generates its pool on the device (eight ranks sharing 16 host cores would spend a minute here otherwise);
    the CPU arms generate theirs on the host.  Same distribution either way."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    ar = torch.arange(hw, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(ar, ar, indexing="ij")
    f = torch.rand(B, 3, 2, generator=g, device=device) * 3.5 + 0.5
    ph = torch.rand(B, 3, 2, generator=g, device=device) * 6.2832
    img = 128 + 90 * torch.sin(6.2832 * f[..., 0, None, None] * xx / hw + ph[..., 0, None, None]) * \
        torch.cos(6.2832 * f[..., 1, None, None] * yy / hw + ph[..., 1, None, None])
    img = img + 20 * torch.randn(B, 3, hw, hw, generator=g, device=device)
    return img.clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()  # [B,H,W,3]


def device_state_dict(cfg, seed):
    """Seeded random weights of the architecture, generated ON THE DEVICE (the CPU initialiser of vit_arch takes seconds
    per tower and every rank would run it on the shared host cores).  Same scales as vit_arch.random_state_dict."""
    import math
    import torch
    from clip_assisted_data_labeling_b200.vit_arch import state_dict_shapes
    g = torch.Generator(device="cuda").manual_seed(seed)
    d = cfg["width"]
    out = {}
    for k, shape in state_dict_shapes(cfg).items():
        r = torch.randn(shape, generator=g, device="cuda")
        if k.endswith("ln_1.weight") or k.endswith("ln_2.weight") or k in ("ln_pre.weight", "ln_post.weight"):
            t = 1.0 + 0.1 * r
        elif k.endswith("bias"):
            t = 0.05 * r
        elif k in ("class_embedding", "positional_embedding", "proj"):
            t = d ** -0.5 * r
        else:
            t = math.prod(shape[1:]) ** -0.5 * r
        out[k] = t
    return out


# --------------------------------------------------------------------------------------------- reference runner (subprocess)
def run_reference_runner(extra, timeout=900):
    """oracle/reference_runner.py in a subprocess (it must decide whether CUDA is visible before torch is imported)."""
    cmd = [sys.executable, "-m", "oracle.reference_runner"] + [str(x) for x in extra]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": "reference runner failed: " + (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"reference runner did not finish: {e}"}


def reference_sample_text(model, n, batch, workers, cores):
    return (f"{n} synthetic 512x512 PNG files x 4 crops per step through the reference's own Feature_Dataset('{model}', batch_size={batch}, "
            f"num_workers={workers}).process() (_1_embed_with_CLIP.py:36-184 + utils/embedder.py + utils/image_features.py, verbatim "
            f"copy; open_clip = oracle/reference_shim.py: random-init tower + open_clip val transform), torch CPU fp32, {cores} threads; "
            "DataLoader workers are forked (the CLI's spawn start-up is not part of a bounded sample)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count()
    n = max(4, args.ref_images)
    res = run_reference_runner(["--mode", "loop", "--model", MODEL, "--device", "cpu", "--images", n, "--steps", args.steps,
                                "--warmup", args.warmup, "--batch", 8, "--workers", 4])
    if "unavailable" in res:
        _emit(args.out_fd, {"impl": "reference", "unavailable": res["unavailable"]})
        return
    dt = sum(res["step_s"])
    v = n * len(res["step_s"]) / dt
    # configs[0]: the reference's own CPU-runnable case (ViT-B/32, batch 8, 4 workers), bounded sample of its 1k images
    c1 = run_reference_runner(["--mode", "loop", "--model", "ViT-B-32/openai", "--device", "cpu", "--images", args.config1_images,
                               "--steps", 1, "--warmup", 0, "--batch", 8, "--workers", 4])
    config1 = c1 if "unavailable" in c1 else {
        "workload": "configs[0]: ViT-B/32 (openai arch, random-init) embedding of synthetic 512x512 images x 4 crops, batch 8, 4 workers, on CPU "
                    "via the reference _1_embed_with_CLIP.py path",
        "value": args.config1_images / c1["step_s"][0], "unit": "images/s", "cores": c1["cores"],
        "sample": f"{args.config1_images} of the config's 1000 images (whole process() loop incl. ImageFeaturizer and .pt writes)"}
    sample = reference_sample_text(MODEL, n, 8, 4, res["cores"])
    _emit(args.out_fd, {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dt / len(res["step_s"]), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": n,
                   "note": "bounded sample of the same workload per step (the B200 arm's 256-image step would take ~2 min per step here); "
                           "the reference's code is run as is, only open_clip (not installable offline) is the repo's shim",
                   "reference_root": res.get("reference_root")},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": res["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "config1": config1,
    })


# --------------------------------------------------------------------------------------------- B200 arm helpers
GEMM_STAGES = ("in_proj", "out_proj", "c_fc", "c_proj")


def stage_profile(step_fn, cfg, n_crops, peaks, steps=2):
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
        step_fn(i)
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


def oracle_from_state_dict(arch, act, sd):
    """fp32 CPU oracle tower carrying exactly the weights the CUDA tower was given (rank 0 only)."""
    import torch
    from oracle import vit_oracle
    with torch.device("meta"):
        m = vit_oracle.VisionTransformer(act=act, **vit_oracle.ARCHS[arch])
    m.load_state_dict({k: v.detach().float().cpu() for k, v in sd.items()}, assign=True)
    return m.eval()


def tower_parity(enc, oracle, images_u8, got, n_images):
    """`got` = f32 [B,4,E] the timed pass produced for `images_u8`; the first n_images images (4 crops each) are pushed
    through the reference pipeline on the CPU: K0's bit-exact f32 crops -> fp32 oracle tower -> L2 normalise."""
    import torch
    from clip_assisted_data_labeling_b200.vit import preprocess_u8
    from oracle import vit_oracle
    px = preprocess_u8(images_u8[:n_images], enc.img_resolution, enc.model.cfg["patch"], "nchw").cpu()
    t0 = time.perf_counter()
    with all_host_cores():
        ref = vit_oracle.encode_image_oracle(oracle, px.view(-1, 3, enc.img_resolution, enc.img_resolution))
    g = got[:n_images].reshape(-1, got.shape[-1]).float().cpu()
    cos = float(torch.nn.functional.cosine_similarity(ref, g, dim=-1).min())
    mx = float((ref - g).abs().max())
    return {"crops": 4 * n_images, "min_cos": cos, "max_abs": mx, "tol": {"min_cos": COS_MIN, "max_abs": MAX_ABS},
            "pass": bool(cos >= COS_MIN and mx <= MAX_ABS), "oracle": "oracle/vit_oracle.py fp32 on the host (rank 0), same weights, "
            "crops = K0's f32 output for the same images", "oracle_seconds": time.perf_counter() - t0}


def timed_steps(step_fn, steps, warmup, barrier, world, dist):
    """W untimed + K timed steps bracketed by barrier + synchronize, CUDA events, max over ranks.  Returns
    (ms_total max over ranks, this rank's ms_total, last output)."""
    import torch
    out = None
    for i in range(warmup):
        out = step_fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = step_fn(i)
    e1.record()
    barrier()
    mine = e0.elapsed_time(e1)
    ms = torch.tensor([mine], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), mine, out


def e2e_host_batches(enc, pool_host, steps, warmup, barrier, world, dist, post=None):
    """Same metric through the public bulk API from pinned host memory: H2D of every batch and D2H of its result inside
    the timed region.  `post` (optional) consumes each yielded [B,4,E] host block (config 5: nothing to do, scores are
    produced on the device by the same call chain)."""
    import torch
    for _ in enc.encode_host_batches(pool_host[i % len(pool_host)] for i in range(min(warmup, 3))):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_out = 0
    for res in enc.encode_host_batches(pool_host[i % len(pool_host)] for i in range(steps)):
        n_out += res.shape[0]
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), n_out


_ALL_CORES = None


@contextlib.contextmanager
def all_host_cores():
    """The CPU-side checks of rank 0 (fp32 oracle tower) use every host core, not the slice the rank's launch thread is
    pinned to (the other ranks are idle at a barrier meanwhile)."""
    import torch
    if _ALL_CORES is None:
        yield
        return
    mine = os.sched_getaffinity(0)
    nthr = torch.get_num_threads()
    try:
        os.sched_setaffinity(0, _ALL_CORES)
        torch.set_num_threads(len(_ALL_CORES))
        yield
    finally:
        torch.set_num_threads(nthr)
        os.sched_setaffinity(0, mine)


def pin_to_local_cores(local, n_local):
    """One rank per GPU shares the host cores with its siblings: give every rank its own contiguous slice so the launch
    threads do not migrate across each other (the reference has no multi-process path; this is launch hygiene only)."""
    global _ALL_CORES
    try:
        cores = sorted(os.sched_getaffinity(0))
        _ALL_CORES = set(cores)
        per = max(1, len(cores) // max(1, n_local))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:  # noqa: BLE001
        return None


# --------------------------------------------------------------------------------------------- dedup (configs[3])
def dedup_dataset(n_total, dim, seed=7, n_planted=20000):
    """The SAME n_total x dim set on every rank and for every N: unit-norm Gaussian rows from one seeded device
    generator, then n_planted rows overwritten by noisy copies of other rows with target cosine U[0.90, 0.999] (so the
    similarities straddle 0.96).  Sources and copies are disjoint row sets drawn over the WHOLE range, so most planted
    pairs cross shard boundaries at N > 1.  Returns (f16 [n_total, dim], src, dst, fp32 cosine of each planted pair)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    e = torch.nn.functional.normalize(torch.randn(n_total, dim, device="cuda", generator=g), dim=1)
    perm = torch.randperm(n_total, device="cuda", generator=g)
    src, dst = perm[:n_planted], perm[n_planted:2 * n_planted]
    c = torch.empty(n_planted, device="cuda").uniform_(0.90, 0.999, generator=g)
    noise = torch.randn(n_planted, dim, device="cuda", generator=g) / dim ** 0.5
    e[dst] = torch.nn.functional.normalize(e[src] + (1 / c ** 2 - 1).sqrt()[:, None] * noise, dim=1)
    e16 = e.to(torch.float16)
    en = torch.nn.functional.normalize(e16.float(), dim=1)  # what the search sees: fp16 rows, re-normalised (_2:38,67)
    cos = (en[src] * en[dst]).sum(1)
    return e16, src, dst, cos


def check_pairs(pairs, src, dst, cos, thr, band=1e-3):
    """north_star: the pair set must be identical except for pairs whose similarity lies within `band` of the threshold."""
    lo = [min(a, b) for a, b in zip(src, dst)]
    hi = [max(a, b) for a, b in zip(src, dst)]
    must = {(a, b) for a, b, s in zip(lo, hi, cos) if s > thr + band}
    may = {(a, b) for a, b, s in zip(lo, hi, cos) if s > thr - band}
    got = set(map(tuple, pairs.tolist()))
    return {"pairs_expected": len(must), "pairs_in_band": len(may) - len(must), "pairs_found": len(got),
            "missing": len(must - got), "unexpected": len(got - may), "pass": bool(not (must - got) and not (got - may)),
            "rule": "every planted pair with fp32 cosine > thr + 1e-3 is found, nothing with cosine < thr - 1e-3 is (random rows of "
                    "dimension 768 never come near 0.96)"}


def run_dedup(args, world, rank, barrier, dist, peaks):
    import numpy as np
    import torch
    from clip_assisted_data_labeling_b200 import _lib
    from clip_assisted_data_labeling_b200.dedup import (duplicate_pairs, duplicate_pairs_distributed, find_near_duplicates_in_store,
                                                         find_near_duplicates_in_store_distributed)
    n_tot = args.dedup_n // (1000 * world) * (1000 * world)
    n_local = n_tot // world
    thr = 0.96
    e16, src, dst, cos = dedup_dataset(n_tot, 768, n_planted=max(8, n_tot // 50))
    shard = e16[rank * n_local:(rank + 1) * n_local].contiguous()
    src_h, dst_h, cos_h = src.cpu().tolist(), dst.cpu().tolist(), cos.cpu().tolist()
    cross = sum(1 for a, b in zip(src_h, dst_h) if a // n_local != b // n_local)
    fn = (lambda: duplicate_pairs_distributed(shard, thr)) if world > 1 else (lambda: duplicate_pairs(e16, thr))
    fn()  # warm-up (NCCL channels for these shapes, pair-buffer sizing)
    barrier()
    l0 = _lib.launch_count()
    _lib.prof_enable(True)
    t0 = time.perf_counter()
    pairs, _ = fn()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    rec = _lib.prof_read()
    _lib.prof_enable(False)
    launches = _lib.launch_count() - l0
    kern_ms = torch.tensor([rec.get("dedup", (0.0, 0))[0]], device="cuda")
    kern_all = [kern_ms.item()]
    if world > 1:
        ka = torch.empty(world, device="cuda")
        dist.all_gather_into_tensor(ka, kern_ms.float())
        kern_all = ka.cpu().tolist()
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    npairs = n_tot * (n_tot - 1) / 2
    flops = 2.0 * 768 * npairs
    out = {"metric": "dedup sim-pairs/sec", "workload": "configs[3]: cosine-similarity duplicate search over %d ViT-L/14-sized embeddings "
           "(768-d) at threshold 0.96" % n_tot, "value": npairs / dt.item(), "unit": "pairs/s", "n_embeddings": n_tot, "dim": 768,
           "threshold": thr, "seconds": dt.item(), "gpu_launches": int(launches),
           "data": "the same seeded set for every N (sliced per rank); %d planted near-duplicate pairs, %d of them across shards" % (
               len(src_h), cross),
           "timing": "host wall clock around the whole call (normalise + all-gather + kernels + count/pair exchange + D2H + sort), max over ranks",
           "parallelism": "own-shard block under the all-gather, then the ranks draw the remaining bands (largest first) from a shared "
                          "counter, so the work follows each GPU's power-capped speed; one all-gather of the shards + two fixed-shape gathers of counts / pairs" if world > 1 else "single GPU, bands of 2048 rows"}
    if rank == 0:
        out["parity"] = check_pairs(pairs, src_h, dst_h, cos_h, thr)
        k_s = kern_ms.item() / 1e3
        out["roofline"] = {"bound": "tensor", "achieved": flops / world / k_s / 1e12 if k_s > 0 else None, "peak": peaks["tf_burst"],
                           "unit": "TFLOP/s", "frac": (flops / world / k_s / 1e12 / peaks["tf_burst"]) if k_s > 0 else None,
                           "kernel": "umma2_tile_kernel<DedupPolicy> (fp16 x fp16 -> fp32, threshold + pair emission in the epilogue)",
                           "kernel_seconds": k_s, "kernel_seconds_per_rank": [round(v / 1e3, 5) for v in kern_all],
                           "algorithmic_flops": flops,
                           "how": "CUDA events around this rank's band launches (library stage timer, max over ranks); 2*E FLOP per "
                                  "unordered pair, per-GPU share; peak = burst (kernel timed alone)", "traffic": None}
    # ---- e2e: the packed store on disk -> pair list of paths on the host
    tmp = None
    try:
        from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter
        base = os.environ.get("B2C_BENCH_TMP") or tempfile.gettempdir()
        tmp = os.path.join(base, "b2c_bench_store")
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
            os.makedirs(tmp)
        barrier()
        with PackedWriter(tmp, MODEL, 768, ["square_padded_crop"], shard=rank, dtype="float16") as w:
            blk = 50000
            for b0 in range(0, n_local, blk):
                rows = shard[b0:b0 + blk].cpu().numpy()
                feats = np.zeros((rows.shape[0], 4, 768), np.float16)
                feats[:, 1] = rows
                w.append(feats, [f"/data/d/{rank * n_local + b0 + i:08d}.jpg" for i in range(rows.shape[0])])
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            res = find_near_duplicates_in_store_distributed(tmp, thr, model_name=MODEL)
            got_paths = res[0]
        else:
            res = find_near_duplicates_in_store(PackedStore(tmp, MODEL), thr, per_directory=False)
            got_paths = res[0][0]
        barrier()
        de = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(de, op=dist.ReduceOp.MAX)
        if rank == 0:
            idx = np.asarray([[int(os.path.basename(a)[:8]), int(os.path.basename(b)[:8])] for a, b in got_paths], np.int64).reshape(-1, 2)
            out["e2e"] = {"value": npairs / de.item(), "unit": "pairs/s", "seconds": de.item(),
                          "h2d_bytes_per_step": n_tot * 768 * 2, "d2h_bytes_per_step": int(len(got_paths)) * 12,
                          "api": "find_near_duplicates_in_store%s: packed fp16 store on disk (one shard per rank, page cache warm) -> "
                                 "mmap -> %s -> (path_i, path_j, sim) lists on the host" % (
                                     "_distributed" if world > 1 else "",
                                     "H2D of the rank's shard -> all-gather -> search" if world > 1 else
                                     "chunks through pinned memory, H2D and search of each arriving column block overlapped"),
                          "same_pairs_as_device_run": bool(np.array_equal(idx, pairs))}
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            out["e2e"] = {"unavailable": repr(e)[:300]}
    finally:
        barrier()
        if rank == 0 and tmp:
            shutil.rmtree(tmp, ignore_errors=True)
    del e16, shard
    torch.cuda.empty_cache()
    return out if rank == 0 else None


# --------------------------------------------------------------------------------------------- secondary embedding configs
def run_secondary(name, model_name, batch, args, world, rank, barrier, dist, peaks, with_fc):
    """config5 / l14_336: device-resident images/s (aggregate over ranks), step-level tensor roofline, stage shares,
    e2e from pinned host memory, parity of the last timed pass against the oracle (rank 0)."""
    import torch
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    from clip_assisted_data_labeling_b200.scorer import FCScorer, SimpleFC
    from clip_assisted_data_labeling_b200.vit_arch import ARCHS, activation_for, flops_per_crop
    arch, pretrained = model_name.split("/")
    cfg = ARCHS[arch]
    sd = device_state_dict(cfg, seed=1)
    with contextlib.redirect_stdout(sys.stderr):
        enc = CLIP_Encoder(model_name, device="cuda", state_dict=sd)
    enc.model.set_lanes(args.lanes)
    scorer, fc = None, None
    if with_fc:
        torch.manual_seed(0)
        fc = SimpleFC(4 * cfg["embed"], [264, 128, 64], 1, [model_name]).eval()
        scorer = FCScorer(fc, "cuda")
    pool = [synth_batch(batch, 1000 * rank + 7 * i + 1, device="cuda") for i in range(3)]
    state = {}

    def step(i):
        emb = enc.encode_images_u8(pool[i % 3])
        if scorer is not None:
            state["scores"] = scorer.score_embeddings(emb)
        state["i"] = i % 3
        return emb

    steps = max(2, args.steps // 2)
    ms_total, _, out = timed_steps(step, steps, 3, barrier, world, dist)
    value = world * batch * steps / (ms_total / 1e3)
    F = flops_per_crop(cfg) * 4
    res = {"model": model_name, "value": value, "unit": "images/s", "ms_per_step": ms_total / steps, "steps": steps, "per_gpu_batch": batch,
           "crops_per_step_per_gpu": 4 * batch, "dtype": "bf16", "fc_scoring": bool(with_fc),
           "step_roofline": {"bound": "tensor", "achieved": value / world * F / 1e12, "unit": "TFLOP/s", "peak": peaks["tf_sustained"],
                             "frac": value / world * F / 1e12 / peaks["tf_sustained"], "flops_per_image": F}}
    if rank == 0:
        roof = stage_profile(step, cfg, 4 * batch, peaks, steps=1)
        res["roofline"] = {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "frac_of_burst", "gemm_share_of_step")}
        res["stage_ms_per_step"] = {k: round(v["ms_per_step"], 3) for k, v in roof["stage_shares"].items()}
    # e2e from pinned host memory
    pool_host = [b.cpu().pin_memory() for b in pool[:2]]
    if with_fc:
        # config 5 through the public call a user of _1 + _5 would make: per step H2D of the images, embed_and_score, D2H of the
        # embeddings (what _1 saves) and of the scores (what _5 writes)
        from clip_assisted_data_labeling_b200.scorer import embed_and_score
        emb_h = torch.empty(batch, 4, cfg["embed"], dtype=torch.float32, pin_memory=True)
        sc_h = torch.empty(batch, 1, dtype=torch.float32, pin_memory=True)

        def e2e_step(i):
            emb, sc = embed_and_score(enc, scorer, pool_host[i % 2].to("cuda", non_blocking=True))
            emb_h.copy_(emb, non_blocking=True)
            sc_h.copy_(sc, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller reads the scores before the next batch
            return emb
        ms2, _, _ = timed_steps(e2e_step, steps, 2, barrier, world, dist)
        api, d2h = "scorer.embed_and_score per pinned host batch (H2D, 4-crop embed + FC on the device, D2H of embeddings and scores, synchronised per step)", batch * (4 * cfg["embed"] + 1) * 4
    else:
        ms2, n_out = e2e_host_batches(enc, pool_host, steps, 2, barrier, world, dist)
        api, d2h = "CLIP_Encoder.encode_host_batches", batch * 4 * cfg["embed"] * 4
    res["e2e"] = {"value": world * batch * steps / (ms2 / 1e3), "unit": "images/s", "h2d_bytes_per_step": batch * IMG_HW * IMG_HW * 3,
                  "d2h_bytes_per_step": d2h, "api": api}
    if rank == 0:
        n_par = 2 if cfg["image"] <= 224 else 1
        last = step(state["i"])  # the batch of the last timed step again (bit-identical pass), scores included
        oracle = oracle_from_state_dict(arch, activation_for(pretrained), sd)
        par = tower_parity(enc, oracle, pool[state["i"]], last, n_par)
        if with_fc:
            from oracle.mlp_oracle import simple_fc_forward
            lin = [m for m in fc.layers if isinstance(m, torch.nn.Linear)]
            feats = last[:n_par].reshape(n_par, -1).float().cpu().numpy()
            want = simple_fc_forward(feats, [l.weight.detach().numpy() for l in lin], [l.bias.detach().numpy() for l in lin])
            got = state["scores"][:n_par].cpu().numpy()
            par["fc_max_abs"] = float(abs(want - got).max())
            par["fc_tol"] = 3e-6
            par["pass"] = bool(par["pass"] and par["fc_max_abs"] <= 3e-6)
        res["parity"] = par
    del enc, pool, sd
    torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from clip_assisted_data_labeling_b200 import _lib
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    from clip_assisted_data_labeling_b200.vit_arch import ARCHS, flops_per_crop

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cores_mine = pin_to_local_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if world > 1 else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    B = args.batch
    cfg = ARCHS["ViT-L-14"]

    sd = device_state_dict(cfg, seed=0)
    with contextlib.redirect_stdout(sys.stderr):  # stdout carries exactly one JSON line
        enc = CLIP_Encoder(MODEL, device="cuda", state_dict=sd)
    enc.weights_source = "random-init on the device (bench.device_state_dict, seed 0)"
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
    sampler = ClockSampler(local)
    for i in range(args.warmup):
        enc.encode_images_u8(pool_dev[i % args.pool])
    barrier()
    sampler.start()
    l0 = _lib.launch_count()
    ms_total, ms_mine, out = timed_steps(lambda i: enc.encode_images_u8(pool_dev[i % args.pool]), args.steps, 0, barrier, world, dist)
    launches = _lib.launch_count() - l0
    value = world * B * args.steps / (ms_total / 1e3)
    norms_ok = bool(torch.allclose(out.norm(dim=-1), torch.ones_like(out[..., 0]), atol=1e-4))
    last_idx = (args.steps - 1) % args.pool

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
    roof = stage_profile(lambda i: enc.encode_images_u8(pool_dev[i % args.pool]), cfg, 4 * B, peaks) if rank == 0 else None

    # ---------------- e2e: host buffers, H2D + D2H inside the timed region, through the public bulk API
    ms2, n_out = e2e_host_batches(enc, pool_host, args.steps, args.warmup, barrier, world, dist)
    assert n_out == B * args.steps
    clocks = sampler.stop()
    e2e_value = world * B * args.steps / (ms2 / 1e3)

    # ---------------- per-rank view of the timed region (N > 1): who was the slowest rank, and at what clock
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_mine / args.steps, float(clocks.get("sm_mhz") or 0.0), float(clocks.get("power_w_max") or 0.0)], device="cuda")
        allr = torch.empty(world, 3, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().tolist()
        ms_r = [r[0] for r in allr]
        per_rank = {"ms_per_step": {"min": min(ms_r), "median": statistics.median(ms_r), "max": max(ms_r), "all": [round(x, 3) for x in ms_r]},
                    "sm_mhz_median": [r[1] for r in allr], "power_w_max": [r[2] for r in allr],
                    "host_cores_of_rank0": cores_mine,
                    "note": "value uses the MAX over ranks (slowest rank); every rank runs the same work on its own images, no collective "
                            "inside the timed region, so max/min is the spread of the GPUs' power-capped clocks plus host launch jitter"}

    # ---------------- parity of the LAST timed pass against the oracle (rank 0): 8 of its 1024 crops
    parity = None
    if rank == 0 and not args.no_parity:
        oracle = oracle_from_state_dict("ViT-L-14", "quick_gelu", sd)
        parity = tower_parity(enc, oracle, pool_dev[last_idx], out, 2)
        parity["from"] = "the last timed %d-crop pass of `value` (lanes=%d, fused LayerNorm=%s)" % (4 * B, args.lanes, not args.standalone_layernorm)
        del oracle
    weights_source = enc.weights_source
    del enc, pool_dev, pool_host, out
    torch.cuda.empty_cache()
    barrier()

    # ---------------- the other embedding configs at this N
    config5 = l336 = None
    if not args.no_secondary:
        config5 = run_secondary("config5", MODEL_H, args.batch_h, args, world, rank, barrier, dist, peaks, with_fc=True)
        config5["workload"] = "configs[4]: ViT-H/14 (LAION arch, exact GELU, hd 80) 4-crop embedding + SimpleFC(4096,[264,128,64],1) regressor " \
                              "scoring (_5_predict_labels path) on %d B200" % world
        barrier()
        l336 = run_secondary("l14_336", MODEL_336, args.batch_336, args, world, rank, barrier, dist, peaks, with_fc=False)
        l336["workload"] = "ViT-L/14-336 (openai arch; the reference's CLI default, _1_embed_with_CLIP.py:190) 4-crop embedding, T = 577"
        barrier()

    # ---------------- dedup (configs[3])
    dedup = run_dedup(args, world, rank, barrier, dist, peaks) if args.dedup_n > 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- rank 0: roofline of the dominant kernel, CPU baseline, library bar, JSON
    prof = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(prof):
        tr = json.load(open(prof))
        roof["traffic"] = tr.get("dram_bytes_per_launch")
        roof["traffic_detail"] = tr
    F = flops_per_crop(cfg) * 4  # FLOPs per image
    step_tf = value / world * F / 1e12
    cpu_base = lib_bar = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_images
        r = run_reference_runner(["--mode", "loop", "--model", MODEL, "--device", "cpu", "--images", n_cpu, "--steps", 1, "--warmup", 0,
                                  "--batch", 8, "--workers", 4])
        if "unavailable" in r:
            cpu_base = {"unavailable": r["unavailable"]}
        else:
            cpu_base = {"value": n_cpu / r["step_s"][0], "unit": "images/s", "cores": r["cores"], "kind": "reference",
                        "sample": reference_sample_text(MODEL, n_cpu, 8, 4, r["cores"]) + " (%.1f s)" % r["step_s"][0]}
    if world == 1 and not args.no_library_bar:
        loop = run_reference_runner(["--mode", "loop", "--model", MODEL, "--device", "cuda", "--images", 64, "--steps", 1, "--warmup", 1,
                                     "--batch", 8, "--workers", 4])
        e8 = run_reference_runner(["--mode", "encode", "--model", MODEL, "--device", "cuda", "--images", 8, "--steps", 10, "--warmup", 3])
        e256 = run_reference_runner(["--mode", "encode", "--model", MODEL, "--device", "cuda", "--images", 64, "--steps", 5, "--warmup", 2])
        lib_bar = {"what": "INFORMATIONAL: the reference's own code in its native CUDA mode on this B200 (precision 'fp16', eager torch: cuBLAS + "
                           "nn.MultiheadAttention library kernels; open_clip = the repo's shim) — what a user of the reference gets on this box",
                   "process_loop": loop if "unavailable" in loop else {
                       "value": 64 / loop["step_s"][0], "unit": "images/s", "note": "Feature_Dataset(batch 8, 4 workers).process() on 64 PNG "
                       "files: host-bound (PIL crops + ImageFeaturizer in 4 workers, two torch.load + one torch.save per image)"},
                   "encode_image_batch8": e8 if "unavailable" in e8 else {
                       "value": 8 * len(e8["step_s"]) / sum(e8["step_s"]), "unit": "images/s", "crops_per_call": 32, "precision": e8.get("precision")},
                   "encode_image_256crops": e256 if "unavailable" in e256 else {
                       "value": 64 * len(e256["step_s"]) / sum(e256["step_s"]), "unit": "images/s", "crops_per_call": 256,
                       "precision": e256.get("precision"), "note": "forward only on resident fp32 crops, the most favourable use of the reference tower"}}
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD + ", per-GPU batch %d images = %d crops per step" % (B, 4 * B),
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
        "parity": parity,
        "per_rank": per_rank,
        "cpu_baseline": cpu_base,
        "gpu_library_bar": lib_bar,
        "config5": config5,
        "l14_336": l336,
        "dedup": dedup,
        "variants": variant,
        "checks": {"unit_norm": norms_ok, "weights": weights_source,
                   "parity_pass": {"headline": (parity or {}).get("pass"), "config5": ((config5 or {}).get("parity") or {}).get("pass"),
                                   "l14_336": ((l336 or {}).get("parity") or {}).get("pass"),
                                   "dedup": ((dedup or {}).get("parity") or {}).get("pass")}},
    }
    _emit(args.out_fd, line)
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (headline, ViT-L/14)")
    ap.add_argument("--batch-h", type=int, default=128, help="images per GPU per step of the config-5 block (ViT-H/14)")
    ap.add_argument("--batch-336", type=int, default=64, help="images per GPU per step of the ViT-L/14-336 block")
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches cycled through")
    ap.add_argument("--dedup-n", type=int, default=1_000_000, help="embeddings in the dedup measurement (0 = skip)")
    ap.add_argument("--cpu-images", type=int, default=16, help="bounded CPU-baseline sample (~10-20 s of CPU work on 16 cores)")
    ap.add_argument("--ref-images", type=int, default=16, help="--impl reference: images per step (bounded sample)")
    ap.add_argument("--config1-images", type=int, default=64, help="--impl reference: images of the configs[0] (ViT-B/32) sample")
    ap.add_argument("--lanes", type=int, default=2, help="independent sub-batches (own stream each) per pass, b2c_vit_set_lanes")
    ap.add_argument("--standalone-layernorm", action="store_true", help="A/B: stand-alone LayerNorm kernels instead of the fused epilogues")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-bar", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config5 / l14_336 blocks")
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
