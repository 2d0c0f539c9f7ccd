#!/usr/bin/env python
"""Power / clock / overlap probe for the stages of the ViT-L/14 step (operator-level C-ABI entry points).
For each stage alone (>= 1.5 s sustained) prints ms per launch, median board power and SM clock; then runs the c_fc GEMM
on one stream with LayerNorm (and, separately, attention) on a second stream to see whether they overlap and what the
pair costs.  Answers: is the step energy-bound under the 1000 W cap, and which stages are worth hiding vs removing.
    python tools/power_probe.py [n_crops]
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from clip_assisted_data_labeling_b200 import _lib  # noqa: E402


class Sampler:
    def __init__(self):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.run = False
        self.p, self.c = [], []

    def __enter__(self):
        self.run = True
        self.p, self.c = [], []
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def _loop(self):
        while self.run:
            try:
                self.p.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                self.c.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __exit__(self, *a):
        self.run = False
        self.t.join()

    def med(self):
        # drop the first third (power ramps / averaging window)
        k = len(self.p) // 3
        p, c = sorted(self.p[k:]), sorted(self.c[k:])
        return (p[len(p) // 2] if p else None), (c[len(c) // 2] if c else None)


def sustained(fns, streams, min_s=1.5, reps=10):
    """fns[i] is enqueued `reps` times on streams[i] per round; returns ms per round-of-one-launch-each."""
    for f, s in zip(fns, streams):
        with torch.cuda.stream(s):
            f(s.cuda_stream)
    torch.cuda.synchronize()
    sm = Sampler()
    n, total = 0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sm:
        while total < min_s * 1e3:
            e0.record()
            for s in streams:
                s.wait_stream(torch.cuda.current_stream())
            for _ in range(reps):
                for f, s in zip(fns, streams):
                    f(s.cuda_stream)
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
            n += reps
    p, c = sm.med()
    return {"ms": total / n, "power_w": p, "sm_mhz": c}


def main():
    n_crops = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    d, mlp, T, heads = 1024, 4096, 257, 16
    M = n_crops * T
    dev = "cuda"
    h = (torch.randn(M, d, device=dev) * 0.5).to(torch.bfloat16)
    big = (torch.randn(M, mlp, device=dev) * 0.5).to(torch.bfloat16)
    x = torch.randn(M, d, device=dev)
    x2 = torch.randn(M, d, device=dev)
    h2 = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    qkv2 = (torch.randn(M, 3 * d, device=dev)).to(torch.bfloat16)
    w_fc = (torch.randn(mlp, d, device=dev) * 0.03).to(torch.bfloat16)
    w_pr = (torch.randn(d, mlp, device=dev) * 0.03).to(torch.bfloat16)
    w_qkv = (torch.randn(3 * d, d, device=dev) * 0.03).to(torch.bfloat16)
    w_out = (torch.randn(d, d, device=dev) * 0.03).to(torch.bfloat16)
    b4 = torch.zeros(mlp, device=dev)
    g = torch.ones(d, device=dev)
    lib = _lib.load()
    import ctypes as C

    def gemm(A, W, out, N, K, mode):
        return lambda st: _lib.call("b2c_gemm_bf16", A.data_ptr(), W.data_ptr(), b4.data_ptr(), out.data_ptr(), M, N, K, mode, st)

    def ln(xx, yy):
        return lambda st: _lib.check(lib.b2c_layernorm_bf16(xx.data_ptr(), g.data_ptr(), g.data_ptr(), yy.data_ptr(), M, d,
                                                            C.c_float(1e-5), C.c_void_p(st)), "ln")

    def attn(q, o):
        return lambda st: _lib.check(lib.b2c_attention_bf16(q.data_ptr(), o.data_ptr(), n_crops, T, heads, 64, C.c_void_p(st)), "attn")

    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    stages = {
        "c_fc": gemm(h, w_fc, big, mlp, d, _lib.EPI_BIAS_QGELU_BF16),
        "c_proj": gemm(big, w_pr, x, d, mlp, _lib.EPI_BIAS_RESID_F32),
        "in_proj": gemm(h, w_qkv, big, 3 * d, d, _lib.EPI_BIAS_BF16),
        "out_proj": gemm(h, w_out, x, d, d, _lib.EPI_BIAS_RESID_F32),
        "layernorm": ln(x2, h2),
        "attention": attn(qkv2, h2),
    }
    res = {"n_crops": n_crops}
    for name, f in stages.items():
        res[name] = sustained([f], [s0])
        print(name, res[name], file=sys.stderr, flush=True)
    idle = Sampler()
    with idle:
        time.sleep(1.0)
    res["idle_power_w"] = idle.med()[0]
    res["c_fc || layernorm"] = sustained([stages["c_fc"], stages["layernorm"]], [s0, s1])
    res["c_fc || attention"] = sustained([stages["c_fc"], stages["attention"]], [s0, s1])
    res["c_fc ; layernorm (one stream)"] = sustained([lambda st: (stages["c_fc"](st), stages["layernorm"](st))], [s0])
    for k, v in res.items():
        if isinstance(v, dict):
            v["joules_per_launch"] = None if v["power_w"] is None else v["power_w"] * v["ms"] / 1e3
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
