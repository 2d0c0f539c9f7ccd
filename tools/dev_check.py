"""Developer smoke checks run on a GPU box: each case runs in its own subprocess under a timeout so a
trap or hang in one kernel cannot take the others (or the box) down.  Not part of the test suite.

    python tools/dev_check.py [case ...]      # parent: runs every case (or the named ones)
    python tools/dev_check.py --child <case>  # child: one case
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lib():
    L = C.CDLL(os.path.join(ROOT, "clip_assisted_data_labeling_b200", "libb2c.so"))
    L.b2c_last_error.restype = C.c_char_p
    return L


def _gemm_case(M, N, K, mode, seed=0, time_it=False):
    import torch
    L = lib()
    torch.manual_seed(seed)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    ref = A.float() @ W.float().t() + bias
    if mode == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif mode == 2:
        ref = torch.nn.functional.gelu(ref)
    if mode == 3:
        out = torch.randn(M, N, device=dev)
        ref = ref + out
    else:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.b2c_gemm_bf16(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()),
                         C.c_void_p(out.data_ptr()), C.c_int64(M), C.c_int(N), C.c_int(K), C.c_int(mode),
                         C.c_void_p(st))
    assert rc == 0, L.b2c_last_error()
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    res = {"M": M, "N": N, "K": K, "mode": mode, "max_abs_err": err, "ref_max": scale}
    if time_it:
        tmp = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            L.b2c_gemm_bf16(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()),
                            C.c_void_p(tmp.data_ptr()), C.c_int64(M), C.c_int(N), C.c_int(K), C.c_int(0), C.c_void_p(st))
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        reps = 10
        for _ in range(reps):
            L.b2c_gemm_bf16(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()),
                            C.c_void_p(tmp.data_ptr()), C.c_int64(M), C.c_int(N), C.c_int(K), C.c_int(0), C.c_void_p(st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        # cuBLAS for comparison
        for _ in range(3):
            torch.matmul(A, W.t())
        e0.record()
        for _ in range(reps):
            torch.matmul(A, W.t())
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / reps) / 1e9
    tol = 0.02 * max(scale, 1.0) if mode != 3 else 2e-3 * max(scale, 1.0)
    res["ok"] = bool(err <= tol)
    return res


def case_gemm_tiny():
    return [_gemm_case(128, 256, 64, 3), _gemm_case(128, 256, 256, 3), _gemm_case(100, 256, 128, 0)]


def case_gemm_shapes():
    out = []
    for (M, N, K, mode) in [(257 * 4, 1024, 1024, 3), (257 * 8, 3072, 1024, 0), (257 * 8, 4096, 1024, 1),
                            (257 * 8, 1024, 4096, 3), (50 * 32, 768, 3072, 2), (257 * 3, 1280, 5120, 3),
                            (128 * 148 * 2 + 77, 1024, 1024, 0)]:
        out.append(_gemm_case(M, N, K, mode))
    return out


def case_gemm_perf():
    return [_gemm_case(257 * 512, 1024, 1024, 0, time_it=True), _gemm_case(257 * 512, 4096, 1024, 0, time_it=True),
            _gemm_case(257 * 512, 1024, 4096, 0, time_it=True), _gemm_case(257 * 512, 3072, 1024, 0, time_it=True)]


def case_layernorm():
    import torch
    L = lib()
    out = []
    for (M, d) in [(1000, 1024), (777, 768), (257 * 5, 1280), (64, 256)]:
        torch.manual_seed(1)
        x = torch.randn(M, d, device="cuda") * 2 + 0.3
        g = torch.randn(d, device="cuda")
        b = torch.randn(d, device="cuda")
        y = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
        rc = L.b2c_layernorm_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(b.data_ptr()),
                                  C.c_void_p(y.data_ptr()), C.c_int64(M), C.c_int(d), C.c_float(1e-5),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, L.b2c_last_error()
        ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)
        err = (y.float() - ref).abs().max().item()
        out.append({"M": M, "d": d, "max_abs_err": err, "ok": bool(err < 0.05)})
    return out


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        res = CASES[sys.argv[2]]()
        print("RESULT " + json.dumps(res))
        return
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    summary = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--child", n], capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            summary[n] = {"rc": r.returncode, "s": round(time.time() - t0, 1),
                          "result": json.loads(lines[-1][7:]) if lines else None,
                          "stderr": r.stderr[-1500:] if r.returncode != 0 else ""}
        except subprocess.TimeoutExpired:
            summary[n] = {"rc": "timeout", "s": round(time.time() - t0, 1)}
        print(n, json.dumps(summary[n]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "dev_check.json"), "w") as fh:
        json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
