"""Our tcgen05 GEMM next to the library bar (cuBLAS via torch.matmul / F.linear) on the four ViT-L/14 block shapes,
same process, interleaved, CUDA events, sustained (>= 1 s per measurement).  Informational: printed to stdout as JSON.
    python tools/gemm_compare.py [n_crops]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from clip_assisted_data_labeling_b200 import _lib  # noqa: E402


def timed(fn, min_s=0.6):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, total = 0, 0.0
    while total < min_s * 1e3:
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
        n += 10
    return total / n


def main():
    n_crops = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    d, mlp, T = 1024, 4096, 257
    M = n_crops * T
    st = torch.cuda.current_stream().cuda_stream
    shapes = [("in_proj", 3 * d, d, _lib.EPI_BIAS_BF16), ("out_proj", d, d, _lib.EPI_BIAS_RESID_F32),
              ("c_fc", mlp, d, _lib.EPI_BIAS_QGELU_BF16), ("c_proj", d, mlp, _lib.EPI_BIAS_RESID_F32)]
    res = {}
    for name, N, K, mode in shapes:
        A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
        b = torch.zeros(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if mode == _lib.EPI_BIAS_RESID_F32 else torch.bfloat16)
        ours = timed(lambda: _lib.call("b2c_gemm_bf16", A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, mode, st))
        ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        bb = b.to(torch.bfloat16)
        lib_plain = timed(lambda: torch.matmul(A, W.t(), out=ob))            # no epilogue at all
        lib_bias = timed(lambda: torch.nn.functional.linear(A, W, bb))      # bias epilogue, bf16 out
        fl = 2.0 * M * N * K / 1e9
        res[name] = {"M": M, "N": N, "K": K, "ours_ms": ours, "ours_tflops": fl / ours, "cublas_plain_tflops": fl / lib_plain,
                     "cublas_bias_tflops": fl / lib_bias}
        del A, W, out, ob
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
