"""Attention-only timing (CUDA events, 1024 crops of ViT-L/14: T=257, 16 heads x 64): python tools/bench_attn.py [n] [T] [heads] [hd]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
L = C.CDLL(os.environ.get("B2C_LIB") or os.path.join(ROOT, "clip_assisted_data_labeling_b200", "libb2c.so"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 257
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 16
hd = int(sys.argv[4]) if len(sys.argv) > 4 else 64
qkv = torch.randn(n * T, 3 * heads * hd, device="cuda").to(torch.bfloat16)
o = torch.zeros(n * T, heads * hd, device="cuda", dtype=torch.bfloat16)
run = lambda: L.b2c_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), n, T, heads, hd, C.c_void_p(0))
for _ in range(5):
    assert run() == 0
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
fl = 4.0 * T * T * heads * hd * n
print(json.dumps({"mode": os.environ.get("B2C_ATTN", "default"), "n": n, "T": T, "heads": heads, "hd": hd, "ms": ms, "tflops": fl / ms / 1e9}))
