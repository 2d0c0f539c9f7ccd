"""Standalone attention launch for ncu: python tools/profile_attn.py [n_crops]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
L = C.CDLL(os.path.join(ROOT, "clip_assisted_data_labeling_b200", "libb2c.so"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T, heads, hd = 257, 16, 64
qkv = torch.randn(n * T, 3 * heads * hd, device="cuda").to(torch.bfloat16)
o = torch.zeros(n * T, heads * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    L.b2c_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), n, T, heads, hd, C.c_void_p(0))
torch.cuda.synchronize()
print("ok")
