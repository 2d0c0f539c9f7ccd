"""Development: phase timeline of attention_umma5_kernel (B2C_ATTN5_VAR=17: the default variant + trace) on one SM, eight
iterations from K0 (env), plus per-CTA durations in cycles and nanoseconds.  REPS launches first (clocks under load)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
L = C.CDLL(os.path.join(ROOT, "clip_assisted_data_labeling_b200", "libb2c.so"))
n, T, heads, hd = int(os.environ.get('N', 512)), 257, 16, 64
L.b2c_debug_attn_trace_start(int(os.environ.get('K0', 4)))
qkv = torch.randn(n * T, 3 * heads * hd, device="cuda").to(torch.bfloat16)
o = torch.zeros(n * T, heads * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(int(os.environ.get('REPS', 3))):
    assert L.b2c_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), n, T, heads, hd, C.c_void_p(0)) == 0
torch.cuda.synchronize()
v5 = True
assert int(os.environ.get("B2C_ATTN5_VAR", "-1")) & 16, "run with B2C_ATTN5_VAR=17 (a trace build)"
tr = np.zeros((22, 8, 16), np.int64)
assert L.b2c_debug_attn5_trace(C.c_void_p(tr.ctypes.data), C.c_size_t(tr.nbytes)) == 0
t0 = tr[tr > 0].min()
print("var", os.environ.get("B2C_ATTN5_VAR"), "period (softmax warp 2, ev0):", np.diff(tr[2, :, 0]).tolist())
names_mma = ["top", "qk_ready", "S0_issued", "pv1prev_issued", "S1_issued", "v_ready", "PV0_issued", "end"]
names_sm = ["top", "S_ready", "max_done", "max_xchg", "P_done", "clskey", "clsrow", "bar7", "pre_O", "O_ready", "epi_done", "end"]
for it in (2, 3):
    print("--- iteration", it + 4, "(cycles since first event)")
    print("mma :", " ".join(f"{nm}={tr[1, it, e] - t0}" for e, nm in enumerate(names_mma)))
    for wp in (2, 6, 10, 14):
        print(f"w{wp:2d} g{(wp - 2) >> 3} h{((wp - 2) >> 2) & 1}:", " ".join(f"{nm}={tr[wp, it, e] - t0}" for e, nm in enumerate(names_sm) if tr[wp, it, e]))
    if v5:
        print("cls :", " ".join(f"{nm}={tr[18, it, e] - t0}" for e, nm in enumerate(["top", "qk_ready", "s0_done", "softmax_done", "end"])))

if v5:
    ct = np.zeros((160, 4), np.int64)
    assert L.b2c_debug_attn5_cta(C.c_void_p(ct.ctypes.data), C.c_size_t(ct.nbytes)) == 0
    ct = ct[:148]
    dur_ns = ct[:, 1] - ct[:, 0]
    dur_clk = ct[:, 3] - ct[:, 2]
    print("per-CTA duration us: min %.1f median %.1f max %.1f | kernel span %.1f us | cycles min %d median %d max %d | MHz median %.0f" % (
        dur_ns.min() / 1e3, np.median(dur_ns) / 1e3, dur_ns.max() / 1e3, (ct[:, 1].max() - ct[:, 0].min()) / 1e3,
        dur_clk.min(), np.median(dur_clk), dur_clk.max(), np.median(dur_clk / (dur_ns / 1e3))))
    order = np.argsort(dur_clk)
    print("fastest CTAs", order[:6].tolist(), "slowest", order[-6:].tolist(), "start spread us %.1f" % ((ct[:, 0].max() - ct[:, 0].min()) / 1e3))
