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


def case_attention():
    import torch
    L = lib()
    out = []
    for (n, T, heads, hd) in [(3, 257, 16, 64), (2, 50, 12, 64), (2, 257, 16, 80), (1, 577, 16, 64)]:
        torch.manual_seed(2)
        d = heads * hd
        qkv = (torch.randn(n * T, 3 * d, device="cuda") * 1.0).to(torch.bfloat16)
        o = torch.zeros(n * T, d, device="cuda", dtype=torch.bfloat16)
        rc = L.b2c_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), n, T, heads, hd,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, L.b2c_last_error()
        torch.cuda.synchronize()
        q, k, v = qkv.float().view(n, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(n * T, d)
        err = (o.float() - ref).abs().max().item()
        out.append({"n": n, "T": T, "heads": heads, "hd": hd, "max_abs_err": err, "ok": bool(err < 0.03)})
    return out


def case_attention_perf():
    import torch
    L = lib()
    n, T, heads, hd = 512, 257, 16, 64
    d = heads * hd
    qkv = torch.randn(n * T, 3 * d, device="cuda").to(torch.bfloat16)
    o = torch.zeros(n * T, d, device="cuda", dtype=torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        assert L.b2c_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), n, T, heads, hd, st) == 0, L.b2c_last_error()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        L.b2c_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), n, T, heads, hd, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    q, k, v = qkv[:4 * T].float().view(4, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(4 * T, d)
    err = (o[:4 * T].float() - ref).abs().max().item()
    return [{"ms": ms, "tflops": 4.0 * T * T * d * n / ms / 1e9, "max_abs_err": err, "path": os.environ.get("B2C_ATTN", "umma"), "ok": bool(err < 0.02)}]


def case_dedup_perf():
    import torch
    L = lib()
    out = []
    for N in (200_000, 1_000_000):
        E = 768
        e = torch.nn.functional.normalize(torch.randn(N, E, device="cuda"), dim=1).half()
        cap = 1 << 20
        pairs = torch.zeros(cap, 3, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        def run():
            rc = L.b2c_dedup_pairs(C.c_void_p(e.data_ptr()), C.c_int64(N), E, C.c_int64(0), C.c_int64(N), C.c_float(0.96), 1,
                                   C.c_void_p(pairs.data_ptr()), C.c_ulonglong(cap), C.c_void_p(cnt.data_ptr()), st)
            assert rc == 0, L.b2c_last_error()
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.time()
        e0.record()
        run()
        e1.record()
        t_launch = time.time() - t0
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out.append({"N": N, "kernel_ms": ms, "host_issue_s": t_launch, "pairs_per_s": N * (N - 1) / 2 / (ms / 1e3),
                    "tflops": N * (N - 1) * E / (ms / 1e3) / 1e12, "ok": True})
    return out


def case_preprocess():
    import numpy as np
    import torch
    from clip_assisted_data_labeling_b200.vit import preprocess_u8
    from oracle.preprocess_oracle import four_crop_preprocess, synthetic_image
    out = []
    rng = np.random.default_rng(0)
    sizes = [(512, 512), (768, 512), (512, 768), (1024, 256), (100, 1000), (64, 64), (513, 512), (333, 517)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    got = preprocess_u8([torch.from_numpy(im).cuda() for im in imgs], 224, 14, "nchw").cpu().numpy()
    for (w, h), im, g in zip(sizes, imgs, got):
        ref = four_crop_preprocess(im, 224)
        out.append({"W": w, "H": h, "bit_exact": bool(np.array_equal(ref, g)), "max_abs": float(np.abs(ref - g).max()),
                    "n_diff": int((ref != g).sum())})
    # uniform batch + patch layout
    batch = np.stack([synthetic_image(k) for k in range(4)])
    gp = preprocess_u8(torch.from_numpy(batch).cuda(), 224, 14, "patch").float().cpu()
    gn = preprocess_u8(torch.from_numpy(batch).cuda(), 224, 14, "nchw").cpu()
    x = gn.view(16, 3, 16, 14, 16, 14).permute(0, 2, 4, 1, 3, 5).reshape(16, 256, 588).to(torch.bfloat16).float()
    out.append({"patch_layout_equal": bool(torch.equal(gp[:, :, :588], x)), "pad_zero": bool((gp[:, :, 588:] == 0).all().item())})
    ref0 = four_crop_preprocess(batch[0], 224)
    out.append({"uniform_bit_exact": bool(np.array_equal(ref0, gn[0].numpy()))})
    for o in out:
        o["ok"] = all(v for k, v in o.items() if isinstance(v, bool))
    return out


def _vit_case(arch, n, seed=0):
    import torch
    from oracle import vit_oracle
    from clip_assisted_data_labeling_b200.vit import VisionTower
    m = vit_oracle.build_visual(arch, "openai" if arch != "ViT-H-14" else "laion2b_s32b_b79k", seed=seed)
    sd = vit_oracle.visual_state_dict(m)
    tower = VisionTower(vit_oracle.ARCHS[arch], m.cfg["act"], "cuda")
    tower.load_state_dict(sd)
    torch.manual_seed(seed + 1)
    R = m.cfg["image"]
    px = torch.randn(n, 3, R, R)
    t0 = time.time()
    ref = vit_oracle.encode_image_oracle(m, px)
    t_cpu = time.time() - t0
    got = tower.forward_pixels(px.cuda()).cpu()
    cos = torch.nn.functional.cosine_similarity(ref, got, dim=-1)
    return {"arch": arch, "n": n, "min_cos": cos.min().item(), "max_abs": (ref - got).abs().max().item(),
            "cpu_s": round(t_cpu, 2), "ok": bool(cos.min().item() >= 0.9995 and (ref - got).abs().max().item() <= 2e-3)}


def case_vit_b32():
    return [_vit_case("ViT-B-32", 8)]


def case_vit_l14():
    return [_vit_case("ViT-L-14", 4)]


def case_vit_h14():
    return [_vit_case("ViT-H-14", 2)]


def case_dedup():
    import torch
    L = lib()
    out = []
    for (N, E, thr) in [(300, 768, 0.96), (5000, 768, 0.96), (3000, 512, 0.9), (10000, 768, 0.96)]:
        torch.manual_seed(3)
        e = torch.nn.functional.normalize(torch.randn(N, E), dim=1)
        ndup = N // 50
        src = torch.randint(0, N, (ndup,))
        dst = torch.randperm(N)[:ndup]
        c = torch.empty(ndup).uniform_(0.90, 0.999)
        sigma = (1 / c ** 2 - 1).sqrt()
        e[dst] = torch.nn.functional.normalize(e[src] + sigma[:, None] * torch.randn(ndup, E) / E ** 0.5, dim=1)
        e16 = e.to(torch.float16).cuda()
        # reference semantics (_2_remove_duplicates.py:67-80) on the GPU in fp16
        nrm = e16 / torch.norm(e16, dim=1, keepdim=True)
        S = nrm @ nrm.T
        idx = torch.where(torch.triu(S, diagonal=1) > thr)
        ref = set(zip(idx[0].tolist(), idx[1].tolist()))
        S32 = (nrm.float() @ nrm.float().T)
        E_pad = (E + 63) // 64 * 64
        buf = torch.zeros(N, E_pad, dtype=torch.float16, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.b2c_normalize_rows_f16(C.c_void_p(e16.data_ptr()), 1, C.c_int64(N), E, C.c_void_p(buf.data_ptr()), st)
        assert rc == 0, L.b2c_last_error()
        cap = 1 << 20
        pairs = torch.zeros(cap, 3, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        rc = L.b2c_dedup_pairs(C.c_void_p(buf.data_ptr()), C.c_int64(N), E_pad, C.c_int64(0), C.c_int64(N), C.c_float(thr), 1,
                               C.c_void_p(pairs.data_ptr()), C.c_ulonglong(cap), C.c_void_p(cnt.data_ptr()), st)
        assert rc == 0, L.b2c_last_error()
        torch.cuda.synchronize()
        k = int(cnt.item())
        got = set((int(a), int(b)) for a, b in pairs[:k, :2].tolist())
        diff = ref ^ got
        # pairs may differ only where the similarity is within 1e-3 of the threshold
        bad = [(i, j) for (i, j) in diff if abs(S32[i, j].item() - thr) > 1e-3]
        out.append({"N": N, "E": E, "ref": len(ref), "got": len(got), "sym_diff": len(diff), "bad": len(bad), "ok": len(bad) == 0 and k > 0})
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
