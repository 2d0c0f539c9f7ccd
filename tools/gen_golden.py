"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported verbatim from
/root/reference through oracle/reference_shim.py) in the build container.  The reference tree does not
exist on the GPU box, so these small fixtures are what pins the oracle (and through it the CUDA path)
there.  Re-run:  python tools/gen_golden.py
"""
import hashlib
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shim as rs  # noqa: E402
from oracle import vit_oracle  # noqa: E402
from oracle.dedup_oracle import synthetic_embeddings  # noqa: E402
from oracle.preprocess_oracle import CROP_NAMES, OPENAI_MEAN, OPENAI_STD, synthetic_image  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen_preprocess():
    """Reference CustomImageDataset.extract_crops + its preprocess transform on synthetic images."""
    from PIL import Image
    emb = rs.import_reference("utils.embedder")
    enc = emb.CLIP_Encoder("ViT-B-32/openai", device="cpu")
    tf = enc.get_preprocess_transform()
    ds = emb.CustomImageDataset([], CROP_NAMES, tf)
    sizes = [(512, 512), (300, 200), (97, 260), (64, 64), (640, 128), (513, 512), (1000, 100)]  # (W, H)
    in_digests, digests, crop_sizes, full = [], [], [], {}
    mean = np.asarray(OPENAI_MEAN, np.float32)[:, None, None]
    std = np.asarray(OPENAI_STD, np.float32)[:, None, None]
    for k, (W, H) in enumerate(sizes):
        img = synthetic_image(k, H, W)  # regenerated from the seed by the tests (numpy Generator streams are stable)
        pil = Image.fromarray(img)
        raw, names = ds.extract_crops(pil)
        assert names == CROP_NAMES
        ref = torch.stack([tf(c) for c in raw]).numpy()  # f32 [4,3,224,224] — utils/embedder.py:173
        u8 = np.rint((ref * std + mean) * 255.0).astype(np.uint8)
        back = ((u8.astype(np.float32) / np.float32(255.0)) - mean) / std
        assert np.array_equal(back, ref), "u8 round trip must reproduce the reference tensor exactly"
        in_digests.append(hashlib.sha256(img.tobytes()).hexdigest())
        digests.append(hashlib.sha256(ref.tobytes()).hexdigest())
        crop_sizes.append([c.size for c in raw])
        if (W, H) == (97, 260):  # one full example (small input) for debugging a digest mismatch
            full = {"full_index": np.asarray(k), "full_img": img, "full_crops_u8": u8[2:4]}  # subcrop1, subcrop2
    np.savez_compressed(os.path.join(OUT, "preprocess_ref.npz"), sizes=np.asarray(sizes), in_sha256=np.asarray(in_digests),
                        sha256=np.asarray(digests), crop_sizes=np.asarray(crop_sizes), **full)
    print("preprocess_ref.npz", [os.path.getsize(os.path.join(OUT, "preprocess_ref.npz"))])


def gen_geometry():
    """Crop rectangles of the reference for many sizes (only sizes, cheap)."""
    from PIL import Image
    emb = rs.import_reference("utils.embedder")
    ds = emb.CustomImageDataset([], CROP_NAMES, None)
    rng = np.random.default_rng(7)
    sizes = [(512, 512), (768, 512), (512, 768), (1024, 256), (100, 1000), (64, 64), (513, 512), (333, 517), (1000, 100),
             (20, 20), (1, 50), (7, 3)]  # (W*H >= 10: below that the reference itself raises NameError at embedder.py:247)
    sizes += [tuple(int(v) for v in rng.integers(8, 1400, 2)) for _ in range(60)]
    rows = []
    for (W, H) in sizes:
        raw, names = ds.extract_crops(Image.new("RGB", (W, H)))
        got = {n: c.size for n, c in zip(names, raw)}
        rows.append([W, H] + [v for n in CROP_NAMES for v in got.get(n, (0, 0))])
    np.savez_compressed(os.path.join(OUT, "crop_geometry_ref.npz"), rows=np.asarray(rows, np.int64))
    print("crop_geometry_ref.npz", len(rows))


def gen_dedup():
    """Unmodified _2_remove_duplicates.get_paths_and_embeddings + find_near_duplicates core on synthetic .pt dirs."""
    ref = rs.import_reference("_2_remove_duplicates")
    cases = {}
    for name, (n, d, seed, thr) in {"a": (512, 256, 11, 0.96), "b": (1500, 768, 12, 0.96), "c": (700, 512, 13, 0.9)}.items():
        e = synthetic_embeddings(n, d, seed)
        tmp = tempfile.mkdtemp()
        root = os.path.join(tmp, "data")
        os.makedirs(root)
        for i in range(n):
            open(os.path.join(root, f"{i:06d}.jpg"), "wb").close()
            torch.save({"ViT-L-14/openai": {"square_padded_crop": e[i:i + 1].clone()}}, os.path.join(root, f"{i:06d}.pt"))
        args = types.SimpleNamespace(root_dir=root, threshold=thr, mode="copy", clip_model_to_use=None, chunk_size=10000, test=True)
        pairs, vals = [], []
        # replay of find_near_duplicates' body (:63-80) on what the reference generator yields, with the reference's ops
        for paths, embs in ref.get_paths_and_embeddings(args, "square_padded_crop"):
            idx_of = [int(os.path.basename(p)[:6]) for p in paths]
            E = torch.stack(embs)
            nE = E / torch.norm(E, dim=1, keepdim=True)
            S = torch.matmul(nE, nE.T)
            w = torch.where(torch.triu(S, diagonal=1) > args.threshold)
            for i, j in zip(w[0].tolist(), w[1].tolist()):
                pairs.append((idx_of[i], idx_of[j], i, j))
                vals.append(S[i, j].item())
        # and the real function end to end (test=True: no file operations) to make sure it runs on this layout
        ref.find_near_duplicates(args)
        order = np.asarray([int(os.path.basename(p)[:6]) for p in paths])
        shutil.rmtree(tmp)
        cases[f"{name}_meta"] = np.asarray([n, d, seed], np.int64)
        cases[f"{name}_thr"] = np.asarray(thr)
        cases[f"{name}_order"] = order  # os.walk order in which the reference stacked the rows
        cases[f"{name}_pairs"] = np.asarray(pairs, np.int64).reshape(-1, 4)
        cases[f"{name}_vals"] = np.asarray(vals, np.float32)
        if name == "a":
            cases["a_emb_f16"] = e.to(torch.float16).numpy()
        print("dedup case", name, "pairs", len(pairs))
    np.savez_compressed(os.path.join(OUT, "dedup_ref.npz"), **cases)


def gen_mlp():
    """Reference utils/nn_model.SimpleFC forward (seeded small instance) + the shipped checkpoint on seeded inputs."""
    nn_model = rs.import_reference("utils.nn_model")
    torch.manual_seed(5)
    m = nn_model.SimpleFC(96, [264, 128, 64], 1, clip_models=["ViT-L-14/openai"], crop_names=["centre_crop"],
                          dropout_prob=0.5).eval()
    x = torch.randn(33, 96)
    with torch.no_grad():
        y = m(x)
    lin = [l for l in m.layers if isinstance(l, torch.nn.Linear)]
    out = {"x": x.numpy(), "y": y.numpy()}
    for i, l in enumerate(lin):
        out[f"w{i}"] = l.weight.detach().numpy()
        out[f"b{i}"] = l.bias.detach().numpy()
    # shipped checkpoint: outputs only (weights stay in the reference tree)
    ck = os.path.join(rs.REFERENCE_ROOT, "models", "single_crop_regression_9.4k_imgs_80_epochs.pth")
    import collections
    with torch.serialization.safe_globals([nn_model.SimpleFC, torch.nn.ModuleList, torch.nn.Linear, torch.nn.LeakyReLU,
                                           torch.nn.Sigmoid, torch.nn.Dropout, set, collections.OrderedDict]):
        shipped = torch.load(ck, map_location="cpu", weights_only=True).eval()
    g = torch.Generator().manual_seed(9)
    xs = torch.nn.functional.normalize(torch.randn(16, 768, generator=g), dim=1)
    with torch.no_grad():
        ys = shipped(xs)
    out["shipped_x"] = xs.numpy()
    out["shipped_y"] = ys.numpy()
    out["shipped_crop_names"] = np.asarray(shipped.crop_names)
    out["shipped_clip_models"] = np.asarray(shipped.clip_models)
    np.savez_compressed(os.path.join(OUT, "mlp_ref.npz"), **out)
    print("mlp_ref.npz ok", float(ys.mean()))


def gen_vit():
    """Oracle tower outputs (the open_clip architecture restated) cross-checked against transformers' CLIP here."""
    out = {}
    for arch, pretrained, n in [("ViT-B-32", "openai", 4)]:
        m = vit_oracle.build_visual(arch, pretrained, seed=0)
        g = torch.Generator().manual_seed(1)
        px = torch.randn(n, 3, m.cfg["image"], m.cfg["image"], generator=g)
        with torch.no_grad():
            raw = m(px)
            hf = vit_oracle.to_hf_clip(m)(pixel_values=px).image_embeds
        assert (raw - hf).abs().max().item() < 2e-5, "oracle tower disagrees with transformers' CLIP"
        out[f"{arch}_emb"] = vit_oracle.encode_image_oracle(m, px).numpy()
        out[f"{arch}_hf_maxabs"] = np.asarray((raw - hf).abs().max().item())
        print(arch, "oracle vs HF max abs", (raw - hf).abs().max().item())
    np.savez_compressed(os.path.join(OUT, "vit_ref.npz"), **out)


def gen_similar():
    """Unmodified tools/find_similar_imgs.py (create_context_embedding + find_similar_imgs, both measures) and
    _3_label_images.diversity_ordered_image_files (tkinter / natsort / cv2 GUI imports stubbed) on synthetic .pt dirs."""
    import random
    from oracle.similar_oracle import synthetic_clusters
    for name in ("tkinter", "tkinter.ttk", "natsort"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tkinter"].ttk = sys.modules["tkinter.ttk"]
    sys.modules["natsort"].natsorted = sorted
    sys.modules["natsort"].ns = types.SimpleNamespace(IGNORECASE=0)
    rs.import_reference("utils.nn_model")  # puts the reference root on sys.path
    import importlib
    fs = importlib.import_module("tools.find_similar_imgs")
    fs.device = torch.device("cpu")
    out = {}
    n, E, n_ctx, seed = 400, 96, 7, 21
    emb = synthetic_clusters(n, E, seed)
    tmp = tempfile.mkdtemp()
    ctx_dir, search_dir = os.path.join(tmp, "ctx"), os.path.join(tmp, "search")
    os.makedirs(ctx_dir)
    os.makedirs(search_dir)
    for i in range(n_ctx):
        torch.save({"M/x": {"square_padded_crop": torch.from_numpy(emb[i:i + 1].copy())}}, os.path.join(ctx_dir, f"{i:05d}.pt"))
    for i in range(n_ctx, n):
        torch.save({"M/x": {"square_padded_crop": torch.from_numpy(emb[i:i + 1].copy())}}, os.path.join(search_dir, f"{i:05d}.pt"))
        if i % 17 != 0:  # samples without a .jpg are skipped by the reference (:106-110)
            open(os.path.join(search_dir, f"{i:05d}.jpg"), "wb").close()
    out["sim_meta"] = np.asarray([n, E, n_ctx, seed], np.int64)
    for measure in ("l2", "cosine"):
        args = types.SimpleNamespace(clip_models_to_use=["all"], crop_name_to_use="square_padded_crop",
                                     similarity_measure=measure, top_n=25, search_dir=search_dir, output_dir=tmp)
        ctx, names = fs.create_context_embedding(args, ctx_dir)
        top = fs.find_similar_imgs(args, ctx, names)
        ids = [int(os.path.basename(p)[:5]) for p in top.best_img_paths]
        d = [float(x) for x in top.best_distances]
        order = sorted(range(len(ids)), key=lambda t: (d[t], ids[t]))
        out[f"sim_{measure}_ctx"] = ctx.numpy()
        out[f"sim_{measure}_idx"] = np.asarray([ids[t] for t in order], np.int64)
        out[f"sim_{measure}_dist"] = np.asarray([d[t] for t in order], np.float32)
        print("similar", measure, out[f"sim_{measure}_idx"][:5], out[f"sim_{measure}_dist"][:3])
    shutil.rmtree(tmp)
    # diversity ordering: the reference reads d['square_padded_crop'] from the top level of the .pt (:141,:153)
    lab = importlib.import_module("_3_label_images")
    n2, E2, seed2, steps, S = 300, 64, 22, 40, 30
    emb2 = synthetic_clusters(n2, E2, seed2, n_clusters=9)
    tmp = tempfile.mkdtemp()
    files = []
    for i in range(n2):
        torch.save({"square_padded_crop": torch.from_numpy(emb2[i:i + 1].copy())}, os.path.join(tmp, f"{i:05d}.pt"))
        files.append(os.path.join(tmp, f"{i:05d}.jpg"))
    random.seed(1234)
    ordered = lab.diversity_ordered_image_files(files, tmp, total_n_ordered_imgs=steps, sample_size=S)
    shutil.rmtree(tmp)
    out["div_meta"] = np.asarray([n2, E2, seed2, steps, S, 1234], np.int64)
    out["div_order"] = np.asarray([int(os.path.basename(f)[:5]) for f in ordered], np.int64)
    print("diversity head", out["div_order"][:10])
    np.savez_compressed(os.path.join(OUT, "similar_ref.npz"), **out)


def gen_imgstats():
    """The reference's own ImageFeaturizer (utils/image_features.py; cv2 4.13 + numpy here) on synthetic images of sizes
    that exercise every cv::resize(INTER_AREA) branch: enlarged (512^2), mixed (one axis up, one down), fractional
    down-scale, integer 2x and 3x box filters, extreme aspect."""
    import importlib
    rs.import_reference("utils.nn_model")
    imf = importlib.import_module("utils.image_features")
    F = imf.ImageFeaturizer()
    sizes = [(512, 512), (300, 200), (97, 260), (1000, 700), (1536, 1536), (2304, 2304), (640, 128), (50, 50), (1100, 1000),
             (33, 700)]  # (W, H)
    rows = []
    for k, (W, H) in enumerate(sizes):
        d = F.process(synthetic_image(k, H, W))
        rows.append([float(d[n]) for n in d])
        names = list(d.keys())
    np.savez_compressed(os.path.join(OUT, "imgstats_ref.npz"), sizes=np.asarray(sizes), stats=np.asarray(rows, np.float64),
                        names=np.asarray(names))
    print("imgstats_ref.npz", len(rows), names[:3])


def make_labelled_dir(root, name, n, E, seed, crop_names=("centre_crop", "subcrop2"), model="M/x", n_nan=3):
    """Synthetic labelled dataset in the reference's layout: <root>/<name>.csv (uuid,label) + <root>/<name>/<uuid>.pt.
    The label is a smooth function of the embedding plus noise so that training has something to fit."""
    import pandas as pd
    g = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, name), exist_ok=True)
    wtrue = g.standard_normal(E * len(crop_names)).astype(np.float32)
    rows = []
    for i in range(n):
        fd = {c: torch.from_numpy(g.standard_normal((1, E)).astype(np.float32)) for c in crop_names}
        torch.save({model: fd}, os.path.join(root, name, f"u{i:05d}.pt"))
        x = np.concatenate([fd[c].numpy().ravel() for c in crop_names])
        label = float(np.tanh(x @ wtrue / np.sqrt(len(x))) * 2 + 3 + 0.1 * g.standard_normal())
        rows.append((f"u{i:05d}", label if i >= n_nan else np.nan))
    rows.append(("missing_file", 1.0))  # no .pt: the reference skips it (_4_train_model.py:72-74)
    pd.DataFrame(rows, columns=["uuid", "label"]).to_csv(os.path.join(root, name + ".csv"), index=False)


def train_args(root, **kw):
    a = dict(train_data_dir=root, train_data_names=["setA"], model_name="regressor", dont_save=False, clip_models_to_use=["all"],
             test_fraction=0.25, n_epochs=10, batch_size=16, lr=0.0002, min_lr=1e-6, restart_epochs=2, weight_decay=0.0006,
             dropout_prob=0.0, hidden_sizes=[24, 12], print_network_layout=False, random_seed=42)
    a.update(kw)
    return types.SimpleNamespace(**a)


def gen_train():
    """Unmodified _4_train_model.train() on CPU (dropout_prob = 0: deterministic) over a synthetic labelled directory;
    the saved whole-module pickle gives the final weights our trainer must reproduce from the same seed."""
    import glob
    import importlib
    rs.import_reference("utils.nn_model")
    t4 = importlib.import_module("_4_train_model")
    nn_model = importlib.import_module("utils.nn_model")
    t4.device = nn_model.device = torch.device("cpu")
    out = {}
    for case, kw in {"a": dict(), "b": dict(n_epochs=12, batch_size=7, hidden_sizes=[16], lr=0.01, weight_decay=0.0, test_fraction=0.2)}.items():
        tmp = tempfile.mkdtemp()
        n, E, seed = (150, 20, 31) if case == "a" else (90, 12, 32)
        make_labelled_dir(tmp, "setA", n, E, seed)
        args = train_args(tmp, **kw)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            t4.train(args, ["centre_crop", "subcrop2"], 0)
            pth = glob.glob(os.path.join(tmp, "models", "*.pth"))[0]
            import collections
            with torch.serialization.safe_globals([nn_model.SimpleFC, torch.nn.ModuleList, torch.nn.Linear, torch.nn.LeakyReLU,
                                                   torch.nn.Sigmoid, torch.nn.Dropout, set, collections.OrderedDict]):
                m = torch.load(pth, map_location="cpu", weights_only=True)
        finally:
            os.chdir(cwd)
        lin = [l for l in m.layers if isinstance(l, torch.nn.Linear)]
        out[f"{case}_meta"] = np.asarray([n, E, seed], np.int64)
        for i, l in enumerate(lin):
            out[f"{case}_w{i}"] = l.weight.detach().numpy()
            out[f"{case}_b{i}"] = l.bias.detach().numpy()
        out[f"{case}_final_mse"] = np.asarray(float(os.path.basename(pth).split("_epochs_")[1].split("_mse")[0]))
        shutil.rmtree(tmp)
        print("train case", case, os.path.basename(pth))
    np.savez_compressed(os.path.join(OUT, "train_ref.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["preprocess", "geometry", "dedup", "mlp", "vit", "similar", "train", "imgstats"]
    for w in which:
        globals()["gen_" + w]()
