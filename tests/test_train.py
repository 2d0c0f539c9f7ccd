"""Regressor training.

CPU: the oracle (oracle/train_oracle.py) driven by the product's host logic (trainer.train with an injected CPU engine)
reproduces the final weights of the UNMODIFIED reference `_4_train_model.train()` (tests/golden/train_ref.npz), which
pins data order, split, initialisation, schedule, loss and Adam semantics; the Philox stream is checked against known
answers.  GPU: the CUDA trainer against the same golden weights and, step by step with dropout on, against the oracle."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle.train_oracle import SimpleFCOracle, dropout_scale, philox4x32_10, train_epoch_oracle  # noqa: E402


def _gen():
    import gen_golden
    return gen_golden


class OracleEngine:
    """CPU step engine with DeviceTrainer's interface, built from the oracle (tests only)."""

    def __init__(self, model, max_batch=16, dropout_p=0.0, seed=0, device="cpu"):
        self.model = model
        lin = [m for m in model.layers if isinstance(m, torch.nn.Linear)]
        with torch.random.fork_rng():  # building nn.Linear layers draws from the global generator: keep the host sequence intact
            self.o = SimpleFCOracle(lin[0].in_features, [l.out_features for l in lin[:-1]], lin[-1].out_features)
        for a, b in zip(self.o.linears, lin):
            a.weight.data.copy_(b.weight.data)
            a.bias.data.copy_(b.bias.data)
        self.lin, self.p, self.seed, self.steps, self.opt = lin, dropout_p, seed, 0, None

    def epoch(self, feats, labels, order, batch, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        if self.opt is None:
            self.opt = torch.optim.Adam(self.o.parameters(), lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
        for g in self.opt.param_groups:
            g["lr"] = lr
        total, self.steps = train_epoch_oracle(self.o, self.opt, feats, labels, order, batch, self.p, self.seed, self.steps)
        return total

    @torch.no_grad()
    def predict(self, feats):
        return self.o(feats)

    def pull(self):
        for a, b in zip(self.o.linears, self.lin):
            b.weight.data.copy_(a.weight.data)
            b.bias.data.copy_(a.bias.data)
        return self.model


CASES = {"a": dict(), "b": dict(n_epochs=12, batch_size=7, hidden_sizes=[16], lr=0.01, weight_decay=0.0, test_fraction=0.2)}


def _run_case(tmp_path, golden, case, **train_kw):
    from clip_assisted_data_labeling_b200.trainer import train
    g = golden("train_ref.npz")
    n, E, seed = g[f"{case}_meta"].tolist()
    gg = _gen()
    gg.make_labelled_dir(str(tmp_path), "setA", n, E, seed)
    args = gg.train_args(str(tmp_path), dont_save=True, **CASES[case])
    model, losses, lrs = train(args, ["centre_crop", "subcrop2"], 0, verbose=False, **train_kw)
    lin = [l for l in model.layers if isinstance(l, torch.nn.Linear)]
    return g, lin, losses, lrs, args


def test_philox_known_answers():
    # Random123 known-answer tests for Philox4x32-10
    assert [int(x) for x in philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(x) for x in philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    s = dropout_scale(7, 3, 1, 64, 264, 0.5)
    assert s.shape == (64, 264) and set(np.unique(s).tolist()) == {0.0, 2.0} and 0.45 < (s > 0).mean() < 0.55
    assert not np.array_equal(s, dropout_scale(7, 4, 1, 64, 264, 0.5))


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_and_host_logic_reproduce_reference_training(tmp_path, golden, case):
    g, lin, losses, lrs, args = _run_case(tmp_path, golden, case, device="cpu", engine_cls=OracleEngine)
    for i, l in enumerate(lin):
        np.testing.assert_allclose(l.weight.detach().numpy(), g[f"{case}_w{i}"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(l.bias.detach().numpy(), g[f"{case}_b{i}"], rtol=0, atol=2e-6)
    assert abs(losses[1][-1] - float(g[f"{case}_final_mse"])) < 6e-5  # the reference prints 4 decimals into the file name
    assert len(lrs) == args.n_epochs


def test_lr_schedule_matches_torch():
    from clip_assisted_data_labeling_b200.trainer import cosine_warm_restarts_lr
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=2e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=10, T_mult=1, eta_min=1e-6)
    for e in range(35):
        assert abs(opt.param_groups[0]["lr"] - cosine_warm_restarts_lr(2e-4, 1e-6, 10, e)) < 1e-12
        opt.step()
        sch.step()


def test_trainer_fails_loudly_without_gpu():
    from clip_assisted_data_labeling_b200 import _lib
    from clip_assisted_data_labeling_b200.scorer import SimpleFC
    from clip_assisted_data_labeling_b200.trainer import DeviceTrainer
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(_lib.B2CError):
        DeviceTrainer(SimpleFC(8, [4], 1, ["M/x"]))


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_device_training_reproduces_reference(tmp_path, golden, lib, case):
    """End to end on the device from the same seed: final weights of the unmodified reference within fp32 round-off
    accumulated over ~100 optimiser steps (different summation order): 2e-5 absolute."""
    g, lin, losses, lrs, args = _run_case(tmp_path, golden, case)
    for i, l in enumerate(lin):
        np.testing.assert_allclose(l.weight.detach().cpu().numpy(), g[f"{case}_w{i}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(l.bias.detach().cpu().numpy(), g[f"{case}_b{i}"], rtol=0, atol=2e-5)
    assert abs(losses[1][-1] - float(g[f"{case}_final_mse"])) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("D,hidden,batch,p,n", [(96, [264, 128, 64], 16, 0.5, 100), (4096, [264, 128, 64], 16, 0.5, 70),
                                                (70, [33], 5, 0.3, 23), (256, [64, 32], 40, 0.0, 130), (40, [], 64, 0.0, 200)])
def test_device_steps_vs_oracle_with_dropout(lib, D, hidden, batch, p, n):
    """Same batches, same Philox dropout masks: weights after two epochs (incl. a partial last batch) agree with the torch
    CPU autograd + torch.optim.Adam restatement."""
    from clip_assisted_data_labeling_b200.scorer import SimpleFC
    from clip_assisted_data_labeling_b200.trainer import DeviceTrainer
    torch.manual_seed(D + n)
    model = SimpleFC(D, hidden, 1, ["M/x"], dropout_prob=p)
    feats = torch.randn(n, D)
    labels = torch.rand(n)
    eng = OracleEngine(model, batch, p, seed=99)
    tr = DeviceTrainer(model, max_batch=batch, dropout_p=p, seed=99)
    fd, ld = feats.cuda(), labels.cuda()
    for ep in range(2):
        order = torch.randperm(n).tolist()
        lo = eng.epoch(feats, labels, order, batch, lr=1e-3 * (ep + 1), weight_decay=6e-4)
        ld_ = tr.epoch(fd, ld, order, batch, lr=1e-3 * (ep + 1), weight_decay=6e-4)
        assert abs(lo - ld_) < 1e-4 * max(1.0, abs(lo))
    assert tr.steps == eng.steps == 2 * ((n + batch - 1) // batch)
    ref = [(l.weight.detach().clone(), l.bias.detach().clone()) for l in eng.o.linears]
    tr.pull()
    lin = [m for m in model.layers if isinstance(m, torch.nn.Linear)]
    for (w, b), l in zip(ref, lin):
        np.testing.assert_allclose(l.weight.detach().numpy(), w.numpy(), rtol=0, atol=3e-5)
        np.testing.assert_allclose(l.bias.detach().numpy(), b.numpy(), rtol=0, atol=3e-5)
    x = torch.randn(9, D)
    np.testing.assert_allclose(tr.predict(x).cpu().numpy(), eng.predict(x).numpy(), rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_saved_regressor_loads_like_a_reference_pickle(tmp_path, golden, lib, monkeypatch):
    from clip_assisted_data_labeling_b200.scorer import load_regressor
    from clip_assisted_data_labeling_b200.trainer import save_regressor
    g, lin, losses, lrs, args = _run_case(tmp_path, golden, "b")
    from clip_assisted_data_labeling_b200.scorer import SimpleFC
    m = SimpleFC(24, [16], 1, ["M/x"], crop_names=["centre_crop", "subcrop2"])
    path = save_regressor(m, args, 72, losses, out_dir=str(tmp_path / "models"))
    with open(path, "rb") as fh:
        assert b"utils.nn_model" in fh.read()  # the class path the reference's _5_predict_labels.py unpickles
    back = load_regressor(path)
    assert back.crop_names == ["centre_crop", "subcrop2"] and back.clip_models == ["M/x"]
    assert SimpleFC.__module__ == "clip_assisted_data_labeling_b200.scorer"


def test_store_backed_features_resolve_the_model_and_match_rows_by_path(tmp_path):
    """trainer.load_labelled_features(store=...): the saved regressor's clip_models must name the store's model (the
    reference's consumers build one encoder per entry, 'all' is not a model), other / several models cannot be served
    from one store, and a CSV row is matched by <train_data_dir>/<name>/<uuid>, not by the bare uuid."""
    from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter
    from clip_assisted_data_labeling_b200.trainer import load_labelled_features
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    import pandas as pd
    E = 8
    root = tmp_path / "data"
    g = torch.Generator().manual_seed(0)
    feats, paths = [], []
    for name in ("setA", "setB"):
        (root / name).mkdir(parents=True)
        for u in ("u0", "u1", "u2"):  # the same uuids in both datasets
            feats.append(torch.randn(4, E, generator=g))
            paths.append(str(root / name / f"{u}.jpg"))
        pd.DataFrame([("u0", 1.0), ("u1", 2.0), ("u2", 3.0), ("absent", 4.0)], columns=["uuid", "label"]).to_csv(root / f"{name}.csv", index=False)
    with PackedWriter(str(tmp_path / "store"), "ViT-L-14/openai", E) as w:
        w.append(torch.stack(feats), paths)
    store = PackedStore(str(tmp_path / "store"))
    crops = ["centre_crop", "subcrop2"]
    cols = [CROP_NAMES.index(c) for c in crops]

    def args(models, names):
        return types.SimpleNamespace(train_data_dir=str(root), train_data_names=names, clip_models_to_use=models)

    a = args(["all"], ["setB"])
    x, y = load_labelled_features(a, crops, store)
    assert a.clip_models_to_use == ["ViT-L-14/openai"]
    assert x.shape == (3, 2 * E) and sorted(y.tolist()) == [1.0, 2.0, 3.0]
    want = {float(l): feats[3 + i][cols].flatten() for i, l in enumerate((1.0, 2.0, 3.0))}  # setB's rows, not setA's
    for row, label in zip(x, y.tolist()):
        assert torch.equal(row, want[label])
    x2, _ = load_labelled_features(args(["ViT-L-14/openai"], ["setA", "setB"]), crops, store)
    assert x2.shape == (6, 2 * E)
    with pytest.raises(ValueError, match="holds 'ViT-L-14/openai' only"):
        load_labelled_features(args(["ViT-H-14/laion2b_s32b_b79k"], ["setA"]), crops, store)
    with pytest.raises(ValueError):
        load_labelled_features(args(["ViT-L-14/openai", "ViT-B-32/openai"], ["setA"]), crops, store)


def test_img_stat_features_follow_the_crops(tmp_path):
    """use_img_stat_features (_4_train_model.py:60-63): each model's img_stat_* scalars, in file order, are appended after
    its crops — from per-image .pt files and from a packed store that carries the statistics alike."""
    from clip_assisted_data_labeling_b200.imgstats import STAT_NAMES
    from clip_assisted_data_labeling_b200.store import PackedStore, import_pt
    from clip_assisted_data_labeling_b200.trainer import load_labelled_features
    import pandas as pd
    E = 4
    root = tmp_path / "data"
    (root / "set").mkdir(parents=True)
    g = torch.Generator().manual_seed(1)
    want = {}
    for k, u in enumerate(("a", "b", "c")):
        crops = {c: torch.randn(1, E, generator=g) for c in ("centre_crop", "square_padded_crop", "subcrop1", "subcrop2")}
        stats = {n: torch.randn((), generator=g) for n in STAT_NAMES}
        torch.save({"M/x": {**stats, **crops}}, root / "set" / f"{u}.pt")
        open(root / "set" / f"{u}.jpg", "wb").close()
        want[float(k)] = torch.cat([crops["centre_crop"].flatten(), crops["subcrop2"].flatten(), torch.stack(list(stats.values()))])
    pd.DataFrame([("a", 0.0), ("b", 1.0), ("c", 2.0)], columns=["uuid", "label"]).to_csv(root / "set.csv", index=False)

    def args():
        return types.SimpleNamespace(train_data_dir=str(root), train_data_names=["set"], clip_models_to_use=["M/x"])

    crops = ["centre_crop", "subcrop2"]
    x, y = load_labelled_features(args(), crops, use_img_stat_features=1)
    assert x.shape == (3, 2 * E + 22)
    for row, label in zip(x, y.tolist()):
        assert torch.equal(row, want[label])
    x0, _ = load_labelled_features(args(), crops)
    assert x0.shape == (3, 2 * E)
    store = import_pt(str(root), str(tmp_path / "store"), "M/x")
    assert store.stat_names == STAT_NAMES
    xs, ys = load_labelled_features(args(), crops, store, use_img_stat_features=1)
    for row, label in zip(xs, ys.tolist()):
        assert torch.equal(row, want[label])
