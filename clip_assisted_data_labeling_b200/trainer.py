"""Regressor training on the device — mirror of the reference's ``_4_train_model.py``.

``train(args, crop_names, use_img_stat_features)`` keeps the reference's argument object (``--train_data_dir
--train_data_names --clip_models_to_use --test_fraction --n_epochs --batch_size --lr --min_lr --restart_epochs
--weight_decay --dropout_prob --hidden_sizes --random_seed --model_name --dont_save``, _4_train_model.py:241-262) and
its host-side sequence, so that the same seed yields the same shuffle (pandas ``sample``), the same train/test split
(``random_split``), the same initial weights (``nn.Linear`` init of ``SimpleFC``) and the same per-epoch batch order
(``DataLoader(shuffle=True)`` — the index order is drawn by the very same torch sampler machinery, over indices only).
What changes underneath: features live in HBM once, and every optimiser step (forward, MSE, backward, Adam) is 3L-1
kernel launches of ``libb2c.so`` (``b2c_trainer_epoch``) with no host round trip inside an epoch; the learning rate of
``CosineAnnealingWarmRestarts`` is evaluated on the host once per epoch like ``scheduler.step()`` does (:206).
Dropout masks come from a counter-based Philox stream (torch's own CUDA masks are not reproducible across launch
geometries either); with ``dropout_prob = 0`` the result equals the reference's to fp32 round-off.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
from torch import nn
from torch.utils.data import DataLoader, Dataset, random_split

from . import _lib
from .scorer import SimpleFC


class _IndexDataset(Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class DeviceTrainer:
    """SimpleFC + Adam state on the device; one call per epoch."""

    def __init__(self, model: nn.Module, max_batch: int = 16, dropout_p: float | None = None, seed: int = 0, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.B2CError("DeviceTrainer needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.model = model
        self.linears = [m for m in model.layers if isinstance(m, nn.Linear)]
        drops = [m.p for m in model.layers if isinstance(m, nn.Dropout)]
        slopes = [m.negative_slope for m in model.layers if isinstance(m, nn.LeakyReLU)]
        if not isinstance(model.layers[-1], nn.Sigmoid) or len(self.linears) > _lib.MLP_MAX_LAYERS:
            raise ValueError("expected the SimpleFC layout (Linear/LeakyReLU/Dropout ... Linear/Sigmoid)")
        cfg = _lib.TrainerCfg()
        cfg.n_layers = len(self.linears)
        cfg.dims[0] = self.linears[0].in_features
        for i, l in enumerate(self.linears):
            cfg.dims[i + 1] = l.out_features
        cfg.max_batch = int(max_batch)
        cfg.leaky_slope = float(slopes[0]) if slopes else 0.01
        cfg.dropout_p = float(dropout_p if dropout_p is not None else (drops[0] if drops else 0.0))
        cfg.seed = int(seed) & (2 ** 64 - 1)
        self.cfg = cfg
        self.lib = _lib.load()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2c_trainer_create(C.byref(cfg), C.byref(h)), "b2c_trainer_create")
        self._h = h
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.push()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.b2c_trainer_destroy(h)

    def _st(self):
        return C.c_void_p(_lib.current_stream_ptr())

    def push(self):
        """torch module -> device trainer."""
        with torch.cuda.device(self.device):
            for i, l in enumerate(self.linears):
                w = l.weight.detach().to(self.device, torch.float32).contiguous()
                b = l.bias.detach().to(self.device, torch.float32).contiguous()
                _lib.check(self.lib.b2c_trainer_set_layer(self._h, i, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), self._st()),
                           "b2c_trainer_set_layer")
            torch.cuda.current_stream().synchronize()

    def pull(self):
        """device trainer -> torch module (in place)."""
        with torch.cuda.device(self.device):
            for i, l in enumerate(self.linears):
                w = torch.empty(l.weight.shape, dtype=torch.float32, device=self.device)
                b = torch.empty(l.bias.shape, dtype=torch.float32, device=self.device)
                _lib.check(self.lib.b2c_trainer_get_layer(self._h, i, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), self._st()),
                           "b2c_trainer_get_layer")
                with torch.no_grad():
                    l.weight.copy_(w.to(l.weight.device))
                    l.bias.copy_(b.to(l.bias.device))
        return self.model

    @property
    def steps(self) -> int:
        return int(self.lib.b2c_trainer_steps(self._h))

    def epoch(self, feats: torch.Tensor, labels: torch.Tensor, order, batch: int, lr: float, weight_decay: float = 0.0,
              betas=(0.9, 0.999), eps: float = 1e-8) -> float:
        """One pass over ``order`` (sample indices into feats/labels, both resident on the device); returns the sum of the
        per-step batch losses (``train_loss`` of _4_train_model.py:203 before the division by len(train_loader))."""
        assert feats.is_cuda and labels.is_cuda and feats.dtype == torch.float32 and labels.dtype == torch.float32
        assert feats.stride(1) == 1 and labels.is_contiguous()
        order_d = torch.as_tensor(np.asarray(order, dtype=np.int32)).to(self.device)
        hyper = _lib.Adam(float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay))
        self._loss.zero_()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2c_trainer_epoch(self._h, C.c_void_p(feats.data_ptr()), feats.stride(0), C.c_void_p(labels.data_ptr()),
                                                  C.c_void_p(order_d.data_ptr()), order_d.numel(), int(batch), C.byref(hyper),
                                                  C.c_void_p(self._loss.data_ptr()), self._st()), "b2c_trainer_epoch")
        return float(self._loss.item())

    @torch.no_grad()
    def predict(self, feats: torch.Tensor) -> torch.Tensor:
        """Eval-mode forward (dropout off) with the live device weights: f32 [n, out]."""
        w = _lib.MlpWeights()
        _lib.check(self.lib.b2c_trainer_weights(self._h, C.byref(w)), "b2c_trainer_weights")
        feats = feats.to(self.device, torch.float32).contiguous()
        out = torch.empty(feats.shape[0], self.linears[-1].out_features, dtype=torch.float32, device=self.device)
        if feats.shape[0]:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.b2c_mlp_score(C.c_void_p(feats.data_ptr()), feats.shape[0], C.byref(w), C.c_void_p(out.data_ptr()),
                                                  self._st()), "b2c_mlp_score")
        return out


def cosine_warm_restarts_lr(base_lr, eta_min, T_0, epochs_done):
    """CosineAnnealingWarmRestarts(T_0, T_mult=1, eta_min) after ``epochs_done`` scheduler steps (_4_train_model.py:126,206)."""
    t_cur = epochs_done % T_0
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / T_0)) / 2


def load_labelled_features(args, crop_names, store=None, use_img_stat_features=0):
    """_4_train_model.py:27-80: per dataset, labels.csv -> shuffled rows -> per-image feature vector (crops of every
    clip model concatenated; with ``use_img_stat_features`` each model's ``img_stat_*`` scalars follow its crops, :60-63).
    With ``store`` (a packed store, store.PackedStore) the vectors come from the shard instead of one torch.load per
    image; rows whose image is missing are skipped like the reference's ``except: continue``."""
    import pandas as pd
    features, labels = [], []
    by_path = None
    if store is not None:
        # A packed store holds ONE model.  The saved regressor carries clip_models (utils/nn_model.py:15) and _5 / the
        # reference's AestheticRegressor build an encoder per entry, so the list must name that model — never 'all'.
        want = list(args.clip_models_to_use)
        if want == ["all"]:
            args.clip_models_to_use = [store.model_name]
            print(f"\n----> Using the packed store's clip model: {args.clip_models_to_use}")
        elif want != [store.model_name]:
            raise ValueError(f"the packed store holds {store.model_name!r} only; clip_models_to_use={want} cannot be served from it "
                             "(train from the per-image .pt files, or pack one store per model)")
        # rows are matched by the image's location <train_data_dir>/<name>/<uuid>.<ext> (_4_train_model.py:46), not by the
        # bare uuid: the same uuid may exist in two datasets
        by_path = {}
        for i, p in enumerate(store.paths):
            by_path.setdefault(os.path.splitext(os.path.abspath(p))[0], i)
        full = store.features(crop_names, with_stats=bool(use_img_stat_features))
        ok = store.has_all(crop_names)
    for name in args.train_data_names:
        data = pd.read_csv(os.path.join(args.train_data_dir, name + ".csv"))
        data = data.dropna(subset=["label"])
        data = data.sample(frac=1).reset_index(drop=True)
        n_samples = skips = 0
        for _index, row in data.iterrows():
            try:
                uuid, label = row["uuid"], row["label"]
                if by_path is not None:
                    i = by_path[os.path.abspath(os.path.join(args.train_data_dir, name, str(uuid)))]
                    if not ok[i]:
                        raise KeyError("missing crop")
                    vec = full[i]
                else:
                    d = torch.load(f"{args.train_data_dir}/{name}/{uuid}.pt", map_location="cpu")
                    if args.clip_models_to_use[0] == "all":
                        args.clip_models_to_use = list(d.keys())
                        print(f"\n----> Using all found clip models: {args.clip_models_to_use}")
                    parts = []
                    for m in args.clip_models_to_use:
                        fd = d[m]
                        missing = set(crop_names) - set(fd.keys())
                        if missing:
                            raise Exception(f"Missing crops {missing} for {uuid}")
                        parts.append(torch.cat([fd[c] for c in crop_names if c in fd], dim=0).flatten())
                        if use_img_stat_features:
                            parts.append(torch.stack([fd[k] for k in fd.keys() if k.startswith("img_stat_")], dim=0).float())
                    vec = torch.cat(parts, dim=0)
                features.append(vec)
                labels.append(label)
                n_samples += 1
            except Exception:  # noqa: BLE001  (_4_train_model.py:72-74: skip the sample)
                skips += 1
        print(f"Loaded {n_samples} samples from {name}!" + (f" (skipped {skips} samples due to loading errors).." if skips else ""))
    return torch.stack(features, dim=0).float(), torch.tensor(labels).float()


def train(args, crop_names, use_img_stat_features=0, store=None, device="cuda", dropout_seed=None, verbose=True,
          engine_cls=None):
    """_4_train_model.py:16-238.  Returns (model, losses [[train...],[test...]], lrs).  ``engine_cls`` (tests only)
    replaces DeviceTrainer by another step engine with the same interface; the product path never passes it."""
    torch.manual_seed(args.random_seed)
    np.random.seed(args.random_seed)
    features, labels = load_labelled_features(args, crop_names, store, use_img_stat_features)
    labels_min, labels_max = labels.min(), labels.max()
    labels = (labels - labels_min) / (labels_max - labels_min)
    n = len(features)
    train_size = int((1 - args.test_fraction) * n)
    test_size = n - train_size
    if verbose:
        print(f"Training on {train_size} samples, testing on {test_size} samples.")
    train_ds, test_ds = random_split(_IndexDataset(n), [train_size, test_size])
    train_loader = DataLoader(train_ds, batch_size=args.batch_size, shuffle=True)
    test_loader = DataLoader(test_ds, batch_size=args.batch_size, shuffle=False)
    model = SimpleFC(features.shape[1], list(args.hidden_sizes), 1, args.clip_models_to_use, crop_names=crop_names,
                     use_img_stat_features=bool(use_img_stat_features), dropout_prob=args.dropout_prob)
    dev = torch.device(device)
    feats_d, labels_d = features.to(dev), labels.to(dev)
    trainer = (engine_cls or DeviceTrainer)(model, max_batch=args.batch_size, dropout_p=args.dropout_prob,
                            seed=args.random_seed if dropout_seed is None else dropout_seed, device=dev)

    def test_loss():
        """_4_train_model.py:129-166: mean over test batches of the batch MSE, next to the predict-the-mean dummy."""
        if len(test_loader) == 0:
            return -1.0, -1.0
        batches = [torch.as_tensor(b) for b in test_loader]  # also draws the loader's base seed like the reference does
        idx_all = torch.cat(batches).to(dev)
        pred = trainer.predict(feats_d[idx_all]).squeeze(1)
        lab = labels_d[idx_all]
        tl = dl = 0.0
        off = 0
        for b in batches:
            p, y = pred[off:off + len(b)], lab[off:off + len(b)]
            tl += float(((p - y) ** 2).mean())
            dl += float(((y.mean() - y) ** 2).mean())
            off += len(b)
        return tl / len(batches), dl / len(batches)

    losses, lrs = [[], []], []
    tl, dl = test_loss()
    if verbose:
        print(f"\nBefore training, test mse-loss: {tl:.4f} (dummy: {dl:.4f})")
    for epoch in range(args.n_epochs):
        lr = cosine_warm_restarts_lr(args.lr, args.min_lr, args.restart_epochs, epoch)
        order, n_batches = [], 0
        for b in train_loader:  # the same sampler machinery as the reference: identical permutation for the same seed
            order.extend(b.tolist())  # Subset over an index dataset: items ARE the original sample indices
            n_batches += 1
        loss_sum = trainer.epoch(feats_d, labels_d, order, args.batch_size, lr, weight_decay=args.weight_decay)
        current_lr = cosine_warm_restarts_lr(args.lr, args.min_lr, args.restart_epochs, epoch + 1)
        lrs.append(current_lr)
        train_loss = loss_sum / max(n_batches, 1)
        tl, dl = test_loss()
        losses[0].append(train_loss)
        losses[1].append(tl)
        if verbose and epoch % 2 == 0:
            ts = f", test mse: {tl:.4f} (dummy: {dl:.4f})" if tl > 0 else ""
            print(f"Epoch {epoch + 1}/{args.n_epochs}, train-mse: {train_loss:.4f}, lr: {current_lr:.6f}{ts}")
    trainer.pull()
    model.eval()
    if verbose and losses[1] and losses[1][-1] > 0:
        print(f"---> Best test mse loss: {min(losses[1]):.4f} in epoch {int(np.argmin(losses[1])) + 1}")
    if not getattr(args, "dont_save", True):
        save_regressor(model, args, train_size, losses)
    return model, losses, lrs


def save_regressor(model, args, train_size, losses, out_dir="models"):
    """_4_train_model.py:229-238: whole-module pickle named like the reference names it, with the class recorded as
    ``utils.nn_model.SimpleFC`` so the reference's _5_predict_labels.py (``torch.load(model_file)``, :98) can open it."""
    import pandas as pd
    ts = pd.Timestamp.now().strftime("%Y-%m-%d_%H:%M:%S")
    name = f"{args.model_name}_{ts}_{(train_size / 1000):.1f}k_imgs_{args.n_epochs}_epochs_{losses[1][-1]:.4f}_mse"
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, name + ".pth")
    old_mod, old_qual = SimpleFC.__module__, SimpleFC.__qualname__
    import sys
    import types
    created = []
    try:
        SimpleFC.__module__ = "utils.nn_model"
        if "utils.nn_model" not in sys.modules:
            if "utils" not in sys.modules:
                pkg = types.ModuleType("utils")
                pkg.__path__ = []
                sys.modules["utils"] = pkg
                created.append("utils")
            mod = types.ModuleType("utils.nn_model")
            mod.SimpleFC = SimpleFC
            sys.modules["utils.nn_model"] = mod
            created.append("utils.nn_model")
        torch.save(model.cpu(), path)
    finally:
        SimpleFC.__module__, SimpleFC.__qualname__ = old_mod, old_qual
        for k in created:
            sys.modules.pop(k, None)
    print("Final model saved to /model dir as:\n", f"{name}.pth")
    return path
