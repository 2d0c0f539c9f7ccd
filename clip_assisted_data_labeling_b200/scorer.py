"""K10 — the reference's FC regressor forward on the device, while the [B,4,E] embeddings are still in HBM.

Mirrors ``SimpleFC`` (utils/nn_model.py:6-41: Linear -> LeakyReLU(0.01) -> Dropout [eval: identity] per hidden
layer, Linear -> Sigmoid) and the feature assembly of ``_5_predict_labels.py``:77-82 (per CLIP model, the
crops listed in ``model.crop_names`` concatenated in that order)."""
from __future__ import annotations

import ctypes as C
import sys
import types

import torch
from torch import nn

from . import _lib
from .vit_arch import CROP_NAMES


class SimpleFC(nn.Module):
    """Structural twin of utils/nn_model.py:6-41 so whole-module pickles written by the reference's
    ``_4_train_model.py``:237 (class path ``utils.nn_model.SimpleFC``) can be loaded without the reference."""

    def __init__(self, input_size, hidden_sizes, output_size, clip_models,
                 crop_names=("centre_crop", "square_padded_crop", "subcrop1", "subcrop2"),
                 use_img_stat_features=False, dropout_prob=0.0, data_min=None, data_max=None, verbose=0):
        super().__init__()
        self.clip_models = clip_models
        self.crop_names = list(crop_names)
        self.use_img_stat_features = use_img_stat_features
        self.data_min, self.data_max = data_min, data_max
        sizes = [input_size] + list(hidden_sizes) + [output_size]
        layers = []
        for i in range(len(sizes) - 1):
            layers.append(nn.Linear(sizes[i], sizes[i + 1]))
            if i < len(sizes) - 2:
                layers.append(nn.LeakyReLU())
                layers.append(nn.Dropout(p=dropout_prob))
        layers.append(nn.Sigmoid())
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


def load_regressor(path: str) -> nn.Module:
    """torch.load of a reference regressor pickle (utils/nn_model.SimpleFC, tensors saved on cuda:0) onto the
    CPU with an allow-list instead of arbitrary unpickling (torch >= 2.6 refuses whole-module pickles by default)."""
    import collections
    created = []
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
        cls = SimpleFC
        cls.__module__ = "utils.nn_model"
    else:
        cls = sys.modules["utils.nn_model"].SimpleFC
    try:
        allow = [cls, nn.ModuleList, nn.Linear, nn.LeakyReLU, nn.Sigmoid, nn.Dropout, set, collections.OrderedDict]
        with torch.serialization.safe_globals(allow):
            model = torch.load(path, map_location="cpu", weights_only=True)
    finally:
        if "utils.nn_model" in created:
            SimpleFC.__module__ = __name__
        for k in created:
            sys.modules.pop(k, None)
    return model.eval()


class FCScorer:
    """SimpleFC.forward as one sm_100a kernel (b2c_mlp_score)."""

    def __init__(self, model: nn.Module, device="cuda"):
        self.device = torch.device(device)
        self.clip_models = list(getattr(model, "clip_models", []))
        self.crop_names = list(getattr(model, "crop_names", CROP_NAMES))
        # regressors trained with the image statistics behind the crops (_4_train_model.py:60-63; the reference's own _5
        # never feeds them, so it cannot score such a model — score_store does)
        self.use_img_stat_features = bool(getattr(model, "use_img_stat_features", False))
        linears = [m for m in model.layers if isinstance(m, nn.Linear)]
        slopes = [m.negative_slope for m in model.layers if isinstance(m, nn.LeakyReLU)]
        if not linears or len(linears) > _lib.MLP_MAX_LAYERS:
            raise ValueError("unsupported regressor depth")
        if not isinstance(model.layers[-1], nn.Sigmoid):
            raise ValueError("expected the SimpleFC layout (final Sigmoid)")
        self.slope = float(slopes[0]) if slopes else 0.01
        self._w = [l.weight.detach().float().contiguous().to(self.device) for l in linears]
        self._b = [l.bias.detach().float().contiguous().to(self.device) for l in linears]
        self.in_dim = self._w[0].shape[1]
        self.out_dim = self._w[-1].shape[0]
        w = _lib.MlpWeights()
        w.n_layers = len(linears)
        w.dims[0] = self.in_dim
        for i, t in enumerate(self._w):
            w.dims[i + 1] = t.shape[0]
            w.weight[i] = t.data_ptr()
            w.bias[i] = self._b[i].data_ptr()
        w.leaky_slope = self.slope
        self._cw = w

    @torch.no_grad()
    def score(self, feats: torch.Tensor) -> torch.Tensor:
        """feats f32 [B, in_dim] -> f32 [B, out_dim]  (``model(features.float())``, _5_predict_labels.py:135)."""
        feats = feats.to(self.device, torch.float32).contiguous()
        if feats.dim() != 2 or feats.shape[1] != self.in_dim:
            raise ValueError(f"expected [B,{self.in_dim}], got {tuple(feats.shape)}")
        out = torch.empty(feats.shape[0], self.out_dim, dtype=torch.float32, device=self.device)
        if feats.shape[0]:
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().b2c_mlp_score(C.c_void_p(feats.data_ptr()), feats.shape[0], C.byref(self._cw),
                                                     C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr())),
                           "b2c_mlp_score")
        return out

    def assemble(self, emb: torch.Tensor, emb_crop_names=CROP_NAMES) -> torch.Tensor:
        """emb [B, n_crops, E] of ONE clip model -> [B, len(self.crop_names) * E] in the regressor's crop order
        (_5_predict_labels.py:77-82)."""
        idx = [list(emb_crop_names).index(c) for c in self.crop_names]
        return emb[:, idx, :].reshape(emb.shape[0], -1)

    def score_embeddings(self, emb: torch.Tensor, emb_crop_names=CROP_NAMES) -> torch.Tensor:
        return self.score(self.assemble(emb, emb_crop_names))


@torch.no_grad()
def embed_and_score(encoder, scorer: "FCScorer", images_u8):
    """BASELINE config 5 in one pass: 4-crop embeddings (encoder.encode_images_u8) and the regressor score while the
    [B,4,E] block is still in HBM — the fused form of _1_embed_with_CLIP.py + _5_predict_labels.py:77-82,135.
    Returns (embeddings f32 [B,4,E], scores f32 [B,1])."""
    emb = encoder.encode_images_u8(images_u8)
    return emb, scorer.score_embeddings(emb)


@torch.no_grad()
def score_store(store, scorer: "FCScorer", batch: int = 65536):
    """Regressor scores for every image of a packed store (store.PackedStore) — the bulk form of the per-image loop of
    _5_predict_labels.py:69-88,135.  Images missing one of the regressor's crops are skipped like the reference skips
    unreadable samples (:86-88).  Returns (paths, scores f32 [n] on the host)."""
    ok = store.has_all(scorer.crop_names)
    feats = store.features(scorer.crop_names, with_stats=scorer.use_img_stat_features)
    keep = [i for i in range(len(store)) if ok[i]]
    out = []
    for b in range(0, len(keep), batch):
        sel = keep[b:b + batch]
        out.append(scorer.score(feats[sel].pin_memory().to(scorer.device, non_blocking=True))[:, 0].cpu())
    return [store.paths[i] for i in keep], (torch.cat(out) if out else torch.zeros(0))
