"""ORACLE (test infrastructure, never imported by the product path) — import the reference's own Python
modules from /root/reference *verbatim*, supplying stand-ins only for the third-party modules that are
not installed in this image:

  * ``open_clip``  -> a shim whose ``create_model_and_transforms`` builds oracle/vit_oracle.py's tower
                      (seeded random init of the named architecture) and the open_clip val transform
                      (torchvision Resize(BICUBIC) -> CenterCrop -> RGB -> ToTensor -> Normalize);
  * ``core.vision_encoder.{pe,transforms}`` (perception_models; hard import at utils/embedder.py:13-16)
                   -> empty stubs (the PE backend is out of scope);
  * ``matplotlib`` -> stub (only _4/_5 plotting).

The reference exists only in the build container: everything here is used to *generate* golden
fixtures (tests/golden/, scripts committed) and to cross-check the restatements; nothing that runs on
the GPU box imports it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B2C_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "embedder.py"))


def _val_transform(image_size: int):
    from torchvision import transforms
    from oracle.preprocess_oracle import OPENAI_MEAN, OPENAI_STD

    def _to_rgb(im):
        return im.convert("RGB")

    return transforms.Compose([
        transforms.Resize(image_size, interpolation=transforms.InterpolationMode.BICUBIC),
        transforms.CenterCrop(image_size),
        _to_rgb,
        transforms.ToTensor(),
        transforms.Normalize(mean=OPENAI_MEAN, std=OPENAI_STD),
    ])


def make_open_clip_shim(seed: int = 0):
    from oracle import vit_oracle

    mod = types.ModuleType("open_clip")

    def create_model_and_transforms(model_name, pretrained=None, precision="fp32", device="cpu", jit=False,
                                    cache_dir=None, **kw):
        visual = vit_oracle.build_visual(model_name, pretrained or "openai", seed=seed)
        model = vit_oracle.CLIPVisualOnly(visual)
        if precision == "fp16":
            model = model.half()
        model = model.to(device)
        tf = _val_transform(vit_oracle.ARCHS[model_name]["image"])
        return model, tf, tf

    def list_pretrained():
        return [(a, "openai") for a in vit_oracle.ARCHS]

    mod.create_model_and_transforms = create_model_and_transforms
    mod.list_pretrained = list_pretrained
    return mod


def install_stubs(seed: int = 0) -> None:
    sys.modules["open_clip"] = make_open_clip_shim(seed)
    for name in ("core", "core.vision_encoder", "core.vision_encoder.pe", "core.vision_encoder.transforms"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["core"].vision_encoder = sys.modules["core.vision_encoder"]
    sys.modules["core.vision_encoder"].pe = sys.modules["core.vision_encoder.pe"]
    sys.modules["core.vision_encoder"].transforms = sys.modules["core.vision_encoder.transforms"]
    try:
        import matplotlib  # noqa: F401
    except Exception:
        class _NoOpModule(types.ModuleType):  # plt.figure(...), plt.savefig(...) ... all become no-ops
            def __getattr__(self, name):
                if name.startswith("__"):
                    raise AttributeError(name)
                return lambda *a, **k: None

        mpl = _NoOpModule("matplotlib")
        plt = _NoOpModule("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def import_reference(module: str, seed: int = 0):
    """Import e.g. 'utils.embedder', '_2_remove_duplicates', 'utils.nn_model' from the reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs(seed)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's top-level package is called 'utils'; make sure no other 'utils' shadows it
    m = sys.modules.get("utils")
    if m is not None:
        locs = [str(getattr(m, "__file__", None) or "")] + [str(p) for p in getattr(m, "__path__", [])]
        if not any(l.startswith(REFERENCE_ROOT) for l in locs):
            for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
                del sys.modules[k]
    return importlib.import_module(module)
