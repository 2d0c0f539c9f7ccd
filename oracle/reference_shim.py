"""ORACLE (test infrastructure, never imported by the product path) — import the reference's own Python
modules from /root/reference *verbatim*, supplying stand-ins only for the third-party modules that are
not installed in this image:

  * ``open_clip``  -> a shim whose ``create_model_and_transforms`` builds oracle/vit_oracle.py's tower
                      (seeded random init of the named architecture) and the open_clip val transform
                      (torchvision Resize(BICUBIC) -> CenterCrop -> RGB -> ToTensor -> Normalize);
  * ``core.vision_encoder.{pe,transforms}`` (perception_models; hard import at utils/embedder.py:13-16)
                   -> empty stubs (the PE backend is out of scope);
  * ``matplotlib`` -> stub (only _4/_5 plotting).

The reference tree is /root/reference in the build container; on the GPU box it is the verbatim copy
oracle/make_ref.py leaves under the git-ignored baseline/_ref/.  Used to *generate* golden fixtures
(tests/golden/, scripts committed), to cross-check the restatements, and by bench.py's reference arm
(oracle/reference_runner.py) — never by the product path.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _locate() -> str:
    for cand in (os.environ.get("B2C_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "utils", "embedder.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _locate()


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


_VISUAL_CACHE = {}


def make_open_clip_shim(seed: int = 0):
    from oracle import vit_oracle

    mod = types.ModuleType("open_clip")

    def create_model_and_transforms(model_name, pretrained=None, precision="fp32", device="cpu", jit=False,
                                    cache_dir=None, **kw):
        import copy
        key = (model_name, pretrained or "openai", seed)
        if key not in _VISUAL_CACHE:  # the random init of a 300-600 M parameter tower takes seconds: build it once
            _VISUAL_CACHE[key] = vit_oracle.build_visual(model_name, pretrained or "openai", seed=seed)
        visual = copy.deepcopy(_VISUAL_CACHE[key])
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
