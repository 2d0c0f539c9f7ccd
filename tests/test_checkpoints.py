"""Weights handling of CLIP_Encoder (utils/embedder.py:59-88 calls open_clip.create_model_and_transforms, which downloads
pretrained weights; there is no network here): checkpoint discovery in a file / directory / open_clip's caches, the
formats a checkpoint may come in (pickled state dict, {'state_dict': ...}, full-CLIP 'visual.' prefix, safetensors,
TorchScript archive, transformers' CLIP names), and the rule that a missing checkpoint is an error unless random
initialisation is requested explicitly.  Host logic only (no GPU)."""
import os

import pytest
import torch

from clip_assisted_data_labeling_b200 import embedder as emb
from clip_assisted_data_labeling_b200.vit_arch import ARCHS, random_state_dict, state_dict_shapes

TINY = dict(image=32, patch=16, width=256, layers=2, heads=4, mlp=512, embed=64)


def _sd(cfg=TINY, seed=0):
    return random_state_dict(cfg, seed=seed)


def _same(a, b):
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k].float(), b[k].float()), k


def test_find_checkpoint_file_directory_and_name_boundaries(tmp_path, monkeypatch):
    monkeypatch.setenv("HOME", str(tmp_path / "home"))
    monkeypatch.delenv("B2C_CLIP_CACHE", raising=False)
    monkeypatch.delenv("HF_HOME", raising=False)
    d = tmp_path / "ckpts"
    d.mkdir()
    assert emb._find_checkpoint(None, "ViT-L-14", "openai") is None
    assert emb._find_checkpoint(str(d), "ViT-L-14", "openai") is None
    (d / "ViT-L-14-336_openai.pt").write_bytes(b"x")
    # the 336 px checkpoint is not a ViT-L-14 checkpoint
    assert emb._find_checkpoint(str(d), "ViT-L-14", "openai") is None
    assert emb._find_checkpoint(str(d), "ViT-L-14-336", "openai") == str(d / "ViT-L-14-336_openai.pt")
    (d / "ViT-L-14.safetensors").write_bytes(b"x")
    assert emb._find_checkpoint(str(d), "ViT-L-14", "openai") == str(d / "ViT-L-14.safetensors")
    (d / "ViT-L-14_openai.pth").write_bytes(b"x")  # the more specific name wins
    assert emb._find_checkpoint(str(d), "ViT-L-14", "openai") == str(d / "ViT-L-14_openai.pth")
    # an explicit file is taken as is, whatever it is called
    f = tmp_path / "weights.bin"
    f.write_bytes(b"x")
    assert emb._find_checkpoint(str(f), "ViT-H-14", "laion2b_s32b_b79k") == str(f)
    (d / "notes.txt").write_text("ViT-H-14")
    assert emb._find_checkpoint(str(d), "ViT-H-14", "laion2b_s32b_b79k") is None


def test_find_checkpoint_in_open_clip_download_locations(tmp_path, monkeypatch):
    home = tmp_path / "home"
    monkeypatch.setenv("HOME", str(home))
    monkeypatch.delenv("B2C_CLIP_CACHE", raising=False)
    monkeypatch.delenv("HF_HOME", raising=False)
    # OpenAI weights: ~/.cache/clip/<file open_clip downloads>
    clip_cache = home / ".cache" / "clip"
    clip_cache.mkdir(parents=True)
    (clip_cache / "ViT-L-14-336px.pt").write_bytes(b"x")
    assert emb._find_checkpoint(None, "ViT-L-14-336", "openai") == str(clip_cache / "ViT-L-14-336px.pt")
    assert emb._find_checkpoint(None, "ViT-L-14", "openai") is None
    # LAION weights: Hugging Face hub cache, repository name carries architecture and tag
    snap = home / ".cache" / "huggingface" / "hub" / "models--laion--CLIP-ViT-H-14-laion2B-s32B-b79K" / "snapshots" / "abc"
    snap.mkdir(parents=True)
    (snap / "open_clip_pytorch_model.bin").write_bytes(b"x")
    assert emb._find_checkpoint(None, "ViT-H-14", "laion2b_s32b_b79k") == str(snap / "open_clip_pytorch_model.bin")
    assert emb._find_checkpoint(None, "ViT-H-14", "openai") is None
    # transformers-format OpenAI CLIP in the hub cache; patch14 must not match patch14-336
    s336 = home / ".cache" / "huggingface" / "hub" / "models--openai--clip-vit-large-patch14-336" / "snapshots" / "r"
    s336.mkdir(parents=True)
    (s336 / "model.safetensors").write_bytes(b"x")
    assert emb._find_checkpoint(None, "ViT-L-14", "openai") is None
    s224 = home / ".cache" / "huggingface" / "hub" / "models--openai--clip-vit-large-patch14" / "snapshots" / "r"
    s224.mkdir(parents=True)
    (s224 / "pytorch_model.bin").write_bytes(b"x")
    assert emb._find_checkpoint(None, "ViT-L-14", "openai") == str(s224 / "pytorch_model.bin")
    # $B2C_CLIP_CACHE comes before the default locations
    other = tmp_path / "other"
    other.mkdir()
    (other / "ViT-L-14.pt").write_bytes(b"x")
    monkeypatch.setenv("B2C_CLIP_CACHE", str(other))
    assert emb._find_checkpoint(None, "ViT-L-14", "openai") == str(other / "ViT-L-14.pt")


def test_load_checkpoint_formats(tmp_path):
    sd = _sd()
    # plain pickled state dict
    torch.save(sd, tmp_path / "a.pt")
    _same(emb._load_checkpoint(str(tmp_path / "a.pt")), sd)
    # training checkpoint wrapper, DataParallel prefix, full CLIP ('visual.' + text tower + logit_scale)
    full = {"module.visual." + k: v for k, v in sd.items()}
    full["module.token_embedding.weight"] = torch.zeros(4, 4)
    full["module.logit_scale"] = torch.tensor(1.0)
    torch.save({"state_dict": full, "epoch": 3}, tmp_path / "b.pth")
    got = emb._load_checkpoint(str(tmp_path / "b.pth"))
    assert "visual.conv1.weight" in got and "logit_scale" in got and "epoch" not in got
    _same({k[len("visual."):]: v for k, v in got.items() if k.startswith("visual.")}, sd)
    # safetensors, fp16 like the files open_clip publishes
    from safetensors.torch import save_file
    save_file({"visual." + k: v.half().contiguous() for k, v in sd.items()}, str(tmp_path / "c.safetensors"))
    got = emb._load_checkpoint(str(tmp_path / "c.safetensors"))
    assert got["visual.proj"].dtype == torch.float16
    _same({k[len("visual."):]: v for k, v in got.items()}, {k: v.half() for k, v in sd.items()})
    with pytest.raises(Exception):
        (tmp_path / "junk.pt").write_bytes(b"not a checkpoint")
        emb._load_checkpoint(str(tmp_path / "junk.pt"))


def test_load_checkpoint_torchscript_archive(tmp_path):
    """OpenAI's published .pt files are TorchScript archives (torch.load(weights_only=True) refuses them)."""
    class Visual(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.proj = torch.nn.Parameter(torch.randn(8, 4))
            self.conv1 = torch.nn.Conv2d(3, 8, 2, 2, bias=False)

        def forward(self, x):
            return self.conv1(x).flatten(1)[:, :8] @ self.proj

    class Clip(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = Visual()

        def forward(self, x):
            return self.visual(x)

    m = Clip().eval()
    torch.jit.script(m).save(str(tmp_path / "ViT-B-32.pt"))
    got = emb._load_checkpoint(str(tmp_path / "ViT-B-32.pt"))
    assert torch.equal(got["visual.proj"], m.visual.proj.detach()) and "visual.conv1.weight" in got


def test_transformers_clip_names_are_mapped_back(tmp_path):
    """A transformers-format CLIP vision checkpoint (what the hub cache holds for openai/clip-vit-*) loads under
    open_clip's names: the inverse of the mapping the oracle uses for its HF cross-check."""
    from oracle import vit_oracle
    m = vit_oracle.VisionTransformer(act="quick_gelu", **TINY).eval()
    hf = vit_oracle.to_hf_clip(m)
    torch.save(hf.state_dict(), tmp_path / "pytorch_model.bin")
    got = emb._load_checkpoint(str(tmp_path / "pytorch_model.bin"))
    want = vit_oracle.visual_state_dict(m)
    assert set(want) <= set(got)
    for k, v in want.items():
        assert tuple(got[k].shape) == tuple(state_dict_shapes(TINY)[k]) and torch.equal(got[k], v), k


class _StubTower:
    loaded = None

    def __init__(self, cfg, act, device):
        self.cfg, self.act = cfg, act

    def load_state_dict(self, sd):
        _StubTower.loaded = sd


def _patch_tower(monkeypatch):
    import clip_assisted_data_labeling_b200.vit as vit
    monkeypatch.setattr(vit, "VisionTower", _StubTower)
    monkeypatch.setattr(vit, "_require_cuda", lambda d: torch.device("cuda", 0))


def test_missing_checkpoint_raises_unless_random_init_is_requested(tmp_path, monkeypatch, capsys):
    monkeypatch.setenv("HOME", str(tmp_path / "home"))
    monkeypatch.delenv("B2C_CLIP_CACHE", raising=False)
    monkeypatch.delenv("HF_HOME", raising=False)
    _patch_tower(monkeypatch)
    # the reference's default call (model_path=None) can only work where open_clip can download: here it must not
    # silently become random weights
    with pytest.raises(FileNotFoundError, match="allow_random_init"):
        emb.CLIP_Encoder("ViT-B-32/openai")
    with pytest.raises(FileNotFoundError):
        emb.CLIP_Encoder("ViT-B-32/openai", model_path=str(tmp_path))
    with pytest.raises(FileNotFoundError):
        emb.CLIP_Encoder("ViT-B-32/openai", seed=5)  # a seed alone is not consent
    enc = emb.CLIP_Encoder("ViT-B-32/openai", seed=5, allow_random_init=True)
    assert enc.weights_source == "random-init(seed=5)" and "not pretrained" in capsys.readouterr().out
    _same(_StubTower.loaded, random_state_dict(ARCHS["ViT-B-32"], seed=5))
    # a checkpoint in model_path (file or directory) is used and named
    sd = random_state_dict(ARCHS["ViT-B-32"], seed=9)
    torch.save({"visual." + k: v for k, v in sd.items()}, tmp_path / "ViT-B-32_openai.pt")
    for mp in (str(tmp_path), str(tmp_path / "ViT-B-32_openai.pt")):
        enc = emb.CLIP_Encoder("ViT-B-32/openai", model_path=mp)
        assert enc.weights_source == str(tmp_path / "ViT-B-32_openai.pt")
        assert "visual.proj" in _StubTower.loaded  # the tower strips the prefix itself (vit.VisionTower.load_state_dict)
    enc = emb.CLIP_Encoder("ViT-B-32/openai", state_dict=sd)
    assert enc.weights_source == "state_dict" and _StubTower.loaded is sd
    assert (enc.precision, enc.img_resolution, enc.model_architecture, enc.pretrained_dataset) == ("bf16", 224, "ViT-B-32", "openai")


def test_feature_dataset_passes_the_opt_in_and_records_the_source(tmp_path, monkeypatch):
    import json
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    monkeypatch.setenv("HOME", str(tmp_path / "home"))
    _patch_tower(monkeypatch)
    root = tmp_path / "data"
    root.mkdir()
    with pytest.raises(FileNotFoundError):
        Feature_Dataset(str(root), "ViT-B-32/openai", 4)
    ds = Feature_Dataset(str(root), "ViT-B-32/openai", 4, allow_random_init=True)
    assert ds.encoder.weights_source.startswith("random-init")
    from clip_assisted_data_labeling_b200.store import PackedWriter
    with PackedWriter(str(tmp_path / "store"), "ViT-B-32/openai", 512, weights_source=ds.encoder.weights_source) as w:
        w.append(torch.zeros(1, 4, 512), ["a.png"])
    meta = json.load(open(os.path.join(tmp_path / "store", "shard-00000.json")))
    assert meta["weights_source"] == "random-init(seed=0)"
