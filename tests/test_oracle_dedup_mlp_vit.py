"""Pins the dedup, regressor and tower oracles against golden vectors generated from the reference's own code
(tools/gen_golden.py) and against independent implementations installed here."""
import numpy as np
import pytest
import torch

from conftest import reference_present
from oracle import vit_oracle
from oracle.dedup_oracle import duplicate_pairs_oracle, pair_sets_match, synthetic_embeddings
from oracle.mlp_oracle import assemble_features, simple_fc_forward


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_dedup_oracle_vs_reference_golden(golden, case):
    g = golden("dedup_ref.npz")
    n, d, seed = g[f"{case}_meta"].tolist()
    thr = float(g[f"{case}_thr"])
    e = synthetic_embeddings(n, d, seed)
    if case == "a":
        assert np.array_equal(e.to(torch.float16).numpy(), g["a_emb_f16"]), "synthetic embedding generator drifted"
    order = g[f"{case}_order"]  # the os.walk order in which the reference stacked the rows
    pairs, sims, _ = duplicate_pairs_oracle(e[order], thr)
    ref = g[f"{case}_pairs"]
    assert len(ref) > 0
    assert np.array_equal(pairs, ref[:, 2:4]), "oracle pair list (row-major, i<j) != reference"
    assert np.array_equal(order[pairs], ref[:, 0:2])
    assert np.array_equal(sims, g[f"{case}_vals"])


def test_pair_band_rule():
    S = np.array([[1.0, 0.9605, 0.97], [0.9605, 1.0, 0.5], [0.97, 0.5, 1.0]], np.float32)
    ok, bad = pair_sets_match([(0, 1), (0, 2)], [(0, 2)], S, 0.96)
    assert ok and not bad  # (0,1) is within 1e-3 of the threshold: don't-care
    ok, bad = pair_sets_match([(0, 2)], [], S, 0.96)
    assert not ok and bad == [(0, 2)]


def test_mlp_oracle_vs_reference_golden(golden):
    g = golden("mlp_ref.npz")
    W = [g[f"w{i}"] for i in range(4)]
    b = [g[f"b{i}"] for i in range(4)]
    y = simple_fc_forward(g["x"], W, b)
    assert y.shape == g["y"].shape
    np.testing.assert_allclose(y, g["y"], rtol=0, atol=2e-6)
    assert g["shipped_crop_names"].tolist() == ["centre_crop"]
    assert g["shipped_clip_models"].tolist() == ["ViT-L-14-336/openai"]


@pytest.mark.skipif(not reference_present(), reason="shipped checkpoint only exists in the build container")
def test_mlp_oracle_vs_shipped_checkpoint(golden):
    from clip_assisted_data_labeling_b200.scorer import load_regressor
    g = golden("mlp_ref.npz")
    m = load_regressor("/root/reference/models/single_crop_regression_9.4k_imgs_80_epochs.pth")
    lin = [l for l in m.layers if isinstance(l, torch.nn.Linear)]
    y = simple_fc_forward(g["shipped_x"], [l.weight.detach().numpy() for l in lin], [l.bias.detach().numpy() for l in lin])
    np.testing.assert_allclose(y, g["shipped_y"], rtol=0, atol=2e-6)


def test_assemble_features_order():
    d = {"centre_crop": np.ones((1, 3)), "subcrop2": 2 * np.ones((1, 3)), "subcrop1": 3 * np.ones((1, 3))}
    f = assemble_features([d, d], ["subcrop2", "centre_crop"])
    assert f.tolist() == [2, 2, 2, 1, 1, 1] * 2


def test_vit_oracle_golden(golden):
    g = golden("vit_ref.npz")
    m = vit_oracle.build_visual("ViT-B-32", "openai", seed=0)
    px = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    out = vit_oracle.encode_image_oracle(m, px).numpy()
    np.testing.assert_allclose(out, g["ViT-B-32_emb"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
    assert float(g["ViT-B-32_hf_maxabs"]) < 2e-5


def test_vit_oracle_vs_transformers_clip():
    """Independent second statement of the architecture on a narrow random config, both activations."""
    for act in ("quick_gelu", "gelu"):
        torch.manual_seed(0)
        m = vit_oracle.VisionTransformer(image=64, patch=16, width=128, layers=2, heads=2, mlp=512, embed=64, act=act).eval()
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.05 * torch.randn_like(p))
        px = torch.randn(3, 3, 64, 64)
        with torch.no_grad():
            a = m(px)
            b = vit_oracle.to_hf_clip(m)(pixel_values=px).image_embeds
        assert (a - b).abs().max().item() < 1e-4


def test_flops_formula():
    assert abs(vit_oracle.flops_per_crop("ViT-L-14") / 1e9 - 162.03) < 0.05
    assert abs(vit_oracle.flops_per_crop("ViT-H-14") / 1e9 - 334.59) < 0.05
    assert abs(vit_oracle.flops_per_crop("ViT-B-32") / 1e9 - 8.82) < 0.02


@pytest.mark.parametrize("arch,pretrained", [("ViT-B-32", "openai"), ("ViT-L-14", "openai")])
def test_vit_oracle_vs_transformers_clip_full_shape(arch, pretrained):
    """The same second statement at the FULL shape of the named architectures (seeded random init, one crop): the oracle
    the GPU parity tests and bench.py's parity blocks trust agrees with transformers' CLIPVisionModelWithProjection live,
    not only through the stored scalar of tests/golden."""
    m = vit_oracle.build_visual(arch, pretrained, seed=0)
    R = vit_oracle.ARCHS[arch]["image"]
    px = torch.randn(1, 3, R, R, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a = m(px)
        b = vit_oracle.to_hf_clip(m)(pixel_values=px).image_embeds
    a, b = a / a.norm(dim=-1, keepdim=True), b / b.norm(dim=-1, keepdim=True)
    assert (a - b).abs().max().item() < 2e-5
