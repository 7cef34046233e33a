"""GPU parity of the ViT stage (A1-A4) against the CPU oracle and the golden reference outputs.

The CUDA path computes GEMM operands in fp16 with fp32 accumulation and an fp32 residual stream;
the oracle is fp32 throughout, so bit-exactness is not defined for this stage (SURVEY.md §7 "Hard
parts").  Tolerances (stated per test): relative Frobenius error and per-token cosine similarity.
"""
import os

import pytest
import torch

from foundpose_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.pt")


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_layernorm_matches_torch():
    from foundpose_b200 import _native

    g = torch.Generator().manual_seed(0)
    for d in (128, 384, 1024):
        x = (torch.randn(1000, d, generator=g) * 3 + 0.5).cuda()
        w = torch.randn(d, generator=g).cuda()
        b = torch.randn(d, generator=g).cuda()
        y = _native.layernorm_f16(x, w, b, 1e-6).float()
        ref = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-6)
        assert (y - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("batch,tokens,heads", [(1, 128, 1), (2, 17, 2), (3, 901, 16), (2, 905, 6), (1, 257, 2)])
def test_attention_matches_torch(batch, tokens, heads):
    """fp16 Q/K/V, fp32 softmax: |err| <= 2e-3 * max|ref| (P is rounded to fp16 before P@V)."""
    from foundpose_b200 import _native

    g = torch.Generator().manual_seed(1)
    d = heads * 64
    qkv = (torch.randn(batch * tokens, 3 * d, generator=g) * 1.5).half().cuda()
    out = _native.attention_f16(qkv, batch, tokens, heads).float()
    q, k, v = qkv.float().reshape(batch, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    attn = ((q * 64 ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    ref = (attn @ v).transpose(1, 2).reshape(batch * tokens, d)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("arch_name", ["tiny-test", "tiny-test-reg"])
def test_vit_tiny_vs_golden_and_oracle(arch_name):
    from foundpose_b200.utils import dinov2_utils
    from oracle import vit as ovit

    gold = torch.load(GOLD_PATH, weights_only=False)
    meta = gold[f"vit/{arch_name}/meta"]
    arch = synthetic.VIT_ARCHS[arch_name]
    sd = synthetic.make_vit_state_dict(arch, seed=meta["wseed"])
    images = synthetic.make_crops(2, (meta["size"], meta["size"]), seed=meta["iseed"])
    name = f"dinov2_version={arch_name}_stride=14_facet=token_layer={meta['layer']}_norm=1"
    ext = dinov2_utils.DinoFeatureExtractor(name, state_dict=sd).to("cuda")
    out = ext(images.cuda())
    ref = gold[f"vit/{arch_name}/feature_maps"]
    assert out["feature_maps"].shape == ref.shape
    # fp16 operands vs fp32 reference: relative Frobenius error <= 5e-3 after <= 3 blocks.
    assert _rel(out["feature_maps"].cpu(), ref) <= 5e-3
    assert _rel(out["cls_tokens"].cpu(), gold[f"vit/{arch_name}/cls_tokens"]) <= 5e-3
    for facet in ("key", "value"):
        ext.facet = facet
        outf = ext(images.cuda())
        assert _rel(outf["feature_maps"].cpu(), gold[f"vit/{arch_name}/feature_maps_{facet}"]) <= 5e-3
    ext.facet = "token"
    # norm=0 variant against the oracle.
    ext.apply_norm = False
    o2 = ext(images.cuda())
    r2 = ovit.extract(sd, arch, images, layer=meta["layer"], apply_norm=False)
    assert _rel(o2["feature_maps"].cpu(), r2["feature_maps"]) <= 5e-3


@pytest.mark.parametrize("tag,batch", [("vits14-reg", 2), ("vitl14", 2)])
def test_vit_real_arch_vs_oracle(tag, batch):
    """ViT-S/14-reg and ViT-L/14 at 420x420, layer 9, against the fp32 CPU oracle (10 blocks).

    Tolerance: relative Frobenius error <= 1e-2, min per-token cosine similarity >= 0.9995.
    """
    from foundpose_b200.utils import dinov2_utils
    from oracle import vit as ovit

    gold = torch.load(GOLD_PATH, weights_only=False)
    meta = gold[f"vit/{tag}/meta"]
    opts = ovit.parse_extractor_name(meta["name"])
    arch = synthetic.VIT_ARCHS[opts["version"]]
    sd = synthetic.make_vit_state_dict(arch, seed=meta["wseed"], depth=opts["layer"] + 1)
    images = torch.cat([synthetic.make_crops(1, (420, 420), seed=meta["iseed"]),
                        synthetic.make_crops(batch - 1, (420, 420), seed=77)])
    ext = dinov2_utils.DinoFeatureExtractor(meta["name"], state_dict=sd).to("cuda")
    out = ext(images.cuda())
    fm = out["feature_maps"].cpu()
    assert fm.shape == (batch, arch.embed_dim, 30, 30)
    # crop 0 against the golden output of the reference's own modules
    assert _rel(fm[:1, ::8, ::3, ::3], gold[f"vit/{tag}/feature_maps_sub"]) <= 1e-2
    ref = ovit.extract(sd, arch, images, layer=opts["layer"], num_blocks=opts["layer"] + 1)["feature_maps"]
    assert _rel(fm, ref) <= 1e-2
    a = fm.permute(0, 2, 3, 1).reshape(-1, arch.embed_dim)
    b = ref.permute(0, 2, 3, 1).reshape(-1, arch.embed_dim)
    cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
    print(f"{tag}: rel err {_rel(fm, ref):.2e}, min cos {cos.min().item():.6f}")
    assert cos.min().item() >= 0.9995


def test_extractor_errors():
    from foundpose_b200.utils import dinov2_utils, feature_util

    arch = synthetic.VIT_ARCHS["tiny-test"]
    sd = synthetic.make_vit_state_dict(arch, seed=1)
    with pytest.raises(AssertionError):
        dinov2_utils.DinoFeatureExtractor("dino_tiny-test", state_dict=sd)
    with pytest.raises(NotImplementedError):
        feature_util.make_feature_extractor("resnet50")
    ext = dinov2_utils.DinoFeatureExtractor("dinov2_version=tiny-test_layer=1", state_dict=sd).to("cuda")
    with pytest.raises(AssertionError):  # H, W must be multiples of 14 (patch_embed.py:72-73)
        ext(torch.rand(1, 3, 50, 56, device="cuda"))
    with pytest.raises(ValueError):      # no CPU fallback
        ext(torch.rand(1, 3, 56, 56))


@pytest.mark.parametrize("tag", ["vits14-reg", "vitl14"])
def test_vit_numerics_with_pretrained_like_statistics(tag):
    """fp16 operands under DINOv2-checkpoint-like statistics (synthetic.make_vit_state_dict_realistic: residual
    outlier channels at 80-300, LayerScale 1e-5..1, LayerNorm gains 0.1..4, attention logits with std ~10 - the
    lazy O-rescale path of the attention kernel -, fc1 pre-activations with std ~3): outputs must be finite and
    within 1e-2 relative Frobenius error / 0.9995 per-token cosine of the fp32 oracle."""
    from foundpose_b200.utils import dinov2_utils
    from oracle import vit as ovit

    arch = synthetic.VIT_ARCHS[tag]
    layer = 9
    sd = synthetic.make_vit_state_dict_realistic(arch, seed=5, depth=layer + 1)
    images = synthetic.make_crops(2, (420, 420), seed=6)
    name = f"dinov2_version={tag}_stride=14_facet=token_layer={layer}_norm=1"
    ext = dinov2_utils.DinoFeatureExtractor(name, state_dict=sd).to("cuda")
    # the unnormalised block output too: the final LayerNorm would hide a saturated residual stream
    ext.apply_norm = False
    raw = ext(images.cuda())["feature_maps"].cpu()
    ext.apply_norm = True
    fm = ext(images.cuda())["feature_maps"].cpu()
    assert torch.isfinite(raw).all() and torch.isfinite(fm).all()
    ref_raw = ovit.extract(sd, arch, images, layer=layer, num_blocks=layer + 1, apply_norm=False)["feature_maps"]
    ref = ovit.extract(sd, arch, images, layer=layer, num_blocks=layer + 1)["feature_maps"]
    assert ref_raw.abs().max().item() > 50.0                      # the outlier channels really are there
    a = fm.permute(0, 2, 3, 1).reshape(-1, arch.embed_dim)
    b = ref.permute(0, 2, 3, 1).reshape(-1, arch.embed_dim)
    cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
    print(f"{tag} (pretrained-like statistics): rel err raw {_rel(raw, ref_raw):.2e}, normed {_rel(fm, ref):.2e}, "
          f"min cos {cos.min().item():.6f}, max |x| {ref_raw.abs().max().item():.0f}")
    assert _rel(raw, ref_raw) <= 1e-2 and _rel(fm, ref) <= 1e-2
    assert cos.min().item() >= 0.9995


def test_checkpoint_file_is_loaded_through_the_env_hook(tmp_path, monkeypatch):
    """FOUNDPOSE_DINOV2_WEIGHTS=<file in the official checkpoint layout> (utils/dinov2_utils.py:82-84 downloads
    the same state_dict): the extractor built from the file == the extractor given the state_dict directly."""
    from foundpose_b200.utils import dinov2_utils, feature_util

    arch = synthetic.VIT_ARCHS["vits14-reg"]
    sd = synthetic.make_vit_state_dict(arch, seed=3)               # all 12 blocks + mask_token, as published
    path = tmp_path / "dinov2_vits14_reg4_pretrain.pth"
    torch.save(sd, str(path))
    monkeypatch.setenv("FOUNDPOSE_DINOV2_WEIGHTS", str(path))
    name = "dinov2_version=vits14-reg_stride=14_facet=token_layer=9_norm=1"
    from_file = feature_util.make_feature_extractor(name).to("cuda")
    monkeypatch.delenv("FOUNDPOSE_DINOV2_WEIGHTS")
    direct = dinov2_utils.DinoFeatureExtractor(name, state_dict=sd).to("cuda")
    assert set(from_file.model.state_dict()) == set(direct.model.state_dict())
    images = synthetic.make_crops(1, (420, 420), seed=4).cuda()
    assert torch.equal(from_file(images)["feature_maps"], direct(images)["feature_maps"])


@pytest.mark.parametrize("realistic", [False, True])
def test_layernorm_fused_into_the_gemms_matches_the_separate_kernels(monkeypatch, realistic):
    """ViT-L/14: norm1 / norm2 folded into the qkv / fc1 GEMMs (raw fp16 residual rows as the operand, (mu, rstd)
    applied in the epilogue, statistics from the partial sums the proj / fc2 / patch-embed epilogues emit) against
    the separate LayerNorm kernels and against the fp32 oracle - also with pretrained-like statistics, where the
    raw rows carry outlier channels at 80-300."""
    from foundpose_b200.utils import dinov2_utils
    from oracle import vit as ovit

    arch = synthetic.VIT_ARCHS["vitl14"]
    make = synthetic.make_vit_state_dict_realistic if realistic else synthetic.make_vit_state_dict
    sd = make(arch, seed=13, depth=10)
    images = synthetic.make_crops(2, (420, 420), seed=14)
    name = "dinov2_version=vitl14_stride=14_facet=token_layer=9_norm=1"
    monkeypatch.setenv("FOUNDPOSE_FUSE_LAYERNORM", "1")
    fused = dinov2_utils.DinoFeatureExtractor(name, state_dict=sd).to("cuda")
    out_f = fused(images.cuda())
    assert fused._native_cache[(420, 420, "cuda:0")]["fuse_ln"] is True
    monkeypatch.setenv("FOUNDPOSE_FUSE_LAYERNORM", "0")
    plain = dinov2_utils.DinoFeatureExtractor(name, state_dict=sd).to("cuda")
    out_p = plain(images.cuda())
    assert plain._native_cache[(420, 420, "cuda:0")]["fuse_ln"] is False
    ref = ovit.extract(sd, arch, images, layer=9, num_blocks=10)
    e_f = _rel(out_f["feature_maps"].cpu(), ref["feature_maps"])
    e_p = _rel(out_p["feature_maps"].cpu(), ref["feature_maps"])
    print(f"vitl14 realistic={realistic}: rel err fused {e_f:.2e}, separate {e_p:.2e}, fused vs separate "
          f"{_rel(out_f['feature_maps'].cpu(), out_p['feature_maps'].cpu()):.2e}")
    assert torch.isfinite(out_f["feature_maps"]).all()
    assert e_f <= 1e-2 and e_p <= 1e-2
    assert _rel(out_f["cls_tokens"].cpu(), ref["cls_tokens"]) <= 1e-2
    # deterministic: the partial sums are combined in a fixed order (no atomics)
    assert torch.equal(fused(images.cuda())["feature_maps"], out_f["feature_maps"])
