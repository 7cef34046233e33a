"""Seeded synthetic inputs for the FoundPose hot path.

There is no network in the build or bench environment, so neither the DINOv2 checkpoints nor the
BOP datasets nor the hosted `repre.pth` banks are available.  Everything the parity tests and
`bench.py` feed to the path is generated here from `torch.Generator().manual_seed(seed)` on the
CPU (bit-reproducible across machines running the same torch build):

  * ViT weights with the DINOv2 state_dict layout (`blocks.{i}.norm1.weight`, `attn.qkv.weight`,
    `ls1.gamma`, `pos_embed`, ... - reference external/dinov2/dinov2/models/vision_transformer.py:44-170)
  * crops in [0, 1) and object masks
  * a PCA projector (orthonormal components + mean)
  * an object representation ("bank"): feat_vectors, feat_to_template_ids, vertices, visual-word
    centroids; idfs / template descriptors are then produced by `template_util.calc_tfidf_descriptors`.

Feature-like tensors are rounded to fp16-representable values (and kept as fp32) so that the fp16
storage used on the GPU is lossless and the fp32 CPU oracle sees bit-identical inputs.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

import torch


@dataclass(frozen=True)
class VitArch:
    """Architecture constants of a DINOv2 backbone (hub/backbones.py:18-61, vision_transformer.py:340-396)."""

    name: str
    embed_dim: int
    depth: int
    num_heads: int
    num_register_tokens: int = 0
    patch_size: int = 14
    img_size: int = 518
    mlp_ratio: int = 4
    interpolate_antialias: bool = False
    interpolate_offset: float = 0.1

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def num_pos(self) -> int:
        return (self.img_size // self.patch_size) ** 2


VIT_ARCHS: Dict[str, VitArch] = {
    "vits14": VitArch("vits14", 384, 12, 6),
    "vitb14": VitArch("vitb14", 768, 12, 12),
    "vitl14": VitArch("vitl14", 1024, 24, 16),
    "vits14-reg": VitArch("vits14-reg", 384, 12, 6, 4, interpolate_antialias=True, interpolate_offset=0.0),
    "vitb14-reg": VitArch("vitb14-reg", 768, 12, 12, 4, interpolate_antialias=True, interpolate_offset=0.0),
    "vitl14-reg": VitArch("vitl14-reg", 1024, 24, 16, 4, interpolate_antialias=True, interpolate_offset=0.0),
    # Tiny architectures used only by the tests (fast on the CPU oracle, small golden fixtures).
    "tiny-test": VitArch("tiny-test", 128, 3, 2, 0, img_size=98),
    "tiny-test-reg": VitArch("tiny-test-reg", 128, 3, 2, 4, img_size=98, interpolate_antialias=True,
                             interpolate_offset=0.0),
}


def fp16_representable(x: torch.Tensor) -> torch.Tensor:
    """Rounds to the nearest fp16 value but keeps fp32 storage."""
    return x.to(torch.float16).to(torch.float32)


def _gen(seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(int(seed))


def make_vit_state_dict(arch: VitArch, seed: int = 0, depth: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Random DINOv2-layout weights.

    Unlike a fresh `pretrained=False` hub model (zero biases, unit LayerNorm/LayerScale) every
    parameter is non-trivial so that parity tests exercise biases, LayerNorm affine terms and
    LayerScale.  Magnitudes follow the reference init (trunc-normal std 0.02 for Linear weights and
    pos_embed, vision_transformer.py:171-177, 386-391).
    """
    g = _gen(seed)
    d = arch.embed_dim
    hidden = arch.mlp_ratio * d
    n_blocks = arch.depth if depth is None else depth

    def tn(*shape, std=0.02):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=std, a=-2 * std, b=2 * std, generator=g)

    def nrm(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    sd["cls_token"] = nrm(1, 1, d, std=0.02)
    sd["pos_embed"] = tn(1, arch.num_pos + 1, d)
    if arch.num_register_tokens:
        sd["register_tokens"] = nrm(1, arch.num_register_tokens, d, std=0.02)
    sd["mask_token"] = torch.zeros(1, d)
    k = 3 * arch.patch_size * arch.patch_size
    sd["patch_embed.proj.weight"] = nrm(d, 3, arch.patch_size, arch.patch_size, std=1.0 / math.sqrt(k))
    sd["patch_embed.proj.bias"] = nrm(d, std=0.02)
    # Drawn before the blocks so that a depth-truncated state dict is a prefix-consistent subset.
    norm_w = nrm(d, std=0.1, mean=1.0)
    norm_b = nrm(d, std=0.05)
    for i in range(n_blocks):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = nrm(d, std=0.1, mean=1.0)
        sd[p + "norm1.bias"] = nrm(d, std=0.05)
        sd[p + "attn.qkv.weight"] = tn(3 * d, d, std=0.04)
        sd[p + "attn.qkv.bias"] = nrm(3 * d, std=0.02)
        sd[p + "attn.proj.weight"] = tn(d, d)
        sd[p + "attn.proj.bias"] = nrm(d, std=0.02)
        sd[p + "ls1.gamma"] = nrm(d, std=0.1, mean=0.8)
        sd[p + "norm2.weight"] = nrm(d, std=0.1, mean=1.0)
        sd[p + "norm2.bias"] = nrm(d, std=0.05)
        sd[p + "mlp.fc1.weight"] = tn(hidden, d)
        sd[p + "mlp.fc1.bias"] = nrm(hidden, std=0.02)
        sd[p + "mlp.fc2.weight"] = tn(d, hidden)
        sd[p + "mlp.fc2.bias"] = nrm(d, std=0.02)
        sd[p + "ls2.gamma"] = nrm(d, std=0.1, mean=0.8)
    sd["norm.weight"] = norm_w
    sd["norm.bias"] = norm_b
    return sd


def make_vit_state_dict_realistic(arch: VitArch, seed: int = 0, depth: Optional[int] = None
                                  ) -> Dict[str, torch.Tensor]:
    """Random weights with the statistics that make PRETRAINED DINOv2 checkpoints numerically harder than a
    trunc-normal init (there is no network to fetch a real checkpoint):

      * "massive activation" channels: three residual channels sit at magnitude ~80-300 in every token (written by
        the patch-embed bias and re-inforced by fc2 biases), two orders of magnitude above the rest;
      * LayerScale gammas log-uniform in [1e-5, 1] (DINOv2 initialises them at 1e-5 and training spreads them);
      * LayerNorm gains spread over [0.1, 4];
      * peaked attention: q / k projections scaled so that the logits have a standard deviation of ~10 (the row
        maximum moves by far more than 2^8 between key blocks: the lazy-rescale path of the attention kernel);
      * fc1 pre-activations with a standard deviation of ~3 (GELU well outside its linear region).
    """
    sd = make_vit_state_dict(arch, seed=seed, depth=depth)
    g = _gen(seed + 7919)
    d = arch.embed_dim
    n_blocks = arch.depth if depth is None else depth
    outliers = torch.randperm(d, generator=g)[:3]
    sd["patch_embed.proj.bias"][outliers] = torch.tensor([300.0, -150.0, 80.0])
    for i in range(n_blocks):
        p = f"blocks.{i}."
        for ls in ("ls1.gamma", "ls2.gamma"):
            sd[p + ls] = torch.exp(torch.empty(d).uniform_(math.log(1e-5), 0.0, generator=g))
        for nm in ("norm1.weight", "norm2.weight"):
            sd[p + nm] = torch.exp(torch.empty(d).uniform_(math.log(0.1), math.log(4.0), generator=g))
        sd[p + "attn.qkv.weight"][: 2 * d] *= 4.0          # q and k rows: logits ~ 16x larger
        sd[p + "mlp.fc1.weight"] *= 3.0
        sd[p + "mlp.fc2.bias"][outliers] += torch.tensor([20.0, -10.0, 5.0])
    return sd


def make_crops(batch: int, size: Tuple[int, int] = (420, 420), seed: int = 0) -> torch.Tensor:
    """B x 3 x H x W float32 in [0, 1) (scripts/infer.py:398 scales uint8 images the same way)."""
    w, h = size
    return torch.rand(batch, 3, h, w, generator=_gen(seed))


def make_masks(batch: int, size: Tuple[int, int] = (420, 420), seed: int = 0, full: bool = False) -> torch.Tensor:
    """B x H x W bool object masks: all-ones (throughput runs) or seeded ellipses (correctness runs)."""
    w, h = size
    if full:
        return torch.ones(batch, h, w, dtype=torch.bool)
    g = _gen(seed)
    ys = torch.arange(h, dtype=torch.float32).view(1, h, 1) + 0.5
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, w) + 0.5
    cx = (0.35 + 0.3 * torch.rand(batch, 1, 1, generator=g)) * w
    cy = (0.35 + 0.3 * torch.rand(batch, 1, 1, generator=g)) * h
    rx = (0.15 + 0.3 * torch.rand(batch, 1, 1, generator=g)) * w
    ry = (0.15 + 0.3 * torch.rand(batch, 1, 1, generator=g)) * h
    return ((xs - cx) / rx) ** 2 + ((ys - cy) / ry) ** 2 <= 1.0


def make_pca(in_dim: int, out_dim: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Tensordict of a PCA projector (keys of utils/projector_util.py:91-113)."""
    g = _gen(seed)
    q, _ = torch.linalg.qr(torch.randn(in_dim, in_dim, generator=g))
    components = fp16_representable(q[:, :out_dim].t().contiguous())
    mean = fp16_representable(torch.randn(in_dim, generator=g) * 0.1)
    ev = torch.linspace(2.0, 0.1, out_dim)
    return {
        "pca_projector": {
            "components": components,
            "explained_variance": ev,
            "explained_variance_ratio": ev / ev.sum(),
            "singular_values": torch.sqrt(ev * 1000.0),
            "mean": mean,
            "noise_variance": torch.tensor(0.05),
            "whiten": torch.tensor(False),
        }
    }


def make_bank_tensors(
    num_templates: int,
    patches_per_template: int,
    feat_dim: int,
    num_words: int = 2048,
    seed: int = 0,
    ragged: bool = False,
) -> Dict[str, torch.Tensor]:
    """Raw tensors of a synthetic object representation (fields of utils/repre_util.py:34-83).

    feat_to_template_ids is built as contiguous ascending runs exactly as
    scripts/gen_repre.py:187-190, 214 does by concatenation.  With `ragged=True` the runs have
    different lengths (as real eroded-mask templates do).
    """
    g = _gen(seed)
    if ragged:
        lens = torch.randint(max(1, patches_per_template // 2), patches_per_template + 1,
                             (num_templates,), generator=g)
    else:
        lens = torch.full((num_templates,), patches_per_template, dtype=torch.int64)
    total = int(lens.sum())
    feat_vectors = fp16_representable(torch.randn(total, feat_dim, generator=g))
    feat_to_template_ids = torch.repeat_interleave(torch.arange(num_templates, dtype=torch.int32), lens)
    vertices = torch.randn(total, 3, generator=g)
    perm = torch.randperm(total, generator=g)[:num_words]
    centroids = feat_vectors[perm].clone()
    return {
        "feat_vectors": feat_vectors,
        "feat_to_template_ids": feat_to_template_ids,
        "feat_to_vertex_ids": torch.arange(total, dtype=torch.int32),
        "vertices": vertices,
        "feat_cluster_centroids": centroids,
    }


def make_query_features(num_queries: int, feat_dim: int, bank_feats: Optional[torch.Tensor] = None,
                        seed: int = 0, noise: float = 0.25) -> torch.Tensor:
    """Query descriptors: noisy copies of random bank rows (so matching is non-trivial) or randn."""
    g = _gen(seed)
    if bank_feats is None:
        return fp16_representable(torch.randn(num_queries, feat_dim, generator=g))
    ids = torch.randint(0, bank_feats.shape[0], (num_queries,), generator=g)
    q = bank_feats[ids] + noise * torch.randn(num_queries, feat_dim, generator=g)
    return fp16_representable(q)


def _rigid_transform(rng: np.random.RandomState, max_angle: float = 0.6) -> np.ndarray:
    """Seeded 4x4 camera-to-world rigid transform (Rodrigues rotation + translation), float64."""
    axis = rng.randn(3)
    axis /= np.linalg.norm(axis)
    angle = rng.uniform(-max_angle, max_angle)
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + math.sin(angle) * k + (1 - math.cos(angle)) * (k @ k)
    T[:3, 3] = rng.uniform(-0.5, 0.5, size=3)
    return T


def make_scene_image(height: int, width: int, seed: int = 0) -> np.ndarray:
    """Seeded uint8 HWC image: smooth colour gradients plus per-pixel noise (so interpolation matters)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    chans = []
    for c in range(3):
        a, b, ph = rng.uniform(0.02, 0.15, size=3)
        chans.append(127 + 80 * np.sin(a * xx + ph * 10) * np.cos(b * yy) + rng.randint(-30, 31, size=(height, width)))
    return np.clip(np.stack(chans, axis=-1), 0, 255).astype(np.uint8)


def make_instance_mask(height: int, width: int, center: Tuple[float, float], radii: Tuple[float, float],
                       seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Seeded elliptical uint8 {0,1} modal mask with a few holes, and its amodal box (left, top, right, bottom)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    inside = ((xx - center[0]) / radii[0]) ** 2 + ((yy - center[1]) / radii[1]) ** 2 <= 1.0
    holes = rng.rand(height, width) < 0.05
    mask = (inside & ~holes).astype(np.uint8)
    box = np.array([center[0] - radii[0], center[1] - radii[1], center[0] + radii[0], center[1] + radii[1]])
    return mask, box


def make_crop_cases() -> List[Dict[str, object]]:
    """Small crop-stage cases (shared by tests/golden/make_golden_crop.py and the parity tests)."""
    cases: List[Dict[str, object]] = []
    specs = [
        # (center, radii, crop_size (W, H), rel_pad, identity pose?)   what it exercises
        ((60.0, 45.0), (20.0, 15.0), (56, 56), 0.2, True),     # magnification -> INTER_LINEAR
        ((64.0, 48.0), (50.0, 40.0), (28, 28), 0.2, False),    # minification -> INTER_AREA (same remap path)
        ((8.0, 20.0), (18.0, 16.0), (42, 42), 0.3, False),     # box leaves the image -> constant border
        ((100.0, 70.0), (22.0, 24.0), (56, 42), 0.1, False),   # non-square viewport, fx != fy
    ]
    for i, (center, radii, crop_size, pad, identity) in enumerate(specs):
        rng = np.random.RandomState(100 + i)
        mask, box = make_instance_mask(96, 128, center, radii, seed=200 + i)
        cases.append({
            "image": make_scene_image(96, 128, seed=300 + i), "mask": mask, "box": box,
            "f": (110.0, 110.0) if i < 3 else (120.0, 104.0), "c": (63.5, 47.5),
            "T_world_from_eye": np.eye(4) if identity else _rigid_transform(rng),
            "crop_size": crop_size, "crop_rel_pad": pad})
    return cases
