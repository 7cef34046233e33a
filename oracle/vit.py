"""Oracle: DINOv2 ViT patch-feature extraction, fp32 on the CPU.  TEST INFRASTRUCTURE ONLY.

Functional restatement (state_dict in, tensors out) of
  * utils/dinov2_utils.py:32-158, 232-311       DinoFeatureExtractor (name grammar, ImageNet
                                                normalise, hook on block `layer`, final LayerNorm)
  * external/dinov2/dinov2/models/vision_transformer.py:179-232, 254-270   token preparation
  * external/dinov2/dinov2/layers/{patch_embed.py:68-81, block.py:89-114, attention.py:56-69,
    mlp.py:34-40, layer_scale.py:26-27}

The same torch ops are applied in the same order as the reference, so on identical weights the
output is bit-identical to the reference modules (checked by tests/golden/make_golden.py).
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def parse_extractor_name(model_name: str) -> Dict[str, object]:
    """utils/dinov2_utils.py:52-78: defaults and the two supported name formats."""
    opts = {"version": "vits14-reg", "stride": 14, "facet": "token", "layer": 9, "norm": True}
    items = model_name.split("_")
    assert items[0] == "dinov2"
    if len(items) == 2:
        opts["version"] = items[1]
    else:
        for item in items[1:]:
            name, value = item.split("=")
            if name == "version":
                opts["version"] = value
            elif name == "stride":
                opts["stride"] = int(value)
            elif name == "facet":
                opts["facet"] = value
            elif name == "layer":
                opts["layer"] = int(value)
            elif name == "norm":
                opts["norm"] = bool(int(value))
    return opts


def interpolate_pos_encoding(pos_embed: torch.Tensor, w: int, h: int, patch_size: int,
                             interpolate_offset: float, interpolate_antialias: bool) -> torch.Tensor:
    """vision_transformer.py:179-211 (w, h are image dims 2 and 3 as the reference names them)."""
    n = pos_embed.shape[1] - 1
    w0 = w // patch_size
    h0 = h // patch_size
    if w0 * h0 == n and w == h:
        return pos_embed
    pos_embed = pos_embed.float()
    class_pos_embed = pos_embed[:, 0]
    patch_pos_embed = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    m = int(math.sqrt(n))
    assert n == m * m
    kwargs = {}
    if interpolate_offset:
        sx = float(w0 + interpolate_offset) / m
        sy = float(h0 + interpolate_offset) / m
        kwargs["scale_factor"] = (sx, sy)
    else:
        kwargs["size"] = (w0, h0)
    patch_pos_embed = F.interpolate(
        patch_pos_embed.reshape(1, m, m, dim).permute(0, 3, 1, 2),
        mode="bicubic",
        antialias=interpolate_antialias,
        **kwargs,
    )
    assert (w0, h0) == patch_pos_embed.shape[-2:]
    patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).view(1, -1, dim)
    return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)


def prepare_tokens(sd: Dict[str, torch.Tensor], arch, x: torch.Tensor) -> torch.Tensor:
    """vision_transformer.py:213-232 + patch_embed.py:68-81."""
    _, _, w, h = x.shape
    ps = arch.patch_size
    assert w % ps == 0, f"Input image height {w} is not a multiple of patch height {ps}"
    assert h % ps == 0, f"Input image width {h} is not a multiple of patch width: {ps}"
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=ps)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(t.shape[0], -1, -1), t), dim=1)
    t = t + interpolate_pos_encoding(sd["pos_embed"], w, h, ps, arch.interpolate_offset,
                                     arch.interpolate_antialias)
    if arch.num_register_tokens:
        t = torch.cat((t[:, :1], sd["register_tokens"].expand(t.shape[0], -1, -1), t[:, 1:]), dim=1)
    return t


def attention(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """attention.py:56-69 (the non-xformers path; xformers is absent here and on the GPU box)."""
    b, n, c = x.shape
    qkv = F.linear(x, sd[prefix + "qkv.weight"], sd[prefix + "qkv.bias"])
    qkv = qkv.reshape(b, n, 3, num_heads, c // num_heads).permute(2, 0, 3, 1, 4)
    scale = (c // num_heads) ** -0.5
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def block(sd: Dict[str, torch.Tensor], i: int, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """block.py:89-114 eval path: x + ls1(attn(norm1(x))); x + ls2(mlp(norm2(x)))."""
    p = f"blocks.{i}."
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-6)
    y = attention(sd, p + "attn.", y, num_heads)
    x = x + y * sd[p + "ls1.gamma"]
    y = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-6)
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    x = x + y * sd[p + "ls2.gamma"]
    return x


def normalize_images(images: torch.Tensor) -> torch.Tensor:
    """torchvision T.Normalize(mean, std) (dinov2_utils.py:111-113, 123)."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=images.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=images.dtype).view(1, 3, 1, 1)
    return (images - mean) / std


@torch.no_grad()
def extract(sd: Dict[str, torch.Tensor], arch, images: torch.Tensor, layer: int = 9,
            facet: str = "token", apply_norm: bool = True, full_depth: bool = False,
            num_blocks: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """DinoFeatureExtractor.forward (dinov2_utils.py:115-158) on B x 3 x H x W images in [0, 1].

    full_depth=True executes every block after `layer` too, exactly as the reference does
    (its forward hook cannot stop the model - SURVEY.md S4); the result is identical, only the
    cost differs.  It is used for the CPU-baseline timing.
    """
    assert facet in ["key", "query", "value", "token"], f"{facet} is not a supported facet for descriptors."
    x = normalize_images(images)
    bsz, _, w, h = x.shape
    depth = arch.depth if num_blocks is None else num_blocks
    tokens = prepare_tokens(sd, arch, x)
    hooked = None
    for i in range(depth):
        if i == layer and facet != "token":
            # dinov2_utils.py:184-196: hook on block.attn recomputes qkv from the attn input.
            p = f"blocks.{i}."
            d = tokens.shape[-1]
            y = F.layer_norm(tokens, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-6)
            qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
            b, n, _ = y.shape
            qkv = qkv.reshape(b, n, 3, arch.num_heads, d // arch.num_heads).permute(2, 0, 3, 1, 4)
            hooked = qkv[{"query": 0, "key": 1, "value": 2}[facet]]  # B x h x t x d
        tokens = block(sd, i, tokens, arch.num_heads)
        if i == layer and facet == "token":
            hooked = tokens.unsqueeze(1)  # B x 1 x t x d
        if i == layer and not full_depth:
            break
    assert hooked is not None, f"layer {layer} is out of range for depth {depth}"
    xh = hooked
    cls_tokens = xh[:, :, [0], :].permute(0, 2, 3, 1).flatten(start_dim=-2, end_dim=-1)  # B x 1 x (d*h)
    xh = xh[:, :, (arch.num_register_tokens + 1):, :]
    patch_tokens = xh.permute(0, 2, 3, 1).flatten(start_dim=-2, end_dim=-1)  # B x t x (d*h)
    if apply_norm:
        d = patch_tokens.shape[-1]
        toks = torch.cat([cls_tokens, patch_tokens], dim=1)
        toks = F.layer_norm(toks, (d,), sd["norm.weight"], sd["norm.bias"], eps=1e-6)
        cls_tokens = toks[:, :1, :]
        patch_tokens = toks[:, 1:, :]
    d = patch_tokens.shape[-1]
    ps = arch.patch_size
    num_patches = (1 + (h - ps) // ps, 1 + (w - ps) // ps)
    feature_maps = patch_tokens.reshape(bsz, num_patches[1], num_patches[0], d).permute(0, 3, 1, 2)
    return {"cls_tokens": cls_tokens[:, 0, :], "feature_maps": feature_maps}
