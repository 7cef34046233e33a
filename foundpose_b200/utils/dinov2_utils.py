"""DINOv2 feature extractor on the B200 kernels - drop-in for the reference's utils/dinov2_utils.py.

Same class name, constructor grammar, call signature and outputs as
`DinoFeatureExtractor` (reference utils/dinov2_utils.py:25-158):

    extractor = DinoFeatureExtractor("dinov2_vitl14")      # or the key=value form
    extractor.to("cuda")
    out = extractor(images_bchw)      # {"cls_tokens": BxD, "feature_maps": BxDxHpxWp}

The transformer runs through `fp_vit_forward` (hand-written sm_100a kernels: tcgen05 GEMMs with
fused bias/GELU/LayerScale/residual epilogues, tcgen05 flash attention, fused LayerNorm).  Only
blocks 0..layer are executed; the reference executes all blocks and discards the rest
 - the result is identical.

Weights: `self.model` holds parameters with the exact DINOv2 state_dict layout
(external/dinov2/dinov2/models/vision_transformer.py), so an official checkpoint loads with
`extractor.model.load_state_dict(torch.load("dinov2_vitl14_pretrain.pth"))`.  Like the reference
(`pretrained=True`, dinov2_utils.py:82-84) the constructor tries to fetch that checkpoint; in an
offline environment set FOUNDPOSE_DINOV2_WEIGHTS=<file> or pass `state_dict=`.
"""

from __future__ import annotations

import math
import os
import typing as tp
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from foundpose_b200 import _native, synthetic
from foundpose_b200.utils import logging

logger: logging.Logger = logging.get_logger()

_DINOV2_BASE_URL = "https://dl.fbaipublicfiles.com/dinov2"
_FACETS = {"token": 0, "query": 1, "key": 2, "value": 3}
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class _ParamTree(nn.Module):
    """A bare parameter container reproducing dotted state_dict names (`blocks.0.attn.qkv.weight`)."""

    def __init__(self) -> None:
        super().__init__()

    def add(self, dotted: str, value: torch.Tensor) -> None:
        head, _, tail = dotted.partition(".")
        if not tail:
            self.register_parameter(head, nn.Parameter(value, requires_grad=False))
            return
        if head not in self._modules:
            self.add_module(head, _ParamTree())
        self._modules[head].add(tail, value)


class Dinov2Weights(_ParamTree):
    """Stands in for `DinoVisionTransformer` as the owner of the weights (`extractor.model`)."""

    def __init__(self, arch: synthetic.VitArch, state_dict: Dict[str, torch.Tensor]) -> None:
        super().__init__()
        self.arch = arch
        self.embed_dim = self.num_features = arch.embed_dim
        self.num_heads = arch.num_heads
        self.n_blocks = arch.depth
        self.patch_size = arch.patch_size
        self.num_register_tokens = arch.num_register_tokens
        self.interpolate_antialias = arch.interpolate_antialias
        self.interpolate_offset = arch.interpolate_offset
        for name, value in state_dict.items():
            self.add(name, value.detach().clone().to(torch.float32))

    def interpolate_pos_encoding(self, w: int, h: int) -> torch.Tensor:
        """external/dinov2/dinov2/models/vision_transformer.py:179-211 (same torch ops)."""
        pos_embed = self.pos_embed.float()
        n = pos_embed.shape[1] - 1
        w0, h0 = w // self.patch_size, h // self.patch_size
        if w0 * h0 == n and w == h:
            return pos_embed
        class_pos_embed = pos_embed[:, 0]
        patch_pos_embed = pos_embed[:, 1:]
        dim = pos_embed.shape[-1]
        m = int(math.sqrt(n))
        assert n == m * m
        kwargs = {}
        if self.interpolate_offset:
            kwargs["scale_factor"] = (float(w0 + self.interpolate_offset) / m,
                                      float(h0 + self.interpolate_offset) / m)
        else:
            kwargs["size"] = (w0, h0)
        # Init-time only (once per image size); evaluated on the CPU so the result is bit-identical
        # to the reference's CPU path regardless of the device the extractor lives on.
        patch_pos_embed = nn.functional.interpolate(
            patch_pos_embed.detach().cpu().reshape(1, m, m, dim).permute(0, 3, 1, 2),
            mode="bicubic", antialias=self.interpolate_antialias, **kwargs)
        assert (w0, h0) == patch_pos_embed.shape[-2:]
        patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).reshape(1, -1, dim).to(pos_embed.device)
        return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)


def _load_pretrained_state_dict(model_base_name: str, arch: synthetic.VitArch) -> Dict[str, torch.Tensor]:
    path = os.environ.get("FOUNDPOSE_DINOV2_WEIGHTS")
    if path:
        logger.info(f"Loading DINOv2 weights from: {path}")
        return torch.load(path, map_location="cpu")
    seed = os.environ.get("FOUNDPOSE_SYNTHETIC_WEIGHTS")
    if seed is not None:
        logger.info(f"Using seeded synthetic DINOv2 weights (seed {seed}); no checkpoint available offline.")
        return synthetic.make_vit_state_dict(arch, seed=int(seed))
    compact = model_base_name.replace("_reg", "")
    full = compact + ("_reg4" if arch.num_register_tokens else "")
    url = f"{_DINOV2_BASE_URL}/{compact}/{full}_pretrain.pth"
    try:
        return torch.hub.load_state_dict_from_url(url, map_location="cpu")
    except Exception as e:  # no network
        raise RuntimeError(
            f"Cannot fetch {url} ({e}). Set FOUNDPOSE_DINOV2_WEIGHTS=<checkpoint file>, pass "
            "state_dict=..., or set FOUNDPOSE_SYNTHETIC_WEIGHTS=<seed> for seeded random weights."
        ) from e


def parse_model_name(model_name: str) -> Dict[str, tp.Any]:
    """Name grammar of the reference (utils/dinov2_utils.py:52-78): defaults + the two formats."""
    opts: Dict[str, tp.Any] = {"version": "vits14-reg", "stride": 14, "facet": "token", "layer": 9, "norm": True}
    name_items = model_name.split("_")
    assert name_items[0] == "dinov2"
    if len(name_items) == 2:
        # Example: "dinov2_vits14"
        opts["version"] = name_items[1]
    else:
        # Example: "dinov2_version=vitl14_stride=14_facet=key_layer=18_norm=1"
        for item in name_items[1:]:
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


class DinoFeatureExtractor(nn.Module):
    """DINOv2 feature extractor (B200-native)."""

    def __init__(self, model_name: str, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 max_batch: int = 64) -> None:
        super().__init__()
        opts = parse_model_name(model_name)
        self.version: str = opts["version"]
        self.stride: int = opts["stride"]
        self.facet: str = opts["facet"]
        self.layer: int = opts["layer"]
        self.apply_norm: bool = opts["norm"]

        if self.version not in synthetic.VIT_ARCHS:
            raise KeyError(f"Unknown DINOv2 version '{self.version}' (ViT-g / SwiGLU is not on the FoundPose path).")
        self.arch = synthetic.VIT_ARCHS[self.version]
        self.model_base_name: str = f"dinov2_{self.version}".replace("-", "_")
        if state_dict is None:
            state_dict = _load_pretrained_state_dict(self.model_base_name, self.arch)
        self.model = Dinov2Weights(self.arch, state_dict)

        if self.stride != 14:
            # patch_vit_resolution (reference :363-389) re-strides the patch conv, but the method it
            # installs (:324, bound at :386) takes one positional too few and raises TypeError on the
            # first forward, so there is no reference behaviour to reproduce.
            raise NotImplementedError("foundpose_b200 supports the DINOv2 training stride (14) only.")
        self.patch_size: int = self.arch.patch_size
        self.max_batch = max_batch
        self.num_patches: Optional[Tuple[int, int]] = None
        self._native_cache: Dict[Tuple[int, int, str], dict] = {}
        self.eval()

    # -- native handle management ---------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        self._release_native()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._release_native()
        return super().load_state_dict(*args, **kwargs)

    def _release_native(self) -> None:
        for entry in getattr(self, "_native_cache", {}).values():
            _native.vit_destroy(entry["handle"])
        self._native_cache = {}

    def __del__(self):
        try:
            self._release_native()
        except Exception:
            pass

    def _prepare(self, h: int, w: int, device: torch.device) -> dict:
        key = (h, w, str(device))
        if key in self._native_cache:
            return self._native_cache[key]
        ps = self.patch_size
        assert h % ps == 0, f"Input image height {h} is not a multiple of patch height {ps}"
        assert w % ps == 0, f"Input image width {w} is not a multiple of patch width: {ps}"
        m = self.model
        d = self.arch.embed_dim
        n_blocks = self.layer + 1
        assert n_blocks <= self.arch.depth, f"layer {self.layer} is out of range for depth {self.arch.depth}"
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        keep: List[torch.Tensor] = []

        def f32(t: torch.Tensor) -> torch.Tensor:
            t = t.to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t

        def f16(t: torch.Tensor) -> torch.Tensor:
            t = t.to(device=device, dtype=torch.float16).contiguous()
            keep.append(t)
            return t

        kdim = 3 * ps * ps
        kpad = (kdim + 63) // 64 * 64
        patch_w = torch.zeros(d, kpad, dtype=torch.float32)
        patch_w[:, :kdim] = sd["patch_embed.proj.weight"].reshape(d, kdim).cpu()
        # Note the reference's (w, h) naming: `B, nc, w, h = x.shape` (vision_transformer.py:214).
        pos = m.interpolate_pos_encoding(h, w)[0]  # [1 + P, D]
        weights = _native.VitWeights()
        weights.patch_w = f16(patch_w).data_ptr()
        weights.patch_b = f32(sd["patch_embed.proj.bias"]).data_ptr()
        weights.cls_pos = f32(sd["cls_token"].reshape(d) + pos[0].to(sd["cls_token"].device)).data_ptr()
        weights.reg_tokens = (f32(sd["register_tokens"].reshape(-1, d)).data_ptr()
                              if self.arch.num_register_tokens else None)
        weights.pos_patch = f32(pos[1:]).data_ptr()
        weights.norm_w = f32(sd["norm.weight"]).data_ptr()
        weights.norm_b = f32(sd["norm.bias"]).data_ptr()
        # LayerNorm fused into the GEMMs around it (csrc/kernels.h, EPI_LN_*): needs the pair GEMM kernel, i.e.
        # D % 256 == 0 and more than 256 patch rows per call (always true at 420 x 420).
        #   LN(x) W^T + b = rstd (x W'^T - mu colsum(W')) + b',  W' = W diag(gamma), b' = b + W beta
        fuse_ln = (d % 256 == 0 and (h // ps) * (w // ps) > 256
                   and os.environ.get("FOUNDPOSE_FUSE_LAYERNORM", "1") != "0")

        def fold(weight: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
            w64 = weight.double().cpu()
            w16 = (w64 * gamma.double().cpu()[None, :]).to(torch.float32).to(torch.float16)
            colsum = w16.double().sum(dim=1).to(torch.float32)       # of the fp16 values the GEMM multiplies
            b_fold = (bias.double().cpu() + w64 @ beta.double().cpu()).to(torch.float32)
            return f16(w16).data_ptr(), f32(b_fold).data_ptr(), f32(colsum).data_ptr()

        blocks = []
        for i in range(n_blocks):
            p = f"blocks.{i}."
            bw = _native.VitBlockWeights()
            bw.norm1_w = f32(sd[p + "norm1.weight"]).data_ptr()
            bw.norm1_b = f32(sd[p + "norm1.bias"]).data_ptr()
            if fuse_ln:
                bw.qkv_w, bw.qkv_b, bw.qkv_colsum = fold(sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"],
                                                         sd[p + "norm1.weight"], sd[p + "norm1.bias"])
                bw.fc1_w, bw.fc1_b, bw.fc1_colsum = fold(sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"],
                                                         sd[p + "norm2.weight"], sd[p + "norm2.bias"])
            else:
                bw.qkv_w = f16(sd[p + "attn.qkv.weight"]).data_ptr()
                bw.qkv_b = f32(sd[p + "attn.qkv.bias"]).data_ptr()
                bw.fc1_w = f16(sd[p + "mlp.fc1.weight"]).data_ptr()
                bw.fc1_b = f32(sd[p + "mlp.fc1.bias"]).data_ptr()
            bw.proj_w = f16(sd[p + "attn.proj.weight"]).data_ptr()
            bw.proj_b = f32(sd[p + "attn.proj.bias"]).data_ptr()
            bw.ls1 = f32(sd[p + "ls1.gamma"]).data_ptr()
            bw.norm2_w = f32(sd[p + "norm2.weight"]).data_ptr()
            bw.norm2_b = f32(sd[p + "norm2.bias"]).data_ptr()
            bw.fc2_w = f16(sd[p + "mlp.fc2.weight"]).data_ptr()
            bw.fc2_b = f32(sd[p + "mlp.fc2.bias"]).data_ptr()
            bw.ls2 = f32(sd[p + "ls2.gamma"]).data_ptr()
            blocks.append(bw)
        cfg = _native.VitConfig(d, self.arch.num_heads, n_blocks, self.arch.num_register_tokens, ps, h, w,
                                int(fuse_ln))
        with torch.cuda.device(device):
            handle = _native.vit_create(cfg, weights, blocks, self.max_batch)
        entry = {"handle": handle, "keep": keep, "hp": h // ps, "wp": w // ps, "fuse_ln": fuse_ln}
        self._native_cache[key] = entry
        return entry

    # -- forward --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_tokens(self, images: torch.Tensor, want_f16: bool = False, want_cls: bool = True,
                       out_tokens: Optional[torch.Tensor] = None, out_tokens_f16: Optional[torch.Tensor] = None,
                       out_cls: Optional[torch.Tensor] = None):
        """Patch tokens of block `layer` as B x (Hp*Wp) x D (token-major; what the kernels produce).

        Returns (tokens fp32, tokens f16 or None, cls fp32 or None).
        """
        if self.facet not in _FACETS:
            raise AssertionError(f"{self.facet} is not a supported facet for descriptors.")
        if not images.is_cuda:
            raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback); call images.cuda().")
        images = images.to(torch.float32).contiguous()
        b, c, h, w = images.shape
        assert c == 3
        entry = self._prepare(h, w, images.device)
        p = entry["hp"] * entry["wp"]
        d = self.arch.embed_dim
        if out_tokens is None:
            out_tokens = torch.empty((b, p, d), dtype=torch.float32, device=images.device)
        if want_f16 and out_tokens_f16 is None:
            out_tokens_f16 = torch.empty((b, p, d), dtype=torch.float16, device=images.device)
        if want_cls and out_cls is None:
            out_cls = torch.empty((b, d), dtype=torch.float32, device=images.device)
        with torch.cuda.device(images.device):
            for s in range(0, b, self.max_batch):
                e = min(b, s + self.max_batch)
                _native.vit_forward(entry["handle"], images[s:e], self.layer, _FACETS[self.facet], self.apply_norm,
                                    out_tokens[s:e], out_tokens_f16[s:e] if out_tokens_f16 is not None else None,
                                    out_cls[s:e] if out_cls is not None else None)
        self.num_patches = (entry["hp"], entry["wp"])
        return out_tokens, out_tokens_f16, out_cls

    def forward(self, images: torch.Tensor) -> tp.Dict[str, torch.Tensor]:
        tokens, _, cls = self.forward_tokens(images)
        b = images.shape[0]
        hp, wp = self.num_patches
        # Same (non-contiguous) B x D x Hp x Wp view of token-major data as the reference returns.
        feature_maps = tokens.reshape(b, hp, wp, tokens.shape[-1]).permute(0, 3, 1, 2)
        return {"cls_tokens": cls, "feature_maps": feature_maps}
