#!/usr/bin/env python3
"""In-process A/B of the LayerNorm fusion: ViT-L/14 forward (10 blocks, B=64) with norm1/norm2 folded into the
GEMMs vs the separate LayerNorm kernels, alternating on the same GPU so that clocks / power state are shared.

    python tools/ln_fusion_ab.py [--rounds 6] [--iters 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from foundpose_b200 import synthetic
    from foundpose_b200.utils import dinov2_utils

    arch = synthetic.VIT_ARCHS["vitl14"]
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=10)
    images = [synthetic.make_crops(64, (420, 420), seed=s).cuda() for s in range(2)]
    ext = {}
    for mode in ("1", "0"):
        os.environ["FOUNDPOSE_FUSE_LAYERNORM"] = mode
        e = dinov2_utils.DinoFeatureExtractor("dinov2_vitl14", state_dict=sd, max_batch=64).cuda()
        e.forward_tokens(images[0], want_cls=False)          # builds the native handle under this mode
        ext[mode] = e
    out = torch.empty((64, 900, 1024), dtype=torch.float32, device="cuda")
    times = {"1": [], "0": []}
    for r in range(args.rounds):
        for mode in (("1", "0") if r % 2 == 0 else ("0", "1")):
            e = ext[mode]
            for i in range(3):
                e.forward_tokens(images[i % 2], want_cls=False, out_tokens=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                e.forward_tokens(images[i % 2], want_cls=False, out_tokens=out)
            e1.record()
            torch.cuda.synchronize()
            times[mode].append(e0.elapsed_time(e1) / args.iters)
    res = {"fused_ms": sorted(times["1"])[len(times["1"]) // 2], "separate_ms": sorted(times["0"])[len(times["0"]) // 2],
           "fused_all": [round(t, 3) for t in times["1"]], "separate_all": [round(t, 3) for t in times["0"]]}
    res["speedup"] = res["separate_ms"] / res["fused_ms"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
