#!/usr/bin/env python3
"""Condenses a bench.py JSON line from stdin to the few numbers looked at while tuning."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        if line:
            print(line[:200])
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print(f"value={d['value']:.1f} {d['unit']}  ms/step={d['ms_per_step']:.3f}  e2e={d['e2e']['value']:.1f}  "
          f"launches={d.get('gpu_launches')}  clocks={d.get('clocks')}")
    if r:
        print(f"gemm={r['achieved']:.1f} TF/s ({r['frac']:.3f})  vit={r['vit_stage']['tflops']:.1f} TF/s "
              f"({r['vit_stage']['frac_of_tensor_peak']:.3f})  attn={r['attention']['tflops']:.1f} TF/s  "
              f"ln={r['layernorm']['gbs']:.0f} GB/s")
        print("families ms/step:", r["families_ms_per_step"])
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        print(f"cpu_baseline={c['value']:.3f} {c['unit']} on {c['cores']} cores (early exit {c.get('early_exit_value')})")
