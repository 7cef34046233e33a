"""bench.py contract checks that need no GPU: the reference arm runs the oracle port on the host and
prints one JSON line with the required keys (tiny workload so the CPU suite stays fast)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "crops/sec" and d["unit"] == "crops/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_vit_flop_model_matches_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    from foundpose_b200 import synthetic

    # ViT-L/14 @ 420^2: 261.08 GF for blocks 0-9, 625.07 GF for all 24 blocks.
    arch = synthetic.VIT_ARCHS["vitl14"]
    assert abs(bench.vit_flops_per_crop(arch, 9) / 1e9 - 261.08) < 0.01
    assert abs(bench.vit_flops_per_crop(arch, 23) / 1e9 - 625.07) < 0.01
    # ViT-S/14-reg: 45.01 GF for 10 blocks.
    assert abs(bench.vit_flops_per_crop(synthetic.VIT_ARCHS["vits14-reg"], 9) / 1e9 - 45.01) < 0.3
