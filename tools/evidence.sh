#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU).  Writes only small summaries to gpurun_out/ (the .ncu-rep
# files are summarised on the box with tools/ncu_report.py and stay in /tmp: gpurun copies back at most 64 MiB).
#   bash tools/evidence.sh            everything
#   bash tools/evidence.sh ncu        only the ncu captures
#   bash tools/evidence.sh bench      only the bench lines + the GPU test log
set -u
WHAT=${1:-all}
mkdir -p gpurun_out
OUT=gpurun_out
if [ "$WHAT" = all ] || [ "$WHAT" = ncu ]; then
# 1. every launch of one config3 micro-batch (64 crops incl. the full-bank search) with its device time
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/r02_launches_config3.csv python tools/profile_step.py --micro 1 > /dev/null 2>&1
python tools/ncu_report.py launches $OUT/r02_launches_config3.csv > $OUT/r02_launches_config3.md
# 2. ncu --set full: the ViT kernels (first 16 launches = token preparation + one and a half blocks) ...
ncu --profile-from-start off --set full --clock-control none -c 16 -o /tmp/r02_vit -f \
    python tools/profile_step.py --micro 1 > $OUT/r02_ncu_vit.log 2>&1
python tools/ncu_report.py full /tmp/r02_vit.ncu-rep --json $OUT/r02_ncu_vit.json > $OUT/r02_ncu_vit.md
# ... everything after the ViT except the full-bank search (final norm, sampling, PCA GEMM, k-NN searches, tf-idf,
# scoring, cyclic buddies, merges)
ncu --profile-from-start off --set full --clock-control none \
    -k regex:"knn_kernel|knn_merge|tfidf|split_rows|cyclic|sample_features|filter_points|row_sqnorm|final_norm|build_pair|gemm2_tn_kernel<4>|gemm_tn_kernel" \
    -o /tmp/r02_retrieval -f python tools/profile_step.py --micro 1 > $OUT/r02_ncu_retrieval.log 2>&1
python tools/ncu_report.py full /tmp/r02_retrieval.ncu-rep --json $OUT/r02_ncu_retrieval.json > $OUT/r02_ncu_retrieval.md
# ... and the full-bank search itself (one 57 600-query x 10.24 M-row launch)
ncu --profile-from-start off --set full --clock-control none -k regex:"knn_pair_kernel" -c 1 \
    -o /tmp/r02_k4 -f python tools/profile_step.py --micro 1 > $OUT/r02_ncu_k4.log 2>&1
python tools/ncu_report.py full /tmp/r02_k4.ncu-rep --json $OUT/r02_ncu_k4.json > $OUT/r02_ncu_k4.md
# 3. the HBM-bound k-NN pass (128 queries sweeping the configs[2] bank), ncu --set full
ncu --set full --clock-control none -k regex:"knn_kernel|knn_merge" -s 4 -c 2 -o /tmp/r02_knn_hbm -f \
    python tools/knn_hbm_probe.py > $OUT/r02_ncu_knn_hbm.log 2>&1
python tools/ncu_report.py full /tmp/r02_knn_hbm.ncu-rep --json $OUT/r02_ncu_knn_hbm.json > $OUT/r02_ncu_knn_hbm.md
ls -la /tmp/*.ncu-rep > $OUT/r02_ncu_sizes.txt
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
# 4. the bench lines: default run, 30 steps for the sustained-power check, the configs[4] sweep
python bench.py --steps 10 --warmup 3 > $OUT/r02_bench_config3_n1.json 2> $OUT/r02_bench_config3_n1.err
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-spot-check --no-extras \
    > $OUT/r02_bench_config3_n1_steps30.json 2> $OUT/r02_bench_config3_n1_steps30.err
python bench.py --workload config5 > $OUT/r02_bench_config5_n1.json 2> $OUT/r02_bench_config5_n1.err
tail -c 400 $OUT/r02_bench_config3_n1.err
python - <<'PY'
import json
for f in ("r02_bench_config3_n1", "r02_bench_config3_n1_steps30"):
    d = json.load(open("gpurun_out/" + f + ".json"))
    print(f, d["value"], d["e2e"]["value"], d["without_k4"]["value"], d["k4"]["tflops"], d["clocks"])
PY
# 5. the GPU test suite
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/r02_pytest_gpu.log
tail -3 $OUT/r02_pytest_gpu.log
fi
du -sh gpurun_out
