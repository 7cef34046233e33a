#!/usr/bin/env python3
"""Benchmark of the FoundPose per-crop hot path on B200 (contract: see the task brief / DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

Default workload = BASELINE.json configs[2], the configuration the north-star target is quoted on:
one "step" = 512 synthetic 420x420 crops (2 micro-batches of 256) through the whole path

    ViT-L/14 (blocks 0..9) -> mask filter -> sampling -> PCA 1024->384 -> visual-word 3-NN -> tf-idf ->
    cosine (bag-of-words) scores against 10 000 template descriptors -> top-5 templates -> 2 x 5 1-NN searches ->
    cyclic buddies -> 2D-3D gathers,  PLUS  the brute-force 5-NN of all 900 query descriptors of every crop
    against the WHOLE bank (10 000 templates x 1024 patches x 384-d = 10.24 M rows, 7.9 GB fp16): "K4".

`value` includes K4 (it is 27x the ViT's flops and sets the crops/s of this configuration); `without_k4` reports
the path as the reference's scripts/infer.py runs it.  Prints ONE JSON line on rank 0.
Other workloads: --workload config2 (configs[1]), config4 (configs[3]: 8 LM-O-shaped objects, 4096 crops sharded
over the ranks), config5 (configs[4]: bank-size sweep of the k-NN), tiny (functional check).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

VITL = "dinov2_vitl14"
WORKLOADS = {
    # BASELINE.json configs[2]: the north-star configuration (default at N=1 and for the scaling run).
    "config3": dict(batch=256, crops_per_step=512, vit=VITL, templates=10000, patches=1024, dim=384, pca=True,
                    words=2048, top_n=5, top_k=300, k4=5, cpu_k4_fraction=0.1,
                    desc="configs[2]: batch=512 synthetic 420x420 crops per step (2 micro-batches of 256; the full-bank "
                         "search is launched per 64 crops = 57 600 queries), ViT-L/14 "
                         "layer 9 + PCA 1024->384 + tf-idf BoW scoring of 10k templates + top-5 retrieval + cyclic "
                         "buddies + brute-force 5-NN of all 900 queries per crop vs the whole 10k-template x "
                         "1024-patch x 384-d bank (10.24M rows)"),
    # BASELINE.json configs[1].
    "config2": dict(batch=64, crops_per_step=64, vit=VITL, templates=2000, patches=1024, dim=256, pca=True,
                    words=2048, top_n=5, top_k=300, k4=0, cpu_k4_fraction=1.0,
                    desc="configs[1]: batch=64 synthetic 420x420 crops, ViT-L/14 layer 9 + PCA 1024->256 + "
                         "tf-idf top-5 template retrieval + cyclic buddies vs 2000-template x 1024-patch x "
                         "256-d bank"),
    # BASELINE.json configs[3]: 8 LM-O-shaped objects, 4096 crops sharded over the ranks (see run_config4).
    "config4": dict(batch=64, crops_per_step=4096, vit=VITL, templates=800, patches=1200, dim=384, pca=True,
                    words=2048, top_n=5, top_k=300, k4=0, objects=8, cpu_k4_fraction=1.0,
                    desc="configs[3]: 8 objects x 800 templates x 1200 patches x 384-d banks resident per GPU, 4096 "
                         "synthetic 420x420 crops (512 per object) sharded over the ranks, object loop as "
                         "scripts/infer.py:179-239, banks NCCL-broadcast at init"),
    # Small variant for quick functional checks of bench.py itself (not a reported configuration).
    "tiny": dict(batch=8, crops_per_step=16, vit="dinov2_version=vits14-reg_stride=14_facet=token_layer=9_norm=1",
                 templates=64, patches=256, dim=256, pca=True, words=256, top_n=5, top_k=300, k4=5,
                 cpu_k4_fraction=0.5, desc="tiny functional check (not a BASELINE configuration)"),
}

CATEGORY_NAMES = ["gemm", "attention", "layernorm", "vit_misc", "knn", "feature_ops", "retrieval", "knn_full_bank"]


# ------------------------------------------------------------------------------------------------
# Synthetic workload construction (torch ops only, on whatever device is given: both arms build the SAME
# bank from the same seeded device generator when a GPU is present)
# ------------------------------------------------------------------------------------------------
def vit_arch_and_layer(name: str):
    from foundpose_b200 import synthetic
    from foundpose_b200.utils import dinov2_utils

    opts = dinov2_utils.parse_model_name(name)
    return synthetic.VIT_ARCHS[opts["version"]], opts


def build_bank(wl: dict, device: torch.device, seed: int = 0) -> dict:
    """fp16-representable synthetic bank generated chunk-wise on `device`:
    feat16 [F,d] fp16, tpl_ids [F] int32 (contiguous ascending runs, scripts/gen_repre.py:187-214), vertices [F,3],
    centroids [W,d] fp32 = seeded sample of the bank rows."""
    T, P, d, W = wl["templates"], wl["patches"], wl["dim"], wl["words"]
    F = T * P
    g = torch.Generator(device=device).manual_seed(seed)
    feat16 = torch.empty((F, d), dtype=torch.float16, device=device)
    step = 1 << 20
    for s in range(0, F, step):
        n = min(step, F - s)
        feat16[s:s + n] = torch.randn((n, d), generator=g, device=device, dtype=torch.float32).to(torch.float16)
    vertices = torch.randn((F, 3), generator=g, device=device, dtype=torch.float32)
    pick = torch.randperm(F, generator=g, device=device)[:W]
    centroids = feat16[pick].to(torch.float32)
    tpl_ids = torch.arange(T, dtype=torch.int32, device=device).repeat_interleave(P)
    return {"feat16": feat16, "tpl_ids": tpl_ids, "vertices": vertices, "centroids": centroids}


def oracle_bank_descriptors(bank: dict, wl: dict):
    """template_descs / idfs through the oracle's vectorised restatement (setup of the reference arm, untimed;
    runs on the device the bank lives on)."""
    from oracle import template as otemplate

    feat, cent = bank["feat16"], bank["centroids"]
    cn = (cent * cent).sum(dim=1).unsqueeze(0)
    f2w = torch.empty(feat.shape[0], dtype=torch.int64, device=feat.device)
    step = 1 << 18
    for s in range(0, feat.shape[0], step):
        x = feat[s:s + step].to(torch.float32)
        f2w[s:s + step] = torch.argmin((x * x).sum(dim=1, keepdim=True) + cn - 2.0 * (x @ cent.t()), dim=1)
    return otemplate.calc_tfidf_descriptors_vectorised(feat, f2w, bank["tpl_ids"], cent, wl["templates"], 3)


def cpu_bank_dict(bank: dict, descs: torch.Tensor, idfs: torch.Tensor) -> dict:
    """The fp32 CPU tensors the reference's establish_correspondences reads (fields of utils/repre_util.py:34-83)."""
    return {
        "feat_vectors": bank["feat16"].cpu().to(torch.float32),
        "feat_to_template_ids": bank["tpl_ids"].cpu(),
        "vertices": bank["vertices"].cpu(),
        "feat_cluster_centroids": bank["centroids"].cpu(),
        "template_descs": descs.cpu(), "feat_cluster_idfs": idfs.cpu(),
    }


# ------------------------------------------------------------------------------------------------
# Clock / throttle sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index: int) -> None:
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _run(self) -> None:
        n = self._nvml
        names = {
            "hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake_slowdown": getattr(n, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self) -> None:
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ------------------------------------------------------------------------------------------------
# Reference arm / CPU baseline: the oracle restatement of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
class CpuReferencePath:
    """The reference's per-crop path (scripts/infer.py:467-545) restated by oracle/, B=1 per call, plus - for
    workloads that name it - the full-bank k-NN K4 (utils/knn_util.py:65-106 with all feat_vectors as the index)."""

    def __init__(self, wl: dict, bank_cpu: dict, full_depth: bool) -> None:
        from foundpose_b200 import synthetic
        from oracle import feature as ofeature

        self.wl = wl
        self.arch, self.opts = vit_arch_and_layer(wl["vit"])
        depth = None if full_depth else self.opts["layer"] + 1
        self.sd = synthetic.make_vit_state_dict(self.arch, seed=0, depth=depth)
        self.full_depth = full_depth
        self.pdict = synthetic.make_pca(self.arch.embed_dim, wl["dim"], seed=0) if wl["pca"] else None
        self.bank = bank_cpu
        self.grid = ofeature.generate_grid_points((420, 420), 14.0)
        F = bank_cpu["feat_vectors"].shape[0]
        self.k4_rows = max(1024, int(F * wl["cpu_k4_fraction"])) if wl["k4"] else 0
        self.k4_scale = F / self.k4_rows if self.k4_rows else 0.0

    def crop(self, image: torch.Tensor, mask: torch.Tensor):
        """Returns (seconds of the path as scripts/infer.py runs it, seconds of K4 extrapolated to the whole bank)."""
        from oracle import corresp as ocorresp
        from oracle import feature as ofeature
        from oracle import knn as oknn
        from oracle import pca as opca
        from oracle import vit as ovit

        t0 = time.perf_counter()
        fmap = ovit.extract(self.sd, self.arch, image.unsqueeze(0), layer=self.opts["layer"],
                            facet=self.opts["facet"], apply_norm=self.opts["norm"],
                            full_depth=self.full_depth)["feature_maps"][0]
        qp = ofeature.filter_points_by_mask(self.grid, mask)
        feats = ofeature.sample_feature_map_at_points(fmap, qp, (420, 420)).contiguous()
        if self.pdict is not None:
            feats = opca.project_features(feats, [self.pdict]).contiguous()
        ocorresp.establish_correspondences(qp, feats, self.bank, self.wl["top_n"], self.wl["top_k"])
        t1 = time.perf_counter()
        t_k4 = 0.0
        if self.k4_rows:
            # faiss IndexFlatL2.search blocks the database (exhaustive_L2sqr_blas); brute force is linear in the
            # number of bank rows, so a contiguous row sample is timed and scaled to the whole bank.
            fv = self.bank["feat_vectors"]
            blocks = ((s, fv[s:s + 65536]) for s in range(0, self.k4_rows, 65536))
            oknn.knn_l2_blocked(feats, blocks, self.wl["k4"])
            t_k4 = (time.perf_counter() - t1) * self.k4_scale
        return t1 - t0, t_k4


def time_cpu_path(path: CpuReferencePath, n_crops: int, warmup: int, seed: int = 1000):
    from foundpose_b200 import synthetic

    images = synthetic.make_crops(n_crops + warmup, (420, 420), seed=seed)
    mask = torch.ones(420, 420, dtype=torch.bool)
    for i in range(warmup):
        path.crop(images[i], mask)
    return [path.crop(images[i], mask) for i in range(warmup, warmup + n_crops)]


def cpu_sample_text(path: CpuReferencePath, n: int, cores: int) -> str:
    depth = path.arch.depth if path.full_depth else path.opts["layer"] + 1
    txt = (f"{n} crops of the same workload, B=1 per call as scripts/infer.py does, fp32 torch-CPU oracle port of the "
           f"reference path with {depth} ViT blocks executed"
           + (" (the reference's forward hook cannot stop the model)" if path.full_depth else "")
           + f", {cores} threads")
    if path.k4_rows:
        txt += (f"; full-bank 5-NN timed on the first {path.k4_rows} bank rows (blocked like faiss "
                f"exhaustive_L2sqr_blas) and scaled x{path.k4_scale:.1f} to the {path.bank['feat_vectors'].shape[0]} rows "
                "(brute force is linear in the bank size)")
    return txt


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # Untimed setup: the synthetic bank and its tf-idf descriptors (torch ops + the oracle's vectorised
    # descriptor builder), on the GPU when there is one - the same seeded generator as the CUDA arm - else on the CPU.
    setup_dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    bank = build_bank(wl, setup_dev)
    descs, idfs = oracle_bank_descriptors(bank, wl)
    bank_cpu = cpu_bank_dict(bank, descs, idfs)
    del bank, descs, idfs
    if setup_dev.type == "cuda":
        torch.cuda.empty_cache()
    path = CpuReferencePath(wl, bank_cpu, full_depth=True)
    times = time_cpu_path(path, args.steps, args.warmup)
    t_path = sum(t[0] for t in times)
    t_k4 = sum(t[1] for t in times)
    value = len(times) / (t_path + t_k4)
    line = {
        "impl": "reference", "metric": "crops/sec", "value": value, "unit": "crops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (t_path + t_k4) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "crops_per_step": 1, "setup_device": str(setup_dev)},
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(path, len(times), cores)},
        "without_k4": {"value": len(times) / t_path, "unit": "crops/s"},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ------------------------------------------------------------------------------------------------
# This repo's arm
# ------------------------------------------------------------------------------------------------
def load_peaks():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))   # kernels timed inside a long step
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json, sustained)" if peaks else "fallback (B200_PROFILING.md)"
    return tensor_peak, hbm_peak, src


def build_repre_on_device(wl: dict, dev: torch.device, rank: int, world: int, seed: int = 0):
    """Rank 0 builds the object representation, the replicas receive it through
    distributed.broadcast_object_repre (NCCL at init only).  Returns (repre, bank dict or None, broadcast seconds)."""
    from foundpose_b200 import distributed, synthetic
    from foundpose_b200.utils import knn_util, projector_util, repre_util, template_util

    arch, _ = vit_arch_and_layer(wl["vit"])
    pdict = synthetic.make_pca(arch.embed_dim, wl["dim"], seed=0) if wl["pca"] else None
    projectors = [projector_util.projector_from_tensordict(pdict)] if pdict is not None else []
    repre, bank = None, None
    if rank == 0:
        bank = build_bank(wl, dev, seed)
        wk = knn_util.KNN(k=1, metric="l2")
        wk.fit(bank["centroids"])
        f2w = torch.empty(bank["feat16"].shape[0], dtype=torch.int64, device=dev)
        step = 1 << 21
        for s in range(0, f2w.shape[0], step):
            f2w[s:s + step] = wk.search(bank["feat16"][s:s + step].to(torch.float32))[1].flatten()
        descs, idfs = template_util.calc_tfidf_descriptors(bank["feat16"].to(torch.float32), f2w, bank["tpl_ids"],
                                                           bank["centroids"], wl["templates"], 3, False, 10.0)
        del f2w
        torch.cuda.empty_cache()
        # The bank travels and is stored as fp16 rows (lossless: it is fp16-representable by construction).
        repre = repre_util.FeatureBasedObjectRepre(
            vertices=bank["vertices"], feat_vectors=bank["feat16"], feat_to_template_ids=bank["tpl_ids"],
            feat_cluster_centroids=bank["centroids"], feat_cluster_idfs=idfs, template_descs=descs,
            template_desc_opts=repre_util.TemplateDescOpts(), feat_raw_projectors=projectors)
    t_bcast = 0.0
    if world > 1:
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        repre = distributed.broadcast_object_repre(repre, src=0, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        t_bcast = time.perf_counter() - t0
    return repre, bank, pdict, t_bcast


def run_cuda_arm(args) -> None:
    from foundpose_b200 import _native, pipeline, synthetic
    from foundpose_b200.utils import dinov2_utils, knn_util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    lib = _native.load()
    wl = WORKLOADS[args.workload]
    if os.environ.get("FP_BENCH_MICRO_BATCH"):      # experiment: crops per micro-batch (must divide crops_per_step)
        wl = dict(wl, batch=int(os.environ["FP_BENCH_MICRO_BATCH"]))
        wl["desc"] = wl["desc"].replace("2 micro-batches of 256", "%d micro-batches of %d" % (wl["crops_per_step"] // wl["batch"], wl["batch"]))
        assert wl["crops_per_step"] % wl["batch"] == 0
    B = wl["batch"]
    n_micro = wl["crops_per_step"] // B
    arch, opts = vit_arch_and_layer(wl["vit"])
    layer = opts["layer"]
    tensor_peak, hbm_peak, peak_src = load_peaks()

    # ---- init (untimed): weights, PCA, bank (rank 0 builds, NCCL-broadcasts to the replicas) -----
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=layer + 1)
    extractor = dinov2_utils.DinoFeatureExtractor(wl["vit"], state_dict=sd, max_batch=B).to(dev)
    repre, bank, pdict, t_bcast = build_repre_on_device(wl, dev, rank, world)
    index = pipeline.ObjectIndex(repre, dev)
    pipe = pipeline.CropBatchPipeline(extractor, index, repre.feat_raw_projectors, B, crop_size=(420, 420),
                                      grid_cell_size=14.0, top_n_templates=wl["top_n"], top_k_buddies=wl["top_k"])
    stride = pipe.stride
    k4 = wl["k4"]
    k4_index = knn_util.KNN.from_packed(index.bank16, index.bank_sqnorm, k=k4, metric="l2") if k4 else None
    # Query descriptors of a whole step (the K4 search reads a micro-batch's slice right after it is produced).
    q_all = torch.zeros((n_micro, B * stride, index.dim_padded), dtype=torch.float16, device=dev)
    qn_all = torch.zeros((n_micro, B * stride), dtype=torch.float32, device=dev)
    # The full-bank search is launched per K4_CROPS crops (57 600 queries: 3 whole waves of 74 clusters + a split tail
    # wave - the launch shape the ncu evidence describes), whatever the micro-batch of the rest of the path.
    k4_crops = min(B, 64)
    k4_rows = k4_crops * stride
    k4_d = [torch.zeros((B * stride, max(k4, 1)), dtype=torch.float32, device=dev) for _ in range(n_micro)]
    k4_i = [torch.zeros((B * stride, max(k4, 1)), dtype=torch.int64, device=dev) for _ in range(n_micro)]

    # Inputs: one pool of n_micro micro-batches (>= 2), different crops per rank.  512 crops = 1.08 GB per step,
    # far larger than the 126 MB L2.
    n_pool = max(n_micro, 2)
    host_images = [synthetic.make_crops(B, (420, 420), seed=100 + 1000 * rank + s).pin_memory() for s in range(n_pool)]
    host_masks = [torch.ones(B, 420, 420, dtype=torch.uint8).pin_memory() for _ in range(n_pool)]
    dev_images = [h.to(dev) for h in host_images]
    dev_masks = [h.to(dev) for h in host_masks]

    def barrier() -> None:
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps: int) -> float:
        """Device time (ms) of `steps` calls, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def micro_batch(images, masks, m: int, with_k4: bool):
        out = pipe.run(images, masks, desc_out=q_all[m])
        if with_k4:
            # ||q||^2 of this micro-batch's rows were computed by the engine (q_sqnorm); K4 = all queries x all rows.
            for r0 in range(0, B * stride, k4_rows):
                d, i = k4_index.search_packed(q_all[m][r0:r0 + k4_rows], pipe.engine.q_sqnorm[r0:r0 + k4_rows])
                k4_d[m][r0:r0 + k4_rows].copy_(d)      # the search returns its plan's buffers, reused by the next launch
                k4_i[m][r0:r0 + k4_rows].copy_(i)
        return out

    def step_resident(i: int, with_k4: bool = bool(k4)) -> None:
        for m in range(n_micro):
            j = (i * n_micro + m) % n_pool
            micro_batch(dev_images[j], dev_masks[j], m, with_k4)

    # ---- parity spot check of the benchmarked shape against the oracle, before any timing -----------
    spot = None
    micro_batch(dev_images[0], dev_masks[0], 0, bool(k4))
    torch.cuda.synchronize()
    if rank == 0 and not args.no_spot_check:
        spot = spot_check(pipe, index, host_images[0], host_masks[0], sd, arch, layer, pdict, wl,
                          q_all[0], k4_d[0], k4_i[0])

    # ---- warm-up + timed region (inputs resident in HBM) -------------------------------------------
    warmup = max(args.warmup, 3)
    for i in range(warmup):
        step_resident(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.fp_launch_count()
    ms_total = timed(step_resident, args.steps)
    gpu_launches = int(lib.fp_launch_count() - launches0)
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    crops_per_step = B * n_micro
    value = world * crops_per_step * args.steps / (ms_total / 1e3)
    without_k4 = None
    if k4:
        ms_nok4 = timed(lambda i: step_resident(i, False), args.steps)
        without_k4 = {"value": world * crops_per_step * args.steps / (ms_nok4 / 1e3), "unit": "crops/s",
                      "ms_per_step": ms_nok4 / args.steps,
                      "note": "the path as scripts/infer.py runs it (no full-bank search)"}

    # ---- per-kernel-family attribution with CUDA events on the launching stream -------------------
    prof_steps = min(args.steps, 2 if k4 else 5)
    import ctypes

    for c in range(len(CATEGORY_NAMES)):
        lib.fp_profile_read(ctypes.c_int(c), None, None, None, ctypes.c_int(1))
    lib.fp_profile_enable(1)
    prof_total_ms = timed(step_resident, prof_steps)
    lib.fp_profile_enable(0)
    fam = {}
    for c, name in enumerate(CATEGORY_NAMES):
        ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        lib.fp_profile_read(ctypes.c_int(c), ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n), ctypes.c_int(1))
        fam[name] = {"ms_per_step": ms.value / prof_steps, "launches_per_step": n.value / prof_steps,
                     "work_per_step": work.value / prof_steps}
    fam_ms = sum(v["ms_per_step"] for v in fam.values())

    def rate(name, scale):
        f = fam[name]
        return f["work_per_step"] / (f["ms_per_step"] * 1e-3) / scale if f["ms_per_step"] > 0 else 0.0

    gemm_tflops, attn_tflops, ln_gbs = rate("gemm", 1e12), rate("attention", 1e12), rate("layernorm", 1e9)
    vit_flops = vit_flops_per_crop(arch, layer) * crops_per_step
    vit_ms = sum(fam[k]["ms_per_step"] for k in ("gemm", "attention", "layernorm", "vit_misc"))
    vit_tflops = vit_flops / (vit_ms * 1e-3) / 1e12 if vit_ms > 0 else 0.0
    g = fam["gemm"]
    sub = {
        "vit_stage": {"tflops": vit_tflops, "frac_of_tensor_peak": vit_tflops / tensor_peak,
                      "flops_per_crop": vit_flops / crops_per_step, "ms_per_step": vit_ms},
        "gemm": {"tflops": gemm_tflops, "frac_of_tensor_peak": gemm_tflops / tensor_peak,
                 "launches_per_step": g["launches_per_step"],
                 "traffic": read_json_key("r02_gemm_traffic.json", "avg_bytes_per_launch")},
        "attention": {"tflops": attn_tflops, "frac_of_tensor_peak": attn_tflops / tensor_peak},
        "layernorm": {"gbs": ln_gbs, "frac_of_hbm_peak": ln_gbs / hbm_peak},
        "families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in fam.items()},
        "profiled_step_ms": prof_total_ms / prof_steps,
    }
    F = index.bank16.shape[0]
    if k4:
        kf = fam["knn_full_bank"]
        launches = max(kf["launches_per_step"], 1)
        flops_per_launch = 2.0 * k4_rows * F * index.dim_padded
        avg_ms = kf["ms_per_step"] / launches
        k4_tflops = flops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        roofline = {
            "kernel": "knn_pair_kernel<5> (tcgen05 cta_group::2 brute-force k-NN, full-bank search K4)",
            "bound": "tensor", "achieved": k4_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
            "frac": k4_tflops / tensor_peak,
            "traffic": read_json_key("r02_k4_traffic.json", "bytes_per_launch"),
            "traffic_note": "DRAM read+write bytes of one launch (ncu --set full, profiles/r02_k4_traffic.json). "
                            "SURVEY 8(d) counts F*d*2 per 128-query tile (450 tiles per launch = 3.5 TB); with the 74 "
                            "clusters of a wave sweeping in lockstep the bank leaves HBM once per WAVE (3 waves + the "
                            "split tail wave = 4 sweeps = 31.5 GB); algorithmic_bytes_per_launch below = ONE sweep + "
                            "queries + results, the floor of a single-sweep schedule. The kernel is tensor-bound: DRAM "
                            "is 1.4% busy",
            "algorithmic_flops_per_launch": flops_per_launch,
            "algorithmic_bytes_per_launch": F * index.dim_padded * 2 + k4_rows * (index.dim_padded * 2 + k4 * 12),
            "peak_source": peak_src, "launches_per_step": kf["launches_per_step"], "avg_launch_ms": avg_ms,
            "share_of_step": kf["ms_per_step"] / fam_ms if fam_ms > 0 else None,
        }
    else:
        roofline = {
            "kernel": "gemm2_tn_kernel (tcgen05 GEMM, all ViT linear layers + patch embed + PCA)",
            "bound": "tensor", "achieved": gemm_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
            "frac": gemm_tflops / tensor_peak, "traffic": sub["gemm"]["traffic"],
            "traffic_note": "avg DRAM read+write bytes per gemm launch (ncu, profiles/r02_gemm_traffic.json)",
            "peak_source": peak_src, "launches_per_step": g["launches_per_step"],
            "avg_launch_ms": g["ms_per_step"] / max(g["launches_per_step"], 1),
            "share_of_step": g["ms_per_step"] / fam_ms if fam_ms > 0 else None,
        }
    roofline.update(sub)

    # ---- end to end through the public API with HOST buffers --------------------------------------
    out_ref = pipe.engine.out
    d2h_fields = [out_ref.template_ids, out_ref.template_scores, out_ref.count, out_ref.query_ids,
                  out_ref.vertex_ids, out_ref.scores, out_ref.coord_2d, out_ref.coord_3d]
    k4_shapes = [((B * stride, k4), torch.float32), ((B * stride, k4), torch.int64)] if k4 else []
    host_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in d2h_fields]
                + [torch.empty(s, dtype=dt).pin_memory() for s, dt in k4_shapes] for _ in range(2)]
    # Double-buffered staging: the H2D copy of micro-batch i+1 (copy stream) overlaps the kernels of micro-batch i
    # (compute stream); every copy still happens inside the timed region.
    stage_img = [torch.empty_like(dev_images[0]) for _ in range(2)]
    stage_msk = [torch.empty_like(dev_masks[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    d2h_done = [torch.cuda.Event() for _ in range(2)]
    state = {"primed": False}

    def issue_h2d(i: int) -> None:
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])          # micro-batch i-2 has finished reading this buffer
            stage_img[b].copy_(host_images[i % n_pool], non_blocking=True)
            stage_msk[b].copy_(host_masks[i % n_pool], non_blocking=True)
            h2d_done[b].record(copy_stream)

    def micro_e2e(i: int) -> None:
        b = i % 2
        cur = torch.cuda.current_stream()
        if not state["primed"]:
            for e in consumed:
                e.record(cur)
            issue_h2d(i)
            state["primed"] = True
        issue_h2d(i + 1)                                   # prefetch the next micro-batch's inputs
        cur.wait_event(h2d_done[b])
        m = i % n_micro
        out = micro_batch(stage_img[b], stage_msk[b], m, bool(k4))
        consumed[b].record(cur)
        fields = [out.template_ids, out.template_scores, out.count, out.query_ids, out.vertex_ids, out.scores,
                  out.coord_2d, out.coord_3d] + ([k4_d[m], k4_i[m]] if k4 else [])
        for h, t in zip(host_out[b], fields):
            h.copy_(t, non_blocking=True)
        d2h_done[b].record(cur)
        if i >= 1:
            d2h_done[(i - 1) % 2].synchronize()            # the caller reads the previous micro-batch's result

    def step_e2e(i: int) -> None:
        for m in range(n_micro):
            micro_e2e(i * n_micro + m)

    step_e2e(0)
    e2e_steps = min(args.steps, 8) if k4 else args.steps
    barrier()
    t0 = time.perf_counter()
    e2e_ms = timed(lambda i: step_e2e(i + 1), e2e_steps)
    d2h_done[((e2e_steps + 1) * n_micro - 1) % 2].synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_value = world * crops_per_step * e2e_steps / (e2e_ms / 1e3)
    h2d = int(n_micro * (host_images[0].numel() * 4 + host_masks[0].numel()))
    d2h = int(n_micro * sum(t.numel() * t.element_size() for t in host_out[0]))

    # ---- explanatory extras (outside the timed regions, rank 0 at N=1) ------------------------------
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = measure_extras(lib, pipe, index, dev, hbm_peak, B)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ---------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = measure_cpu_baseline(wl, bank, index, arch, layer, args.cpu_crops)
        except Exception as e:   # e.g. not enough host memory for the fp32 bank: the headline line must still appear
            cpu_baseline = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": "crops/sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate, fp32 residual stream)",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "crops_per_step_per_gpu": crops_per_step, "micro_batch": B,
                       "vit_blocks_executed": layer + 1, "queries_per_crop": stride, "bank_rows": F,
                       "parallelism": f"crops sharded over {world} GPU(s), bank replicated"
                       + (f" (distributed.broadcast_object_repre over NCCL at init: {t_bcast:.2f} s)" if world > 1 else ""),
                       "l2_policy": f"inputs larger than L2: {h2d / 1e6:.0f} MB of crops+masks per step, "
                                    f"{n_pool} rotating micro-batch inputs, ~20 MB of activations rewritten per crop" + (f", {F * index.dim_padded * 2 / 1e9:.1f} GB bank swept by K4" if k4 else "")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "wall_s": e2e_wall},
            "gpu_launches": gpu_launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "spot_check": spot,
        }
        if without_k4 is not None:
            line["without_k4"] = without_k4
            kf = fam["knn_full_bank"]
            line["k4"] = {"tflops": roofline["achieved"], "frac_of_tensor_peak": roofline["frac"],
                          "ms_per_step": kf["ms_per_step"], "flops_per_crop": 2.0 * stride * F * index.dim_padded,
                          "k": k4, "queries_per_launch": k4_rows, "bank_rows": F}
        line.update(extras)
        emit_line(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def measure_cpu_baseline(wl: dict, bank: dict, index, arch, layer: int, n_crops: int) -> dict:
    """The oracle port of the reference path on the host cores, on a bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bank_cpu = cpu_bank_dict(bank, index.template_descs, index.idfs)
    shipped = CpuReferencePath(wl, bank_cpu, full_depth=True)
    t_full = time_cpu_path(shipped, n_crops, 1)
    early = CpuReferencePath(wl, bank_cpu, full_depth=False)
    t_early = time_cpu_path(early, max(1, n_crops // 2), 1)
    tot = sum(a + b for a, b in t_full)
    return {
        "value": len(t_full) / tot, "unit": "crops/s", "cores": cores, "kind": "port",
        "sample": cpu_sample_text(shipped, len(t_full), cores),
        "without_k4_value": len(t_full) / sum(a for a, _ in t_full),
        "early_exit_without_k4_value": len(t_early) / sum(a for a, _ in t_early),
        "early_exit_note": f"same without K4, only blocks 0..{layer} executed (the ViT work this repo's path does)",
    }


def read_json_key(name: str, key: str):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)[key]
    except Exception:
        return None


def spot_check(pipe, index, host_images, host_masks, sd, arch, layer, pdict, wl, q_desc16, k4_d, k4_i) -> dict:
    """One crop of the benchmarked shape against the oracle (oracle/check.py), before the timed region:
    descriptor stage vs the fp32 oracle ViT, retrieval + correspondences with chained inputs (bit-exact where the
    margins allow), full-bank k-NN for a few queries, and the end-to-end agreement rate through the fp16 ViT."""
    from oracle import check as ocheck

    t0 = time.perf_counter()
    crop = 0
    res = {}
    try:
        st1 = ocheck.descriptor_stage(pipe, crop, host_images[crop], host_masks[crop].bool(), sd, arch, layer, pdict,
                                      desc=q_desc16)
        res["descriptor_rel_err"] = st1["rel_err"]
        res["query_points_equal"] = st1["points_equal"]
        assert st1["points_equal"] and st1["rel_err"] <= 1e-2, f"descriptor stage off: {st1['rel_err']}"
        n = st1["n_queries"]
        s = pipe.stride
        q = q_desc16[crop * s: crop * s + n].float().cpu()
        st2 = ocheck.retrieval_stage(index, pipe.engine, crop, st1["oracle_points"], q, wl["top_k"])
        res.update({k: st2[k] for k in ("templates_equal", "template_gap", "pairs", "pairs_exact", "nn_checked",
                                        "nn_sure_frac", "nn_equal_frac", "corr_agree", "template_score_err")})
        assert st2["templates_equal"] or not st2["templates_sure"], "template ids differ from the oracle"
        if k4_d is not None:
            nq4 = min(8, n)
            st4 = ocheck.full_bank_knn_stage(index.bank16, q[:nq4], k4_d[crop * s: crop * s + nq4],
                                             k4_i[crop * s: crop * s + nq4], wl["k4"])
            res["k4"] = st4
        e2e = ocheck.end_to_end_agreement(index, pipe.engine.out, crop, st1["oracle_points"], st1["oracle_desc"],
                                          wl["top_k"])
        res["end_to_end"] = e2e
        res["ok"] = True
    except Exception as e:   # reported in the line (ok: false), never fatal for the measurement itself
        res["ok"] = False
        res["error"] = f"{type(e).__name__}: {e}"
    finally:
        res["seconds"] = round(time.perf_counter() - t0, 2)
    return res


def measure_extras(lib, pipe, index, dev, hbm_peak: float, B: int) -> dict:
    """Measurements that explain the headline but are not part of it:

    * `knn_hbm`: the k-NN kernel in its bandwidth-bound pass structure (BASELINE metric "kNN HBM GB/s vs peak"):
      one 128-query tile sweeps a 10k-template x 1024-patch x 384-d fp16 bank (7.9 GB, larger than L2) once;
      algorithmic bytes = F*d*2 per search over the knn_kernel's CUDA-event time.
    * `bow`: the bag-of-words scoring kernel alone: T*W*4 descriptor bytes per 64 crops over its CUDA-event time.
    * `coarse_pose`: fp_pnp_ransac on the correspondences of the last step (B x top-N problems, 400 iterations).
    """
    import ctypes

    from foundpose_b200 import _native
    from foundpose_b200.utils import knn_util, pnp_util

    out = {}

    def reset():
        for c in range(len(CATEGORY_NAMES)):
            lib.fp_profile_read(ctypes.c_int(c), None, None, None, ctypes.c_int(1))

    def read(cat):
        ms, w, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        lib.fp_profile_read(ctypes.c_int(cat), ctypes.byref(ms), ctypes.byref(w), ctypes.byref(n), ctypes.c_int(1))
        return ms.value, w.value, n.value

    try:
        rows, dim, nq, iters = 10000 * 1024, 384, 128, 5
        if index.bank16.shape == (rows, dim):
            bank, bank_n = index.bank16, index.bank_sqnorm          # the benchmarked bank itself
        else:
            bank = torch.empty(rows, dim, device=dev, dtype=torch.float16)
            for s0 in range(0, rows, 1 << 20):
                bank[s0:s0 + (1 << 20)] = torch.randn(min(1 << 20, rows - s0), dim, device=dev, dtype=torch.float16)
            bank_n = _native.row_sqnorm_f16(bank)
        knn = knn_util.KNN.from_packed(bank, bank_n, k=5, metric="l2")
        q = torch.randn(nq, dim, device=dev)
        for _ in range(2):
            knn.search(q)
        torch.cuda.synchronize()
        reset()
        lib.fp_profile_enable(1)
        for _ in range(iters):
            knn.search(q)
        lib.fp_profile_enable(0)
        ms, _, _ = read(4)
        reset()
        t = ms / iters * 1e-3
        gbs = rows * dim * 2 / t / 1e9
        out["knn_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                          "knn_kernel_ms": t * 1e3, "bank_bytes": rows * dim * 2,
                          "workload": "128 queries x (10k templates x 1024 patches x 384-d fp16) bank of configs[2], "
                                      "k=5, bank split over all SMs"}
        del knn, bank, q
        torch.cuda.empty_cache()
    except Exception as e:   # never let an extra break the headline line
        out["knn_hbm"] = {"error": repr(e)}
    try:
        eng = pipe.engine
        for _ in range(2):
            eng.score_templates()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.score_templates()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e-3
        nbytes = index.template_descs.numel() * 4
        out["bow"] = {"bound": "hbm", "achieved": nbytes / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": nbytes / t / 1e9 / hbm_peak, "stage_ms": t * 1e3, "descriptor_bytes": nbytes,
                      "crops": B, "templates": int(index.template_descs.shape[0]),
                      "path": "tensor cores (split-fp16 inner-product top-N through the k-NN kernel)" if eng.bow_tensor
                      else "fp32 CUDA cores", "bytes_read": int(index.split_descs().numel() * 2) if eng.bow_tensor else nbytes,
                      "note": "whole scoring stage (query split + search + merge) for one micro-batch; algorithmic "
                              "bytes = T*W*4 of fp32 template descriptors"}
    except Exception as e:
        out["bow"] = {"error": repr(e)}
    try:
        o = pipe.engine.out
        topn, kk = o.count.shape[1], o.coord_2d.shape[2]
        intr = torch.tensor([[600.0, 600.0, 210.0, 210.0]], dtype=torch.float64, device=dev).expand(B * topn, 4).contiguous()
        args_ = (o.coord_2d.reshape(B * topn, kk, 2), o.coord_3d.reshape(B * topn, kk, 3), o.count.reshape(-1), intr,
                 400, 10.0, 0.99)
        for _ in range(2):
            pnp_util.estimate_poses_batched(*args_)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pnp_util.estimate_poses_batched(*args_)
        e1.record()
        torch.cuda.synchronize()
        out["coarse_pose"] = {"ms_per_micro_batch": e0.elapsed_time(e1) / 5, "problems": B * topn,
                              "correspondences": kk, "ransac_iterations": 400,
                              "note": "fp_pnp_ransac on one micro-batch's correspondences; not inside the timed "
                                      "regions (the north-star path ends at the correspondences)"}
    except Exception as e:
        out["coarse_pose"] = {"error": repr(e)}
    return out


def vit_flops_per_crop(arch, layer: int, size: int = 420) -> float:
    """patch_embed + (layer+1) * block flops for one crop."""
    p = (size // arch.patch_size) ** 2
    n = p + 1 + arch.num_register_tokens
    d = arch.embed_dim
    block = 2 * n * 3 * d * d + 2 * 2 * n * n * d + 2 * n * d * d + 2 * 2 * n * 4 * d * d
    patch = 2 * p * 3 * arch.patch_size ** 2 * d
    return float(patch + (layer + 1) * block)


_REAL_STDOUT = None


def capture_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the
    first collective), so everything but the result line is routed to stderr at the file-descriptor level."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit_line(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["config5"], default="config3")
    ap.add_argument("--cpu-crops", type=int, default=3, help="crops timed for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spot-check", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.workload in ("config4", "config5") and args.impl == "cuda":
        import bench_configs

        bench_configs.bench = sys.modules[__name__]    # this module runs as __main__: share ITS stdout capture
        getattr(bench_configs, "run_" + args.workload)(args)
    elif args.impl == "reference":
        if args.workload == "config5":
            args.workload = "config3"
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
