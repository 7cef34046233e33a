"""The two remaining BASELINE.json configurations of bench.py (same JSON-line contract, same timing rules).

  --workload config4   configs[3]: 8 LM-O-shaped objects (800 templates x 1200 patches x 384-d each) resident on every
                       GPU, 4096 crops (512 per object) sharded over the ranks, the reference's per-object loop
                       (scripts/infer.py:179-239: load the object's representation, build its indices, run its
                       instances), banks replicated with distributed.broadcast_object_repre at init (time reported).
                       Total work is fixed -> "scaling": "strong".
  --workload config5   configs[4]: bank-size sweep 1k -> 50k templates x 1024 patches at d = 384 / 768: the k-NN kernel
                       in its HBM-bound pass (128 queries per bank sweep, GB/s vs the measured copy peak) and in
                       its tensor-bound pass (900 queries per crop, crops/s and TFLOP/s), queries sharded over the
                       ranks, next to the CPU oracle port timed on a bounded bank sample.
"""

from __future__ import annotations

import ctypes
import gc
import os
import time

import torch

import bench


def _dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def _barrier(world: int) -> None:
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def _timed(fn, steps: int, world: int, dev) -> float:
    _barrier(world)
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
    _barrier(world)
    return float(ms.item())


def run_config4(args) -> None:
    from foundpose_b200 import _native, distributed, pipeline, synthetic
    from foundpose_b200.utils import dinov2_utils

    world, rank, local_rank, dev = _dist_setup()
    lib = _native.load()
    wl = bench.WORKLOADS["config4"]
    B, n_obj = wl["batch"], wl["objects"]
    arch, opts = bench.vit_arch_and_layer(wl["vit"])
    layer = opts["layer"]
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=layer + 1)
    extractor = dinov2_utils.DinoFeatureExtractor(wl["vit"], state_dict=sd, max_batch=B).to(dev)

    # ---- init: every object's representation built on rank 0 and replicated (timed, reported) ----
    pipes, t_bcast, bank_bytes = [], 0.0, 0
    for obj in range(n_obj):
        repre, _, _, t = bench.build_repre_on_device(wl, dev, rank, world, seed=obj)
        t_bcast += t
        index = pipeline.ObjectIndex(repre, dev)
        bank_bytes += index.bank16.numel() * 2 + index.template_descs.numel() * 4
        pipes.append(pipeline.CropBatchPipeline(extractor, index, repre.feat_raw_projectors, B, crop_size=(420, 420),
                                                grid_cell_size=14.0, top_n_templates=wl["top_n"],
                                                top_k_buddies=wl["top_k"]))
        del repre
    torch.cuda.empty_cache()

    # ---- this rank's share: crops of every object dealt round-robin to the ranks ------------------
    per_obj = wl["crops_per_step"] // n_obj
    mine = len(distributed.shard_round_robin(per_obj, rank, world))      # crops per object on this rank
    micro_per_obj = (mine + B - 1) // B
    n_pool = 8
    host_images = [synthetic.make_crops(B, (420, 420), seed=100 + 1000 * rank + s).pin_memory() for s in range(n_pool)]
    host_masks = [torch.ones(B, 420, 420, dtype=torch.uint8).pin_memory() for _ in range(n_pool)]
    dev_images = [h.to(dev) for h in host_images]
    dev_masks = [h.to(dev) for h in host_masks]

    def step_resident(i: int) -> None:
        # scripts/infer.py:179-239: objects one after the other, each with its own indices; a ragged last
        # micro-batch is padded with masked-out crops (zero masks -> no query points, no correspondences).
        j = i
        for obj in range(n_obj):
            for _ in range(micro_per_obj):
                pipes[obj].run(dev_images[j % n_pool], dev_masks[j % n_pool])
                j += 1

    warmup = max(args.warmup, 3)
    for i in range(warmup):
        step_resident(i)
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.fp_launch_count()
    ms_total = _timed(step_resident, args.steps, world, dev)
    gpu_launches = int(lib.fp_launch_count() - launches0)
    clocks = sampler.stop()
    total_crops = wl["crops_per_step"]
    value = total_crops * args.steps / (ms_total / 1e3)

    # ---- end to end with host buffers ------------------------------------------------------------
    out_ref = pipes[0].engine.out
    d2h_fields = [out_ref.template_ids, out_ref.template_scores, out_ref.count, out_ref.query_ids,
                  out_ref.vertex_ids, out_ref.scores, out_ref.coord_2d, out_ref.coord_3d]
    host_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in d2h_fields] for _ in range(2)]
    stage_img = [torch.empty_like(dev_images[0]) for _ in range(2)]
    stage_msk = [torch.empty_like(dev_masks[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    d2h_done = [torch.cuda.Event() for _ in range(2)]
    cur = torch.cuda.current_stream()
    for e in consumed:
        e.record(cur)
    counter = {"i": 0}

    def issue_h2d(i: int) -> None:
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            stage_img[b].copy_(host_images[i % n_pool], non_blocking=True)
            stage_msk[b].copy_(host_masks[i % n_pool], non_blocking=True)
            h2d_done[b].record(copy_stream)

    issue_h2d(0)

    def step_e2e(_: int) -> None:
        for obj in range(n_obj):
            for _m in range(micro_per_obj):
                i = counter["i"]
                b = i % 2
                issue_h2d(i + 1)
                cur.wait_event(h2d_done[b])
                out = pipes[obj].run(stage_img[b], stage_msk[b])
                consumed[b].record(cur)
                fields = [out.template_ids, out.template_scores, out.count, out.query_ids, out.vertex_ids,
                          out.scores, out.coord_2d, out.coord_3d]
                for h, t in zip(host_out[b], fields):
                    h.copy_(t, non_blocking=True)
                d2h_done[b].record(cur)
                if i >= 1:
                    d2h_done[(i - 1) % 2].synchronize()
                counter["i"] = i + 1

    step_e2e(0)
    e2e_ms = _timed(step_e2e, args.steps, world, dev)
    e2e_value = total_crops * args.steps / (e2e_ms / 1e3)
    n_micro = n_obj * micro_per_obj
    h2d = int(n_micro * (host_images[0].numel() * 4 + host_masks[0].numel()))
    d2h = int(n_micro * sum(t.numel() * t.element_size() for t in d2h_fields))

    tensor_peak, hbm_peak, peak_src = bench.load_peaks()
    vit_flops = bench.vit_flops_per_crop(arch, layer) * n_micro * B
    if rank == 0:
        line = {
            "metric": "crops/sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 (fp32 accumulate, fp32 residual stream)", "data": "synthetic",
            "config": {"workload": wl["desc"], "objects": n_obj, "crops_total": total_crops,
                       "crops_per_object_per_rank": mine, "micro_batch": B, "micro_batches_per_rank": n_micro,
                       "bank_broadcast_s": t_bcast, "bank_bytes_per_gpu": bank_bytes,
                       "parallelism": f"crops of every object dealt round-robin to {world} rank(s); all {n_obj} banks "
                                      "resident on every GPU (distributed.broadcast_object_repre at init)",
                       "l2_policy": "inputs larger than L2: 135 MB of crops per micro-batch, 8 rotating inputs"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": gpu_launches,
            "roofline": {"kernel": "whole ViT stage (flops required for layer 9 over the step time incl. retrieval)",
                         "bound": "tensor", "achieved": vit_flops / (ms_total / args.steps * 1e-3) / 1e12,
                         "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": vit_flops / (ms_total / args.steps * 1e-3) / 1e12 / tensor_peak, "traffic": None,
                         "peak_source": peak_src},
            "cpu_baseline": None,
        }
        bench.emit_line(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def run_config5(args) -> None:
    from foundpose_b200 import _native
    from foundpose_b200.utils import knn_util
    from oracle import knn as oknn

    world, rank, local_rank, dev = _dist_setup()
    lib = _native.load()
    tensor_peak, hbm_peak, peak_src = bench.load_peaks()
    templates = [int(t) for t in os.environ.get("FP_SWEEP_TEMPLATES", "1000,2000,5000,10000,20000,50000").split(",")]
    dims = [int(d) for d in os.environ.get("FP_SWEEP_DIMS", "384,768").split(",")]
    crops = 64                                   # crops per rank per search (900 queries each)
    nq = crops * 900
    rows_out = []
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.fp_launch_count()
    t_begin = time.perf_counter()

    def knn_ms(cat: int) -> float:
        ms = ctypes.c_double()
        lib.fp_profile_read(ctypes.c_int(cat), ctypes.byref(ms), None, None, ctypes.c_int(1))
        return ms.value

    for d in dims:
        g = torch.Generator(device=dev).manual_seed(d)
        q = torch.randn((nq, d), generator=g, device=dev).to(torch.float16)
        qn = _native.row_sqnorm_f16(q)
        q128 = q[:128].float()
        for T in templates:
            F = T * 1024
            bank = torch.empty((F, d), dtype=torch.float16, device=dev)
            for s in range(0, F, 1 << 20):
                n = min(1 << 20, F - s)
                bank[s:s + n] = torch.randn((n, d), generator=g, device=dev).to(torch.float16)
            index = knn_util.KNN.from_packed(bank, _native.row_sqnorm_f16(bank), k=5, metric="l2")
            # HBM-bound pass: 128 queries, one bank sweep
            # Two figures (tools/knn_hbm_series.py, profiles/r02_knn_hbm_series.md): "burst" = the best of 20 single
            # searches after a 2-search warm-up - how MEASURED_PEAKS.json's copy figure (best of 10) was taken;
            # "sustained" = >= 0.1 s of searches after >= 0.25 s of back-to-back searches,
            # when the board has lowered the SM clock to 0.7-0.9 GHz (HBM at full rate + the tensor pipe).
            def window(n):
                for c in (4, 7):
                    knn_ms(c)
                lib.fp_profile_enable(1)
                for _ in range(n):
                    index.search(q128)
                torch.cuda.synchronize()
                lib.fp_profile_enable(0)
                return knn_ms(4) / n * 1e-3

            for _ in range(2):
                index.search(q128)
            torch.cuda.synchronize()
            _barrier(world)
            t_hbm = min(window(1) for _ in range(20))   # best of 20 single searches, like the copy peak's best of 10
            for _ in range(max(3, int(0.25 / t_hbm))):
                index.search(q128)
            t_hbm_sustained = window(max(5, int(0.1 / t_hbm)))
            # tensor-bound pass: 900 queries per crop, 64 crops per rank
            index.search_packed(q, qn)
            ms = _timed(lambda i: index.search_packed(q, qn), 2, world, dev) / 2
            flops = 2.0 * nq * F * d
            row = {"templates": T, "dim": d, "bank_gb": F * d * 2 / 1e9,
                   "hbm_pass_gbs": F * d * 2 / t_hbm / 1e9, "hbm_pass_frac": F * d * 2 / t_hbm / 1e9 / hbm_peak,
                   "hbm_pass_ms": t_hbm * 1e3,
                   "hbm_pass_sustained_gbs": F * d * 2 / t_hbm_sustained / 1e9,
                   "hbm_pass_sustained_frac": F * d * 2 / t_hbm_sustained / 1e9 / hbm_peak,
                   "k4_crops_per_s": world * crops / (ms * 1e-3), "k4_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
                   "k4_frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / tensor_peak, "k4_ms": ms}
            if rank == 0 and not args.no_cpu_baseline:
                # CPU oracle port: one crop's 900 queries vs a bounded bank sample, scaled (brute force is linear in F)
                sample = min(F, 262144)
                qc = q[:900].float().cpu()
                xc = bank[:sample].float().cpu()
                t0 = time.perf_counter()
                oknn.knn_l2_blocked(qc, ((s, xc[s:s + 65536]) for s in range(0, sample, 65536)), 5)
                t_cpu = (time.perf_counter() - t0) * (F / sample)
                row["cpu_k4_crops_per_s"] = 1.0 / t_cpu
                row["cpu_sample_rows"] = sample
            rows_out.append(row)
            del index, bank
            gc.collect()                 # KNN.index refers to itself: the bank is only released by the collector
            torch.cuda.empty_cache()
    clocks = sampler.stop()
    gpu_launches = int(lib.fp_launch_count() - launches0)
    if rank == 0:
        best = max(rows_out, key=lambda r: r["templates"] * 1000 + r["dim"])
        line = {
            "metric": "kNN HBM GB/s vs peak; k-NN crops/sec", "value": best["k4_crops_per_s"], "unit": "crops/s",
            "n_gpus": world, "steps": 2, "warmup": 1, "ms_per_step": best["k4_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "configs[4]: bank-size sweep 1k->50k templates x 1024 patches at d in {384, 768}; "
                                   "value = k-NN-only crops/s at the largest bank", "queries_per_rank": nq,
                       "l2_policy": "banks of 0.8-78.6 GB, far larger than L2"},
            "clocks": clocks, "gpu_launches": gpu_launches, "sweep": rows_out, "cpu_cores": cores,
            "peaks": {"hbm_gbs": hbm_peak, "tensor_tflops": tensor_peak, "source": peak_src},
            "wall_s": time.perf_counter() - t_begin,
            "roofline": {"bound": "hbm", "achieved": best["hbm_pass_gbs"], "peak": hbm_peak, "unit": "GB/s",
                         "frac": best["hbm_pass_frac"], "traffic": None,
                         "sustained_achieved": best["hbm_pass_sustained_gbs"], "sustained_frac": best["hbm_pass_sustained_frac"],
                         "note": "achieved / frac: best of 20 single searches (the copy peak is a best-of-10 burst figure "
                                 "too); sustained_*: after 0.25 s of back-to-back searches, when the board has lowered the SM clock "
                                 "(128 queries per sweep at the copy peak are also 819 TFLOP/s of tensor work)",
                         "kernel": "knn_kernel<5>, 128 queries per bank sweep, largest bank of the sweep"},
            "cpu_baseline": None, "e2e": None,
        }
        bench.emit_line(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
