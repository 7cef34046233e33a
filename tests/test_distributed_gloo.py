"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: bank broadcast + crop sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, tmpdir: str) -> None:
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    from foundpose_b200 import distributed, synthetic
    from foundpose_b200.utils import projector_util, repre_util

    r, w, _ = distributed.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    repre = None
    if rank == 0:
        bank = synthetic.make_bank_tensors(5, 12, 64, num_words=8, seed=7, ragged=True)
        repre = repre_util.FeatureBasedObjectRepre(
            vertices=bank["vertices"], feat_vectors=bank["feat_vectors"],
            feat_to_template_ids=bank["feat_to_template_ids"], feat_cluster_centroids=bank["feat_cluster_centroids"],
            feat_cluster_idfs=torch.arange(8, dtype=torch.float32), template_descs=torch.rand(5, 8),
            template_desc_opts=repre_util.TemplateDescOpts(tfidf_knn_k=3),
            feat_opts=repre_util.FeatureOpts("dinov2_vits14-reg"),
            feat_raw_projectors=[projector_util.projector_from_tensordict(synthetic.make_pca(128, 64, 1))])
    got = distributed.broadcast_object_repre(repre, src=0)
    # every rank now holds a bit-identical replica
    ref = synthetic.make_bank_tensors(5, 12, 64, num_words=8, seed=7, ragged=True)
    assert torch.equal(got.feat_vectors, ref["feat_vectors"])
    assert torch.equal(got.feat_to_template_ids, ref["feat_to_template_ids"])
    assert got.feat_to_template_ids.dtype == torch.int32
    assert torch.equal(got.feat_cluster_idfs, torch.arange(8, dtype=torch.float32))
    assert got.template_desc_opts.tfidf_knn_k == 3 and got.feat_opts.extractor_name == "dinov2_vits14-reg"
    assert got.feat_raw_projectors[0].pca.components_.shape == (64, 128)
    # crops are sharded without overlap; results gathered on rank 0 in rank order
    n_crops = 11
    s, e = distributed.shard_range(n_crops, rank, world)
    local = torch.arange(s, e, dtype=torch.int64).reshape(-1, 1) * 10
    counts = [distributed.shard_range(n_crops, q, world)[1] - distributed.shard_range(n_crops, q, world)[0]
              for q in range(world)]
    outs = distributed.gather_int_results(local, counts)
    if rank == 0:
        assert torch.equal(torch.cat(outs).flatten(), torch.arange(n_crops) * 10)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")


def test_two_rank_gloo_bank_broadcast_and_sharding(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
