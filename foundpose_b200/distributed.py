"""Multi-GPU plumbing of the hot path: one process per GPU, crops sharded, bank replicated.

The reference is single-process.  Crops are independent given a read-only object
representation (scripts/infer.py:368 carries no state between instances), so the path shards by
dealing crops to ranks; there is NO collective on the per-crop path.  The only exchange is at init:
rank 0 loads / builds the object representation and broadcasts every tensor field to the replicas
(NCCL over NVLink on the GPU box; the same code runs over gloo on CPU tensors in the tests).
"""

from __future__ import annotations

import dataclasses
import os
from typing import Any, List, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialises torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def shard_range(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(num_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_round_robin(num_items: int, rank: int, world: int) -> List[int]:
    """Item ids dealt round-robin (item i goes to rank i % world)."""
    return list(range(rank, num_items, world))


def broadcast_object_repre(repre: Any, src: int = 0, device: Optional[torch.device] = None) -> Any:
    """Replicates a FeatureBasedObjectRepre from rank `src` to all ranks.

    Tensor fields are broadcast one by one (shape/dtype metadata first through broadcast_object_list),
    everything else (option tuples, projector tensordicts, camera dicts) travels as a pickled object.
    Returns the representation every rank should use (rank `src` returns its input, moved to `device`).
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return repre
    from foundpose_b200.utils import projector_util, repre_util

    if device is None and dist.get_backend() == "nccl":
        # An NCCL-only group cannot broadcast CPU tensors: stage them on this rank's GPU.
        device = torch.device("cuda", torch.cuda.current_device())
    rank = dist.get_rank()
    is_src = rank == src
    fields = [f.name for f in dataclasses.fields(repre_util.FeatureBasedObjectRepre)]
    meta: List[Any] = [None]
    if is_src:
        tensor_meta, other = {}, {}
        for name in fields:
            value = getattr(repre, name)
            if torch.is_tensor(value):
                tensor_meta[name] = (tuple(value.shape), value.dtype)
            elif name in ("feat_raw_projectors", "feat_vis_projectors"):
                other[name] = [projector_util.projector_to_tensordict(p) for p in value]
            else:
                other[name] = value
        meta = [(tensor_meta, other)]
    dist.broadcast_object_list(meta, src=src)
    tensor_meta, other = meta[0]
    out = repre if is_src else repre_util.FeatureBasedObjectRepre()
    for name, (shape, dtype) in tensor_meta.items():
        if is_src:
            t = getattr(repre, name)
            t = t.to(device) if device is not None else t
            t = t.contiguous()
        else:
            t = torch.empty(shape, dtype=dtype, device=device if device is not None else "cpu")
        dist.broadcast(t, src=src)
        setattr(out, name, t)
    if not is_src:
        for name, value in other.items():
            if name in ("feat_raw_projectors", "feat_vis_projectors"):
                value = [projector_util.projector_from_tensordict(d) for d in value]
            setattr(out, name, value)
    return out


def gather_int_results(local: torch.Tensor, world_counts: List[int]) -> Optional[List[torch.Tensor]]:
    """Gathers per-rank integer result tensors (e.g. retrieved template ids) on rank 0."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size()
    # dist.gather needs equal shapes: pad every rank's rows to the largest shard, trim on rank 0.
    rows = max(world_counts)
    padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    outs = [torch.empty_like(padded) for _ in range(world)] if dist.get_rank() == 0 else None
    dist.gather(padded, outs, dst=0)
    if outs is None:
        return None
    return [o[: world_counts[r]] for r, o in enumerate(outs)]
