"""Row sharding of the similarity rows across the GPUs of one box.

Every rank holds the full (replicated) rating matrix and owns a contiguous block of left rows;
it computes that block against all N rows and selects top-k locally.  The only exchange is one
all-gather of the fixed-size neighbour lists (int32 idx, float64 sim)[rows][k] over NCCL/NVLink
(gloo in the CPU tests).  torch.distributed is the plumbing, not the product."""
from __future__ import annotations


def shard_rows(n: int, world: int, rank: int, align: int = 128):
    """[begin,end) of the left rows owned by `rank`: contiguous, `align`-row aligned boundaries
    (the tensor-core tile height) except for the last shard, sizes differing by <= align."""
    blocks = (n + align - 1) // align
    b0 = blocks * rank // world
    b1 = blocks * (rank + 1) // world
    return min(b0 * align, n), min(b1 * align, n)


def allgather_topk(idx_local, sim_local, n: int, k: int, group=None, align: int = 128):
    """All-gather the per-shard neighbour lists into the full (n, k) lists on every rank.
    idx_local/sim_local: torch tensors (rows_of_this_rank, k) on the collective's device."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    bounds = [shard_rows(n, world, r, align) for r in range(world)]
    max_rows = max(b - a for a, b in bounds)
    pad_i = torch.full((max_rows, k), -1, dtype=torch.int32, device=idx_local.device)
    pad_s = torch.full((max_rows, k), float("nan"), dtype=torch.float64, device=sim_local.device)
    pad_i[: idx_local.shape[0]] = idx_local
    pad_s[: sim_local.shape[0]] = sim_local
    all_i = torch.empty((world * max_rows, k), dtype=torch.int32, device=idx_local.device)
    all_s = torch.empty((world * max_rows, k), dtype=torch.float64, device=sim_local.device)
    dist.all_gather_into_tensor(all_i, pad_i, group=group)
    dist.all_gather_into_tensor(all_s, pad_s, group=group)
    out_i = torch.cat([all_i[r * max_rows: r * max_rows + (b - a)] for r, (a, b) in enumerate(bounds)])
    out_s = torch.cat([all_s[r * max_rows: r * max_rows + (b - a)] for r, (a, b) in enumerate(bounds)])
    return out_i, out_s


def slab_owner(slab: int, world: int) -> int:
    """Which shard computes slab `slab` under symmetric-slab sharding: slabs are dealt in snake order
    (0..world-1, world-1..0, ...) because slab s costs ~ (n_slabs - s) — the same rule as csrc/api.cu."""
    rnd, pos = divmod(slab, world)
    return world - 1 - pos if rnd & 1 else pos


def allgather_partial_topk(idx_part, sim_part, group=None):
    """Symmetric-slab sharding (rs_knn_params.shard_count >= 1): every rank holds PARTIAL neighbour
    lists for all n rows, (n, k).  Returns the stacked (world, n, k) tensors on every rank — the one
    exchange step of the sharded Fit, an all-gather over NCCL/NVLink (gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n, k = idx_part.shape
    all_i = torch.empty((world * n, k), dtype=idx_part.dtype, device=idx_part.device)
    all_s = torch.empty((world * n, k), dtype=sim_part.dtype, device=sim_part.device)
    dist.all_gather_into_tensor(all_i, idx_part.contiguous(), group=group)
    dist.all_gather_into_tensor(all_s, sim_part.contiguous(), group=group)
    return all_i.view(world, n, k), all_s.view(world, n, k)


def union_topk_device(all_i, all_s):
    """Unites stacked partial lists (n_lists, n, k) into the final (n, k) lists on the device
    (rs_knn_topk_union_device; CUDA tensors, the current torch stream)."""
    import torch

    from . import core

    n_lists, n, k = all_i.shape
    out_i = torch.empty((n, k), dtype=torch.int32, device=all_i.device)
    out_s = torch.empty((n, k), dtype=torch.float64, device=all_s.device)
    core._check(core.knn_lib().rs_knn_topk_union_device(
        n_lists, n, k, all_i.data_ptr(), all_s.data_ptr(), out_i.data_ptr(), out_s.data_ptr(),
        torch.cuda.current_stream(all_i.device).cuda_stream))
    return out_i, out_s


def allgather_predictions(pred_local, counts, group=None):
    """All-gather per-rank prediction vectors of (known) lengths `counts` into one vector."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    m = max(counts)
    pad = torch.zeros(m, dtype=torch.float64, device=pred_local.device)
    pad[: pred_local.shape[0]] = pred_local
    out = torch.empty(world * m, dtype=torch.float64, device=pred_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m: r * m + c] for r, c in enumerate(counts)])
