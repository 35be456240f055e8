"""Multi-GPU plumbing (one process per GPU, torch.distributed): template sharding by tmpl_idx, scene
sharding, and the merge of per-rank top-K match lists with one small all-gather.
The data path has no other collective: every rank builds (or receives) the scene's feature map and
searches its own shard.  Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import MATCH_DTYPE


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [begin, end) of n_items for `rank` of `world`."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_templates(templates, rank, world):
    """Contiguous shard of the template list; returns (templates_of_rank, tmpl_idx_base)."""
    b, e = shard_range(len(templates), rank, world)
    return templates[b:e], b


def shard_scenes(n_scenes, rank, world):
    """Scene s goes to rank s % world (config 5: multi-view batches)."""
    return list(range(rank, n_scenes, world))


def merge_topk_records(per_rank, k):
    """Merge per-rank top-k MATCH_DTYPE arrays: ascending score, ties by (rank order, local order)."""
    allm = np.concatenate([np.asarray(r, MATCH_DTYPE) for r in per_rank]) if per_rank else np.zeros(0, MATCH_DTYPE)
    order = np.argsort(allm["score"], kind="stable")
    return allm[order[:k]]


def allgather_topk(local_topk, k, device=None, group=None):
    """All-gather the ranks' top-k lists (k x 32 B each, padded) and merge them locally on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = np.zeros(k, MATCH_DTYPE)
    n = min(k, len(local_topk))
    local[:n] = local_topk[:n]
    local["score"][n:] = np.inf
    local["tmpl_idx"][n:] = -1
    if world == 1:
        return merge_topk_records([local[:n]], k)
    buf = torch.from_numpy(local.view(np.uint8).copy())
    if device is not None:
        buf = buf.to(device, non_blocking=False)
    out = torch.empty(world * buf.numel(), dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    rec = out.cpu().numpy().view(MATCH_DTYPE).reshape(world, k)
    return merge_topk_records([r[r["tmpl_idx"] >= 0] for r in rec], k)


def broadcast_featuremap(featuremap, src=0, group=None):
    """NCCL-broadcast the planes of `featuremap` from rank `src` into the same-shaped maps of the other ranks
    (alternative to every rank rebuilding the scene's map; SURVEY.md §8e — use whichever measures faster)."""
    ptr, nbytes = featuremap.device_ptr()
    n = nbytes // 4

    class _Blob:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}

    t = torch.as_tensor(_Blob(), device=f"cuda:{featuremap.device}")
    dist.broadcast(t, src=src, group=group)
    torch.cuda.synchronize(featuremap.device)
