"""Multi-GPU partitioning of focal-stack synthesis.

The path has no exchange step: every (image n, slice s) work item depends only on that image,
its depth map and one focus distance.  Items are dealt to ranks in contiguous, balanced runs of
the n-major item list, so a rank touches as few distinct images as possible; the PSFNet weights
(2.3 MB) are replicated.  The single collective, ``all_gather`` of the rendered items, exists for
verification / for callers that want the whole stack on every rank -- it is never on the hot path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def item_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced split: the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_runs(N: int, S: int, world: int, rank: int) -> list[tuple[int, int, int]]:
    """This rank's items as (image n, first slice, end slice) runs."""
    lo, hi = item_range(N * S, world, rank)
    runs = []
    while lo < hi:
        n, s0 = divmod(lo, S)
        s1 = min(S, s0 + (hi - lo))
        runs.append((n, s0, s1))
        lo += s1 - s0
    return runs


@torch.no_grad()
def render_stack_sharded(lens, img, depth, foc_dists, rank: int, world: int, gather: bool = True,
                         group=None, mode=None):
    """Render this rank's share of the [N,C,S,H,W] stack with ``lens.render_stack`` and, if
    ``gather``, all-gather the shares so every rank returns the full stack (bit-identical to the
    single-GPU result: the kernel is deterministic per item).  Returns (stack_or_local, runs)."""
    N, C, H, W = img.shape
    S = foc_dists.shape[1]
    runs = local_runs(N, S, world, rank)
    kw = {} if mode is None else {"mode": mode}
    parts = [lens.render_stack(img[n:n + 1], depth[n:n + 1], foc_dists[n:n + 1, s0:s1], **kw)[0].transpose(0, 1)
             for (n, s0, s1) in runs]                      # each [s1-s0, C, H, W]
    local = torch.cat(parts, 0) if parts else img.new_zeros((0, C, H, W))
    if not gather:
        return local, runs
    per = -(-N * S // world)                               # ceil: pad every share to the same length
    padded = local.new_zeros((per, C, H, W))
    padded[:local.shape[0]] = local
    shares = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(shares, padded, group=group)
    items = torch.cat([shares[r][:item_range(N * S, world, r)[1] - item_range(N * S, world, r)[0]]
                       for r in range(world)], 0)           # [N*S, C, H, W], n-major
    return items.view(N, S, C, H, W).permute(0, 2, 1, 3, 4).contiguous(), runs
