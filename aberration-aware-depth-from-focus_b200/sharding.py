"""Multi-GPU partitioning of focal-stack synthesis.

The path has no exchange step: every output pixel of every (image n, slice s) depends only on that
image, its depth map and one focus distance.  The unit of the partition is the **tile row** -- 8 image
rows of one (image, slice), the row of tiles the fused kernel walks -- numbered in (image, slice, row)
order, R = (n*S + s)*ceil(H/8) + h/8.  Every rank gets one contiguous, balanced run of tile rows and
renders it with ONE launch (``aadff_render_stack_rows_f32``), whatever the shape of the batch:

    c3 (16 images x 5 slices x 256 rows)   8 GPUs: 320 tile rows each = 2 images, all slices
    c2 ( 1 image  x 5 slices x 512 rows)   8 GPUs:  40 tile rows each = 5/8 of a slice -> slice x row-band split
    c4 ( 1 image  x 10 slices x 1080 rows) 8 GPUs: 168-169 tile rows each (1.25 slices)

so no GPU idles when there are fewer (image, slice) items than ranks.  Inputs (16 B per pixel) and the
2.3 MB PSFNet are replicated.  The single collective, ``all_gather`` of the rendered rows, exists for
verification / for callers that want the whole stack on every rank -- it is never on the hot path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE_ROW_H = 8          # == aadff_tile_row_height(); checked against the library in tests


def tile_rows_per_slice(H: int) -> int:
    return -(-H // TILE_ROW_H)


def item_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced split: the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_row_range(N: int, S: int, H: int, world: int, rank: int) -> tuple[int, int]:
    """This rank's run [R0, R1) of the N*S*ceil(H/8) tile rows."""
    return item_range(N * S * tile_rows_per_slice(H), world, rank)


def flat_row(R: int, H: int) -> int:
    """First image row of tile row R in the flattened (image, slice, row) list of N*S*H rows."""
    ty = tile_rows_per_slice(H)
    return (R // ty) * H + min((R % ty) * TILE_ROW_H, H)


def local_runs(N: int, S: int, H: int, world: int, rank: int) -> list[tuple[int, int, int, int]]:
    """The same run as (image n, slice s, first row, end row) pieces -- what the rank actually renders."""
    R0, R1 = tile_row_range(N, S, H, world, rank)
    ty = tile_rows_per_slice(H)
    runs = []
    while R0 < R1:
        ns, t0 = divmod(R0, ty)
        t1 = min(ty, t0 + (R1 - R0))
        runs.append((ns // S, ns % S, t0 * TILE_ROW_H, min(t1 * TILE_ROW_H, H)))
        R0 += t1 - t0
    return runs


@torch.no_grad()
def render_stack_sharded(lens, img, depth, foc_dists, rank: int, world: int, gather: bool = True,
                         group=None, mode=None):
    """Render this rank's tile rows of the [N,C,S,H,W] stack with ``lens.render_stack_rows`` (one launch) and, if
    ``gather``, all-gather the shares so that every rank returns the full stack -- bit-identical to the
    single-GPU ``render_stack``: a pixel's value does not depend on which tile or launch computes it.
    Returns (stack, (R0, R1)) or, with gather=False, (rows [flat_rows, C, W], (R0, R1))."""
    N, C, H, W = img.shape
    S = foc_dists.shape[1]
    R0, R1 = tile_row_range(N, S, H, world, rank)
    kw = {} if mode is None else {"mode": mode}
    local = lens.render_stack_rows(img, depth, foc_dists, R0, R1, **kw)         # [flat rows, C, W]
    if not gather:
        return local, (R0, R1)
    counts = [flat_row(item_range(N * S * tile_rows_per_slice(H), world, r)[1], H) -
              flat_row(item_range(N * S * tile_rows_per_slice(H), world, r)[0], H) for r in range(world)]
    per = max(counts)                                        # pad every share to the same length
    padded = local.new_zeros((per, C, W))
    padded[:local.shape[0]] = local
    shares = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(shares, padded, group=group)
    rows = torch.cat([shares[r][:counts[r]] for r in range(world)], 0)          # [N*S*H, C, W]
    return rows.view(N, S, H, C, W).permute(0, 3, 1, 2, 4).contiguous(), (R0, R1)
