"""Multi-GPU plumbing: one process per GPU, sample-pass sharding, one film reduction.

The reference has no distributed path.  Path samples are independent and the film
is additive - per bin it stores (sum of value*weight, sum of weight) and is developed as a ratio
afterwards (pyrite/src/film.rs:132-185) - so the render shards with no data-path collective:
rank r of W renders path samples r, r+W, r+2W, ... of every tile (`pyr_render_params.sample_offset /
sample_stride`; per-path RNG streams are keyed by (seed, tile, sample), so the union over ranks is
exactly the single-GPU job) and the films are summed ONCE at the end with an NCCL reduce over
NVLink, *before* developing (the ratio is not linear).  Tile sharding is not used because the
bidirectional integrator splats light-traced samples to arbitrary pixels
(pyrite/src/renderer/bidirectional.rs:253-306).

The collective itself lives behind the C ABI (`pyr_comm_init` / `pyr_film_reduce`, include/pyrite_b200.h): the library owns
the NCCL communicator and a host in any language only carries the 128-byte id from rank 0 to the others.  Here that
carrier is torch.distributed's object broadcast (any backend); `reduce_film` on a tensor remains for hosts that already
hold the film as a tensor and for the gloo tests of the host logic.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_for_rank(rank: int, world: int) -> Tuple[int, int]:
    """(sample_offset, sample_stride) of `rank` in a `world`-rank job."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    return rank, world


def samples_of_shard(iterations: int, offset: int, stride: int) -> int:
    """How many of a tile's `iterations` path samples (= area * spp) the shard renders."""
    return 0 if offset >= iterations else (iterations - offset + stride - 1) // stride


def reduce_film(film: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor:
    """Sum the (accumulator, weight) films of all ranks into rank `dst` (in place).  Works on the
    CUDA alias of the context's film (NCCL) and on CPU tensors (gloo, used by the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def all_reduce_film(film: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    return film


def init_film_comm(renderer, group=None, wait: bool = True) -> None:
    """Create the renderer's own NCCL communicator over the ranks of the (already initialised) torch.distributed group:
    rank 0 draws the id (`pyr_comm_unique_id`), the object broadcast carries it, every rank calls `pyr_comm_init`.
    `wait=False`: `pyr_comm_init_async` - the library sets the communicator up on its own thread and stream while the
    caller renders (seconds on an 8-GPU box), and `film_reduce` waits for it."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    box = [type(renderer).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    renderer.comm_init(world, rank, box[0], wait=wait)
    renderer._film_comm = True


def render_sharded(renderer, seed: int = 0, spp: int = 0, develop: bool = True, **kw):
    """Render this rank's share of the job, sum the films on rank 0 (`pyr_film_reduce`: NCCL inside the library) and
    develop there.  `renderer` is a pyrite_b200.api.Renderer with a project loaded; with world > 1 a torch.distributed
    process group must be initialised (it only carries the communicator id).  Returns (xyz, srgb) on rank 0 and
    (None, None) elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    offset, stride = shard_for_rank(rank, world)
    if world > 1 and not getattr(renderer, "_film_comm", False):
        init_film_comm(renderer, wait=False)   # set up under the render
    renderer.render(seed=seed, spp=spp, sample_offset=offset, sample_stride=stride, **kw)
    if world > 1:
        renderer.film_reduce(0)
    if develop and rank == 0:
        return renderer.develop()
    return None, None
