"""Multi-rank host logic on CPU: gloo, world_size 2.  Each rank renders its sample-pass shard with the
CPU emulation of the device code and the films are reduced exactly as bench.py / render_sharded do."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import scene_ir
    from emu_lib import Emu

    from pyrite_b200.distributed import reduce_film, samples_of_shard, shard_for_rank

    emu = Emu(scene_ir("cornell"), (64, 64, 64))
    offset, stride = shard_for_rank(rank, world)
    emu.render(seed=21, spp=4, sample_offset=offset, sample_stride=stride)
    film = torch.from_numpy(emu.film())
    mine = torch.tensor([float(film[..., 1].sum())], dtype=torch.float64)
    reduce_film(film, dst=0)
    total = mine.clone()
    dist.all_reduce(total)
    expected = sum(samples_of_shard(32 * 32 * 4, *shard_for_rank(r, world)) for r in range(world)) * 4
    assert expected == 64 * 64 * 4
    if rank == 0:
        np.save(Path(out_dir) / "reduced.npy", film.numpy())
        np.save(Path(out_dir) / "weights.npy", np.array([float(total[0])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single(tmp_path):
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import scene_ir
    from emu_lib import Emu

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    emu = Emu(scene_ir("cornell"), (64, 64, 64))
    emu.render(seed=21, spp=4)
    whole = emu.film()
    assert np.array_equal(whole[..., 1], reduced[..., 1])
    assert np.allclose(whole[..., 0], reduced[..., 0], rtol=1e-5, atol=1e-6)
    assert float(np.load(tmp_path / "weights.npy")[0]) == float(whole[..., 1].sum())


def test_shard_plan_partitions_every_tile():
    from pyrite_b200.distributed import samples_of_shard, shard_for_rank

    for world in (1, 2, 3, 4, 8):
        for iterations in (0, 1, 7, 1024, 32 * 32 * 256 + 5):
            assert sum(samples_of_shard(iterations, *shard_for_rank(r, world)) for r in range(world)) == iterations
    with pytest.raises(ValueError):
        shard_for_rank(2, 2)
