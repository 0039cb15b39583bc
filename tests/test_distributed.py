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


class _RecordingRenderer:
    """Stands in for api.Renderer in the host-logic test below: records what the plumbing calls, in order."""
    ids_drawn = 0

    def __init__(self):
        self.calls = []

    @staticmethod
    def comm_unique_id():
        _RecordingRenderer.ids_drawn += 1
        return bytes([7] * 64 + [os.getpid() % 251] * 64)

    def comm_init(self, n_ranks, rank, comm_id, wait=True):
        self.calls.append(("comm_init", n_ranks, rank, bytes(comm_id), wait))

    def render(self, **kw):
        self.calls.append(("render", kw["sample_offset"], kw["sample_stride"], kw["seed"], kw["spp"]))

    def film_reduce(self, root):
        self.calls.append(("film_reduce", root))

    def develop(self):
        self.calls.append(("develop",))
        return "xyz", "srgb"


def _plumbing_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json

    from pyrite_b200.distributed import render_sharded

    r = _RecordingRenderer()
    first = render_sharded(r, seed=5, spp=8)
    second = render_sharded(r, seed=6, spp=8, develop=False)   # the communicator exists now: no second set-up
    calls = [[c[0]] + [x.hex() if isinstance(x, bytes) else x for x in c[1:]] for c in r.calls]
    (Path(out_dir) / f"calls{rank}.json").write_text(json.dumps({"calls": calls, "ids_drawn": _RecordingRenderer.ids_drawn,
                                                                  "first": list(first), "second": list(second)}))
    dist.barrier()
    dist.destroy_process_group()


def test_render_sharded_plumbing_on_two_ranks(tmp_path):
    """`render_sharded` on two gloo ranks with a recording renderer: rank 0 draws ONE communicator id, both ranks hand the same 128 bytes
    to `comm_init` without waiting for the set-up (`pyr_comm_init_async`: it runs under the render), every rank renders its own sample
    passes, the reduce comes after the render and only rank 0 develops."""
    import json

    world = 2
    mp.spawn(_plumbing_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [json.loads((tmp_path / f"calls{r}.json").read_text()) for r in range(world)]
    assert got[0]["ids_drawn"] == 1 and got[1]["ids_drawn"] == 0
    ids = set()
    for rank, g in enumerate(got):
        names = [c[0] for c in g["calls"]]
        assert names == (["comm_init", "render", "film_reduce"] + (["develop"] if rank == 0 else []) + ["render", "film_reduce"]), names
        init = g["calls"][0]
        assert init[1] == world and init[2] == rank and init[4] is False and len(bytes.fromhex(init[3])) == 128
        ids.add(init[3])
        assert g["calls"][1][1:] == [rank, world, 5, 8]
        assert g["first"] == (["xyz", "srgb"] if rank == 0 else [None, None]) and g["second"] == [None, None]
    assert len(ids) == 1
