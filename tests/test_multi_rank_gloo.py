"""N>1 host logic on CPU: two gloo ranks shard the env batch, step their slices with the CPU oracle
standing in for a device (no GPU here), all-gather the publish slab and compare with the unsharded run.
The GPU path uses the same partition + slab layout with NCCL (b2mj_allgather_publish)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, model_path


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, nsteps, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from mujoco_ros_pkgs_b200 import _capi, shard
    from oracle import binding as ob

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _capi.Model.from_xml_file(model_path("panda_like.xml"))
    lo, hi = shard.env_range(total, world, rank)
    rng = np.random.default_rng(5)
    qpos_all = np.tile(model.qpos0, (total, 1)) + rng.uniform(-0.1, 0.1, (total, model.nq))
    ctrl_all = rng.uniform(model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1], (total, model.nu))
    _, nmax, cnt = shard.gathered_slab_shape(total, world, model.nq + model.nv)
    local = torch.zeros(nmax, cnt, dtype=torch.float64)
    for k, e in enumerate(range(lo, hi)):
        o = ob.Oracle(model)
        o.set("qpos", qpos_all[e])
        o.set("ctrl", ctrl_all[e])
        o.step(nsteps)
        local[k, :model.nq] = torch.from_numpy(o.get("qpos"))
        local[k, model.nq:] = torch.from_numpy(o.get("qvel"))
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.stack(gathered).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partition_properties():
    from mujoco_ros_pkgs_b200 import shard

    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = shard.env_range(total, world, r)
                covered += list(range(lo, hi))
                assert 0 <= hi - lo <= -(-total // world) if total else hi == lo
            assert covered == list(range(total))
            for e in (0, total // 2, total - 1):
                if 0 <= e < total:
                    r, k = shard.owner_of(e, total, world)
                    lo, hi = shard.env_range(total, world, r)
                    assert lo + k == e and e < hi
    with pytest.raises(ValueError):
        shard.env_range(4, 2, 2)


def test_two_rank_shard_and_publish(tmp_path):
    from mujoco_ros_pkgs_b200 import _capi, shard
    from oracle import binding as ob

    total, world, nsteps = 7, 2, 15
    mp.spawn(_worker, args=(world, _free_port(), total, nsteps, str(tmp_path)), nprocs=world, join=True)
    g = np.load(tmp_path / "gathered.npy")
    assert np.load(tmp_path / "tmax.npy")[0] == world
    model = _capi.Model.from_xml_file(model_path("panda_like.xml"))
    rng = np.random.default_rng(5)
    qpos_all = np.tile(model.qpos0, (total, 1)) + rng.uniform(-0.1, 0.1, (total, model.nq))
    ctrl_all = rng.uniform(model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1], (total, model.nu))
    assert g.shape == shard.gathered_slab_shape(total, world, model.nq + model.nv)
    for e in range(total):
        o = ob.Oracle(model)
        o.set("qpos", qpos_all[e])
        o.set("ctrl", ctrl_all[e])
        o.step(nsteps)
        r, k = shard.owner_of(e, total, world)
        np.testing.assert_array_equal(g[r, k, :model.nq], o.get("qpos"))
        np.testing.assert_array_equal(g[r, k, model.nq:], o.get("qvel"))
