"""The one exchange step of the path on real GPUs: two ranks, one GPU each, NCCL all-gather of the packed
qpos | qvel | sensordata slab through b2mj_allgather_publish_multi (raw ncclComm_t), checked against the CPU oracle.
Skips cleanly on a box with fewer than 2 GPUs; the packing kernel alone is checked on one GPU."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from mujoco_ros_pkgs_b200 import _capi, shard
from mujoco_ros_pkgs_b200.batch import BatchSim
from mujoco_ros_pkgs_b200.nccl_comm import NcclComm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = NcclComm(rank, world)
model = _capi.Model.from_xml_file(os.path.join(%(root)r, "mujoco_ros_pkgs_b200", "models", "humanoid_like.xml"))
total, nsteps = 48 * world, 30
rng = np.random.default_rng(5)
qpos_all = np.tile(model.qpos0, (total, 1))
qpos_all[:, 7:] += rng.uniform(-0.05, 0.05, (total, model.nq - 7))
ctrl_all = rng.uniform(-1, 1, (total, model.nu))
lo, hi = shard.env_range(total, world, rank)
sim = BatchSim(model, hi - lo, device=local)
sim.set("qpos", qpos_all[lo:hi]); sim.set("ctrl", ctrl_all[lo:hi])
row = model.nq + model.nv + model.nsensordata
dst = torch.zeros(world, hi - lo, row, dtype=torch.float64, device="cuda")
for s in range(nsteps):
    sim.step(1)
    sim.allgather_publish_multi(["qpos", "qvel", "sensordata"], comm.ptr, dst.data_ptr())
sim.sync(); torch.cuda.synchronize()
g = dst.cpu().numpy()
ok = True
if rank == 0:
    from oracle import binding as ob
    worst = 0.0
    for e in range(0, total, 5):
        o = ob.Oracle(model)
        o.set("qpos", qpos_all[e]); o.set("ctrl", ctrl_all[e])
        o.step(nsteps)
        r, k = shard.owner_of(e, total, world)
        want = np.concatenate([o.get("qpos"), o.get("qvel"), o.get("sensordata")])
        worst = max(worst, float(np.max(np.abs(g[r, k] - want) / (1 + np.abs(want)))))
    ok = worst < 1e-5
    print(f"publish check: world={world} worst rel err {worst:.3e}", flush=True)
dist.barrier()
comm.destroy()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


FUSED_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from mujoco_ros_pkgs_b200 import _capi, shard
from mujoco_ros_pkgs_b200.batch import BatchSim
from mujoco_ros_pkgs_b200.nccl_comm import NcclComm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = NcclComm(rank, world)
model = _capi.Model.from_xml_file(os.path.join(%(root)r, "mujoco_ros_pkgs_b200", "models", "humanoid_like.xml"))
per, nsteps = 300, 25
total = per * world
rng = np.random.default_rng(5)
qpos_all = np.tile(model.qpos0, (total, 1))
qpos_all[:, 7:] += rng.uniform(-0.05, 0.05, (total, model.nq - 7))
ctrl_all = rng.uniform(-1, 1, (total, model.nu))
lo, hi = rank * per, (rank + 1) * per
fields = ["qpos", "qvel", "sensordata"]
a, b = BatchSim(model, per, device=local), BatchSim(model, per, device=local)
for s_ in (a, b):
    s_.set("qpos", qpos_all[lo:hi]); s_.set("ctrl", ctrl_all[lo:hi])
handle = a.publish_fused_create(world, rank, fields)
handles = [None] * world
dist.all_gather_object(handles, handle)
a.publish_fused_connect(handles)
dist.barrier()
row = model.nq + model.nv + model.nsensordata
ref = torch.zeros(world, per, row, dtype=torch.float64, device="cuda")
ok = True
for s in range(nsteps):
    a.step_publish()
    ptr, cnt = a.publish_fused_wait()
    assert cnt == row
    a.sync()
    class Slab:
        __cuda_array_interface__ = {"shape": (world, per, row), "typestr": "<f8", "data": (ptr, False), "version": 3}
    got = torch.as_tensor(Slab(), device="cuda").clone()
    b.step(1)
    b.allgather_publish_multi(fields, comm.ptr, ref.data_ptr())
    b.sync(); torch.cuda.synchronize()
    if not torch.equal(got, ref):
        ok = False
        print(f"rank {rank} step {s}: fused publish differs from the NCCL all-gather, max abs {(got - ref).abs().max().item():.3e}", flush=True)
        break
print(f"fused publish check: rank {rank} world {world} ok={ok}", flush=True)
dist.barrier()
del a, b
comm.destroy()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_nccl_publish_matches_oracle(capi, tmp_path):
    if capi.lib.b2mj_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "publish_worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "publish check" in r.stdout


def test_two_rank_fused_publish_equals_nccl_all_gather(capi, tmp_path):
    """b2mj_step_publish: the step kernel stores every finished env's row into every rank's slab over peer memory; the
    gathered slab must be bitwise what step + pack + ncclAllGather produces, at every step (double-buffered slabs)."""
    if capi.lib.b2mj_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "fused_worker.py"
    script.write_text(FUSED_WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok=True") == 2, r.stdout[-2000:]


def test_fused_publish_single_rank(load_model, capi):
    """world = 1: the fused stores land in the local slab; equals the packed slab of the plain step."""
    import torch

    from mujoco_ros_pkgs_b200.batch import BatchSim

    model = load_model("panda_like.xml")
    nenv = 130
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    rng = np.random.default_rng(2)
    ctrl = rng.uniform(-1, 1, (nenv, model.nu))
    for s_ in (a, b):
        s_.set("ctrl", ctrl)
    h = a.publish_fused_create(1, 0, ["qpos", "qvel", "sensordata", "time"])
    a.publish_fused_connect([h])
    row = model.nq + model.nv + model.nsensordata + 1
    for s in range(15):
        a.step_publish()
        ptr, cnt = a.publish_fused_wait()
        a.sync()
        assert cnt == row

        class Slab:
            __cuda_array_interface__ = {"shape": (nenv, row), "typestr": "<f8", "data": (ptr, False), "version": 3}

        got = torch.as_tensor(Slab(), device="cuda").cpu().numpy()
        b.step(1)
        want = np.concatenate([b.get("qpos"), b.get("qvel"), b.get("sensordata"), b.get("time")], axis=1)
        np.testing.assert_array_equal(got, want)
    with pytest.raises(capi.B2mjError, match="state record"):
        a.publish_fused_create(1, 0, ["xpos"])


def test_publish_pack_layout_single_gpu(load_model, capi):
    """b2mj_publish_pack: the packed slab is [nenv][qpos | qvel | sensordata] exactly."""
    import torch

    from mujoco_ros_pkgs_b200.batch import BatchSim

    model = load_model("humanoid_like.xml")
    nenv = 37
    sim = BatchSim(model, nenv)
    rng = np.random.default_rng(1)
    sim.set("ctrl", rng.uniform(-1, 1, (nenv, model.nu)))
    sim.step(12)
    ptr, row = sim.publish_pack(["qpos", "qvel", "sensordata"])
    assert row == model.nq + model.nv + model.nsensordata
    sim.sync()
    class Slab:  # the slab pointer is a raw device address: view it through the CUDA array interface
        __cuda_array_interface__ = {"shape": (nenv, row), "typestr": "<f8", "data": (ptr, False), "version": 3}

    got = torch.as_tensor(Slab(), device="cuda").cpu().numpy()
    want = np.concatenate([sim.get("qpos"), sim.get("qvel"), sim.get("sensordata")], axis=1)
    np.testing.assert_array_equal(got, want)
