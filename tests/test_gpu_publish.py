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
