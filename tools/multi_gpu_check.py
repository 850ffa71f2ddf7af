"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs):
each rank owns a contiguous slice of envs (mujoco_ros_pkgs_b200.shard), steps it on its GPU, then the
qpos slab is all-gathered with b2mj_allgather_publish over a raw NCCL communicator (created here through
ctypes from the libnccl that torch ships) and rank 0 checks every env against the CPU oracle.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import ctypes as C
import glob
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_ros_pkgs_b200 import _capi, shard  # noqa: E402
from mujoco_ros_pkgs_b200.batch import BatchSim, lib  # noqa: E402


def load_nccl():
    import nvidia.nccl  # torch's bundled wheel

    cands = glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so*"))
    return C.CDLL(cands[0], mode=C.RTLD_GLOBAL)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nccl = load_nccl()

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]

    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

    total, nsteps = 64 * world, 25
    model = _capi.Model.from_xml_file(os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
    rng = np.random.default_rng(5)
    qpos_all = np.tile(model.qpos0, (total, 1)) + rng.uniform(-0.1, 0.1, (total, model.nq))
    ctrl_all = rng.uniform(model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1], (total, model.nu))
    lo, hi = shard.env_range(total, world, rank)
    sim = BatchSim(model, hi - lo, device=local)
    sim.set("qpos", qpos_all[lo:hi])
    sim.set("ctrl", ctrl_all[lo:hi])
    sim.step(nsteps)
    dst = torch.zeros(world, hi - lo, model.nq, dtype=torch.float64, device="cuda")
    rc = lib.b2mj_allgather_publish(sim.handle, _capi.field_id("qpos"), comm, C.c_void_p(dst.data_ptr()))
    assert rc == 0, _capi.last_error()
    sim.sync()
    torch.cuda.synchronize()
    g = dst.cpu().numpy()
    ok = True
    if rank == 0:
        from oracle import binding as ob

        worst = 0.0
        for e in range(0, total, 7):
            o = ob.Oracle(model)
            o.set("qpos", qpos_all[e])
            o.set("ctrl", ctrl_all[e])
            o.step(nsteps)
            r, k = shard.owner_of(e, total, world)
            worst = max(worst, float(np.max(np.abs(g[r, k] - o.get("qpos")))))
        ok = worst < 1e-9
        print(f"multi_gpu_check: world={world} total_envs={total} worst |qpos - oracle| over sampled envs = {worst:.3e} -> {'OK' if ok else 'FAIL'}")
    dist.barrier()
    nccl.ncclCommDestroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
