"""Env sharding across ranks: contiguous env ranges, one process per GPU, no data-path
collective.  The only exchange is the publish all-gather of a per-rank [nenv_local][count] slab; the
layout helpers here are what bench.py and the C-ABI's b2mj_allgather_publish agree on."""
from __future__ import annotations


def env_range(total_envs: int, world_size: int, rank: int) -> tuple[int, int]:
    """[start, stop) of the envs rank owns: contiguous, sizes differ by at most one (first ranks larger)."""
    if not (0 <= rank < world_size) or total_envs < 0:
        raise ValueError("bad rank / size")
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def owner_of(env: int, total_envs: int, world_size: int) -> tuple[int, int]:
    """(rank, local index) of a global env id under env_range's partition."""
    base, rem = divmod(total_envs, world_size)
    cut = rem * (base + 1)
    if env < cut:
        return env // (base + 1), env % (base + 1)
    if base == 0:
        raise ValueError("env out of range")
    return rem + (env - cut) // base, (env - cut) % base


def gathered_slab_shape(total_envs: int, world_size: int, count: int) -> tuple[int, int, int]:
    """Shape of the all-gather destination [world][max_local][count] (ranks are padded to max_local)."""
    return world_size, -(-total_envs // world_size), count
