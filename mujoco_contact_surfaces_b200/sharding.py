"""Env-index sharding across the GPUs of one box.

Environments are independent (one env = one mjData of the reference, mujoco_contact_surfaces_plugin.cpp:88),
so the batch is cut into contiguous blocks [start, start+count) per rank, static geometry is replicated
on every GPU at finalize, and NOTHING is exchanged on the step path.  torch.distributed (NCCL on the GPUs,
gloo in the CPU tests) is used only to launch one process per GPU and to gather results / timings.
"""
import numpy as np


def shard_range(n_envs_total, rank, world_size):
    """Contiguous block of env indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_envs_total, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class ShardedBatch:
    """Runs the shard of a global env batch that belongs to this rank.

    `make_engine(n_local_envs)` must return an object with the engine surface (step / pair_results /
    geom_wrenches / sensor_image): the CUDA HydroelasticEngine in production.
    """

    def __init__(self, scene, n_envs_total, make_engine, rank=0, world_size=1):
        self.scene, self.n_total, self.rank, self.world = scene, n_envs_total, rank, world_size
        self.start, self.count = shard_range(n_envs_total, rank, world_size)
        self.engine = make_engine(self.count) if self.count > 0 else None

    def local_poses(self, seed):
        """Per-env RNG streams make the shard's poses identical to the same envs of the full batch."""
        return self.scene.poses(self.count, seed, env_offset=self.start)

    def step(self, xpos, xmat, vel, with_sensors=False):
        """xpos/xmat/vel are either this rank's shard [count, ...] or the full batch [n_total, ...]."""
        if self.engine is None:
            return
        if len(xpos) == self.n_total and self.n_total != self.count:
            sl = slice(self.start, self.start + self.count)
            xpos, xmat, vel = xpos[sl], xmat[sl], vel[sl]
        self.engine.step(xpos, xmat, vel, with_sensors=with_sensors)

    def gather_pair_results(self):
        """All ranks' per-pair results in global env order (collective; the same array on every rank)."""
        local = self.engine.pair_results() if self.engine is not None else None
        return gather_env_major(local, self.count, self.world)


def gather_env_major(local, count, world_size):
    """Concatenate per-rank env-major arrays in rank order.  world_size == 1 needs no process group."""
    if world_size == 1:
        return local
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world_size)]
    dist.all_gather(counts, torch.tensor([count], dtype=torch.int64, device=dev))
    counts = [int(c) for c in counts]
    # the row layout (dtype + trailing shape) comes from the first rank that owns envs
    meta = [None] * world_size
    dist.all_gather_object(meta, None if local is None else (local.dtype.descr if local.dtype.names else local.dtype.str,
                                                             local.shape[1:]))
    descr, tail = next(m for m in meta if m is not None)
    dtype = np.dtype(descr, align=True) if isinstance(descr, list) else np.dtype(descr)
    row = int(np.prod(tail, dtype=np.int64)) * dtype.itemsize if tail else dtype.itemsize
    bufs = [torch.zeros(max(c, 0) * row, dtype=torch.uint8, device=dev) for c in counts]
    mine = np.zeros(0, dtype=np.uint8) if local is None else np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    send = torch.from_numpy(mine.copy()).to(dev)
    # all_gather needs equal sizes: pad to the largest shard
    biggest = max(counts) * row
    padded = torch.zeros(biggest, dtype=torch.uint8, device=dev)
    padded[:send.numel()] = send
    recv = [torch.zeros(biggest, dtype=torch.uint8, device=dev) for _ in range(world_size)]
    dist.all_gather(recv, padded)
    parts = []
    for c, r in zip(counts, recv):
        if c > 0:
            parts.append(r[:c * row].cpu().numpy().view(dtype).reshape((c,) + tuple(tail)))
    del bufs
    return np.concatenate(parts, axis=0)
