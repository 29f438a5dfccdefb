"""Multi-GPU host logic: framesets are independent, so a batch is partitioned by contiguous
ranges of frameset indices, one process per GPU, with NO collective on the data path. The only exchange step
is the optional gather of the fused frames to one output rank (the rank that would publish
sensor_msgs/Image, fusion_node.cpp:799-808), done with torch.distributed (NCCL on the GPUs, gloo in the CPU
tests)."""
from typing import List, Optional, Tuple


def shard_range(n_framesets: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of frameset indices owned by `rank`; the remainder goes to the first ranks so the
    shard sizes differ by at most one and the concatenation over ranks restores stream order."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if n_framesets < 0:
        raise ValueError("negative frameset count")
    base, rem = divmod(n_framesets, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_framesets: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(n_framesets, world, r) for r in range(world))]


class FrameGather:
    """Gathers every rank's fused frames `[n_local, H, W, 3] uint8` to `dst` in stream order.

    Shards may be ragged (sizes differ by one): every rank sends a buffer padded to the largest shard, the
    destination trims. `start()` is asynchronous so that the next batch's fusion overlaps the transfer; the
    caller keeps `local` alive and unmodified until `finish()`."""

    def __init__(self, n_framesets: int, frame_shape, dst: int = 0, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dst = dst
        self.sizes = shard_sizes(n_framesets, self.world)
        self.cap = max(self.sizes) if self.sizes else 0
        self.frame_shape = tuple(frame_shape)
        self._work = None
        self._recv: Optional[list] = None
        self._pad = None

    def start(self, local):
        import torch

        if local.shape[0] != self.sizes[self.rank] or tuple(local.shape[1:]) != self.frame_shape:
            raise ValueError(f"rank {self.rank}: expected {self.sizes[self.rank]} frames of {self.frame_shape}, got {tuple(local.shape)}")
        send = local
        if local.shape[0] != self.cap:  # ragged shard: pad to the common size
            if self._pad is None:
                self._pad = torch.zeros((self.cap,) + self.frame_shape, dtype=local.dtype, device=local.device)
            self._pad[: local.shape[0]].copy_(local)
            send = self._pad
        if self.rank == self.dst:
            if self._recv is None:
                self._recv = [torch.empty((self.cap,) + self.frame_shape, dtype=local.dtype, device=local.device)
                              for _ in range(self.world)]
            self._work = self.dist.gather(send, self._recv, dst=self.dst, group=self.group, async_op=True)
        else:
            self._work = self.dist.gather(send, None, dst=self.dst, group=self.group, async_op=True)
        return self

    def finish(self):
        """Waits for the transfer; on `dst` returns the per-rank views `[sizes[r], H, W, 3]` in stream order
        (valid until the next start()), elsewhere None."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self.rank != self.dst:
            return None
        return [buf[:n] for buf, n in zip(self._recv, self.sizes)]
