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


class PeerFrameBuffer:
    """The gather WITHOUT a second pass: the output rank owns one device buffer `[n_framesets, H, W, 3]`, exports it
    (CUDA IPC through the C ABI: dfb_ipc_export / dfb_ipc_open) and every rank maps it. A rank fuses with
    `d_out = slice_ptr()` — its contiguous range of the stream, see shard_range — so the stream kernel's resolve writes
    the fused frames straight into the output rank's memory with 16-byte peer stores over NVLink while it computes;
    `publish()` only has to tell the output rank that the launch has finished. One process per GPU, all on one node
    (NVLink / NVSwitch peers). torch.distributed carries the 64-byte handle and the completion barrier, no data."""

    def __init__(self, n_framesets: int, frame_shape, dst: int = 0, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import capi

        self.dist, self.group, self.dst = dist, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.sizes = shard_sizes(n_framesets, self.world)
        self.frame_shape = tuple(frame_shape)
        self.frame_bytes = 1
        for d in self.frame_shape:
            self.frame_bytes *= int(d)
        self.n_framesets = n_framesets
        L = capi.load()
        self._L, self._owned, self._mapped = L, C.c_void_p(), C.c_void_p()
        handle = [None]
        if self.rank == dst:
            capi.check(L.dfb_device_alloc(C.byref(self._owned), max(1, n_framesets * self.frame_bytes)))
            raw = C.create_string_buffer(64)
            capi.check(L.dfb_ipc_export(self._owned, raw))
            handle = [raw.raw]
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank == dst:
            self.base = self._owned.value
        else:
            capi.check(L.dfb_ipc_open(handle[0], C.byref(self._mapped)))
            self.base = self._mapped.value

    def slice_ptr(self, rank=None) -> int:
        """device address (valid in THIS process) of the first frame `rank` owns"""
        r = self.rank if rank is None else rank
        lo, _ = shard_range(self.n_framesets, self.world, r)
        return self.base + lo * self.frame_bytes

    def publish(self):
        """every rank: call after the fuse launch(es) that wrote slice_ptr() have been synchronised; returns when all
        ranks have done so (the frames are then complete in the output rank's buffer)"""
        self.dist.barrier(group=self.group)

    def frames(self):
        """output rank: the gathered frames as a torch uint8 tensor view `[n_framesets, *frame_shape]`"""
        import torch

        if self.rank != self.dst:
            return None

        class _Ext:  # __cuda_array_interface__ view of the buffer we own
            pass

        ext = _Ext()
        ext.__cuda_array_interface__ = {"shape": (self.n_framesets,) + self.frame_shape, "typestr": "|u1",
                                        "data": (self.base, False), "version": 3}
        return torch.as_tensor(ext, device="cuda")

    def close(self):
        if self._mapped:
            self._L.dfb_ipc_close(self._mapped)
            self._mapped.value = None
        self.dist.barrier(group=self.group)  # nobody maps the buffer any more
        if self._owned:
            self._L.dfb_device_free(self._owned)
            self._owned.value = None
