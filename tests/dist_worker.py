"""Worker of tests/test_sharding_cpu.py: one rank of a world_size-N gloo job. Each rank produces the fused
frames of ITS shard of a small synthetic stream (the CPU oracle stands in for the GPU fuser — this test is
about the host-side sharding/gather logic, not the kernels), the frames are gathered to rank 0 with
dfb_b200.sharding.FrameGather, and rank 0 checks that the result equals the whole stream fused in one process."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(REPO, "pointcloud-depthfusion_b200"), os.path.join(REPO, "oracle")]


def main():
    import dfb_b200
    import pyoracle
    from dfb_b200.sharding import FrameGather, shard_range

    n_framesets = int(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = dfb_b200.make_config(96, 64, hole_fraction=0.05)
    lo, hi = shard_range(n_framesets, world, rank)

    def fuse(first, count):
        frames = np.zeros((count, cfg.out_height, cfg.out_width, 3), np.uint8)
        for i in range(count):
            ds, cs = pyoracle.synth_frameset(cfg, first + i)
            frames[i] = pyoracle.fuse_frameset(cfg.desc, ds, cs, mode=1)["image"]
        return frames

    gather = FrameGather(n_framesets, (cfg.out_height, cfg.out_width, 3), dst=0)
    for rep in range(2):  # twice: the receive buffers are reused
        local = torch.from_numpy(fuse(lo, hi - lo))
        parts = gather.start(local).finish()
        if rank == 0:
            got = torch.cat(parts).numpy()
            assert got.shape[0] == n_framesets
            assert np.array_equal(got, fuse(0, n_framesets)), "gathered stream differs from the single-process result"
        else:
            assert parts is None
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: shard [{lo},{hi}) ok")


if __name__ == "__main__":
    main()
