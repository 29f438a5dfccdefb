"""Host-side multi-GPU logic on CPU: shard assignment and the world_size-2 gloo gather."""
import os
import socket
import subprocess
import sys

import pytest

from dfb_b200.sharding import shard_range, shard_sizes


def test_shard_ranges_cover_the_stream_in_order():
    for n in (0, 1, 2, 5, 512, 513, 1000):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert lo == nxt and hi >= lo
                nxt = hi
            assert nxt == n
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_framesets", [4, 5], ids=["even", "ragged"])
def test_gloo_world2_gather_matches_single_process(repo_root, n_framesets):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(repo_root, "tests", "dist_worker.py"), str(n_framesets)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out}"
        assert f"rank {rank}: shard" in out
