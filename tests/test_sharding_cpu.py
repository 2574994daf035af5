"""CPU: the N>1 host logic.  Shard arithmetic, and a world_size-2 gloo run of decode_sharded in which each rank
decodes its shard with the CPU oracle (test-only stand-in for the per-rank GPU decoder) and rank 0 gathers."""
import os
import socket
import sys

import numpy as np
import pytest

from ldpc_b200 import codes
from ldpc_b200.parallel import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for total in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from ldpc_b200.parallel import decode_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = codes.regular_ldpc(120, 3, 6, seed=5)
    syn = codes.bsc_syndromes(H, 0.05, 101, seed=3)  # odd size: ragged shards
    port_oracle = oracle.PortOracle()
    kw = dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.625)

    def decode_fn(shard):
        d, c, i, _ = port_oracle.decode_batch(H, shard, 0.05, want_llr=False, **kw)
        return d, c, i

    dec, conv, its = decode_sharded(decode_fn, syn, gather_to=0)
    if rank == 0:
        want = port_oracle.decode_batch(H, syn, 0.05, want_llr=False, **kw)
        ok = np.array_equal(dec, want[0]) and np.array_equal(conv, want[1]) and np.array_equal(its, want[2])
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
