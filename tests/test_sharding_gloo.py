"""Host-side logic of the N > 1 path, on CPU with the gloo backend and world_size 2: the range split tiles the
batch, per-range input synthesis equals the global synthesis, and the reductions bench.py relies on
(max-over-ranks timing, AND of the correctness flags, gathered shard sizes) behave."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (os.path.join(ROOT, "zk-nullifier-sig_b200"), os.path.join(ROOT, "oracle"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import plume_b200
    import bench
    import c_oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = plume_b200.shard_range(n_total, rank, world)
    msgs, sk, r = bench.synth_inputs(7, first, last - first)
    # this rank's share of the work, computed by the checker (the CUDA path needs a GPU; here only the
    # plumbing around it is under test)
    o = c_oracle.sign_batch(2, msgs, sk, r)
    ok = c_oracle.verify_batch(2, msgs, o["pk"], o["nullifier"], o["c"], o["s"])
    mx = plume_b200.reduce_max([float(rank + 1), 10.0 - rank], dist)
    assert mx == [float(world), 10.0]
    assert plume_b200.all_ranks_true(bool(ok.all()), dist)
    assert plume_b200.all_ranks_true(rank != 1, dist) is False
    counts = plume_b200.gather_counts(last - first, dist)
    assert sum(counts) == n_total and len(counts) == world
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), msgs=msgs, sk=sk, r=r, c=o["c"], first=first, last=last)
    dist.barrier()
    dist.destroy_process_group()


def test_range_split_unit():
    import plume_b200
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 8):
            edges = [plume_b200.shard_range(n, g, world) for g in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        plume_b200.shard_range(10, 2, 2)


def test_library_range_split_matches():
    """The split the multi-device context applies inside libplume_b200.so (plume_shard_range, no GPU needed) is the same
    rule: it tiles the batch in order for any device count, including batches smaller than the device count."""
    import ctypes
    import plume_b200
    lib = plume_b200.load()
    for n in (0, 1, 5, 7, 8, 1000, (1 << 20) + 1, (1 << 24) + 3, (1 << 63) + 12345):
        for parts in (1, 2, 3, 5, 8, 16):
            pos = 0
            for g in range(parts):
                f, c = ctypes.c_size_t(0), ctypes.c_size_t(0)
                assert lib.plume_shard_range(n, g, parts, ctypes.byref(f), ctypes.byref(c)) == 0
                assert f.value == pos
                if n < (1 << 60):
                    assert (f.value, f.value + c.value) == plume_b200.shard_range(n, g, parts)
                pos += c.value
            assert pos == n
    f, c = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert lib.plume_shard_range(10, 2, 2, ctypes.byref(f), ctypes.byref(c)) == -1


def test_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    import bench
    import c_oracle
    n_total, world = 37, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % g)) for g in range(world)]
    assert parts[0]["first"] == 0 and parts[0]["last"] == parts[1]["first"] and parts[1]["last"] == n_total
    msgs, sk, r = bench.synth_inputs(7, 0, n_total)
    for key, whole in (("msgs", msgs), ("sk", sk), ("r", r)):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), whole)
    whole_c = c_oracle.sign_batch(2, msgs, sk, r)["c"]
    assert np.array_equal(np.concatenate([p["c"] for p in parts]), whole_c)
