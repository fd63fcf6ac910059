"""The N > 1 path on CPU: two gloo ranks shard the seed-cell units, each produces its shard's count tensor and the
single reduce (sum, int64) on rank 0 must equal the unsharded run.  On the GPU box the shard tensor comes from the
CUDA kernel (bench.py, test_gpu_parity.py::test_shards_sum_to_whole); here the oracle stands in for it so that the
host-side plumbing (cuda_pro_cell_b200.dist) is what is under test."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path, shard_level=0):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from cuda_pro_cell_b200 import dist as pdist
    from cuda_pro_cell_b200 import synth
    w = synth.workload(2, 0.004)
    plan = oracle_lib.OraclePlan(w.values, w.freqs, w.phi)
    shard = pdist.shard_spec(unit=64)
    assert shard == (rank, world, 64)
    part = oracle_lib.simulate(plan, w.types, w.t_max, w.seed, shard=shard, n_threads=2, shard_level=shard_level)
    n_sets, n_keys, n_types = part["counts"].shape
    buf = pdist.packed_buffer(n_sets, n_keys, n_types, "cpu")
    counts, div = pdist.unpack(buf, n_sets, n_keys, n_types)
    counts.copy_(torch.from_numpy(part["counts"]))
    div.copy_(torch.from_numpy(part["divisions"]))
    pdist.reduce_packed(buf, dst=0)
    if rank == 0:
        whole = oracle_lib.simulate(plan, w.types, w.t_max, w.seed, n_threads=2)
        ok = np.array_equal(counts.numpy(), whole["counts"]) and np.array_equal(div.numpy(), whole["divisions"])
        Path(out_path).write_text("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_unsharded(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_two_rank_reduce_equals_unsharded_with_subtree_sharding(tmp_path):
    """the same exchange step when the ranks share the first tree levels and own subtrees at level 3"""
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out), 3), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_subtree_owner_functions():
    sys.path.insert(0, str(ROOT))
    from cuda_pro_cell_b200 import dist as pdist
    for world in (2, 3, 8):
        kids = [pdist.owner_of_subtree(r, h, world) for r in (0, 5, 4096) for h in range(64, 128)]
        assert set(kids) == set(range(world))                      # the 64 subtrees of a lineage spread over all ranks
        assert pdist.owner_of_subtree(7, 64, world) == (7 + 64) % world
        assert pdist.owner_of_subtree(0xFFFFFFFF, 2, world) == 1 % world      # 32-bit wrap, as on the device
        assert pdist.owner_of_shared_node(13, world) == 13 % world


def test_owner_function_matches_the_shard_rule():
    sys.path.insert(0, str(ROOT))
    from cuda_pro_cell_b200 import dist as pdist
    for world in (1, 2, 8):
        owners = [pdist.owner_of_seed(r, world, 256) for r in range(0, 5000, 97)]
        assert all(0 <= o < world for o in owners)
        assert pdist.owner_of_seed(255, world, 256) == 0 and pdist.owner_of_seed(256, world, 256) == 1 % world
