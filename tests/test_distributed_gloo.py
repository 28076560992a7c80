"""world_size-2 gloo test of the multi-GPU host logic: tree sharding + merge of the root arrays.

Each rank computes the root arrays of ITS tree shard with the oracle (standing in, as the checker,
for the device call the ranks make on the GPU box), merges them with the package's all-reduce
helper and must obtain exactly the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from ismcsolver_b200 import distributed as D


def test_shard_trees_partitions_the_range():
    for total in (2, 7, 64, 2368):
        for world in (1, 2, 4, 8):
            if total < world:
                with pytest.raises(ValueError):
                    D.shard_trees(total, world, 0)
                continue
            spans = [D.shard_trees(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (a, n), (b, _) in zip(spans, spans[1:]):
                assert a + n == b
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_best_moves_first_maximum():
    v = torch.tensor([[0, 5, 5, 1], [0, 0, 0, 0], [3, 2, 9, 9]])
    assert D.best_moves(v).tolist() == [1, -1, 2]


def _worker(rank, world, port, trees_total, iterations, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, trees = D.shard_trees(trees_total, world, rank)
        rows = []
        for pos in range(3):
            s = O.kw_deal(5, pos, 4)
            v, s2, av, _ = O.root_parallel(O.GAME_KW, s, iterations, trees, seed=9, pos=pos, tree_base=base,
                                           trees_total=trees_total)
            rows.append(np.stack([v, s2, av]).astype(np.int64))
        t = torch.from_numpy(np.stack(rows))
        D.merge_root_arrays(t)
        if rank == 0:
            out.put(t.numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_merge_to_the_unsharded_result():
    trees_total, iterations, world = 6, 6 * 150 + 4, 2
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, trees_total, iterations, out)) for r in range(world)]
    for p in procs:
        p.start()
    merged = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for pos in range(3):
        s = O.kw_deal(5, pos, 4)
        v, s2, av, best = O.root_parallel(O.GAME_KW, s, iterations, trees_total, seed=9, pos=pos)
        assert (merged[pos, 0] == v.astype(np.int64)).all()
        assert (merged[pos, 1] == s2.astype(np.int64)).all()
        assert (merged[pos, 2] == av.astype(np.int64)).all()
        assert int(D.best_moves(torch.from_numpy(merged[pos, 0]))) == best
