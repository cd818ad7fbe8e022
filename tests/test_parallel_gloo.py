"""CPU: the N > 1 path (query sharding + final gather of placements) with world_size 2 over gloo.

The per-rank worker is the oracle here (test infrastructure); what is under test is the sharding and the gather:
the 2-rank result must equal the 1-rank result in input order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util
from apples_b200 import parallel


def test_shard_bounds():
    assert parallel.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert parallel.shard_bounds(8, 8) == [(i, i + 1) for i in range(8)]
    assert parallel.shard_bounds(3, 8)[3:] == [(3, 3)] * 5
    assert parallel.shard_bounds(1000000, 8)[0] == (0, 125000)
    assert parallel.shard_bounds(0, 2) == [(0, 0), (0, 0)]


def _oracle_block(workdir, b, e):
    from oracle import apples_oracle as orc
    ci = util.CaseInputs('c2_matrix_FM_MLSE', workdir)
    ctx = ci.oracle_context()
    edge, err, dis, pen, st = [], [], [], [], []
    for q in ci.queries[b:e]:
        r, status = ctx.runquery(q[0], q[1], dict(q[2]))
        p = r['placements'][0]['p'][0]
        edge.append(p[0]); err.append(p[1]); dis.append(p[3]); pen.append(p[4]); st.append(status)
    return (np.array(edge, np.int32), np.array(err, np.float64), np.array(dis, np.float64), np.array(pen, np.float64),
            np.array(st, np.int32))


def _worker(rank, world, port, workdir, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n = 10
        res = parallel.place_sharded(lambda b, e: _oracle_block(workdir, b, e), n)
        q.put((rank, [a.tolist() for a in res]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [2, 3])
def test_two_ranks_equal_one_rank(world, workdir):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, workdir, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    single = [a.tolist() for a in _oracle_block(workdir, 0, 10)]
    for rank, res in got:
        assert res == single, rank
