"""world_size-2 CPU (gloo) test of the multi-GPU host logic: ranks own disjoint key sets (key id * world + rank), every
rank runs an independent operator (here: the oracle stands in for the per-rank operator), the union of the per-rank
outputs equals the output of one operator over the union of the inputs, and the max-over-ranks timing reduction works.
This mirrors RepartitionExec(Hash(group keys), n) -> n independent GroupedWindowAggStreams
(physical_optimizer/coalesce_before_streaming_window_aggregate.rs:63-73, streaming_window.rs:470-481)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import OracleWindow, synth_batch
    from tests.helpers import rows_to_batch, batch_to_rows
    G, n, nb, rpm = 50, 2048, 6, 4
    o = OracleWindow(1000, 0, ("max", ">", 100))
    keys = set()
    for b in range(nb):
        batch = synth_batch(b * n, n, seed=42 + rank, groups=G, rows_per_ms=rpm, key_mul=world, key_add=rank)
        keys |= {r[2] for r in batch_to_rows(batch)}
        o.push(batch)
    o.push(rows_to_batch([(1_700_000_000_000 + nb * n // rpm + 5000, 1.0, b"sensor_%d" % rank)]))
    rows = o.results()
    # key ownership: every key id of this rank is congruent to rank mod world
    assert all(int(k[7:]) % world == rank for k in keys)
    gathered = [None] * world
    dist.all_gather_object(gathered, (sorted(keys), rows))
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # bench.py: step time = max over ranks
    dist.barrier()
    if rank == 0:
        q.put((gathered, float(t.item())))
    dist.destroy_process_group()


def test_key_sharded_operators_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 11.0
    k0, k1 = set(gathered[0][0]), set(gathered[1][0])
    assert k0 and k1 and not (k0 & k1)
    # one operator over the union of both ranks' inputs emits exactly the union of the per-rank outputs
    from oracle import OracleWindow, synth_batch
    from tests.helpers import assert_rows_equal, rows_to_batch
    G, n, nb, rpm = 50, 2048, 6, 4
    o = OracleWindow(1000, 0, ("max", ">", 100))
    for b in range(nb):
        for rank in range(world):
            o.push(synth_batch(b * n, n, seed=42 + rank, groups=G, rows_per_ms=rpm, key_mul=world, key_add=rank))
    for rank in range(world):
        o.push(rows_to_batch([(1_700_000_000_000 + nb * n // rpm + 5000, 1.0, b"sensor_%d" % rank)]))
    want = o.results()
    got = gathered[0][1] + gathered[1][1]
    assert len(want) > 20
    assert_rows_equal(got, want, rel=0.0)
