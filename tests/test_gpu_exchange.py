"""GPU: the multi-GPU pane exchange on ONE device -- `world` operators in exchange mode, batches dealt round-robin (so
every key's rows are spread over all ranks), packets moved by LocalTransport.  The union of what the ranks emit must equal
ONE operator (the oracle) over the whole stream: count/min/max bit-exact, avg within 1e-9 (north_star)."""
import numpy as np
import pytest

from tests.helpers import assert_rows_equal, record_batch_rows, rows_to_batch, run_oracle_batches, to_record_batch

pytestmark = pytest.mark.gpu
T0 = 1_700_000_000_000


def _stream(rng, nb, n, n_keys, span_ms, long_keys):
    out, t = [], T0
    for _ in range(nb):
        ks = rng.integers(0, n_keys, n)
        rows = []
        for i in range(n):
            k = int(ks[i])
            key = (b"uuid-%032d" % k) if (long_keys and k % 3 == 0) else (b"sensor_%d" % k)
            val = float(rng.random() * 115.0) if k % 11 else None          # some groups only ever see NULL readings
            rows.append((t + int(rng.integers(0, span_ms)), val, key if k != 5 else None))
        out.append(rows_to_batch(rows))
        t += span_ms
    return out, t


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("L,S,filt", [(1000, 0, None), (4000, 1000, ("max", ">", 100))])
def test_exchange_equals_one_operator(world, L, S, filt):
    from denormalized_b200.exchange import LocalTransport
    from tests.helpers import gpu_window
    rng = np.random.default_rng(7 * world + L + S)
    batches, t_end = _stream(rng, 45, 1500, 700, 200, long_keys=True)
    close = (t_end // 1000 + 1) * 1000 + 2 * L
    want = run_oracle_batches(batches + [rows_to_batch([(close, 1.0, b"sentinel")])], L, S, filt)
    wins = [gpu_window(L, S, filt, expected_groups=64) for _ in range(world)]          # small tables: merges must grow them
    for r, w in enumerate(wins):
        w.set_exchange(r, world)
    lt = LocalTransport(wins)
    got, steps_with_rows = [], 0
    for i, b in enumerate(batches):
        wins[i % world].push(to_record_batch(b))
        if i % (3 * world) == 3 * world - 1:
            res = lt.exchange_all()
            steps_with_rows += any(rb.num_rows for rb in res)
            for rb in res:
                got += record_batch_rows(rb)
    for w in wins:                                     # every rank sees the end-of-stream marker
        w.push(to_record_batch(rows_to_batch([(close, 1.0, b"sentinel")])))
    for rb in lt.exchange_all():
        got += record_batch_rows(rb)
    st = [w.stats() for w in wins]
    for w in wins:
        w.close()
    # the sentinel key was pushed on every rank: its owner merged `world` rows; the oracle saw one.  Windows still open at the
    # end of the stream are never emitted (grouped_window_agg_stream.rs:343-348), so the sentinel's window does not appear.
    assert steps_with_rows >= 3 and len(want) > 500
    assert sum(s["exchanged_out"] for s in st) == sum(s["exchanged_in"] for s in st) > 1000
    assert_rows_equal(got, want)
    # ownership: a (window, key) row comes from exactly one rank
    assert len({(r[0], r[2]) for r in got}) == len(got)


def test_exchange_rejects_late_batches_for_exchanged_panes():
    from denormalized_b200 import DnzError
    from denormalized_b200.exchange import LocalTransport
    from tests.helpers import gpu_window
    wins = [gpu_window(1000) for _ in range(2)]
    for r, w in enumerate(wins):
        w.set_exchange(r, 2)
    lt = LocalTransport(wins)
    for w in wins:
        w.push(to_record_batch(rows_to_batch([(T0 + 10, 1.0, b"a"), (T0 + 20, 2.0, b"b")])))
        w.push(to_record_batch(rows_to_batch([(T0 + 3500, 3.0, b"a")])))
    rows = [r for rb in lt.exchange_all() for r in record_batch_rows(rb)]
    assert sorted((r[2], r[3]) for r in rows) == [(b"a", 2), (b"b", 2)]
    wins[0].push(to_record_batch(rows_to_batch([(T0 + 30, 9.0, b"a")])))        # pane 0 was exchanged and emitted
    with pytest.raises(DnzError):
        wins[0].process()
    for w in wins:
        w.close()


@pytest.mark.timeout(600, method="thread")
def test_exchange_over_nccl():
    """Two ranks on two GPUs: the host-driven exchange over NCCL, then the fused exchange over CUDA IPC / P2P (skipped on
    single-GPU boxes)."""
    import os
    import socket
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(root, "tests", "mgpu_exchange_worker.py")],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and "NCCL exchange ok" in out.stdout and "fused P2P exchange ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.timeout(120, method="thread")
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("L,S,filt", [(1000, 0, None), (4000, 1000, ("max", ">", 100))])
def test_fused_group_exchange_equals_one_operator(world, L, S, filt):
    """The library-owned communicator (dnz_group, all ranks in this process on one GPU): packets are written straight into the
    owners' receive rings by the pack kernel (remote-atomic reservation + P2P stores + flags), merged by the owner, emitted by the
    owner -- ONE collective call per step, no host-driven transfer."""
    from denormalized_b200 import ExchangeGroup
    from tests.helpers import gpu_window
    rng = np.random.default_rng(11 * world + L + S)
    batches, t_end = _stream(rng, 48, 1500, 700, 200, long_keys=True)
    close = (t_end // 1000 + 1) * 1000 + 2 * L
    want = run_oracle_batches(batches + [rows_to_batch([(close, 1.0, b"sentinel")])], L, S, filt)
    groups = ExchangeGroup.create_local([0] * world, ring_entries=1 << 16, ring_key_bytes=4 << 20)
    wins = [gpu_window(L, S, filt, expected_groups=2048) for _ in range(world)]
    for g, w in zip(groups, wins):
        g.attach(w)
    got = []

    def step(force=False):
        for g, w in zip(groups, wins):
            if force:
                w.process()                              # the step covers everything that was pushed
            g.step_begin(w)
        for g, w in zip(groups, wins):
            g.step_pack(w)
        gw = [g.step_finish(w) for g, w in zip(groups, wins)]
        assert len(set(gw)) == 1                       # every rank computed the same global watermark
        n = 0
        for w in wins:
            rb = w.poll()
            n += rb.num_rows
            got.extend(record_batch_rows(rb))
        return n
    steps_with_rows = 0
    for i, b in enumerate(batches):
        wins[i % world].push(to_record_batch(b))
        if i % (3 * world) == 3 * world - 1:
            steps_with_rows += step() > 0
    for w in wins:                                     # every rank sees the end-of-stream marker
        w.push(to_record_batch(rows_to_batch([(close, 1.0, b"sentinel")])))
    for _ in range(3):                                 # the protocol is pipelined over three steps: publish | pack | merge + emit
        step(force=True)
    step()                                             # one more (empty) step makes the packet counters visible
    st = [w.stats() for w in wins]
    for w in wins:
        w.close()
    for g in groups:
        g.close()
    assert steps_with_rows >= 1 and len(want) > 500
    assert sum(s["exchanged_out"] for s in st) == sum(s["exchanged_in"] for s in st) > 1000
    assert_rows_equal(got, want)
    assert len({(r[0], r[2]) for r in got}) == len(got)
