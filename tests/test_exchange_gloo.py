"""world_size-2 CPU (gloo) test of the pane-exchange DRIVER (denormalized_b200/exchange.py): watermark all-reduce, split-size
exchange, the two payload all-to-alls, pane-range agreement, owner merge, emission of owned keys only.  The CUDA operator
cannot run here, so a numpy stand-in with the same methods and the same 64 B packet format (PartialEntry) plays the
per-rank operator; the union of the ranks' rows must equal ONE oracle operator over the whole stream."""
import os
import socket
import struct
import sys
import zlib

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T0 = 1_700_000_000_000
SIGN = 1 << 63
ORD_MAX = 0x7FEFFFFFFFFFFFFF | SIGN
ORD_MIN = (~ORD_MAX) & (2 ** 64 - 1)


def _ord(v):
    b = struct.unpack("<Q", struct.pack("<d", v))[0]
    return (~b) & (2 ** 64 - 1) if b & SIGN else b | SIGN


def _unord(o):
    b = o & ~SIGN if o & SIGN else (~o) & (2 ** 64 - 1)
    return struct.unpack("<d", struct.pack("<Q", b))[0]


class PaneWindowStandIn:
    """Pane states in Python dicts; packets in numpy; the method surface of GpuStreamingWindow that exchange_step uses."""

    def __init__(self, L, rank, world):
        from denormalized_b200.exchange import PARTIAL_DTYPE
        self.L, self.rank, self.world, self.dt = L, rank, world, PARTIAL_DTYPE
        self.panes, self.lwm, self.exported, self.emitted, self.pending, self.out = {}, None, None, None, [], []

    def owner(self, key):
        return zlib.crc32(key) % self.world

    def push(self, rows):
        self.pending.append(rows)

    def _add(self, pane, key, cnt, s, mn, mx):
        st = self.panes.setdefault(pane, {}).setdefault(key, [0, 0.0, 0, 0])
        st[0] += cnt; st[1] += s; st[2] = max(st[2], mn); st[3] = max(st[3], mx)

    def process(self):
        for rows in self.pending:
            for ts, v, key in rows:
                self._add(ts // self.L, key, 1, v, ORD_MAX - _ord(v), _ord(v) - ORD_MIN)
            mn = min(r[0] for r in rows)
            self.lwm = mn if self.lwm is None else max(self.lwm, mn)
        self.pending = []
        return self.lwm

    def export_partials(self, gw):
        res = dict(entries=(0, 0), keys=(0, 0), owner_counts=[0] * self.world, owner_key_bytes=[0] * self.world, pane_lo=0, pane_hi=-1)
        if gw is None:
            return res
        hi = gw // self.L - 1
        lo = (min(self.panes) if self.panes else hi + 1) if self.exported is None else self.exported + 1
        if hi < lo:
            return res
        self.exported = hi
        per_owner = [[] for _ in range(self.world)]
        for p in sorted(self.panes):
            if lo <= p <= hi:
                for key, st in self.panes[p].items():
                    if self.owner(key) != self.rank:
                        per_owner[self.owner(key)].append((p, key, st))
        n = sum(len(x) for x in per_owner)
        self._ent = np.zeros(n, self.dt); kb = bytearray(); i = 0
        for o in range(self.world):
            seg0 = len(kb)
            for p, key, st in per_owner[o]:
                self._ent[i] = (p, st[0], st[1], st[2], st[3], 0, 2 ** 64 - 1, len(kb) - seg0, len(key)); i += 1
                kb += key + b"\0" * ((-len(key)) % 8)
            res["owner_counts"][o] = len(per_owner[o]); res["owner_key_bytes"][o] = len(kb) - seg0
        self._kb = np.frombuffer(bytes(kb) + b"\0", np.uint8).copy()
        res.update(entries=(self._ent.ctypes.data, n), keys=(self._kb.ctypes.data, len(kb)), pane_lo=lo, pane_hi=hi)
        return res

    def import_partials(self, eptr, src_counts, kptr, src_kbytes, lo, hi):
        import ctypes
        n = sum(src_counts)
        ent = np.frombuffer((ctypes.c_uint8 * (n * 64)).from_address(eptr), self.dt)
        kb = bytes((ctypes.c_uint8 * max(sum(src_kbytes), 1)).from_address(kptr)) if kptr else b""
        i, kbase = 0, 0
        for src in range(self.world):
            for _ in range(src_counts[src]):
                e = ent[i]; i += 1
                key = kb[kbase + int(e["key_off"]):kbase + int(e["key_off"]) + int(e["key_len"])]
                assert lo <= int(e["pane"]) <= hi and self.owner(key) == self.rank
                self._add(int(e["pane"]), key, int(e["cnt"]), float(e["sum"]), int(e["minkey"]), int(e["maxkey"]))
            kbase += src_kbytes[src]

    def flush(self, gw):
        for p in sorted(self.panes):
            if (p + 1) * self.L <= gw and (self.emitted is None or p > self.emitted):
                for key, st in self.panes[p].items():
                    if self.owner(key) == self.rank:
                        self.out.append((p * self.L, (p + 1) * self.L, key, st[0], _unord(ORD_MAX - st[2]), _unord(st[3] + ORD_MIN), st[1] / st[0], 0))
        self.emitted = max(self.emitted or -1, gw // self.L - 1)

    def poll(self):
        o, self.out = self.out, []
        return o


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _stream(world):
    rng = np.random.default_rng(11)
    batches, t = [], T0
    for b in range(24):
        rows = [(t + int(rng.integers(0, 300)), float(rng.random() * 115), b"sensor_%d" % int(rng.integers(0, 40))) for _ in range(200)]
        batches.append(rows); t += 300
    return batches, ((t // 1000) + 3) * 1000


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from denormalized_b200.exchange import TorchTransport, exchange_step
    tr = TorchTransport(device="cpu")
    w = PaneWindowStandIn(1000, rank, world)
    batches, close = _stream(world)
    rows = []
    assert exchange_step(w, tr) == []                       # no watermark anywhere yet: nothing moves, nothing is emitted
    for i, b in enumerate(batches):
        if i % world == rank:
            w.push(b)
        if i % 6 == 5:
            rows += exchange_step(w, tr)
    w.push([(close, 1.0, b"end_%d" % rank)])
    rows += exchange_step(w, tr)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_driver_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from tests.helpers import assert_rows_equal, rows_to_batch, run_oracle_batches
    batches, close = _stream(world)
    want = run_oracle_batches([rows_to_batch(b) for b in batches] + [rows_to_batch([(close, 1.0, b"end")])], 1000)
    got = [tuple(r) for part in gathered for r in part]
    assert len(want) > 200 and all(len(part) > 0 for part in gathered)
    assert_rows_equal(got, want)


def test_packet_layout_matches_the_c_struct():
    from denormalized_b200.exchange import PARTIAL_BYTES, PARTIAL_DTYPE, plan_splits
    assert PARTIAL_DTYPE.itemsize == PARTIAL_BYTES == 64
    assert [PARTIAL_DTYPE.fields[n][1] for n in ("pane", "cnt", "sum", "minkey", "maxkey", "nullrows", "fz", "key_off", "key_len")] == [0, 8, 16, 24, 32, 40, 48, 56, 60]
    assert plan_splits([3, 0, 2], [24, 0, 16]) == ([192, 0, 128], [24, 0, 16])
    hdr = open(os.path.join(ROOT, "include", "dnz_gpu.h")).read()
    assert "#define DNZ_PARTIAL_BYTES 64" in hdr
