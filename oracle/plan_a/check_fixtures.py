#!/usr/bin/env python
"""Compare the CPU oracle with a fixture written by the reference itself (oracle/plan_a/emit_fixtures.rs).

usage: check_fixtures.py <cfg1..cfg5> <rows> <fixture.jsonl>
Exit code 0: every row the reference emitted is reproduced by the oracle (window bounds, key, count exact; min/max bit-exact;
average within 1e-9 relative)."""
import json
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

CFG = {"cfg1": (1_000, 1_000, 1_000, 0, None, False), "cfg2": (100_000, 10_000, 1_000, 0, None, False),
       "cfg3": (1_000_000, 10_000, 10_000, 1_000, None, False), "cfg4": (100_000, 10_000, 1_000, 0, ("max", ">", 113.0), False),
       "cfg5": (10_000_000, 8_000, 60_000, 5_000, None, True)}


def load_fixture(path):
    rows = {}
    with open(path) as f:
        for line in f:
            r = json.loads(line)
            f64 = lambda h: struct.unpack("<d", struct.pack("<Q", int(h, 16)))[0]
            rows[(r["ws"], r["key"].encode())] = (r["we"], r["count"], f64(r["min"]), f64(r["max"]), f64(r["avg"]))
    return rows


def check(cfg, n_rows, path):
    from oracle import OracleWindow, synth_batch
    from tests.helpers import rows_to_batch
    groups, rpm, win, slide, filt, uuid = CFG[cfg]
    want = load_fixture(path)
    o = OracleWindow(win, slide, filt)
    for r0 in range(0, n_rows, 65536):
        o.push(synth_batch(r0, min(65536, n_rows - r0), groups=groups, rows_per_ms=rpm, uuid_keys=uuid))
    last = 1_700_000_000_000 + (n_rows - 1) // rpm
    o.push(rows_to_batch([((last // 1000 + 1) * 1000 + 2 * win, 1.0, b"sentinel")]))
    got = {(r[0], r[2]): (r[1], r[3], r[4], r[5], r[6]) for r in o.results()}
    bad = 0
    for k, w in want.items():
        g = got.get(k)
        ok = g is not None and g[0] == w[0] and g[1] == w[1] and struct.pack("<d", g[2]) == struct.pack("<d", w[2]) and \
            struct.pack("<d", g[3]) == struct.pack("<d", w[3]) and abs(g[4] - w[4]) <= 1e-9 * max(abs(w[4]), 1e-300)
        if not ok:
            bad += 1
            if bad <= 10:
                print("MISMATCH", k, "reference", w, "oracle", g)
    print(f"{len(want)} reference rows, {bad} mismatches, oracle emitted {len(got)} rows in total")
    return bad == 0


if __name__ == "__main__":
    sys.exit(0 if check(sys.argv[1], int(sys.argv[2]), sys.argv[3]) else 1)
