"""Generates tests/golden/*.json from the CPU oracle (oracle/dnz_oracle.c).  The reference holds no golden vectors for this
path and cannot be built here, so these fixtures pin OUR restatement: they freeze the oracle's
answers on small seeded streams so that neither the oracle nor the CUDA path can drift silently.
    python tests/golden/make_golden.py        (re-run only when the semantics are deliberately changed)
Floats are stored as hex strings (bit exact); keys as latin-1 strings; None = SQL NULL."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import OracleWindow  # noqa: E402
from tests.helpers import random_stream, rows_to_batch  # noqa: E402

T0 = 1_700_000_000_000
CASES = {
    "tumbling_1s": dict(L=1000, S=0, filt=None, seed=11, kw=dict(span_ms=300, ragged=True)),
    "tumbling_1s_filter_max_gt_113": dict(L=1000, S=0, filt=["max", ">", 113], seed=12, kw=dict(span_ms=300)),
    "sliding_5s_1s": dict(L=5000, S=1000, filt=None, seed=13, kw=dict(span_ms=400, ragged=True)),
    "sliding_4s_2s_late": dict(L=4000, S=2000, filt=None, seed=14, kw=dict(span_ms=350, jitter_ms=600, late_every=4, late_shift_ms=5000)),
    "nulls_and_specials": dict(L=2000, S=0, filt=None, seed=15, kw=dict(span_ms=300, null_frac=0.12, special_vals=True)),
}


def enc_f(x):
    return None if x is None else float(x).hex()


def main():
    for name, c in CASES.items():
        rng = np.random.default_rng(c["seed"])
        batches = random_stream(rng, 24, 48, 7, **c["kw"])
        batches.append([(T0 + 24 * 400 + 4 * c["L"], 1.0, b"sentinel")])
        o = OracleWindow(c["L"], c["S"], tuple(c["filt"]) if c["filt"] else None)
        out = []
        for i, rows in enumerate(batches):
            o.push(rows_to_batch(rows))
            out += [[r[0], r[1], None if r[2] is None else r[2].decode("latin-1"), r[3], enc_f(r[4]), enc_f(r[5]), enc_f(r[6]), r[7]] for r in o.results()]
        doc = {"window_ms": c["L"], "slide_ms": c["S"], "filter": c["filt"],
               "batches": [[[r[0], enc_f(r[1]), None if r[2] is None else r[2].decode("latin-1")] for r in rows] for rows in batches],
               "expected": out}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, len(out), "rows")


if __name__ == "__main__":
    main()
