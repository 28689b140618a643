"""SURVEY.md §8 f2 -- ungrouped windows `.window([], aggs, len, slide)`: the device reduces rows into panes (WindowAggStream,
Partial), the library's host side replays the per-batch emission schedule and runs the Final stage (FullWindowAggStream) --
against the oracle's restatement of one partition of that chain, batch by batch: totalOrder min/max (row accumulators, not the
grouped ones), emission lag of the Final stage, late partial results dropped."""
import numpy as np
import pytest

from oracle import OracleUngrouped
from tests.helpers import DEFAULT_AGGS, assert_rows_equal, random_stream, rows_to_batch, to_record_batch

pytestmark = pytest.mark.gpu
T0 = 1_700_000_000_000


def rows_of(rb, seq):
    cols = {name: rb.column(i) for i, name in enumerate(rb.schema.names)}
    ws = cols["window_start_time"].cast("int64").to_pylist(); we = cols["window_end_time"].cast("int64").to_pylist()
    return [(ws[i], we[i], None, cols["count"][i].as_py(), cols["min"][i].as_py(), cols["max"][i].as_py(), cols["average"][i].as_py(), seq)
            for i in range(rb.num_rows)]


def run_both(batches, L, S, **kw):
    from denormalized_b200 import GpuStreamingWindow, canonical_schema
    w = GpuStreamingWindow(canonical_schema(), None, DEFAULT_AGGS, L, S, None, **kw)
    o = OracleUngrouped(L, S)
    got, want = [], []
    for i, b in enumerate(batches):
        w.push(to_record_batch(b)); rb = w.poll()
        assert rb.schema.names == ["count", "min", "max", "average", "window_start_time", "window_end_time"]
        got += rows_of(rb, i)
        o.push(b); want += o.results()
    st = w.stats()
    w.close()
    return got, want, st


@pytest.mark.parametrize("L,S", [(1000, 0), (2000, 0), (4000, 1000), (6000, 2000)])
def test_ungrouped_in_order_stream(L, S):
    rng = np.random.default_rng(L + S)
    raw = random_stream(rng, 60, 300, 5, span_ms=300, null_frac=0.1, ragged=True)
    batches = [rows_to_batch(r) for r in raw]
    got, want, st = run_both(batches, L, S)
    assert len(want) > 2 and st["agg_launches"] >= 50
    assert_rows_equal(got, want, check_seq=True)


def test_ungrouped_late_batches_and_queued_mode():
    rng = np.random.default_rng(9)
    raw = random_stream(rng, 50, 200, 3, span_ms=300, jitter_ms=500, late_every=6, late_shift_ms=2500)
    batches = [rows_to_batch(r) for r in raw]
    got, want, st = run_both(batches, 1000, 0)
    assert st["late_batches"] > 0 and len(want) > 5
    assert_rows_equal(got, want, check_seq=True)
    # everything queued, one poll at the end: the same rows
    from denormalized_b200 import GpuStreamingWindow, canonical_schema
    w = GpuStreamingWindow(canonical_schema(), None, DEFAULT_AGGS, 1000, 0, None, max_rows_per_launch=1500)
    for b in batches:
        w.push(to_record_batch(b))
    got2 = rows_of(w.poll(), 0)
    w.close()
    assert_rows_equal(got2, want)
