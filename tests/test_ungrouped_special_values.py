"""Ungrouped windows, special Float64 values: totalOrder min / max over NaN, +-0.0, +-inf (the row accumulators of the ungrouped
path, streaming_window.rs:640-828 -> DataFusion MinAccumulator / MaxAccumulator), a window that only sees zeros and windows that
only see NULL values (count 0, min / max / average NULL).  Compared with the oracle's restatement, batch by batch."""
import pytest

from tests.test_gpu_ungrouped import T0, run_both
from tests.helpers import assert_rows_equal, rows_to_batch

pytestmark = pytest.mark.gpu


def test_ungrouped_total_order_min_max_and_special_values():
    vals = [float("nan"), -0.0, 0.0, float("inf"), float("-inf"), 5.0, -3.0, None]
    batches = []
    for b in range(16):
        rows = [(T0 + b * 500 + i, vals[(b + i) % len(vals)], b"x") for i in range(40)]
        if b == 4:
            rows = [(T0 + b * 500 + i, -0.0 if i % 2 else 0.0, b"x") for i in range(40)]      # a window of zeros only: min -0.0, max +0.0
        if b in (6, 7):
            rows = [(T0 + b * 500 + i, None, b"x") for i in range(10)]                         # a window that only sees NULL values
        batches.append(rows_to_batch(rows))
    got, want, _ = run_both(batches, 1000, 0)
    assert any(r[4] is None for r in want) and len(want) >= 3
    assert_rows_equal(got, want, check_seq=True)
