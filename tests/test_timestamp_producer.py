"""SURVEY.md §8 f1 -- the input-contract producer: canonical event time derived from a raw column, as the reference's Kafka
reader does before the window operator (kafka_stream_read.rs:236-268, utils/time.rs:59-94).
CPU: the oracle restatement (oracle.ts_convert) against Python's datetime.  GPU: RAW batches (no `_streaming_internal_metadata`)
through the operator == the oracle window over the converted timestamps."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pytest

from oracle import Batch, OracleWindow, ts_convert
from tests.helpers import assert_rows_equal, record_batch_rows

T0 = 1_700_000_000_000
UTC = dt.timezone.utc


def iso(ms, fmt):
    d = dt.datetime.fromtimestamp(ms // 1000, UTC).replace(tzinfo=None)
    s = d.strftime(fmt.replace("%.f", "").replace("%.3f", "").replace("%3f", ""))
    frac = ms % 1000
    if fmt.endswith("%.f"):
        s += (".%03d" % frac).rstrip("0").rstrip(".") if frac else ""
    elif fmt.endswith("%.3f"):
        s += ".%03d" % frac
    elif fmt.endswith("%3f"):
        s += "%03d" % frac
    return s


@pytest.mark.parametrize("fmt", ["%Y-%m-%dT%H:%M:%S%.f", "%Y-%m-%d %H:%M:%S%.3f", "%d/%m/%Y %H:%M:%S", "%Y%m%dT%H%M%S%3f"])
def test_oracle_iso8601_against_datetime(fmt):
    rng = np.random.default_rng(1)
    ms = np.concatenate([rng.integers(-2_000_000_000_000, 4_000_000_000_000, 500), [0, -1, 951782400000, 1709251199999]]).astype(np.int64)
    if "%f" not in fmt.replace("%.f", "").replace("%.3f", "").replace("%3f", "") and not any(x in fmt for x in ("%.f", "%.3f", "%3f")):
        ms = ms // 1000 * 1000
    strings = [iso(int(m), fmt) for m in ms]
    out = ts_convert(3, strings=strings, fmt=fmt)
    assert np.array_equal(out, ms), [(s, int(a), int(b)) for s, a, b in zip(strings, out, ms) if a != b][:5]


def test_oracle_units_and_errors():
    assert list(ts_convert(1, values=[5, -7])) == [5, -7]
    assert list(ts_convert(2, values=[1_700_000_000, -2])) == [1_700_000_000_000, -2000]
    for bad in ["2023-02-29T00:00:00", "2023-13-01T00:00:00", "2023-01-01T00:00", "2023-01-01T00:00:00Z", ""]:
        with pytest.raises(ValueError):
            ts_convert(3, strings=[bad], fmt="%Y-%m-%dT%H:%M:%S%.f")
    assert ts_convert(3, strings=["2016-12-31T23:59:60"], fmt="%Y-%m-%dT%H:%M:%S")[0] == 1483228799000 + 1000    # leap second: chrono keeps it as 59 s + 1e9 ns


def _raw_stream(seed, unit):
    rng = np.random.default_rng(seed)
    out, t = [], T0
    for _ in range(20):
        n = 1500
        ts = t + rng.integers(0, 400, n)
        if unit == 2:
            ts = ts // 1000 * 1000
        out.append((ts.astype(np.int64), rng.random(n) * 115.0, [b"sensor_%d" % int(k) for k in rng.integers(0, 60, n)]))
        t += 400 if unit != 2 else 1000
    out.append((np.array([t + 9000 - (t + 9000) % 1000], np.int64), np.array([1.0]), [b"sentinel"]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("unit,fmt", [(1, None), (2, None), (3, "%Y-%m-%dT%H:%M:%S%.f"), (3, "%F %T%.3f")])
def test_gpu_derives_canonical_timestamps_from_raw_batches(unit, fmt):
    from denormalized_b200 import GpuStreamingWindow
    from tests.helpers import DEFAULT_AGGS
    py_fmt = (fmt or "").replace("%F", "%Y-%m-%d").replace("%T", "%H:%M:%S")
    ts_type = pa.utf8() if unit == 3 else pa.int64()
    schema = pa.schema([pa.field("occurred_at", ts_type), pa.field("reading", pa.float64()), pa.field("sensor_name", pa.utf8())])
    w = GpuStreamingWindow(schema, "sensor_name", DEFAULT_AGGS, 2000, 1000, None, expected_groups=128, timestamp=(unit, "occurred_at", fmt))
    o = OracleWindow(2000, 1000)
    got = []
    for i, (ts, val, keys) in enumerate(_raw_stream(3 + unit, unit)):
        if unit == 1:
            raw, canon = pa.array(ts), ts_convert(1, values=ts)
        elif unit == 2:
            raw, canon = pa.array(ts // 1000), ts_convert(2, values=ts // 1000)
        else:
            strings = [iso(int(m), py_fmt) for m in ts]
            raw, canon = pa.array(strings, pa.utf8()), ts_convert(3, strings=strings, fmt=py_fmt)
        w.push(pa.RecordBatch.from_arrays([raw, pa.array(val), pa.array([k.decode() for k in keys], pa.utf8())], schema=schema))
        got += record_batch_rows(w.poll(), i)
        off = np.zeros(len(keys) + 1, np.int32); off[1:] = np.cumsum([len(k) for k in keys])
        o.push(Batch(ts=np.ascontiguousarray(canon, np.int64), val=np.ascontiguousarray(val, np.float64), key_off=off,
                     key_bytes=np.frombuffer(b"".join(keys) + b"\0" * 16, np.uint8).copy()))
    want = o.results()
    w.close()
    assert len(want) > 500
    assert_rows_equal(got, want, check_seq=True)


@pytest.mark.gpu
def test_gpu_rejects_unparsable_timestamp_strings():
    from denormalized_b200 import DnzError, GpuStreamingWindow
    from tests.helpers import DEFAULT_AGGS
    schema = pa.schema([pa.field("occurred_at", pa.utf8()), pa.field("reading", pa.float64()), pa.field("sensor_name", pa.utf8())])
    w = GpuStreamingWindow(schema, "sensor_name", DEFAULT_AGGS, 1000, 0, None, timestamp=(3, "occurred_at", "%Y-%m-%dT%H:%M:%S"))
    w.push(pa.RecordBatch.from_arrays([pa.array(["2023-11-14T22:13:20", "yesterday"]), pa.array([1.0, 2.0]), pa.array(["a", "b"])], schema=schema))
    with pytest.raises(DnzError):
        w.poll()
    with pytest.raises(DnzError):        # chrono specifiers the device parser does not implement are refused at creation
        GpuStreamingWindow(schema, "sensor_name", DEFAULT_AGGS, 1000, 0, None, timestamp=(3, "occurred_at", "%A %B %e"))
    w.close()
