"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, against the CPU oracle
on identical inputs.  count/min/max bit-exact, avg within 1e-9 relative (north_star); rows compared as a multiset sorted
by (window_start, key), plus the emitting push where polls are per batch."""
import math

import numpy as np
import pytest

from oracle import synth_batch
from tests.helpers import (assert_rows_equal, random_stream, rows_to_batch, run_gpu, run_oracle_batches)

pytestmark = pytest.mark.gpu
T0 = 1_700_000_000_000


def sentinel(ts):
    return rows_to_batch([(ts, 1.0, b"sentinel")])


@pytest.mark.parametrize("L,S", [(1000, 0), (2000, 0), (5000, 1000), (10_000, 1000), (4000, 2000)])
@pytest.mark.parametrize("flags", [0, 2, 4])
def test_in_order_stream_per_batch(L, S, flags):
    rng = np.random.default_rng(L + S)
    batches = [rows_to_batch(r) for r in random_stream(rng, 30, 200, 17, span_ms=350, ragged=True)]
    batches.append(sentinel(T0 + 30 * 350 + 3 * L))
    got, st = run_gpu(batches, L, S, flags=flags)
    want = run_oracle_batches(batches, L, S)
    assert len(want) > 20
    assert_rows_equal(got, want, check_seq=True)
    assert st["late_batches"] == 0


@pytest.mark.parametrize("L,S", [(1000, 0), (3000, 1000), (6000, 2000)])
def test_late_rows_reopen_windows_exactly(L, S):
    rng = np.random.default_rng(99 + L)
    batches = [rows_to_batch(r) for r in random_stream(rng, 40, 60, 9, span_ms=350, jitter_ms=700, ragged=True,
                                                       late_every=5, late_shift_ms=4200)]
    batches.append(sentinel(T0 + 40 * 350 + 3 * L))
    want = run_oracle_batches(batches, L, S)
    seen = {}
    for r in want:                      # the stream really re-opens windows: some (window, key) is emitted more than once
        seen[(r[0], r[2])] = seen.get((r[0], r[2]), 0) + 1
    assert max(seen.values()) > 1
    got, st = run_gpu(batches, L, S)
    assert st["late_batches"] > 0
    assert_rows_equal(got, want, check_seq=True)
    # queued mode (one poll for everything): same multiset of rows
    got2, _ = run_gpu(batches, L, S, per_batch_poll=False)
    assert_rows_equal(got2, want)


@pytest.mark.parametrize("flags", [0, 2])
def test_nulls_and_special_values(flags):
    rng = np.random.default_rng(5)
    raw = random_stream(rng, 25, 80, 6, span_ms=300, null_frac=0.15, special_vals=True)
    raw[0] = raw[0] + [(T0 + 5, None, b"onlynull"), (T0 + 6, None, b"onlynull")]
    batches = [rows_to_batch(r) for r in raw] + [sentinel(T0 + 25 * 300 + 9000)]
    for L, S in [(1000, 0), (3000, 1000)]:
        want = run_oracle_batches(batches, L, S)
        got, _ = run_gpu(batches, L, S, flags=flags)
        assert any(r[2] is None for r in want) and any(r[4] is None for r in want)
        assert_rows_equal(got, want, check_seq=True)


def test_min_max_quirks_and_filter_total_order():
    rows = [(T0, float("nan"), b"nan"), (T0, -0.0, b"z1"), (T0 + 1, 0.0, b"z1"), (T0, 0.0, b"z2"), (T0 + 1, -0.0, b"z2"),
            (T0, float("inf"), b"pinf"), (T0, float("-inf"), b"ninf"), (T0, 5.0, b"mix"), (T0, float("nan"), b"mix"),
            (T0, 114.0, b"hi"), (T0, 112.0, b"lo"), (T0, None, b"null"), (T0, 113.0, b"eq")]
    batches = [rows_to_batch(rows), sentinel(T0 + 3000)]
    for filt in [None, ("max", ">", 113), ("average", ">", 113), ("count", ">=", 1), ("min", "<=", 0.0), ("max", "!=", 113)]:
        want = run_oracle_batches(batches, 1000, 0, filt)
        got, _ = run_gpu(batches, 1000, 0, filt)
        assert_rows_equal(got, want, check_seq=True)
    got = {r[2]: r for r in run_gpu(batches, 1000)[0]}
    assert math.copysign(1, got[b"z1"][4]) == -1 and math.copysign(1, got[b"z2"][4]) == 1


@pytest.mark.parametrize("L,S", [(1000, 0), (4000, 1000)])
@pytest.mark.parametrize("late", [False, True])
@pytest.mark.parametrize("expected_groups", [0, 16])      # 16: small tables => per-CTA private pane copies (low-cardinality path)
def test_minmax_hints_never_lose_an_extreme(L, S, late, expected_groups):
    """Few keys x many rows per pane, fast (TMA-staged) tiles, one launch per batch: the per-group min/max reduction filter
    (DictSlot::hint) is hit hard -- trends, sign changes, tiny/huge magnitudes, values equal in their top 16 key bits,
    panes revisited by late batches (new pane instances => new hint tags)."""
    rng = np.random.default_rng(1234 + L + S + late)
    batches, t = [], T0
    for b in range(36):
        n = 2500
        k = rng.integers(0, 12, n)
        mode = b % 6
        if mode == 0:
            v = rng.random(n) * 115.0
        elif mode == 1:
            v = 100.0 + np.arange(n) * 1e-9 + rng.random(n) * 1e-7          # same top-16 bucket, creeping upwards
        elif mode == 2:
            v = -100.0 - np.arange(n) * 1e-9 - rng.random(n) * 1e-7         # creeping downwards
        elif mode == 3:
            v = (rng.random(n) - 0.5) * 10.0 ** rng.integers(-300, 300, n).astype(np.float64)
        elif mode == 4:
            v = np.where(rng.random(n) < 0.5, 113.99999, 114.00001) * np.where(rng.random(n) < 0.1, -1.0, 1.0)
        else:
            v = rng.standard_normal(n) * 1e-3
        back = 2600 if (late and b % 7 == 6) else 0
        ts = t - back + rng.integers(0, 300, n)
        rows = [(int(ts[i]), float(v[i]), b"g%d" % k[i]) for i in range(n)]
        batches.append(rows_to_batch(rows))
        t += 300
    batches.append(sentinel(t + 3 * L))
    want = run_oracle_batches(batches, L, S)
    got, st = run_gpu(batches, L, S, expected_groups=expected_groups)
    assert st["fast_tiles"] > 30 and (st["late_batches"] > 0) == late
    assert_rows_equal(got, want, check_seq=True)
    got2, _ = run_gpu(batches, L, S, per_batch_poll=False, expected_groups=expected_groups)
    assert_rows_equal(got2, want)


def test_long_and_empty_keys_and_dictionary_growth():
    rng = np.random.default_rng(3)
    rows_all = []
    for b in range(12):
        rows = []
        for i in range(3000):
            k = int(rng.integers(0, 5000))
            key = (b"%05d" % k) * (1 + k % 11)            # lengths 5..55: inline and arena keys
            if k == 7:
                key = b""
            rows.append((T0 + b * 400 + int(rng.integers(0, 400)), float(rng.random()), key))
        rows_all.append(rows)
    batches = [rows_to_batch(r) for r in rows_all] + [sentinel(T0 + 20_000)]
    want = run_oracle_batches(batches, 2000, 1000)
    got, st = run_gpu(batches, 2000, 1000, expected_groups=64, per_batch_poll=False)   # forces table growth + deferred rows
    assert st["deferred_rows"] > 0
    assert_rows_equal(got, want)


def test_empty_batches_and_errors():
    from denormalized_b200 import DnzError
    from tests.helpers import gpu_window, to_record_batch, record_batch_rows
    w = gpu_window(1000)
    w.push(to_record_batch(rows_to_batch([(T0 + 10, 1.0, b"a")])))
    w.push(to_record_batch(rows_to_batch([])))
    assert w.poll().num_rows == 0 and w.watermark == T0 + 10
    w.push(to_record_batch(rows_to_batch([(T0 + 1500, 2.0, b"a")])))
    rb = w.poll()
    assert record_batch_rows(rb) == [(T0, T0 + 1000, b"a", 1, 1.0, 1.0, 1.0, 0)]
    assert rb.schema.names == ["sensor_name", "count", "min", "max", "average", "window_start_time", "window_end_time"]
    with pytest.raises(DnzError):     # the reference panics on an all-null timestamp column
        w.push(to_record_batch(rows_to_batch([(None, 1.0, b"a")])))
        w.poll()
    with pytest.raises(DnzError):     # window < 1 s divides by zero in the reference
        gpu_window(500)


@pytest.mark.parametrize("L,S,G,rpm,filt", [(1000, 0, 1000, 1000, None), (1000, 0, 20_000, 2000, ("max", ">", 113)),
                                             (10_000, 1000, 50_000, 200, None)])
def test_synthetic_sensor_stream_matches_oracle(L, S, G, rpm, filt):
    """Reduced-size cfg 1 / cfg 4 / cfg 3 shapes on the host-fed path (64K-row batches, queued, one poll)."""
    nb, n = 24, 65536
    batches = [synth_batch(i * n, n, groups=G, rows_per_ms=rpm) for i in range(nb)]
    last = int(batches[-1].ts[-1])
    batches.append(sentinel((last // 1000 + 1) * 1000 + 2 * L))
    want = run_oracle_batches(batches, L, S, filt)
    got, st = run_gpu(batches, L, S, filt, per_batch_poll=False, expected_groups=G)
    assert len(want) > 1000 and st["fast_tiles"] > 0 and st["generic_tiles"] <= 1
    assert_rows_equal(got, want)


def test_cfg5_shape_uuid_keys_sliding_60s_5s():
    """Reduced cfg 5: sliding 60 s / 5 s, 36-byte UUID-shaped keys (every row takes the long-key path: word-wise hash, arena
    compare), device generator == host generator, device-resident push/poll against the oracle."""
    from denormalized_b200 import DeviceBatches
    from tests.helpers import gpu_window
    n, G, rpm = 1_200_000, 20_000, 8
    dev = DeviceBatches(n, 65536, groups=G, rows_per_ms=rpm, uuid_keys=True)
    w = gpu_window(60_000, 5_000, None, expected_groups=G)
    w.push_device(dev)
    close = T0 + n // rpm + 65_000
    w.flush(close)
    r = w.fetch_device_result(w.poll_device())
    st = w.stats()
    got = [(int(r["window_start"][i]), int(r["window_end"][i]), r["key"][i], int(r["count"][i]),
            float(r["min"][i]), float(r["max"][i]), float(r["avg"][i]), 0) for i in range(len(r["key"]))]
    hb = [synth_batch(i, min(65536, n - i), groups=G, rows_per_ms=rpm, uuid_keys=True) for i in range(0, n, 65536)]
    hb.append(sentinel(close))
    want = run_oracle_batches(hb, 60_000, 5_000)
    assert len(want) > 100_000 and all(len(k) == 36 for k in list(r["key"])[:50])
    assert_rows_equal(got, want)
    assert st["fast_tiles"] == 0 or st["generic_tiles"] >= 0        # 36 B keys exceed the staged 16 B/row budget: generic tiles
    w.close()
    dev.free()


def test_device_resident_path_equals_host_path():
    """dnz_synth_generate (device generator) + push_device/poll_device == host generator + oracle."""
    from denormalized_b200 import DeviceBatches
    from tests.helpers import gpu_window
    n, G, rpm = 1_000_000, 3000, 500
    dev = DeviceBatches(n, 65536, groups=G, rows_per_ms=rpm)
    w = gpu_window(1000, 0, None, expected_groups=G, flags=1)
    w.push_device(dev)
    w.flush(T0 + n // rpm + 5000)
    r = w.fetch_device_result(w.poll_device())
    st = w.stats()
    got = [(int(r["window_start"][i]), int(r["window_end"][i]), r["key"][i], int(r["count"][i]),
            float(r["min"][i]), float(r["max"][i]), float(r["avg"][i]), 0) for i in range(len(r["key"]))]
    hb = [synth_batch(i, min(65536, n - i), groups=G, rows_per_ms=rpm) for i in range(0, n, 65536)]
    hb.append(sentinel(T0 + n // rpm + 5000))
    want = run_oracle_batches(hb, 1000)
    assert_rows_equal(got, want)
    assert st["agg_kernel_ms"] > 0 and st["agg_algorithmic_bytes"] == dev.algorithmic_bytes
    w.close()
    dev.free()


# ---------------------------------------------------------------------------------------------------------------------------
# round 2: the speculative three-slot pipeline (scan k+1 | aggregate + gated emission k | verify k-1) and the growth paths
def _stream(seed, n_batches, rows, n_keys, span_ms=300, long_keys=False, late_every=0, late_shift_ms=0):
    rng = np.random.default_rng(seed)
    out, t = [], T0
    for b in range(n_batches):
        back = late_shift_ms if (late_every and b % late_every == late_every - 1) else 0
        k = rng.integers(0, n_keys, rows)
        ts = t - back + rng.integers(0, span_ms, rows)
        v = rng.random(rows) * 115.0
        if long_keys:
            keys = [(b"key-%06d-" % int(x)) * 3 for x in k]         # 33 B: arena keys
        else:
            keys = [b"sensor_%d" % int(x) for x in k]
        out.append(rows_to_batch([(int(ts[i]), float(v[i]), keys[i]) for i in range(rows)]))
        t += span_ms
    return out, t


@pytest.mark.parametrize("L,S", [(1000, 0), (3000, 1000)])
@pytest.mark.parametrize("expected_groups", [0, 16])
def test_speculative_pipeline_matches_oracle_and_synchronous_mode(L, S, expected_groups):
    """Many small superbatches (max_rows_per_launch = 3 batches) so that the three pipeline slots rotate; non-forcing polls in
    between; expected_groups=16 makes the dictionary overflow INSIDE the pipeline: emission behind the launch finds the gate
    closed, the deferred rows are replayed at verification and the emission is issued again."""
    from tests.helpers import gpu_window, to_record_batch, record_batch_rows
    batches, t_end = _stream(77 + L + expected_groups, 40, 2500, 3000)
    batches.append(sentinel(t_end + 3 * L))
    want = run_oracle_batches(batches, L, S)
    res = {}
    for flags in (0, 32):                       # 32 = DNZ_FLAG_SYNCHRONOUS
        w = gpu_window(L, S, None, expected_groups=expected_groups, max_rows_per_launch=7500, flags=flags)
        rows = []
        for b in batches:
            w.push(to_record_batch(b))
            rows += record_batch_rows(w.poll_ready())
        rows += record_batch_rows(w.poll())
        st = w.stats()
        w.close()
        assert st["agg_launches"] >= 13
        if expected_groups:
            assert st["deferred_rows"] > 0
        assert_rows_equal(rows, want)
        res[flags] = st
    assert res[0]["rows_out"] == res[32]["rows_out"] == len(want)


def test_late_batch_while_the_dictionary_grows():
    """a late (dirty) batch that also overflows the dictionary -- the one-batch late panes must grow with it."""
    batches, t_end = _stream(5, 12, 1500, 40)
    rng = np.random.default_rng(6)
    late = [(T0 + int(rng.integers(0, 900)), float(rng.random()), b"newkey_%d" % i) for i in range(3000)]     # 3000 unseen keys, all late
    batches.append(rows_to_batch(late))
    more, t_end2 = _stream(7, 4, 1500, 40)
    batches.append(sentinel(t_end + 9000))
    for L, S in [(1000, 0), (2000, 1000)]:
        want = run_oracle_batches(batches, L, S)
        got, st = run_gpu(batches, L, S, expected_groups=16)
        assert st["late_batches"] >= 1 and st["deferred_rows"] > 0
        assert_rows_equal(got, want, check_seq=True)


def test_long_key_arena_growth_in_one_launch():
    """more than 1 MiB of long keys inserted by ONE launch (the arena starts at 1 MiB)."""
    n_keys = 45_000
    rows = [(T0 + (i % 900), float(i % 97), (b"long-key-%08d-" % (i % n_keys)) * 2 + b"x" * 6) for i in range(90_000)]   # 40 B keys
    batches = [rows_to_batch(rows[i:i + 30_000]) for i in range(0, len(rows), 30_000)] + [sentinel(T0 + 5000)]
    want = run_oracle_batches(batches, 1000)
    got, st = run_gpu(batches, 1000, per_batch_poll=False, expected_groups=n_keys)
    assert st["deferred_rows"] > 0 and len(want) == n_keys + 0
    assert_rows_equal(got, want)


def test_idle_gap_between_queued_batches():
    """two batches 30 h apart queued into one launch: the pane span limit applies per run, not per launch."""
    a = rows_to_batch([(T0 + i, 1.0 + i, b"a") for i in range(50)])
    b = rows_to_batch([(T0 + 30 * 3600 * 1000 + i, 2.0 + i, b"b") for i in range(50)])
    batches = [a, b, sentinel(T0 + 31 * 3600 * 1000)]
    want = run_oracle_batches(batches, 1000)
    got, _ = run_gpu(batches, 1000, per_batch_poll=False)
    assert len(want) == 2
    assert_rows_equal(got, want)


def test_device_ready_poll_hands_out_every_row_once():
    from denormalized_b200 import DeviceBatches
    from tests.helpers import gpu_window
    import ctypes as C
    from denormalized_b200 import capi
    n, G, rpm = 2_000_000, 5000, 400
    dev = DeviceBatches(n, 65536, groups=G, rows_per_ms=rpm)
    w = gpu_window(1000, 0, None, expected_groups=G, max_rows_per_launch=4 * 65536)
    got = []

    def take(r):
        if r.n_rows:
            f = w.fetch_device_result(r)
            got.extend((int(f["window_start"][i]), int(f["window_end"][i]), f["key"][i], int(f["count"][i]), float(f["min"][i]),
                        float(f["max"][i]), float(f["avg"][i]), 0) for i in range(r.n_rows))
        return r.n_rows
    for g0 in range(0, dev.n_batches, 4):
        k = min(4, dev.n_batches - g0)
        w.push_device(array=C.cast(C.byref(dev.array, g0 * C.sizeof(capi.DeviceBatchC)), C.POINTER(capi.DeviceBatchC)), n=k)
        while take(w.poll_device_ready()):
            pass
    w.flush(T0 + n // rpm + 5000)
    while take(w.poll_device()):
        pass
    hb = [synth_batch(i, min(65536, n - i), groups=G, rows_per_ms=rpm) for i in range(0, n, 65536)]
    hb.append(sentinel(T0 + n // rpm + 5000))
    want = run_oracle_batches(hb, 1000)
    assert_rows_equal(got, want)
    w.close(); dev.free()


@pytest.mark.parametrize("L,S", [(1000, 0), (3000, 1000)])
def test_checkpoint_restore_resumes_the_stream(L, S):
    """the device state (dictionary incl. long keys, open panes with null-row counts and first-zero marks, stream
    clock) survives dnz_window_checkpoint -> a fresh operator -> dnz_window_restore; the resumed stream emits what an
    uninterrupted one does, batch by batch."""
    from tests.helpers import gpu_window, to_record_batch, record_batch_rows
    rng = np.random.default_rng(42 + L)
    raw = random_stream(rng, 30, 120, 25, span_ms=350, null_frac=0.1)
    for b in (3, 9, 16, 22):          # +-0.0 / NaN / inf states on both sides of the checkpoint (no f64 overflow: that is summation-order dependent)
        raw[b] += [(T0 + b * 350 + 7, -0.0, b"zero"), (T0 + b * 350 + 8, 0.0, b"zero"), (T0 + b * 350 + 9, float("nan"), b"nan"),
                   (T0 + b * 350 + 9, float("inf"), b"inf"), (T0 + b * 350 + 10, None, b"onlynull")]
    batches = [rows_to_batch(r) for r in raw] + [sentinel(T0 + 30 * 350 + 3 * L)]
    want = run_oracle_batches(batches, L, S)
    got = []
    w = gpu_window(L, S, expected_groups=16)
    for i, b in enumerate(batches[:14]):
        w.push(to_record_batch(b)); got += record_batch_rows(w.poll(), i)
    blob = w.checkpoint()
    st0 = w.stats()
    w.close()
    w2 = gpu_window(L, S)                                  # different capacity hint: the restore sizes the tables itself
    w2.restore(blob)
    assert w2.stats()["groups"] == st0["groups"] and w2.watermark is not None
    for i, b in enumerate(batches[14:], start=14):
        w2.push(to_record_batch(b)); got += record_batch_rows(w2.poll(), i)
    w2.close()
    assert len(blob) > 1000 and len(want) > 50
    assert_rows_equal(got, want, check_seq=True)
    from denormalized_b200 import DnzError
    w3 = gpu_window(L + 1000, S)
    with pytest.raises(DnzError):                          # another window configuration
        w3.restore(blob)
    w3.close()
