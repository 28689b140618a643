//! Plan-A harness: run the UNMODIFIED reference (`denormalized` crate, DataFusion fork d812edc) over the
//! synthetic sensor stream and write what its own `GroupedWindowAggStream` (+ `FilterExec`) emits, so that the CPU oracle of
//! this repository (oracle/dnz_oracle.c) can be pinned against the real thing.
//!
//! NOT COMPILED IN THE AUTHORING IMAGE (no cargo/rustc, no vendored crates): this is the recipe a maintainer with a Rust
//! toolchain runs once; see oracle/plan_a/README.md.  Drop this file into the reference checkout as
//! `examples/examples/emit_fixtures.rs` and `cargo run --release --example emit_fixtures -- <cfg> <rows> <out.jsonl>`.
//!
//! Shape of the query = `examples/examples/simple_aggregation.rs:45-55` (+ `.filter(col("max").gt(lit(113)))` as in
//! `examples/examples/udf_example.rs:62` for cfg 4).  Kafka is replaced by an in-memory `PartitionStream` that yields the
//! synthetic batches -- each carrying `_streaming_internal_metadata { barrier_batch, canonical_timestamp }` exactly as
//! `kafka_stream_read.rs:222-271` attaches it -- followed by ONE sentinel batch and then stays `Pending` forever (the window
//! stream never returns `None`, `grouped_window_agg_stream.rs:343-348`).  `target_partitions = 1`, so the output is what a
//! single partition's stream emits and `emit_seq` (index of the poll that produced a row) is meaningful.
use std::sync::Arc;
use std::time::Duration;

use datafusion::arrow::array::{ArrayRef, Float64Array, Int64Array, StringArray, StructArray, TimestampMillisecondArray};
use datafusion::arrow::datatypes::{DataType, Field, Schema, SchemaRef, TimeUnit};
use datafusion::arrow::record_batch::RecordBatch;
use datafusion::datasource::streaming::StreamingTable;
use datafusion::execution::{SendableRecordBatchStream, TaskContext};
use datafusion::functions_aggregate::expr_fn::{avg, count, max, min};
use datafusion::logical_expr::{col, lit};
use datafusion::physical_plan::stream::RecordBatchStreamAdapter;
use datafusion::physical_plan::streaming::PartitionStream;
use denormalized::context::Context;
use denormalized::datastream::DataStream;
use futures::StreamExt;

const T0: i64 = 1_700_000_000_000;
const BATCH_ROWS: usize = 65_536;

fn splitmix64(x: u64) -> u64 {
    let x = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
    let mut z = x;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}

/// Row `i` of the stream: identical to oracle/dnz_oracle.c `orc_synth_fill` and denormalized_b200/csrc/dnz_synth.cu.
fn row(i: u64, seed: u64, groups: u64, rows_per_ms: u64, uuid: bool) -> (i64, f64, String) {
    let r = splitmix64(seed ^ i);
    let key_id = (r >> 11) % groups;
    let r2 = splitmix64(r);
    let reading = ((r2 >> 11) as f64) * (115.0 / 9_007_199_254_740_992.0);
    let ts = T0 + (i / rows_per_ms) as i64;
    let key = if uuid {
        let (a, b) = (splitmix64(key_id), splitmix64(splitmix64(key_id)));
        let h = format!("{:016x}{:016x}", a, b);
        format!("{}-{}-{}-{}-{}", &h[0..8], &h[8..12], &h[12..16], &h[16..20], &h[20..32])
    } else {
        format!("sensor_{}", key_id)
    };
    (ts, reading, key)
}

fn schema() -> SchemaRef {
    let meta = DataType::Struct(
        vec![
            Field::new("barrier_batch", DataType::Utf8, false),
            Field::new("canonical_timestamp", DataType::Timestamp(TimeUnit::Millisecond, None), true),
        ]
        .into(),
    );
    Arc::new(Schema::new(vec![
        Field::new("occurred_at_ms", DataType::Int64, true),
        Field::new("reading", DataType::Float64, true),
        Field::new("sensor_name", DataType::Utf8, true),
        Field::new("_streaming_internal_metadata", meta, false),
    ]))
}

fn batch(rows: &[(i64, f64, String)]) -> RecordBatch {
    let ts: Vec<i64> = rows.iter().map(|r| r.0).collect();
    let meta = StructArray::from(vec![
        (
            Arc::new(Field::new("barrier_batch", DataType::Utf8, false)),
            Arc::new(StringArray::from(vec!["no_barrier"; rows.len()])) as ArrayRef,
        ),
        (
            Arc::new(Field::new("canonical_timestamp", DataType::Timestamp(TimeUnit::Millisecond, None), true)),
            Arc::new(TimestampMillisecondArray::from(ts.clone())) as ArrayRef,
        ),
    ]);
    RecordBatch::try_new(
        schema(),
        vec![
            Arc::new(Int64Array::from(ts)),
            Arc::new(Float64Array::from(rows.iter().map(|r| r.1).collect::<Vec<_>>())),
            Arc::new(StringArray::from(rows.iter().map(|r| r.2.clone()).collect::<Vec<_>>())),
            Arc::new(meta),
        ],
    )
    .unwrap()
}

#[derive(Debug)]
struct Synthetic { rows: u64, seed: u64, groups: u64, rows_per_ms: u64, uuid: bool, sentinel_ts: i64 }

impl PartitionStream for Synthetic {
    fn schema(&self) -> &SchemaRef {
        static S: std::sync::OnceLock<SchemaRef> = std::sync::OnceLock::new();
        S.get_or_init(schema)
    }
    fn execute(&self, _ctx: Arc<TaskContext>) -> SendableRecordBatchStream {
        let (n, seed, g, rpm, uuid, sentinel) = (self.rows, self.seed, self.groups, self.rows_per_ms, self.uuid, self.sentinel_ts);
        let data = futures::stream::iter((0..n).step_by(BATCH_ROWS).map(move |r0| {
            let rows: Vec<_> = (r0..(r0 + BATCH_ROWS as u64).min(n)).map(|i| row(i, seed, g, rpm, uuid)).collect();
            Ok(batch(&rows))
        }));
        let tail = futures::stream::once(async move { Ok(batch(&[(sentinel, 1.0, "sentinel".to_string())])) });
        // never terminate: FilterExec / the window stream would otherwise spin on a finished input
        let stream = data.chain(tail).chain(futures::stream::pending());
        Box::pin(RecordBatchStreamAdapter::new(schema(), stream))
    }
}

#[tokio::main]
async fn main() -> datafusion::error::Result<()> {
    let args: Vec<String> = std::env::args().collect();
    let (cfg, rows, out) = (args[1].as_str(), args[2].parse::<u64>().unwrap(), args[3].clone());
    // (groups, rows_per_ms, window_ms, slide_ms, filter, uuid): SURVEY.md §8d / bench.py WORKLOADS
    let (groups, rpm, win, slide, filter, uuid) = match cfg {
        "cfg1" => (1_000, 1_000, 1_000, None, false, false),
        "cfg2" => (100_000, 10_000, 1_000, None, false, false),
        "cfg3" => (1_000_000, 10_000, 10_000, Some(1_000), false, false),
        "cfg4" => (100_000, 10_000, 1_000, None, true, false),
        "cfg5" => (10_000_000, 8_000, 60_000, Some(5_000), false, true),
        _ => panic!("cfg1..cfg5"),
    };
    let last = T0 + ((rows - 1) / rpm) as i64;
    let sentinel_ts = (last / 1000 + 1) * 1000 + 2 * win as i64;
    let mut config = Context::default_config();
    config.options_mut().execution.target_partitions = 1;
    let ctx = Context::with_config(config).unwrap();
    let table = StreamingTable::try_new(schema(), vec![Arc::new(Synthetic { rows, seed: 42, groups, rows_per_ms: rpm, uuid, sentinel_ts })])?
        .with_infinite_table(true);
    ctx.register_table("synthetic".to_string(), Arc::new(table)).await.unwrap();
    let df = ctx.session_context.table("synthetic").await?;
    let mut ds = DataStream::new(Arc::new(df), Arc::new(ctx.clone())).window(
        vec![col("sensor_name")],
        vec![
            count(col("reading")).alias("count"),
            min(col("reading")).alias("min"),
            max(col("reading")).alias("max"),
            avg(col("reading")).alias("average"),
        ],
        Duration::from_millis(win),
        slide.map(Duration::from_millis),
    ).unwrap();
    if filter { ds = ds.filter(col("max").gt(lit(113))).unwrap(); }
    let mut stream = ds.df.as_ref().clone().execute_stream().await?;
    let last_window_start = (last / 1000) * 1000 - (last / 1000 * 1000 - T0) % slide.unwrap_or(win) as i64;
    let mut w = std::io::BufWriter::new(std::fs::File::create(out)?);
    use std::io::Write;
    let mut poll = 0u64;
    'outer: while let Some(b) = stream.next().await {
        let b = b?;
        poll += 1;
        if b.num_rows() == 0 { continue; }
        let key = b.column(0).as_any().downcast_ref::<StringArray>().unwrap();
        let (c, mn, mx, av) = (
            b.column_by_name("count").unwrap().as_any().downcast_ref::<Int64Array>().unwrap(),
            b.column_by_name("min").unwrap().as_any().downcast_ref::<Float64Array>().unwrap(),
            b.column_by_name("max").unwrap().as_any().downcast_ref::<Float64Array>().unwrap(),
            b.column_by_name("average").unwrap().as_any().downcast_ref::<Float64Array>().unwrap(),
        );
        let ws = b.column_by_name("window_start_time").unwrap().as_any().downcast_ref::<TimestampMillisecondArray>().unwrap();
        let we = b.column_by_name("window_end_time").unwrap().as_any().downcast_ref::<TimestampMillisecondArray>().unwrap();
        for i in 0..b.num_rows() {
            // f64 as bit patterns: the comparison is bit-exact for min/max and 1e-9 relative for average
            writeln!(w, "{{\"ws\":{},\"we\":{},\"key\":{:?},\"count\":{},\"min\":\"{:016x}\",\"max\":\"{:016x}\",\"avg\":\"{:016x}\",\"poll\":{}}}",
                     ws.value(i), we.value(i), key.value(i), c.value(i), mn.value(i).to_bits(), mx.value(i).to_bits(), av.value(i).to_bits(), poll)?;
            if ws.value(i) >= last_window_start { w.flush()?; break 'outer; }     // the window holding the last data row has been emitted
        }
    }
    w.flush()?;
    Ok(())
}
