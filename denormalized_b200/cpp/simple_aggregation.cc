// simple_aggregation.cc -- the reference's examples/examples/simple_aggregation.rs:30-60 (+ the filter of udf_example.rs:62)
// written against the C++ mirror of the operator API.  Kafka is replaced by in-memory synthetic sensor batches of the
// canonical schema (kafka_config.rs:186-214; generator of SURVEY.md §8d).  Prints the emitted rows as CSV so that the
// GPU tests can diff them against the oracle.
//   usage: simple_aggregation <n_batches> <rows_per_batch> <groups> <rows_per_ms> <window_ms> <slide_ms|0> [filter_max_gt]
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "denormalized.hpp"

using namespace denormalized;

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}

// ---- minimal hand-rolled Arrow C-Data producers (what arrow-rs' FFI_ArrowArray::new does for the Rust caller) -------
struct OwnedColumns {            // keeps every buffer of one exported batch alive until the consumer releases it
  std::vector<int64_t> ts, occurred; std::vector<double> val; std::vector<int32_t> key_off, bar_off; std::string key_bytes, bar_bytes;
  ArrowArray children[4]; ArrowArray meta_children[2]; ArrowArray* child_ptrs[4]; ArrowArray* meta_ptrs[2];
  const void* bufs[6][3];
};
static void release_noop(ArrowArray* a) { a->release = nullptr; }
static void release_top(ArrowArray* a) { delete static_cast<OwnedColumns*>(a->private_data); a->release = nullptr; }
static void init_array(ArrowArray* a, int64_t n, int nbuf, const void** bufs, int nchild, ArrowArray** children) {
  memset(a, 0, sizeof *a);
  a->length = n; a->n_buffers = nbuf; a->buffers = bufs; a->n_children = nchild; a->children = children; a->release = release_noop;
}

static void make_batch(int64_t row0, int64_t n, int64_t groups, int64_t rows_per_ms, ArrowArray* out) {
  const int64_t T0 = 1700000000000ll;
  auto* c = new OwnedColumns();
  c->ts.resize(n); c->occurred.resize(n); c->val.resize(n); c->key_off.resize(n + 1); c->bar_off.resize(n + 1);
  c->key_off[0] = 0; c->bar_off[0] = 0;
  for (int64_t k = 0; k < n; k++) {
    uint64_t i = (uint64_t)(row0 + k), r = splitmix64(42 ^ i), r2 = splitmix64(r);
    c->ts[k] = T0 + (int64_t)(i / (uint64_t)rows_per_ms); c->occurred[k] = c->ts[k];
    c->val[k] = ((double)(r2 >> 11) * 0x1.0p-53) * 115.0;
    c->key_bytes += "sensor_" + std::to_string((r >> 11) % (uint64_t)groups); c->key_off[k + 1] = (int32_t)c->key_bytes.size();
    c->bar_bytes += "no_barrier"; c->bar_off[k + 1] = (int32_t)c->bar_bytes.size();
  }
  c->key_bytes.append(16, '\0');
  c->bufs[0][0] = nullptr; c->bufs[0][1] = c->occurred.data();
  c->bufs[1][0] = nullptr; c->bufs[1][1] = c->val.data();
  c->bufs[2][0] = nullptr; c->bufs[2][1] = c->key_off.data(); c->bufs[2][2] = c->key_bytes.data();
  c->bufs[3][0] = nullptr;
  c->bufs[4][0] = nullptr; c->bufs[4][1] = c->bar_off.data(); c->bufs[4][2] = c->bar_bytes.data();
  c->bufs[5][0] = nullptr; c->bufs[5][1] = c->ts.data();
  init_array(&c->meta_children[0], n, 3, c->bufs[4], 0, nullptr);
  init_array(&c->meta_children[1], n, 2, c->bufs[5], 0, nullptr);
  c->meta_ptrs[0] = &c->meta_children[0]; c->meta_ptrs[1] = &c->meta_children[1];
  init_array(&c->children[0], n, 2, c->bufs[0], 0, nullptr);
  init_array(&c->children[1], n, 2, c->bufs[1], 0, nullptr);
  init_array(&c->children[2], n, 3, c->bufs[2], 0, nullptr);
  init_array(&c->children[3], n, 1, c->bufs[3], 2, c->meta_ptrs);
  for (int i = 0; i < 4; i++) c->child_ptrs[i] = &c->children[i];
  static const void* top_bufs[1] = {nullptr};
  init_array(out, n, 1, top_bufs, 4, c->child_ptrs);
  out->private_data = c; out->release = release_top;
}

struct OwnedSchema { ArrowSchema top, c[4], m[2]; ArrowSchema* cp[4]; ArrowSchema* mp[2]; };
static void schema_noop(ArrowSchema* s) { s->release = nullptr; }
static void init_schema(ArrowSchema* s, const char* fmt, const char* name, int nchild, ArrowSchema** ch) {
  memset(s, 0, sizeof *s); s->format = fmt; s->name = name; s->flags = ARROW_FLAG_NULLABLE; s->n_children = nchild; s->children = ch; s->release = schema_noop;
}
static void canonical_schema(OwnedSchema* o) {
  init_schema(&o->m[0], "u", "barrier_batch", 0, nullptr); init_schema(&o->m[1], "tsm:", "canonical_timestamp", 0, nullptr);
  o->mp[0] = &o->m[0]; o->mp[1] = &o->m[1];
  init_schema(&o->c[0], "l", "occurred_at_ms", 0, nullptr); init_schema(&o->c[1], "g", "reading", 0, nullptr);
  init_schema(&o->c[2], "u", "sensor_name", 0, nullptr); init_schema(&o->c[3], "+s", "_streaming_internal_metadata", 2, o->mp);
  for (int i = 0; i < 4; i++) o->cp[i] = &o->c[i];
  init_schema(&o->top, "+s", "", 4, o->cp);
}

static void print_rows(const RecordBatch& rb) {
  if (rb.num_rows() == 0) return;
  const ArrowArray* key = rb.array.children[0];
  const int32_t* off = (const int32_t*)key->buffers[1]; const char* bytes = (const char*)key->buffers[2];
  const int64_t* cnt = (const int64_t*)rb.array.children[1]->buffers[1];
  const double* mn = (const double*)rb.array.children[2]->buffers[1]; const double* mx = (const double*)rb.array.children[3]->buffers[1];
  const double* av = (const double*)rb.array.children[4]->buffers[1];
  const int64_t* ws = (const int64_t*)rb.array.children[5]->buffers[1]; const int64_t* we = (const int64_t*)rb.array.children[6]->buffers[1];
  for (int64_t i = 0; i < rb.num_rows(); i++)
    printf("%" PRId64 ",%" PRId64 ",%.*s,%" PRId64 ",%a,%a,%a\n", ws[i], we[i], off[i + 1] - off[i], bytes + off[i], cnt[i], mn[i], mx[i], av[i]);
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s n_batches rows_per_batch groups rows_per_ms window_ms slide_ms [filter_max_gt]\n", argv[0]); return 2; }
  int64_t nb = atoll(argv[1]), n = atoll(argv[2]), groups = atoll(argv[3]), rpm = atoll(argv[4]), L = atoll(argv[5]), S = atoll(argv[6]);
  OwnedSchema schema; canonical_schema(&schema);
  try {
    DataStream ds = DataStream::from_schema(&schema.top)
                        .window({col("sensor_name")},
                                {count(col("reading")).alias("count"), min(col("reading")).alias("min"), max(col("reading")).alias("max"),
                                 avg(col("reading")).alias("average")},
                                Duration::from_millis(L), S ? std::optional<Duration>(Duration::from_millis(S)) : std::nullopt);
    if (argc > 7) ds = ds.filter(col("max").gt(lit(atof(argv[7]))));
    StreamingWindowExec exec = ds.plan();
    GroupedWindowAggStream stream = exec.execute(0, groups);
    for (int64_t b = 0; b < nb; b++) {
      ArrowArray batch; make_batch(b * n, n, groups, rpm, &batch);
      print_rows(stream.poll_next(&batch));           // the reference polls once per upstream batch
    }
    // closing batch: one row far enough in the future to close every window (the reference never flushes)
    ArrowArray last; make_batch(nb * n + 100 * (L + 1000) * rpm, 1, groups, rpm, &last);
    print_rows(stream.poll_next(&last));
    dnz_stats st = stream.metrics();
    fprintf(stderr, "rows_in=%" PRId64 " rows_out=%" PRId64 " launches=%" PRId64 "\n", st.rows_in, st.rows_out, st.total_launches);
  } catch (const DataFusionError& e) {
    fprintf(stderr, "DataFusionError(%d): %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
