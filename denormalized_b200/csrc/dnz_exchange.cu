// dnz_exchange.cu -- multi-GPU pane exchange: the all-to-all of per-pane partial aggregates that replaces
// RepartitionExec(Hash(group keys)) (physical_optimizer/coalesce_before_streaming_window_aggregate.rs:63-73) when the
// input is NOT key-partitioned: every GPU aggregates the batches it was dealt into its own panes; once a pane can no
// longer change, each GPU sends the states of the keys it does not own to their owners (owner = key hash % world), the
// owner merges them into its pane (count +, sum +, min / max over the ordered keys, null rows +, first-zero min) and
// emits the windows for its keys only (k_emit's owner filter).
#include "dnz_device.cuh"

namespace dnz {

// ---- pack: thread per group id of one pane; two passes (count, then write) over every pane of the export ----------------
__global__ void __launch_bounds__(256) k_pack_partials(const __grid_constant__ PackParams P) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.n_groups) return;
  const GroupState s = P.st[g];
  const unsigned long long nr = P.nullrows ? P.nullrows[g] : 0ull;
  if (s.cnt == 0.0 && nr == 0ull) return;
  const GidKey gk = P.dict.gid_key[g];
  const bool null_key = gk.len == 0xFFFFFFFFu;
  const uint32_t klen = null_key ? 0u : gk.len;
  const int owner = null_key ? 0 : (int)((klen <= (uint32_t)INLINE_KEY ? hash_inline(gk.k0, gk.k1, klen) : gk.k0) % (uint64_t)P.world);
  if (owner == P.rank) return;
  const uint32_t kpad = (klen + 7u) & ~7u;
  const unsigned long long c = atomicAdd(P.owner_cursor + owner, (1ull << 32) | kpad);
  if (P.pass == 0) return;
  const uint64_t row = (P.owner_base[owner] >> 32) + (c >> 32);
  const uint32_t boff = (uint32_t)(c & 0xFFFFFFFFull);                    // inside this owner's key segment
  PartialEntry e;
  e.pane = P.pane; e.cnt = (unsigned long long)s.cnt; e.sum = s.sum; e.minkey = s.minkey; e.maxkey = s.maxkey;
  e.nullrows = nr; e.fz = P.fz ? P.fz[g] : ~0ull;
  e.key_off = boff; e.key_len = null_key ? 0xFFFFFFFFu : klen;
  P.entries[row] = e;
  if (!null_key) {
    uint8_t* dst = P.key_bytes + (P.owner_base[owner] & 0xFFFFFFFFull) + boff;
    if (klen <= (uint32_t)INLINE_KEY) {
      const uint64_t w[2] = {gk.k0, gk.k1};
      for (uint32_t i = 0; i < klen; i++) dst[i] = (uint8_t)(w[i >> 3] >> ((i & 7) * 8));
    } else {
      const uint8_t* src = P.dict.arena + gk.k1;
      for (uint32_t i = 0; i < klen; i++) dst[i] = src[i];
    }
  }
}
cudaError_t launch_pack_partials(const PackParams& p, cudaStream_t s) {
  if (!p.n_groups) return cudaSuccess;
  k_pack_partials<<<(p.n_groups + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ---- merge: thread per received packet ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_merge_partials(const __grid_constant__ MergeParams P) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P.n_entries) return;
  const PartialEntry e = P.entries[i];
  int src = 0;
  while (src + 1 < P.world && i >= P.src_entry_end[src]) src++;
  uint32_t gid;
  if (e.key_len == 0xFFFFFFFFu) gid = dict_lookup_null(P.dict);
  else {
    KeyRef k; load_key<false>(P.key_bytes + P.src_key_base[src] + e.key_off, e.key_len, k);
    gid = dict_lookup(P.dict, k, false);
  }
  const int64_t pi = e.pane - P.panes.pane0;
  if (gid >= GID_DEFER_ARENA || pi < 0 || pi >= P.panes.n_panes || P.panes.main[pi] == nullptr) { atomicOr(P.error, 1u); return; }
  GroupState* s = P.panes.main[pi] + gid;
  if (e.cnt) {
    red_add_f64(&s->cnt, (double)e.cnt); red_add_f64(&s->sum, e.sum);
    red_max_u64(&s->minkey, e.minkey); red_max_u64(&s->maxkey, e.maxkey);
  }
  if (e.nullrows) { if (P.panes.nullrows_main[pi]) red_add_u64(P.panes.nullrows_main[pi] + gid, e.nullrows); else atomicOr(P.error, 2u); }
  if (e.fz != ~0ull) { if (P.panes.fz_main[pi]) red_min_u64(P.panes.fz_main[pi] + gid, e.fz); else atomicOr(P.error, 4u); }
}
cudaError_t launch_merge_partials(const MergeParams& p, cudaStream_t s) {
  if (p.n_entries <= 0) return cudaSuccess;
  k_merge_partials<<<(unsigned)((p.n_entries + 255) / 256), 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace dnz
