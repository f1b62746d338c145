// capi.cu — the C ABI (include/recom_b200.h): plan construction, per-request
// argument marshalling and kernel enqueue. Host code only; kernels live in
// kernels.cu.
//
// Reference counterparts (reference root relative):
//   plan create  <-> generated CreateConstBuffers + the code generator's per-column
//                    bookkeeping (graph_optimizers/cuda_emitter.cc:975-1038,2260-2301)
//   plan run     <-> generated ProcessFeatureColumns + kernel caller
//                    (cuda_emitter.cc:2139-2258,2303-2494) and the op's Compute
//                    (custom_ops/feature_column_process/feature_column_process_op_gpu.cu.cc:65-131)
// Differences by design: no host<->device synchronisation anywhere (the reference
// has three per request), one small H2D upload per request from pinned memory,
// errors are returned instead of exit().
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <queue>
#include <string>
#include <vector>

#include "plan.h"

using namespace recom;

namespace {

thread_local std::string g_last_error;
thread_local int g_last_launches = 0;

recom_status fail(recom_status st, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return st;
}

}  // namespace

namespace recom {
// shared with host_ops.cc
recom_status set_error(recom_status st, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return st;
}
}  // namespace recom

namespace {

#define CUDA_TRY(expr)                                                             \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess)                                                         \
      return fail(RECOM_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

constexpr int kRing = 16;       // pinned staging slots of the upload path
constexpr int kTableSets = 8;   // device-resident table pointer arrays kept per plan
constexpr int kCounterSlots = 64;  // requests of one plan in flight at once (more: the call waits)
constexpr int kCounterInts = 4; // per slot: queue claims, finished warps, transform tasks done, -

struct ArgSlot {
  std::mutex mu;
  uint8_t* pinned = nullptr;
  size_t cap = 0;
  cudaEvent_t done = nullptr;
  bool in_flight = false;
};

// A device-resident copy of a request's table pointer array, keyed by content. Entries are
// immutable once published and live until the plan dies, so kernels in flight on any stream
// may read them without further synchronisation. (Tables are the graph's variables: a serving
// process passes the same addresses on every call.)
struct TableSet {
  std::vector<const void*> host;
  const void** dev = nullptr;
};

int dtype_bytes(int32_t dt) {
  switch (dt) {
    case RECOM_F32: return 4;
    case RECOM_BF16: return 2;
    case RECOM_I32: return 4;
    case RECOM_I64: return 8;
    case RECOM_F64: return 8;
    case RECOM_STR: return 4;  // the offsets slot
    default: return 0;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Schedule of one batch size (see Sched in plan.h).
struct TileTable {
  std::vector<Tile> runs;        // FAST runs grouped by CTA
  std::vector<ColItem> queue;    // column tiles
  std::vector<CtaWork> cta;      // [grid]
  std::vector<ColItem> flat;     // column tiles of the flat kernel (one warp each), by descending cost
  int dense_tiles = 0;           // column tiles of the persistent kernel with one id per row
  int grid = 1;
  int fast_warps = 0;
  int fast_ctas = -1;
  void* d_block = nullptr;  // device copy: runs, queue, cta (cached tables only)
  ~TileTable() {
    if (d_block) cudaFree(d_block);
  }
  size_t runs_bytes() const { return align_up(runs.size() * sizeof(Tile), 16); }
  size_t queue_bytes() const { return align_up(queue.size() * sizeof(ColItem), 16); }
  size_t cta_bytes() const { return align_up(cta.size() * sizeof(CtaWork), 16); }
  size_t bytes() const { return runs_bytes() + queue_bytes() + cta_bytes() + flat.size() * sizeof(ColItem); }
  const ColItem* flat_items(const void* base) const {
    return (const ColItem*)((const uint8_t*)base + runs_bytes() + queue_bytes() + cta_bytes());
  }
  void write(uint8_t* dst) const {
    if (!runs.empty()) memcpy(dst, runs.data(), runs.size() * sizeof(Tile));
    if (!queue.empty()) memcpy(dst + runs_bytes(), queue.data(), queue.size() * sizeof(ColItem));
    memcpy(dst + runs_bytes() + queue_bytes(), cta.data(), cta.size() * sizeof(CtaWork));
    if (!flat.empty())
      memcpy(dst + runs_bytes() + queue_bytes() + cta_bytes(), flat.data(), flat.size() * sizeof(ColItem));
  }
  Sched sched(const void* base) const {
    Sched s{};
    const uint8_t* b = (const uint8_t*)base;
    s.runs = (const Tile*)b;
    s.queue = (const ColItem*)(b + runs_bytes());
    s.cta = (const CtaWork*)(b + runs_bytes() + queue_bytes());
    s.num_queue = (int32_t)queue.size();
    s.fast_warps = fast_warps;
    s.fast_ctas = fast_ctas;
    for (size_t c = 0; c < cta.size() && c < (size_t)kMaxParamCtas; ++c) s.cta_param[c] = cta[c];
    return s;
  }
};

}  // namespace

struct recom_plan {
  recom_plan_config cfg{};
  std::vector<recom_column_desc> descs;  // as given (for serialisation)
  std::vector<float> boundaries_in;      // as given
  std::vector<recom_column_filter> filters_in;  // as given, one per column (empty: none)
  std::vector<DevFilter> dev_filters;    // distinct filters in use
  // host mirrors of the device arrays
  std::vector<DevColumn> cols;
  std::vector<DevSlab> slabs;
  std::vector<DevChunk> chunks;
  std::vector<DevSlabCol> slab_cols;
  std::vector<DevSlabBnd> slab_bnds;
  std::vector<float> boundary_pool;
  std::map<int32_t, int32_t> set_words;  // pool offset of a boundary set -> its words incl. grid
  std::vector<int32_t> sparse_cols;  // column index of each non-DENSE column
  std::vector<ColItem> units;        // queue units: distinct (input, transform, segments) triples
  std::vector<ColUser> users;        // the columns reading each unit (ColItem::first_user)
  std::vector<double> unit_cost;     // estimated cost of a full tile of a unit (a FAST tile = 2)
  std::vector<char> unit_hbm;        // a user's table is HBM resident: smaller tiles
  std::vector<DevCoo> coos;          // distinct COO inputs
  std::vector<DevSplit> splits;      // distinct string inputs split on the device
  std::vector<int32_t> split_of_col; // [num_cols] index into splits, -1
  // device
  void* d_block = nullptr;
  DevPlan dev{};
  const int32_t* d_out_offsets = nullptr;  // [num_cols] element offset of every column's slice
  int elem_bytes = 4;
  int num_sms = 148;
  int ctas_per_sm = 2;
  int shard_world = 1, shard_rank = 0;  // Mode B: rows of every table sharded over ranks
  bool exclusive = false;      // FAST runs and column tiles on disjoint CTAs
  double col_weight = 1.0;     // scales the column tiles' cost in that split
  int force_fast_warps = -1;  // >= 0: override the cost-proportional split (experiments)
  int col_rows_dense = kColTileRowsDense;    // rows of a column tile, one id per row (<= 128)
  int col_rows_sparse = kColTileRowsSparse;  // rows of a multi-hot column tile (<= 32)
  // The flat kernel of the column tiles (one warp per tile, in front of the persistent kernel,
  // which then keeps the FAST runs only): plain, Mode-B and Mode-A requests alike (sharded output:
  // the warps of both launches count themselves, the last one signals the peers).
  int flat_mode = -1;          // RECOM_B200_FLAT: 1 always, 0 never, unset: by cost (prefer_flat)
  bool flat_pdl = true;        // the persistent kernel starts while the flat one drains
  int flat_ctas_per_sm = 8;    // resident CTAs (4 warps each) of the flat kernel
  std::atomic<bool> request_overlap{true};  // see recom_plan_set_request_overlap
  int defer_ids = kDeferIds;   // RECOM_B200_DEFER_IDS (>= kDeferIds: the workspace bound assumes it)
  int long_ctas_per_sm = 8;    // RECOM_B200_LONG_CTAS (measured on config 3, B = 2048: 2 / 4 / 8 CTAs per SM 177 / 156 / 148 us)
  int flat_warps_target = 0;   // tiles are cut finer until there are about this many (0: resident warps)
  DevPlan dev_flat{};          // = dev with the flat kernel's warp scratch
  // tile tables depend only on the batch size: cached (read-mostly)
  std::mutex tiles_mu;
  std::map<int64_t, std::shared_ptr<const TileTable>> tiles_by_batch;  // key: (batch * 2 + flat) * 4 + density class
  bool stage_tables = true;
  // profiling builds only (RECOM_B200_PROF_DUMP=<file>): per-CTA phase timers of the last run
  uint64_t* d_prof = nullptr;
  int prof_grid = 0;
  std::string prof_path;
  // device-resident table pointer arrays, keyed by content
  std::mutex tables_mu;
  std::vector<std::unique_ptr<TableSet>> table_sets;
  // queue counters: a ring of slots in device memory, one per request in flight. The kernel's
  // last warp resets the slot and stores the request's sequence number into the slot's word of
  // `done_seq` (host-mapped): a slot is handed out again only when its previous user is done.
  int32_t* d_counters = nullptr;
  uint32_t* h_done_seq = nullptr;  // pinned, mapped
  uint32_t* d_done_seq = nullptr;  // device address of the same words
  std::atomic<uint32_t> counter_next{0};
  // Blocking uploads of the caches (first use of a batch size / second sight of an argument
  // block) go through this stream and are complete when the device pointer is published: a
  // plain cudaMemcpy from pageable memory is ordered on the legacy stream only, which does not
  // order it before kernels on the caller's non-blocking streams.
  cudaStream_t up_stream = nullptr;
  // pinned argument ring of the upload path (first sight of a block, COO inputs, full caches)
  ArgSlot ring[kRing];
  std::atomic<uint32_t> ring_next{0};
};

namespace {

recom_status validate_columns(const recom_column_desc* cols, int32_t n, const recom_plan_config& cfg,
                              int64_t num_boundaries) {
  if (cfg.out_dtype != RECOM_F32 && cfg.out_dtype != RECOM_BF16)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "out_dtype must be F32 or BF16");
  if (cfg.out_width <= 0) return fail(RECOM_ERR_INVALID_ARGUMENT, "out_width must be > 0");
  for (int32_t i = 0; i < n; ++i) {
    const recom_column_desc& c = cols[i];
    if (c.dim <= 0 || c.out_offset < 0 || (int64_t)c.out_offset + c.dim > cfg.out_width)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: slice [%d,+%d) outside out_width %d", i,
                  c.out_offset, c.dim, cfg.out_width);
    if (c.table < 0 || c.table >= cfg.num_tables)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: table index %d out of range", i, c.table);
    if (c.table_dtype != cfg.out_dtype)
      return fail(RECOM_ERR_UNSUPPORTED,
                  "column %d: table dtype must equal the concat dtype (got %d vs %d)", i,
                  c.table_dtype, cfg.out_dtype);
    if (c.table_rows <= 0 || c.table_rows > 0x7fffffffLL)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: table_rows %lld unsupported", i,
                  (long long)c.table_rows);
    if ((int64_t)c.dim * dtype_bytes(c.table_dtype) > (1 << 20))
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: dim too large", i);
    if (c.value_slot < 0 || c.value_slot >= cfg.num_slots)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: value_slot out of range", i);
    switch (c.transform) {
      case RECOM_XFORM_NONE:
        if (c.value_dtype != RECOM_I32 && c.value_dtype != RECOM_I64)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: ids must be I32 or I64", i);
        break;
      case RECOM_XFORM_BUCKETIZE:
        if (c.value_dtype != RECOM_F32 && c.value_dtype != RECOM_I32 &&
            c.value_dtype != RECOM_I64 && c.value_dtype != RECOM_F64)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: bad bucketize input dtype", i);
        if (c.num_boundaries < 0 || c.boundaries_offset < 0 ||
            c.boundaries_offset + c.num_boundaries > num_boundaries)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: boundaries out of range", i);
        // ids are 0..num_boundaries (bucketize_op.cc:42-44): they must index the table
        if ((int64_t)c.num_boundaries + 1 > c.table_rows)
          return fail(RECOM_ERR_INVALID_ARGUMENT,
                      "column %d: %d boundaries give ids up to %d, table has %lld rows", i,
                      c.num_boundaries, c.num_boundaries, (long long)c.table_rows);
        break;
      case RECOM_XFORM_HASH_STR:
        if (c.value_dtype != RECOM_STR || c.bytes_slot < 0 || c.bytes_slot >= cfg.num_slots)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: HASH_STR needs STR values + bytes_slot", i);
        if (c.num_buckets <= 0 || c.num_buckets > 0x7fffffffLL)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: num_buckets unsupported", i);
        if (c.num_buckets > c.table_rows)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: num_buckets %lld exceeds table_rows %lld",
                      i, (long long)c.num_buckets, (long long)c.table_rows);
        break;
      case RECOM_XFORM_HASH_INT:
        if (c.value_dtype != RECOM_I32 && c.value_dtype != RECOM_I64)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: HASH_INT needs I32/I64 values", i);
        if (c.num_buckets <= 0 || c.num_buckets > 0x7fffffffLL)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: num_buckets unsupported", i);
        if (c.num_buckets > c.table_rows)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: num_buckets %lld exceeds table_rows %lld",
                      i, (long long)c.num_buckets, (long long)c.table_rows);
        break;
      default:
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: unknown transform %d", i, c.transform);
    }
    if (c.lookup == RECOM_LOOKUP_DENSE) {
      // nothing
    } else if (c.lookup == RECOM_LOOKUP_GATHER_SCATTER || c.lookup == RECOM_LOOKUP_SEGMENT) {
      if (c.seg_format == RECOM_SEG_SPLIT_STR) {
        if (c.transform != RECOM_XFORM_HASH_STR || c.seg_slot != -1 || ((c.seg_stride >> 24) & 15) > 3)
          return fail(RECOM_ERR_INVALID_ARGUMENT,
                      "column %d: SPLIT_STR needs transform HASH_STR, seg_slot -1 and a RECOM_SPLIT_SPEC", i);
      } else if (c.seg_format == RECOM_SEG_FIXED) {
        if (c.seg_slot != -1 || c.seg_stride < 0 || c.lookup != RECOM_LOOKUP_SEGMENT)
          return fail(RECOM_ERR_INVALID_ARGUMENT,
                      "column %d: SEG_FIXED needs lookup SEGMENT, seg_slot -1 and seg_stride >= 0", i);
      } else if (c.seg_slot < 0 || c.seg_slot >= cfg.num_slots)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: seg_slot out of range", i);
      if (c.seg_format == RECOM_SEG_SPLIT_STR || c.seg_format == RECOM_SEG_FIXED) {
        // checked above
      } else if (c.seg_format == RECOM_SEG_COO_RESHAPE_I64) {
        const int rank = c.seg_stride & 255, extra = (c.seg_stride & 256) ? 2 : 1;
        if (rank < 1 || rank > 8 || (c.seg_stride & ~511) || c.seg_slot + extra >= cfg.num_slots)
          return fail(RECOM_ERR_INVALID_ARGUMENT,
                      "column %d: COO_RESHAPE needs RECOM_RESHAPE_SPEC(rank 1..8, has_new_shape) and the shape "
                      "slots behind seg_slot", i);
      } else if (c.seg_format == RECOM_SEG_COO_I64) {
        if (c.seg_stride < 1)
          return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: COO seg_stride must be >= 1", i);
      } else if (c.seg_format != RECOM_SEG_ROW_SPLITS_I32) {
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: unknown seg_format", i);
      }
      if (c.lookup == RECOM_LOOKUP_SEGMENT && c.combiner != RECOM_COMBINE_SUM &&
          c.combiner != RECOM_COMBINE_MEAN && c.combiner != RECOM_COMBINE_SQRTN)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: unknown combiner", i);
    } else {
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: unknown lookup kind", i);
    }
  }
  // columns that share a table must describe the same table
  {
    std::vector<int32_t> first(cfg.num_tables, -1);
    for (int32_t i = 0; i < n; ++i) {
      const int32_t f = first[cols[i].table];
      if (f < 0) {
        first[cols[i].table] = i;
      } else if (cols[f].table_rows != cols[i].table_rows || cols[f].dim != cols[i].dim ||
                 cols[f].table_dtype != cols[i].table_dtype) {
        return fail(RECOM_ERR_INVALID_ARGUMENT,
                    "columns %d and %d share table %d but declare different shapes", f, i,
                    cols[i].table);
      }
    }
  }
  // output slices must not overlap
  std::vector<std::pair<int32_t, int32_t>> spans(n);
  for (int32_t i = 0; i < n; ++i) spans[i] = {cols[i].out_offset, cols[i].dim};
  std::sort(spans.begin(), spans.end());
  for (int32_t i = 1; i < n; ++i)
    if (spans[i].first < spans[i - 1].first + spans[i - 1].second)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "output slices overlap at offset %d", spans[i].first);
  return RECOM_OK;
}

// ---- bucketize acceleration grid -------------------------------------------------------
// A boundary set is followed in the pool by [b0, inv_w, G, 0 | G words]. cell(v) =
// clamp(int((v - b0) * inv_w), 0, G - 1) is monotone in v, so the floats that fall in one cell
// form an interval; word c = lo | hi << 16 holds upper_bound() of the interval's smallest and
// largest float: for every float v of the cell the exact TF bucket lies in [lo, hi] (usually
// one candidate), and the kernel finishes with the same `value < boundary` comparisons as
// bucketize_op.cc:42-44 over that range. The host evaluates cell() with the same IEEE
// single-precision operations as the device (no fused multiply-add on either side).
uint32_t float_key(float f) {  // monotone in f (NaN excluded)
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
float key_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
int grid_cell(float v, float b0, float inv, int G) {
  volatile float d = v - b0;
  volatile float x = d * inv;
  const float xx = x;
  int c;
  if (xx != xx) c = 0;  // cvt.rzi.s32.f32 of NaN
  else if (xx >= 2147483648.f) c = 0x7fffffff;
  else if (xx <= -2147483648.f) c = (int)0x80000000;
  else c = (int)xx;
  return std::min(std::max(c, 0), G - 1);
}
// Appends header + grid of the set `b` to the pool.
void append_grid(std::vector<float>& pool, const std::vector<float>& b) {
  const int nb = (int)b.size();
  int G = 1;
  float b0 = 0.f, inv = 0.f;
  if (nb > 65535) {
    // lo | hi << 16 cannot describe such a set: header with cells = 0, the kernel falls back to
    // the plain upper_bound over the whole set (bucketize_grid in kernels.cu)
    pool.push_back(0.f);
    pool.push_back(0.f);
    pool.push_back(0.f);
    pool.push_back(0.f);
    return;
  }
  bool ok = nb >= 8;
  if (ok) {
    for (float x : b) ok = ok && std::isfinite(x);
    ok = ok && b[nb - 1] > b[0];
  }
  if (ok) {
    G = 16;
    while (G < 2 * nb && G < 1024) G *= 2;
    b0 = b[0];
    volatile float w = b[nb - 1] - b[0];
    volatile float q = (float)G / w;
    inv = q;
    ok = std::isfinite(inv) && inv > 0.f;
    if (!ok) {
      G = 1;
      b0 = inv = 0.f;
    }
  }
  auto ans = [&](float v) { return (uint32_t)(std::upper_bound(b.begin(), b.end(), v) - b.begin()); };
  const uint32_t kmin = float_key(-INFINITY), kmax = float_key(INFINITY);
  // first[c] = smallest key whose cell is >= c (kmax + 1 when there is none)
  std::vector<uint64_t> first(G + 1);
  for (int c = 0; c <= G; ++c) {
    if (c == 0) {
      first[c] = kmin;
      continue;
    }
    if (c == G || grid_cell(INFINITY, b0, inv, G) < c) {
      first[c] = (uint64_t)kmax + 1;
      continue;
    }
    uint64_t lo = kmin, hi = kmax;  // invariant: cell(hi) >= c
    while (lo < hi) {
      const uint64_t mid = (lo + hi) / 2;
      if (grid_cell(key_float((uint32_t)mid), b0, inv, G) >= c) hi = mid;
      else lo = mid + 1;
    }
    first[c] = lo;
  }
  auto as_float = [](uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  pool.push_back(b0);
  pool.push_back(inv);
  pool.push_back(as_float((uint32_t)G));
  pool.push_back(0.f);
  for (int c = 0; c < G; ++c) {
    uint32_t lo = 0, hi = 0;
    if (first[c] < first[c + 1]) {
      lo = ans(key_float((uint32_t)first[c]));
      hi = ans(key_float((uint32_t)(first[c + 1] - 1)));
    }
    pool.push_back(as_float(lo | (hi << 16)));
  }
  while (pool.size() % 4) pool.push_back(0.f);
}

// Build DevColumn/DevSlab/DevChunk arrays from the descriptors.
void build_layout(recom_plan* p) {
  const int n = (int)p->descs.size();
  const int es = p->elem_bytes;
  std::map<std::vector<float>, int32_t> pool_index;  // dedup identical boundary sets
  p->cols.resize(n);
  p->split_of_col.assign(n, -1);
  for (int i = 0; i < n; ++i) {
    const recom_column_desc& d = p->descs[i];
    DevColumn& c = p->cols[i];
    memset(&c, 0, sizeof c);
    c.table_rows = d.table_rows;
    c.num_buckets = d.num_buckets;
    if (d.num_buckets > 1) {
      const uint64_t nb = (uint64_t)d.num_buckets;
      c.bucket_magic = ~0ull / nb + ((nb & (nb - 1)) == 0 ? 1 : 0);  // floor(2^64 / nb)
    }
    c.table = d.table;
    c.dim = d.dim;
    c.row_bytes = d.dim * es;
    c.transform = d.transform;
    c.value_dtype = d.value_dtype;
    c.value_slot = d.value_slot;
    c.bytes_slot = d.bytes_slot;
    c.seg_slot = d.seg_slot;
    c.seg_format = d.seg_format;
    c.seg_stride = d.seg_stride;
    c.lookup = d.lookup;
    // GATHER_SCATTER == SEGMENT(sum) with <=1 id per row (lookup_optimizer.cc:324-440)
    c.combiner = d.lookup == RECOM_LOOKUP_SEGMENT ? d.combiner : RECOM_COMBINE_SUM;
    c.sparse_index = -1;
    c.coo_index = -1;
    if (d.lookup != RECOM_LOOKUP_DENSE) {
      c.sparse_index = (int32_t)p->sparse_cols.size();
      p->sparse_cols.push_back(i);
      if (d.seg_format == RECOM_SEG_COO_I64 || d.seg_format == RECOM_SEG_COO_RESHAPE_I64 ||
          d.seg_format == RECOM_SEG_FIXED) {
        DevCoo dc{};
        dc.seg_slot = d.seg_slot;
        dc.seg_stride = d.seg_stride;
        if (d.seg_format == RECOM_SEG_COO_RESHAPE_I64) {
          dc.kind = 1;
          dc.seg_stride = d.seg_stride & 255;
          dc.has_new_shape = (d.seg_stride & 256) ? 1 : 0;
        } else if (d.seg_format == RECOM_SEG_FIXED) {
          dc.kind = 2;
          dc.value_slot = d.value_slot;
          dc.value_str = d.value_dtype == RECOM_STR ? 1 : 0;
          // the units below read "multi-hot" off seg_slot >= 0; the row_splits come from coo_index
          c.seg_slot = 0;
        }
        int32_t k = 0;
        for (; k < (int32_t)p->coos.size(); ++k)
          if (memcmp(&p->coos[k], &dc, sizeof dc) == 0) break;
        if (k == (int32_t)p->coos.size()) p->coos.push_back(dc);
        c.coo_index = k;
      } else if (d.seg_format == RECOM_SEG_SPLIT_STR) {
        DevSplit sp{c.bucket_magic, (int32_t)d.num_buckets, d.value_slot, d.bytes_slot, d.seg_stride};
        int32_t k = 0;
        for (; k < (int32_t)p->splits.size(); ++k)
          if (memcmp(&p->splits[k], &sp, sizeof sp) == 0) break;
        if (k == (int32_t)p->splits.size()) p->splits.push_back(sp);
        p->split_of_col[i] = k;
      }
    }
    if (d.transform == RECOM_XFORM_BUCKETIZE) {
      std::vector<float> b(p->boundaries_in.begin() + d.boundaries_offset,
                           p->boundaries_in.begin() + d.boundaries_offset + d.num_boundaries);
      auto it = pool_index.find(b);
      if (it == pool_index.end()) {
        // every set starts 16-byte aligned and is padded to 16 bytes (TMA bulk copies)
        it = pool_index.emplace(b, (int32_t)p->boundary_pool.size()).first;
        p->boundary_pool.insert(p->boundary_pool.end(), b.begin(), b.end());
        while (p->boundary_pool.size() % 4) p->boundary_pool.push_back(0.f);
        append_grid(p->boundary_pool, b);
        p->set_words[it->second] = (int32_t)p->boundary_pool.size() - it->second;
      }
      c.boundaries_offset = it->second;
      c.num_boundaries = d.num_boundaries;
    }
  }
  while (p->boundary_pool.size() < 4) p->boundary_pool.push_back(0.f);
  // the row_splits of split string inputs are built in the workspace, behind the COO inputs'
  for (int i = 0; i < n; ++i)
    if (p->split_of_col[i] >= 0) p->cols[i].coo_index = (int32_t)p->coos.size() + p->split_of_col[i];

  // slabs: walk columns in concat order, group same-class neighbours up to 32 chunks
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    return p->descs[a].out_offset < p->descs[b].out_offset;
  });
  const bool row_aligned = ((int64_t)p->cfg.out_width * es) % 16 == 0;
  auto column_class = [&](const recom_column_desc& d) {
    if (d.lookup != RECOM_LOOKUP_DENSE) return (int)kSlabSparse;
    const int64_t bytes = d.table_rows * d.dim * es;
    const bool stage = p->stage_tables && d.table_rows <= kStageTableRows &&
                       bytes <= kStageTableBytes && bytes % 16 == 0;
    return stage ? (int)kSlabDenseStaged : (int)kSlabDense;
  };
  auto open_slab = [&](int cls, int vec) {
    DevSlab s{};
    s.cls = cls;
    s.vec_bytes = vec;
    s.first_chunk = (int32_t)p->chunks.size();
    s.first_col = (int32_t)p->slab_cols.size();
    s.first_bnd = (int32_t)p->slab_bnds.size();
    s.first_bnd_pool = -1;
    s.tile_rows = kDenseTileRows;
    p->slabs.push_back(s);
  };
  // smem offset of a table already staged by the open slab, or -1
  auto staged_offset = [&](const DevSlab& s, int table) {
    for (int k = 0; k < s.num_cols; ++k) {
      const DevSlabCol& sc = p->slab_cols[s.first_col + k];
      if (sc.stage_off >= 0 && p->descs[sc.col].table == table) return sc.stage_off;
    }
    return -1;
  };
  int prev_end = -1;
  for (int oi = 0; oi < n; ++oi) {
    const int i = order[oi];
    const recom_column_desc& d = p->descs[i];
    const int cls = column_class(d);
    const int64_t table_bytes = d.table_rows * d.dim * es;
    const bool vec16 = row_aligned && (d.dim * es) % 16 == 0 && ((int64_t)d.out_offset * es) % 16 == 0;
    const int vec = vec16 ? 16 : es;
    const int nchunks = d.dim * es / vec;
    int done = 0;
    while (done < nchunks) {
      bool need_new = p->slabs.empty();
      if (!need_new) {
        const DevSlab& s = p->slabs.back();
        need_new = s.cls != cls || s.vec_bytes != vec || s.num_chunks >= kMaxChunks ||
                   (done == 0 && d.out_offset != prev_end) ||
                   // keep a column inside one slab when it fits in an empty one
                   (done == 0 && nchunks <= kMaxChunks && s.num_chunks + nchunks > kMaxChunks) ||
                   (cls == kSlabDenseStaged && staged_offset(s, d.table) < 0 &&
                    s.stage_bytes + table_bytes > kStageSlabBytes);
      }
      if (need_new) open_slab(cls, vec);
      DevSlab& s = p->slabs.back();
      const int take = std::min(nchunks - done, kMaxChunks - s.num_chunks);
      const int col_local = s.num_cols;
      DevSlabCol sc{};
      sc.col = i;
      sc.stage_off = -1;
      sc.stage_copy = 0;
      sc.bnd_off = -1;
      {
        const DevColumn& c = p->cols[i];
        sc.transform = c.transform;
        sc.value_dtype = c.value_dtype;
        sc.value_slot = c.value_slot;
        sc.bytes_slot = c.bytes_slot;
        sc.table = c.table;
        sc.row_bytes = c.row_bytes;
        sc.table_rows = (int32_t)c.table_rows;
        sc.num_boundaries = c.num_boundaries;
        sc.boundaries_offset = c.boundaries_offset;
        sc.num_buckets = (int32_t)c.num_buckets;
        sc.bucket_magic = c.bucket_magic;
        sc.seg_slot = c.seg_slot;
        sc.coo_index = c.coo_index;
        sc.sparse_index = c.sparse_index;
        sc.combiner = c.combiner;
      }
      if (cls == kSlabDenseStaged) {
        const int have = staged_offset(s, d.table);
        if (have >= 0) {
          sc.stage_off = have;
        } else {
          sc.stage_off = s.stage_bytes;
          sc.stage_copy = 1;
          s.stage_bytes += (int32_t)table_bytes;
        }
      }
      if (cls != kSlabSparse && d.transform == RECOM_XFORM_BUCKETIZE && d.num_boundaries > 0) {
        const int32_t pool_off = p->cols[i].boundaries_offset;
        const int32_t floats = p->set_words.at(pool_off);  // boundaries + header + grid
        for (int k = 0; k < s.num_bnds; ++k)
          if (p->slab_bnds[s.first_bnd + k].pool_off == pool_off)
            sc.bnd_off = p->slab_bnds[s.first_bnd + k].smem_off;
        if (sc.bnd_off < 0 && s.bnd_bytes / 4 + floats <= kStageBndFloats) {
          DevSlabBnd sb{};
          sb.pool_off = pool_off;
          sb.floats = floats;
          sb.smem_off = s.bnd_bytes / 4;
          p->slab_bnds.push_back(sb);
          if (s.num_bnds == 0) s.first_bnd_pool = pool_off;
          s.num_bnds++;
          s.bnd_bytes += floats * 4;
          sc.bnd_off = sb.smem_off;
        }
      }
      p->slab_cols.push_back(sc);
      s.num_cols++;
      for (int k = 0; k < take; ++k) {
        DevChunk ch;
        ch.col = i;
        ch.col_local = col_local;
        ch.table_byte = (done + k) * vec;
        ch.out_byte = d.out_offset * es + (done + k) * vec;
        p->chunks.push_back(ch);
      }
      s.num_chunks += take;
      done += take;
    }
    prev_end = d.out_offset + d.dim;
  }
  // per-slab derived fields: rows of a warp tile, fast groups
  for (size_t si = 0; si < p->slabs.size(); ++si) {
    DevSlab& sl = p->slabs[si];
    if (sl.cls == kSlabSparse) {
      sl.tile_rows = kColTileRowsSparse;
      continue;
    }
    sl.tile_rows = kDenseTileRows;
    for (int g = 0; g * 8 + 8 <= sl.num_cols; ++g) {
      bool fast = true;
      for (int k = 0; k < 8; ++k) {
        const DevSlabCol& sc = p->slab_cols[sl.first_col + g * 8 + k];
        fast = fast && sc.transform == RECOM_XFORM_BUCKETIZE && sc.value_dtype == RECOM_F32 &&
               sc.bnd_off >= 0 && sc.num_boundaries > 0;  // (float values: no filter applies)
      }
      if (fast) {
        sl.fast8_mask |= 1u << g;
        bool uni = true;
        for (int k = 1; k < 8; ++k)
          uni = uni && p->slab_cols[sl.first_col + g * 8 + k].bnd_off ==
                           p->slab_cols[sl.first_col + g * 8].bnd_off &&
                p->slab_cols[sl.first_col + g * 8 + k].num_boundaries ==
                    p->slab_cols[sl.first_col + g * 8].num_boundaries;
        if (uni) sl.uni8_mask |= 1u << g;
      }
    }
    if (p->shard_world <= 1 && sl.cls == kSlabDenseStaged && sl.num_cols >= 8 && sl.num_cols % 8 == 0 &&
        sl.fast8_mask == (uint32_t)((1ull << (sl.num_cols / 8)) - 1ull)) {
      sl.all_fast = 1;  // a FAST slab: scheduled statically, everything else goes to the queue
      if (sl.num_cols >= 16) sl.tile_rows = kFastTileRows;
    } else {
      sl.tile_rows = kDenseTileRows;
    }
  }

  // units of the queue: every distinct (input, transform, segments) triple outside the FAST slabs
  // is one unit - transformed once per tile, whatever number of columns (`users`) read it.
  {
    struct Key {
      int32_t v[11];
      bool operator<(const Key& o) const { return memcmp(v, o.v, sizeof v) < 0; }
    };
    std::map<Key, int32_t> index;
    std::vector<std::vector<ColUser>> unit_users;
    for (const DevSlab& sl : p->slabs) {
      if (sl.all_fast && sl.cls == kSlabDenseStaged) continue;
      for (int k = 0; k < sl.num_cols; ++k) {
        const DevSlabCol& sc = p->slab_cols[sl.first_col + k];
        const recom_column_desc& d = p->descs[sc.col];
        const DevColumn& c = p->cols[sc.col];
        const bool sparse = d.lookup != RECOM_LOOKUP_DENSE;
        const bool bkt = d.transform == RECOM_XFORM_BUCKETIZE;
        const bool hashed = d.transform == RECOM_XFORM_HASH_STR || d.transform == RECOM_XFORM_HASH_INT;
        const int32_t split = p->split_of_col[sc.col];  // >= 0: ids = hashed tokens in the workspace
        Key key{};
        key.v[0] = split >= 0 ? -1 - split : d.value_slot;
        key.v[1] = d.value_dtype == RECOM_STR ? d.bytes_slot : -1;
        key.v[2] = d.transform;
        key.v[3] = d.value_dtype;
        key.v[4] = hashed ? (int32_t)d.num_buckets : 0;
        key.v[5] = bkt ? c.boundaries_offset : -1;
        key.v[6] = bkt ? d.num_boundaries : 0;
        key.v[7] = sparse ? 1 : 0;
        key.v[8] = sparse ? d.seg_slot : -1;
        key.v[9] = sparse ? c.coo_index : -1;
        // the stream's filter: distinct filters are numbered by content
        int32_t fidx = -1;
        if (!p->filters_in.empty() && p->filters_in[sc.col].kind != RECOM_FILTER_NONE) {
          const recom_column_filter& f = p->filters_in[sc.col];
          for (size_t q = 0; q < p->dev_filters.size() && fidx < 0; ++q)
            if (memcmp(&p->dev_filters[q], &f, sizeof f) == 0) fidx = (int32_t)q;
          if (fidx < 0) {
            DevFilter df;
            memcpy(&df, &f, sizeof f);
            fidx = (int32_t)p->dev_filters.size();
            p->dev_filters.push_back(df);
          }
        }
        key.v[10] = fidx;
        auto it = index.find(key);
        if (it == index.end()) {
          ColItem u{};
          u.bucket_magic = c.bucket_magic;
          u.num_buckets = (int32_t)d.num_buckets;
          u.transform = d.transform;
          u.value_dtype = d.value_dtype;
          u.value_slot = d.value_slot;
          u.bytes_slot = d.value_dtype == RECOM_STR ? d.bytes_slot : -1;
          u.seg_slot = !sparse ? -1 : d.seg_format == RECOM_SEG_FIXED ? 0 : d.seg_slot;  // (FIXED: row_splits from coo_index)
          u.coo_index = sparse ? c.coo_index : -1;
          u.boundaries_offset = c.boundaries_offset;
          u.num_boundaries = c.num_boundaries;
          u.filter = fidx;
          if (split >= 0) {
            // the split kernel has hashed the tokens: this stream reads int32 ids from the workspace
            u.transform = RECOM_XFORM_NONE;
            u.value_dtype = RECOM_I32;
            u.value_slot = -1 - split;
            u.bytes_slot = -1;
            u.seg_slot = 0;  // (multi-hot; its row_splits come from coo_index)
            u.num_buckets = 0;
          }
          it = index.emplace(key, (int32_t)p->units.size()).first;
          p->units.push_back(u);
          unit_users.emplace_back();
        }
        ColUser cu{};
        cu.table = d.table;
        cu.row_bytes = d.dim * es;
        cu.table_rows = (int32_t)d.table_rows;
        cu.out_byte = d.out_offset * es;
        cu.combiner = p->shard_world > 1 ? (int32_t)RECOM_COMBINE_SUM : d.combiner;  // Mode B: partial sums
        cu.col = sc.col;
        cu.vec_bytes = sl.vec_bytes;
        cu.hbm = d.table_rows * (int64_t)d.dim * es > kHbmTableBytes ? 1 : 0;
        unit_users[it->second].push_back(cu);
      }
    }
    p->unit_cost.assign(p->units.size(), 0.0);
    p->unit_hbm.assign(p->units.size(), 0);
    for (size_t ui = 0; ui < p->units.size(); ++ui) {
      ColItem& u = p->units[ui];
      u.first_user = (int32_t)p->users.size();
      u.num_users = (int32_t)unit_users[ui].size();
      // estimated cost of a full tile (arbitrary units): the transform plus the gathers of every user
      double xf;
      switch (u.transform) {
        case RECOM_XFORM_HASH_STR: xf = 3.0; break;
        case RECOM_XFORM_HASH_INT: xf = 2.5; break;
        case RECOM_XFORM_BUCKETIZE: xf = 1.5; break;
        default: xf = 0.5; break;
      }
      double g = 0;
      for (const ColUser& cu : unit_users[ui]) {
        if (&cu == &unit_users[ui][0]) u.u0 = cu;
        g += (cu.hbm ? 2.0 : 1.0) * (1.0 + cu.row_bytes / 64.0);
        p->unit_hbm[ui] = p->unit_hbm[ui] || cu.hbm;
        p->users.push_back(cu);
      }
      p->unit_cost[ui] = (u.seg_slot >= 0 ? 1.6 : 1.0) * (xf + g);
    }
  }

  // shared memory layout: [staged tables | staged boundaries | FAST contexts | warp scratch]
  // Only FAST runs stage tables and boundaries; a FAST tile keeps its row ids in the warp's
  // scratch, a column tile its row_splits and row ids.
  int max_stage = 0, max_bnd = 0, max_cols = 1, need = kWarpScratchBytes;
  for (const DevSlab& sl : p->slabs) {
    if (sl.all_fast && sl.cls == kSlabDenseStaged) {
      max_stage = std::max(max_stage, sl.stage_bytes);
      max_bnd = std::max(max_bnd, sl.bnd_bytes);
      max_cols = std::max(max_cols, sl.num_cols);
      // row ids of a tile + two buffers of its raw values (the next tile's are copied ahead)
      need = std::max(need, sl.num_cols * kIdPitch * 4 + 2 * sl.num_cols * sl.tile_rows * 4);
    }
  }
  p->dev.smem_off_bnd = (int32_t)align_up((size_t)max_stage, 128);
  p->dev.smem_off_ctx = p->dev.smem_off_bnd + (int32_t)align_up((size_t)max_bnd, 128);
  p->dev.smem_off_scratch =
      p->dev.smem_off_ctx + (int32_t)align_up((size_t)max_cols * sizeof(ColCtx), 128);
  p->dev.scratch_bytes = (int32_t)align_up((size_t)need, 128);
  p->dev.smem_bytes = p->dev.smem_off_scratch + kWarps * p->dev.scratch_bytes;
}

// Schedule of a batch size (Sched in plan.h): fully static.
//   * FAST slabs (fast bucketize groups only, tables staged in shared memory): the CTAs are
//     apportioned to them by cost, each CTA owns ONE run (a row range of one slab, so the staged
//     tables and column contexts are paid once per CTA); with more FAST slabs than CTAs, whole
//     slabs go to the least loaded CTA instead. Inside a run the warps take 16-row tiles from a
//     shared-memory counter.
//   * Everything else is cut into column tiles (one unit = one id stream, kColTileRowsDense /
//     kColTileRowsSparse rows; fewer for small batches, until there is about one tile per
//     resident warp), sorted by descending cost and dealt to the global warps round robin.
// Flat or fused? (RECOM_B200_FLAT unset.) Measured on B200 at batch 2048: with the column tiles
// in their own one-warp-per-tile kernel, mixed1000 (no FAST slab) runs in 123 instead of 155 us,
// production2000 (column tiles 5.3x the FAST cost) in 102 instead of 111, the 20 multi-hot
// columns of model E alone in 16 instead of 25; models E and F (1.8x / 1.2x) lose 4-5 us, the
// 980 one-id columns of E (0.5x) 9 us: two launches do not overlap the way the fused kernel's
// warps do. Flat when the column tiles weigh at least three times the FAST slabs.
bool prefer_flat(const recom_plan* p, int64_t B, int dens) {
  if (p->flat_mode >= 0) return p->flat_mode != 0;
  double fast_cost = 0, col_cost = 0;
  for (const DevSlab& sl : p->slabs)
    if (sl.cls == kSlabDenseStaged && sl.all_fast)
      fast_cost += (double)((B + sl.tile_rows - 1) / sl.tile_rows) * 2.0 * (sl.tile_rows * sl.num_chunks * sl.vec_bytes) / 8192.0;
  for (size_t ui = 0; ui < p->units.size(); ++ui)
    // (a request of density class c holds ~2^c times the ids per multi-hot row the cost model assumes)
    col_cost += p->unit_cost[ui] * (double)B /
                (p->units[ui].seg_slot >= 0 ? (double)kColTileRowsSparse / (double)(1 << dens) : (double)kColTileRowsDense);
  return col_cost > 0 && col_cost >= 3.0 * fast_cost;
}

// Density class of a request: multi-hot tiles are cut by rows, but their cost goes with their
// ids - with long rows (config 3: zipf lengths, mean 60 ids per row) a 32-row tile holds
// thousands of ids and the launch lasts as long as its fullest tile. Class c halves the rows per
// multi-hot tile c times (32 -> 16 -> 8 -> 4) so that a tile holds ~256 ids on average. The
// schedule cache is keyed by (batch, schedule, class).
int density_class(const recom_plan* p, const recom_request* req) {
  if (req->batch <= 0) return 0;
  int64_t worst = 0;
  for (const ColItem& u : p->units) {
    if (u.seg_slot < 0 || u.value_slot < 0) continue;
    const int64_t nnz = req->slot_elems[u.value_slot] - (u.value_dtype == RECOM_STR ? 1 : 0);
    worst = std::max(worst, nnz / req->batch);
  }
  int c = 0;
  while (c < 3 && worst >= (16ll << c)) ++c;
  return c;
}

std::shared_ptr<TileTable> build_tiles(const recom_plan* p, int64_t B, bool flat, int dens) {
  const int cap_ctas = std::max(1, p->num_sms * p->ctas_per_sm);
  auto tt = std::make_shared<TileTable>();
  auto make = [&](int s, int64_t b, int64_t rows) {
    Tile t{};
    t.slab_id = s;
    t.row_begin = (int32_t)b;
    t.rows = (int32_t)rows;
    t.slab = p->slabs[s];
    t.stage = t.slab.cls == kSlabDenseStaged && rows >= kMinStageRows;
    t.slab.run_tile_rows = (int16_t)t.slab.tile_rows;
    t.slab.run_tiles = (int32_t)((rows + t.slab.tile_rows - 1) / t.slab.tile_rows);
    return t;
  };
  std::vector<int> fast;
  for (int s = 0; s < (int)p->slabs.size(); ++s) {
    const DevSlab& sl = p->slabs[s];
    if (sl.cls == kSlabDenseStaged && sl.all_fast) fast.push_back(s);
  }
  auto tiles_of = [&](int s) { return (B + p->slabs[s].tile_rows - 1) / p->slabs[s].tile_rows; };

  // ---- column tiles ----
  // rows per tile: the fewer tiles than warps, the finer (a launch lasts as long as its longest
  // chain of dependent loads, and a tile's chain does not get shorter with fewer rows - but the
  // transform and gather passes behind it do)
  const int64_t warps = !flat ? (int64_t)cap_ctas * kWarps
                         : p->flat_warps_target > 0 ? p->flat_warps_target
                                                    : (int64_t)p->num_sms * p->flat_ctas_per_sm * 4;
  int rows_dense = std::min(kColTileRowsDense, p->col_rows_dense);
  int rows_sparse = std::max(4, std::min(kColTileRowsSparse, p->col_rows_sparse) >> dens);
  {
    auto count = [&](int rd, int rs) {
      int64_t n = 0;
      for (const ColItem& u : p->units) n += u.seg_slot >= 0 ? (B + rs - 1) / rs : (B + rd - 1) / rd;
      return n;
    };
    while (count(rows_dense, rows_sparse) * 2 <= warps && (rows_dense > 32 || rows_sparse > 4)) {
      if (rows_dense > 32) rows_dense /= 2;
      if (rows_sparse > 4) rows_sparse /= 2;
    }
  }
  struct Costed {
    double cost;
    ColItem it;
  };
  std::vector<Costed> qs;
  double queue_cost = 0, flat_cost = 0;
  bool long_queue;
  {
    int64_t n = 0;
    for (const ColItem& u : p->units)
      n += u.seg_slot >= 0 ? (B + rows_sparse - 1) / rows_sparse : (B + rows_dense - 1) / rows_dense;
    long_queue = n > warps;
  }
  for (size_t ui = 0; ui < p->units.size(); ++ui) {
    const ColItem& u = p->units[ui];
    const int64_t full = u.seg_slot >= 0 ? kColTileRowsSparse : kColTileRowsDense;
    int64_t step = u.seg_slot >= 0 ? rows_sparse : rows_dense;
    // rows gathered from HBM-resident tables: a quarter of the rows, four times the tiles in
    // flight - unless the queue is long anyway (every tile pays its chain of dependent loads)
    if (p->unit_hbm[ui] && !long_queue) step = std::max<int64_t>(u.seg_slot >= 0 ? 4 : 32, step / 4);
    for (int64_t b = 0; b < B; b += step) {
      ColItem it = u;
      it.begin = (int32_t)b;
      it.count = (int32_t)std::min<int64_t>(step, B - b);
      const double c = p->unit_cost[ui] * (0.5 + 0.5 * (double)it.count / (double)full);
      qs.push_back({c, it});
      queue_cost += c;
    }
  }
  std::stable_sort(qs.begin(), qs.end(), [](const Costed& a, const Costed& b) { return a.cost > b.cost; });
  if (flat) {  // every column tile gets its own warp of the flat kernel; the persistent one keeps the FAST runs
    for (const Costed& c : qs) tt->flat.push_back(c.it);
    qs.clear();
    flat_cost = queue_cost;
    queue_cost = 0;
  }

  // ---- FAST runs ----
  std::vector<double> cost(fast.size());
  double fast_cost = 0;
  int64_t fast_tiles = 0;
  for (size_t i = 0; i < fast.size(); ++i) {
    // a FAST tile costs 2 units per 8 KB it stores (16 rows x 512 bytes)
    const DevSlab& sl = p->slabs[fast[i]];
    cost[i] = (double)tiles_of(fast[i]) * 2.0 * (sl.tile_rows * sl.num_chunks * sl.vec_bytes) / 8192.0;
    fast_cost += cost[i];
    fast_tiles += tiles_of(fast[i]);
  }
  // grid: enough CTAs for one warp per item, in whole waves of the SMs
  int64_t units = std::max<int64_t>(1, ((int64_t)qs.size() + fast_tiles + kWarps - 1) / kWarps);
  units = std::max<int64_t>(units, std::min<int64_t>((int64_t)fast.size(), cap_ctas));
  if (units > p->num_sms) units = (units + p->num_sms - 1) / p->num_sms * p->num_sms;
  const int grid = (int)std::min<int64_t>(cap_ctas, units);
  tt->grid = fast.empty() && qs.empty() ? 0 : grid;  // 0: no persistent launch (flat tiles only)

  // Exclusive roles (RECOM_B200_EXCLUSIVE): the FAST runs go to the first `nf` CTAs only, the
  // column tiles to the others, in proportion to the estimated cost of the two kinds of work.
  int nf = grid;
  if (p->exclusive && !fast.empty() && !qs.empty() && grid >= 4) {
    nf = (int)(grid * fast_cost / (fast_cost + queue_cost * p->col_weight) + 0.5);
    nf = std::max(1, std::min(grid - 1, nf));
    tt->fast_ctas = nf;
  }
  std::vector<std::vector<Tile>> per_cta(grid);
  if (!fast.empty() && (size_t)nf >= fast.size()) {
    std::vector<int> n_ctas(fast.size(), 1);
    int given = (int)fast.size();
    std::vector<std::pair<double, int>> rem;
    for (size_t i = 0; i < fast.size(); ++i) {
      const double share = (nf - (double)fast.size()) * cost[i] / fast_cost;
      n_ctas[i] += (int)share;
      given += (int)share;
      rem.push_back({-(share - (int)share), (int)i});
    }
    std::sort(rem.begin(), rem.end());
    for (int k = 0; given < nf && k < (int)rem.size(); ++k, ++given) n_ctas[rem[k].second]++;
    int c = 0;
    for (size_t i = 0; i < fast.size(); ++i) {
      const int64_t T = tiles_of(fast[i]), TR = p->slabs[fast[i]].tile_rows;
      const int64_t n = std::min<int64_t>(n_ctas[i], T);
      for (int64_t j = 0; j < n; ++j) {
        const int64_t t0 = j * T / n, t1 = (j + 1) * T / n;
        const int64_t r0 = t0 * TR, r1 = std::min<int64_t>(B, t1 * TR);
        if (r1 > r0) per_cta[c % nf].push_back(make(fast[i], r0, r1 - r0));
        ++c;
      }
    }
  } else if (!fast.empty()) {
    std::vector<int> idx(fast.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(nf, 0.0);
    for (int i : idx) {
      const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      per_cta[c].push_back(make(fast[i], 0, B));
      load[c] += cost[i];
    }
  }
  tt->cta.assign(grid, CtaWork{0, 0, 0, 0});
  // The cost-sorted column tiles go to the least loaded CTA that holds column tiles (all of
  // them, or those behind the FAST CTAs); a CTA's load starts with its FAST runs (55 slabs over
  // 296 CTAs: some CTAs own a fifth more rows than others), column tiles weighted as in the
  // split of the warps below.
  const int first_c = tt->fast_ctas >= 0 ? tt->fast_ctas : 0;
  std::vector<std::vector<int>> q_of(grid);
  {
    using Load = std::pair<double, int>;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int c = first_c; c < grid; ++c) {
      double l = 0;
      for (const Tile& t : per_cta[c]) {
        const DevSlab& sl = t.slab;
        l += (double)t.slab.run_tiles * 2.0 * (sl.tile_rows * sl.num_chunks * sl.vec_bytes) / 8192.0;
      }
      heap.push({l, c});
    }
    for (size_t i = 0; i < qs.size(); ++i) {
      Load top = heap.top();
      heap.pop();
      q_of[top.second].push_back((int)i);
      top.first += 3.0 * qs[i].cost;
      heap.push(top);
    }
  }
  for (const Costed& c : qs) tt->dense_tiles += c.it.seg_slot < 0 ? 1 : 0;
  for (int c = 0; c < grid; ++c) {
    tt->cta[c].first = (int32_t)tt->runs.size();
    tt->cta[c].count = (int32_t)per_cta[c].size();
    for (Tile& t : per_cta[c]) tt->runs.push_back(t);
    tt->cta[c].q_first = (int32_t)tt->queue.size();
    for (int i : q_of[c]) tt->queue.push_back(qs[i].it);
    tt->cta[c].q_count = (int32_t)q_of[c].size();
  }
  // warps that start with the FAST tiles: in proportion to the estimated cost of the two kinds
  if (fast.empty()) tt->fast_warps = 0;
  else if (qs.empty()) tt->fast_warps = kWarps;
  else
    // (measured on models E / F: a column tile costs a warp about three times its estimate
    // relative to a FAST tile - chains of dependent loads against a store stream)
    tt->fast_warps = std::max(1, std::min(kWarps - 1, (int)(kWarps * fast_cost / (fast_cost + 3.0 * queue_cost) + 0.5)));
  if (getenv("RECOM_B200_DEBUG_SCHED"))
    fprintf(stderr, "[recom sched] batch %lld flat %d: fast slabs %zu tiles %lld cost %.0f | column tiles %zu cost %.0f | grid %d\n",
            (long long)B, (int)flat, fast.size(), (long long)fast_tiles, fast_cost, flat ? tt->flat.size() : qs.size(),
            flat ? flat_cost : queue_cost, tt->grid);
  if (p->force_fast_warps >= 0 && !fast.empty() && !qs.empty())
    tt->fast_warps = std::max(1, std::min(kWarps - 1, p->force_fast_warps));
  return tt;
}

constexpr size_t kMaxCachedTileTables = 512;

// Host -> device copy that has fully landed when this returns, whatever stream the caller will
// launch on (see recom_plan::up_stream).
cudaError_t blocking_upload(recom_plan* p, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, p->up_stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(p->up_stream);
}

// Cached (device resident) schedule of a batch size, or nullptr when the cache is full:
// the caller then ships the schedule inside the request's argument block.
std::shared_ptr<const TileTable> cached_tiles(recom_plan* p, int64_t B, bool flat, int dens, recom_status* st) {
  *st = RECOM_OK;
  std::lock_guard<std::mutex> lock(p->tiles_mu);
  const int64_t key = (B * 2 + (flat ? 1 : 0)) * 4 + dens;
  auto it = p->tiles_by_batch.find(key);
  if (it != p->tiles_by_batch.end()) return it->second;
  if (p->tiles_by_batch.size() >= kMaxCachedTileTables) return nullptr;
  std::shared_ptr<TileTable> tt = build_tiles(p, B, flat, dens);
  std::vector<uint8_t> host(tt->bytes() + 64, 0);
  tt->write(host.data());
  // first use of a batch size (warm-up): a blocking allocation + copy
  if (cudaMalloc(&tt->d_block, host.size()) != cudaSuccess ||
      blocking_upload(p, tt->d_block, host.data(), host.size()) != cudaSuccess) {
    *st = fail(RECOM_ERR_CUDA, "tile table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  p->tiles_by_batch[key] = tt;
  return tt;
}

recom_status upload_plan(recom_plan* p) {
  struct Part {
    const void* src;
    size_t bytes;
    size_t off;
  };
  std::vector<int32_t> out_offsets;
  for (const recom_column_desc& d : p->descs) out_offsets.push_back(d.out_offset);
  Part parts[11] = {
      {p->cols.data(), p->cols.size() * sizeof(DevColumn), 0},
      {p->slabs.data(), p->slabs.size() * sizeof(DevSlab), 0},
      {p->chunks.data(), p->chunks.size() * sizeof(DevChunk), 0},
      {p->slab_cols.data(), p->slab_cols.size() * sizeof(DevSlabCol), 0},
      {p->slab_bnds.data(), p->slab_bnds.size() * sizeof(DevSlabBnd), 0},
      {p->boundary_pool.data(), p->boundary_pool.size() * sizeof(float), 0},
      {p->users.data(), p->users.size() * sizeof(ColUser), 0},
      {p->coos.data(), p->coos.size() * sizeof(DevCoo), 0},
      {out_offsets.data(), out_offsets.size() * sizeof(int32_t), 0},
      {p->dev_filters.data(), p->dev_filters.size() * sizeof(DevFilter), 0},
      {p->splits.data(), p->splits.size() * sizeof(DevSplit), 0},
  };
  size_t total = 0;
  for (Part& pt : parts) {
    pt.off = total;
    total += align_up(pt.bytes + 16, 256);
  }
  std::vector<uint8_t> host(total, 0);
  for (const Part& pt : parts)
    if (pt.bytes) memcpy(host.data() + pt.off, pt.src, pt.bytes);
  CUDA_TRY(cudaMalloc(&p->d_block, total));
  CUDA_TRY(cudaMemcpy(p->d_block, host.data(), total, cudaMemcpyHostToDevice));
  uint8_t* base = (uint8_t*)p->d_block;
  p->dev.cols = (const DevColumn*)(base + parts[0].off);
  p->dev.slabs = (const DevSlab*)(base + parts[1].off);
  p->dev.chunks = (const DevChunk*)(base + parts[2].off);
  p->dev.slab_cols = (const DevSlabCol*)(base + parts[3].off);
  p->dev.slab_bnds = (const DevSlabBnd*)(base + parts[4].off);
  p->dev.boundaries = (const float*)(base + parts[5].off);
  p->dev.users = (const ColUser*)(base + parts[6].off);
  p->dev.coos = (const DevCoo*)(base + parts[7].off);
  p->d_out_offsets = (const int32_t*)(base + parts[8].off);
  p->dev.filters = (const DevFilter*)(base + parts[9].off);
  p->dev.splits = (const DevSplit*)(base + parts[10].off);
  p->dev.num_splits = (int32_t)p->splits.size();
  p->dev.num_coo = (int32_t)p->coos.size();
  p->dev.num_cols = (int32_t)p->cols.size();
  p->dev.num_slabs = (int32_t)p->slabs.size();
  p->dev.elem_bytes = p->elem_bytes;
  p->dev.check_ids = p->cfg.check_ids;
  p->dev.shard_world = p->shard_world;
  p->dev.shard_rank = p->shard_rank;
  return RECOM_OK;
}

// Bytes of the blob the plan's columns reach (slots are 16-byte padded): what the kernel may
// prefetch into L2.
int64_t blob_extent(const recom_plan* p, const recom_request* req) {
  int64_t end = 0;
  for (const recom_column_desc& d : p->descs) {
    end = std::max(end, req->slot_offsets[d.value_slot] +
                            req->slot_elems[d.value_slot] * dtype_bytes(d.value_dtype));
    if (d.value_dtype == RECOM_STR && d.bytes_slot >= 0)
      end = std::max(end, req->slot_offsets[d.bytes_slot] + req->slot_elems[d.bytes_slot]);
    if (d.lookup != RECOM_LOOKUP_DENSE && d.seg_slot >= 0) {
      const bool coo = d.seg_format == RECOM_SEG_COO_I64 || d.seg_format == RECOM_SEG_COO_RESHAPE_I64;
      end = std::max(end, req->slot_offsets[d.seg_slot] + req->slot_elems[d.seg_slot] * (coo ? 8 : 4));
      if (d.seg_format == RECOM_SEG_COO_RESHAPE_I64)
        for (int k = 1; k <= ((d.seg_stride & 256) ? 2 : 1); ++k)
          end = std::max(end, req->slot_offsets[d.seg_slot + k] + req->slot_elems[d.seg_slot + k] * 8);
    }
  }
  return end;
}

int64_t column_num_ids(const recom_plan* p, int ci, const recom_request* req) {
  const recom_column_desc& d = p->descs[ci];
  if (d.lookup == RECOM_LOOKUP_DENSE) return req->batch;
  if (d.seg_format == RECOM_SEG_SPLIT_STR) return 0;  // tokens are counted on the device
  const int64_t e = req->slot_elems[d.value_slot];
  return d.value_dtype == RECOM_STR ? e - 1 : e;
}

// O(1) checks, every call.
recom_status validate_request_head(const recom_plan* p, const recom_request* req) {
  if (!req->slot_offsets || !req->slot_elems || !req->tables)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "request host arrays must not be null");
  if (req->batch < 0 || req->batch > 0x7fffffffLL)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "batch out of range");
  if (req->batch > 0 && (!req->out || (!req->blob && p->cfg.num_slots > 0)))
    return fail(RECOM_ERR_INVALID_ARGUMENT, "blob/out device pointers must not be null");
  if (((uintptr_t)req->blob & 15) || ((uintptr_t)req->out & 15) || ((uintptr_t)req->workspace & 255))
    return fail(RECOM_ERR_INVALID_ARGUMENT,
                "blob/out must be 16-byte aligned and workspace 256-byte aligned");
  return RECOM_OK;
}

// O(columns) checks of the request's shape (slot offsets / element counts). The table pointers
// are checked when a table set is first seen (device_tables).
recom_status validate_request(const recom_plan* p, const recom_request* req) {
  const int n = (int)p->descs.size();
  for (int i = 0; i < n; ++i) {
    const recom_column_desc& d = p->descs[i];
    const int64_t voff = req->slot_offsets[d.value_slot];
    const int valign = dtype_bytes(d.value_dtype);
    if (voff < 0 || voff % valign)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: value slot offset %lld not %d-byte aligned",
                  i, (long long)voff, valign);
    const int64_t velems = req->slot_elems[d.value_slot];
    if (d.lookup == RECOM_LOOKUP_DENSE) {
      const int64_t want = d.value_dtype == RECOM_STR ? req->batch + 1 : req->batch;
      if (velems != want)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: value slot has %lld elements, want %lld",
                    i, (long long)velems, (long long)want);
    } else if (d.seg_format == RECOM_SEG_SPLIT_STR) {
      if (velems != req->batch + 1)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: a split string input holds batch+1 offsets (has %lld)",
                    i, (long long)velems);
    } else if (d.seg_format == RECOM_SEG_FIXED) {
      const int64_t nnz = d.value_dtype == RECOM_STR ? velems - 1 : velems;
      const int64_t T = d.seg_stride > 0 ? d.seg_stride : (req->batch > 0 ? nnz / req->batch : 0);
      if (nnz < 0 || nnz != T * req->batch)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: a [batch, T] input holds %lld values for batch %lld, T %lld",
                    i, (long long)nnz, (long long)req->batch, (long long)T);
    } else {
      const int64_t nnz = d.value_dtype == RECOM_STR ? velems - 1 : velems;
      if (nnz < 0) return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: empty STR offsets slot", i);
      const int64_t soff = req->slot_offsets[d.seg_slot];
      const int64_t selems = req->slot_elems[d.seg_slot];
      if (d.seg_format == RECOM_SEG_ROW_SPLITS_I32) {
        if (soff % 4 || selems != req->batch + 1)
          return fail(RECOM_ERR_INVALID_ARGUMENT,
                      "column %d: row_splits slot must hold batch+1 int32 (has %lld)", i,
                      (long long)selems);
      } else {
        const int64_t stride = d.seg_format == RECOM_SEG_COO_RESHAPE_I64 ? (d.seg_stride & 255) : d.seg_stride;
        if (soff % 8 || selems != nnz * stride)
          return fail(RECOM_ERR_INVALID_ARGUMENT,
                      "column %d: COO slot must hold nnz*stride int64 (has %lld, nnz %lld)", i,
                      (long long)selems, (long long)nnz);
        if (d.seg_format == RECOM_SEG_COO_RESHAPE_I64) {
          // dense_shape (int64[rank]) and, when given, the new shape (int64[1..8]) behind the indices
          if (req->slot_offsets[d.seg_slot + 1] % 8 || req->slot_elems[d.seg_slot + 1] != stride)
            return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: slot %d must hold the int64[%lld] dense_shape",
                        i, d.seg_slot + 1, (long long)stride);
          if ((d.seg_stride & 256) && (req->slot_offsets[d.seg_slot + 2] % 8 || req->slot_elems[d.seg_slot + 2] < 1 ||
                                       req->slot_elems[d.seg_slot + 2] > 8))
            return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: slot %d must hold the int64[1..8] new shape",
                        i, d.seg_slot + 2);
        }
      }
    }
  }
  return RECOM_OK;
}

}  // namespace

extern "C" {

const char* recom_last_error(void) { return g_last_error.c_str(); }
int recom_abi_version(void) { return RECOM_B200_ABI_VERSION; }
int recom_last_launch_count(void) { return g_last_launches; }

recom_status recom_plan_create(const recom_column_desc* cols, int32_t num_cols,
                               const float* boundaries, int64_t num_boundaries,
                               const recom_plan_config* cfg, recom_plan** out) {
  return recom_plan_create_filtered(cols, num_cols, boundaries, num_boundaries, nullptr, cfg, out);
}

recom_status recom_plan_create_filtered(const recom_column_desc* cols, int32_t num_cols,
                                        const float* boundaries, int64_t num_boundaries,
                                        const recom_column_filter* filters,
                                        const recom_plan_config* cfg, recom_plan** out) {
  if (!cols || num_cols <= 0 || !cfg || !out)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "null or empty argument");
  if (num_boundaries < 0 || (num_boundaries > 0 && !boundaries))
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad boundaries array");
  recom_status st = validate_columns(cols, num_cols, *cfg, num_boundaries);
  if (st != RECOM_OK) return st;
  for (int32_t i = 0; filters && i < num_cols; ++i) {
    const recom_column_filter& f = filters[i];
    if (f.kind == RECOM_FILTER_NONE) continue;
    if (f.kind == RECOM_FILTER_KEEP || f.kind == RECOM_FILTER_SELECT) {
      if (f.num_intervals < 1 || f.num_intervals > RECOM_MAX_INTERVALS)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: filter needs 1..%d intervals", i, RECOM_MAX_INTERVALS);
      if (cols[i].value_dtype != RECOM_I32 && cols[i].value_dtype != RECOM_I64)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: KEEP / SELECT filters need I32 or I64 values", i);
    } else if (f.kind == RECOM_FILTER_STR_NE) {
      if (cols[i].value_dtype != RECOM_STR || f.not_equal_len < 0 || f.not_equal_len > 16)
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: STR_NE filter needs STR values and a constant of <= 16 bytes", i);
    } else {
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: unknown filter kind %d", i, f.kind);
    }
    if ((cfg->row_shard & 0xffff) > 1 && f.kind != RECOM_FILTER_SELECT)
      return fail(RECOM_ERR_UNSUPPORTED, "column %d: dropping filters and row-sharded tables (the counts "
                  "recom_plan_scale_rows divides by come from the row_splits)", i);
  }
  for (int32_t i = 0; i < num_cols; ++i) {
    if (cols[i].transform != RECOM_XFORM_BUCKETIZE) continue;
    const float* b = boundaries + cols[i].boundaries_offset;
    for (int32_t k = 1; k < cols[i].num_boundaries; ++k)
      if (b[k] < b[k - 1])  // bucketize_op.cc:57-59 "Expected sorted boundaries"
        return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: boundaries not sorted", i);
  }
  int n_dev = 0;
  CUDA_TRY(cudaGetDeviceCount(&n_dev));
  if (cfg->device < 0 || cfg->device >= n_dev)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "device %d not present (%d visible)", cfg->device, n_dev);
  int prev_dev = 0;
  CUDA_TRY(cudaGetDevice(&prev_dev));
  CUDA_TRY(cudaSetDevice(cfg->device));

  recom_plan* p = new (std::nothrow) recom_plan();
  if (!p) return fail(RECOM_ERR_OUT_OF_MEMORY, "host allocation failed");
  p->cfg = *cfg;
  // row_shard: world | rank << 16 (0: whole tables)
  p->shard_world = std::max(1, cfg->row_shard & 0xffff);
  p->shard_rank = (cfg->row_shard >> 16) & 0xffff;
  if (p->shard_rank >= p->shard_world) {
    delete p;
    cudaSetDevice(prev_dev);
    return fail(RECOM_ERR_INVALID_ARGUMENT, "row_shard: rank %d outside world %d", (cfg->row_shard >> 16) & 0xffff,
                cfg->row_shard & 0xffff);
  }
  p->descs.assign(cols, cols + num_cols);
  if (filters) p->filters_in.assign(filters, filters + num_cols);
  if (num_boundaries) p->boundaries_in.assign(boundaries, boundaries + num_boundaries);
  p->elem_bytes = cfg->out_dtype == RECOM_BF16 ? 2 : 4;
  // tuning knobs (the reference's counterpart: its RECOM_* env flags, fc_optimize_pass.cc:41-69)
  if (const char* e = getenv("RECOM_B200_STAGE_TABLES")) p->stage_tables = atoi(e) != 0;
  auto pow2_env = [](const char* name, int lo, int hi, int dflt) {
    const char* e = getenv(name);
    if (!e) return dflt;
    int v = atoi(e), r = lo;
    while (r * 2 <= v && r * 2 <= hi) r *= 2;
    return r;
  };
  if (const char* e = getenv("RECOM_B200_FAST_WARPS")) p->force_fast_warps = atoi(e);
  if (const char* e = getenv("RECOM_B200_EXCLUSIVE")) p->exclusive = atoi(e) != 0;
  if (const char* e = getenv("RECOM_B200_COL_WEIGHT")) p->col_weight = atof(e);
  p->col_rows_dense = pow2_env("RECOM_B200_COL_ROWS_DENSE", 32, kColTileRowsDense, kColTileRowsDense);
  p->col_rows_sparse = pow2_env("RECOM_B200_COL_ROWS_SPARSE", 4, kColTileRowsSparse, kColTileRowsSparse);
  if (const char* e = getenv("RECOM_B200_FLAT")) p->flat_mode = atoi(e) != 0;
  if (const char* e = getenv("RECOM_B200_FLAT_PDL")) p->flat_pdl = atoi(e) != 0;
  if (const char* e = getenv("RECOM_B200_REQUEST_PDL")) p->request_overlap = atoi(e) != 0;
  if (const char* e = getenv("RECOM_B200_FLAT_WARPS")) p->flat_warps_target = atoi(e);
  if (const char* e = getenv("RECOM_B200_DEFER_IDS")) p->defer_ids = std::max(32, atoi(e));
  if (const char* e = getenv("RECOM_B200_LONG_CTAS")) p->long_ctas_per_sm = std::max(1, std::min(8, atoi(e)));
  build_layout(p);
  st = upload_plan(p);
  if (st == RECOM_OK) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) p->num_sms = prop.multiProcessorCount;
    p->dev_flat = p->dev;
    p->dev_flat.scratch_bytes = kFlatScratchBytes;
    if (p->flat_mode != 0) {
      const int occ_flat = column_tiles_configure(p->elem_bytes == 2, 4 * kFlatScratchBytes);
      if (occ_flat < 1) st = fail(RECOM_ERR_CUDA, "flat kernel does not fit: %s", cudaGetErrorString(cudaGetLastError()));
      else p->flat_ctas_per_sm = occ_flat;
      long_rows_configure(p->elem_bytes == 2, kLongRowWarps * kFlatScratchBytes);
    }
    int occ = fused_kernel_configure(p->elem_bytes == 2, p->dev.smem_bytes);
    if (occ < 1) {
      st = fail(RECOM_ERR_CUDA, "fused kernel does not fit (smem %d bytes): %s", p->dev.smem_bytes,
                cudaGetErrorString(cudaGetLastError()));
      occ = 1;
    }
    p->ctas_per_sm = cfg->max_ctas_per_sm > 0 ? std::min(cfg->max_ctas_per_sm, occ) : occ;
    if (const char* e = getenv("RECOM_B200_PROF_DUMP")) {
      p->prof_path = e;
      if (cudaMalloc((void**)&p->d_prof, (size_t)p->num_sms * 4 * kWarps * 8 * 8) != cudaSuccess) p->d_prof = nullptr;
    }
    if (cudaStreamCreateWithFlags(&p->up_stream, cudaStreamNonBlocking) != cudaSuccess)
      st = fail(RECOM_ERR_CUDA, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    const size_t cbytes = (size_t)kCounterSlots * kCounterInts * sizeof(int32_t);
    if (cudaMalloc((void**)&p->d_counters, cbytes) != cudaSuccess ||
        cudaMemset(p->d_counters, 0, cbytes) != cudaSuccess ||
        cudaHostAlloc((void**)&p->h_done_seq, kCounterSlots * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&p->d_done_seq, p->h_done_seq, 0) != cudaSuccess)
      st = fail(RECOM_ERR_CUDA, "counter ring allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    else
      for (int i = 0; i < kCounterSlots; ++i) p->h_done_seq[i] = (uint32_t)(i - kCounterSlots);
    for (int i = 0; i < kRing && st == RECOM_OK; ++i) {
      cudaError_t e = cudaEventCreateWithFlags(&p->ring[i].done, cudaEventDisableTiming);
      if (e != cudaSuccess) st = fail(RECOM_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    }
  }
  cudaSetDevice(prev_dev);
  if (st != RECOM_OK) {
    recom_plan_destroy(p);
    return st;
  }
  *out = p;
  return RECOM_OK;
}

recom_status recom_plan_destroy(recom_plan* p) {
  if (!p) return RECOM_OK;
  if (p->d_prof && p->prof_grid > 0) {
    cudaDeviceSynchronize();
    std::vector<uint64_t> h((size_t)p->prof_grid * kWarps * 8);
    if (cudaMemcpy(h.data(), p->d_prof, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE* f = fopen(p->prof_path.c_str(), "w")) {
        for (size_t w = 0; w < (size_t)p->prof_grid * kWarps; ++w) {
          for (int i = 0; i < 8; ++i) fprintf(f, "%llu%c", (unsigned long long)h[w * 8 + i], i == 7 ? '\n' : ' ');
        }
        fclose(f);
      }
    }
  }
  if (p->d_prof) cudaFree(p->d_prof);
  for (int i = 0; i < kRing; ++i) {
    if (p->ring[i].done) {
      cudaEventSynchronize(p->ring[i].done);
      cudaEventDestroy(p->ring[i].done);
    }
    if (p->ring[i].pinned) cudaFreeHost(p->ring[i].pinned);
  }
  for (auto& e : p->table_sets)
    if (e->dev) cudaFree((void*)e->dev);
  if (p->d_counters) cudaFree(p->d_counters);
  if (p->h_done_seq) cudaFreeHost(p->h_done_seq);
  if (p->d_block) cudaFree(p->d_block);
  if (p->up_stream) cudaStreamDestroy(p->up_stream);
  delete p;
  return RECOM_OK;
}

// ---- serialisation: the payload that replaces the reference's `dlpath` attr ----
struct BlobHeader {
  char magic[8];  // "RECOMB2\0"
  int32_t version;
  int32_t num_cols;
  int64_t num_boundaries;
  recom_plan_config cfg;
  int32_t desc_bytes;
  int32_t has_filters;  // one recom_column_filter per column follows the boundaries
};

recom_status recom_plan_serialize(const recom_plan* p, void* buf, size_t buf_bytes, size_t* needed) {
  if (!p || !needed) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  const size_t total = sizeof(BlobHeader) + p->descs.size() * sizeof(recom_column_desc) +
                       p->boundaries_in.size() * sizeof(float) +
                       p->filters_in.size() * sizeof(recom_column_filter);
  *needed = total;
  if (!buf) return RECOM_OK;
  if (buf_bytes < total) return fail(RECOM_ERR_INVALID_ARGUMENT, "buffer too small");
  BlobHeader h{};
  memcpy(h.magic, "RECOMB2", 8);
  h.version = RECOM_B200_ABI_VERSION;
  h.num_cols = (int32_t)p->descs.size();
  h.num_boundaries = (int64_t)p->boundaries_in.size();
  h.cfg = p->cfg;
  h.desc_bytes = (int32_t)sizeof(recom_column_desc);
  h.has_filters = p->filters_in.empty() ? 0 : 1;
  uint8_t* o = (uint8_t*)buf;
  memcpy(o, &h, sizeof h);
  o += sizeof h;
  memcpy(o, p->descs.data(), p->descs.size() * sizeof(recom_column_desc));
  o += p->descs.size() * sizeof(recom_column_desc);
  if (!p->boundaries_in.empty())
    memcpy(o, p->boundaries_in.data(), p->boundaries_in.size() * sizeof(float));
  o += p->boundaries_in.size() * sizeof(float);
  if (!p->filters_in.empty())
    memcpy(o, p->filters_in.data(), p->filters_in.size() * sizeof(recom_column_filter));
  return RECOM_OK;
}

recom_status recom_plan_create_from_blob(const void* buf, size_t bytes, int32_t device,
                                         recom_plan** out) {
  if (!buf || bytes < sizeof(BlobHeader) || !out)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "blob too small");
  BlobHeader h;
  memcpy(&h, buf, sizeof h);
  if (memcmp(h.magic, "RECOMB2", 8) != 0 || h.version != RECOM_B200_ABI_VERSION ||
      h.desc_bytes != (int32_t)sizeof(recom_column_desc) || h.num_cols <= 0 || h.num_boundaries < 0)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "not a recom_b200 plan blob (or ABI mismatch)");
  const size_t want = sizeof(BlobHeader) + (size_t)h.num_cols * sizeof(recom_column_desc) +
                      (size_t)h.num_boundaries * sizeof(float) +
                      (h.has_filters ? (size_t)h.num_cols * sizeof(recom_column_filter) : 0);
  if (bytes < want) return fail(RECOM_ERR_INVALID_ARGUMENT, "plan blob truncated");
  std::vector<recom_column_desc> descs(h.num_cols);
  std::vector<float> bnd((size_t)h.num_boundaries);
  const uint8_t* o = (const uint8_t*)buf + sizeof h;
  memcpy(descs.data(), o, descs.size() * sizeof(recom_column_desc));
  o += descs.size() * sizeof(recom_column_desc);
  if (!bnd.empty()) memcpy(bnd.data(), o, bnd.size() * sizeof(float));
  o += bnd.size() * sizeof(float);
  std::vector<recom_column_filter> flt;
  if (h.has_filters) {
    flt.resize(h.num_cols);
    memcpy(flt.data(), o, flt.size() * sizeof(recom_column_filter));
  }
  recom_plan_config cfg = h.cfg;
  cfg.device = device;
  return recom_plan_create_filtered(descs.data(), h.num_cols, bnd.data(), h.num_boundaries,
                                    flt.empty() ? nullptr : flt.data(), &cfg, out);
}

recom_status recom_plan_num_columns(const recom_plan* p, int32_t* out) {
  if (!p || !out) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  *out = (int32_t)p->descs.size();
  return RECOM_OK;
}
recom_status recom_plan_out_width(const recom_plan* p, int32_t* out) {
  if (!p || !out) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  *out = p->cfg.out_width;
  return RECOM_OK;
}

}  // extern "C"
namespace {
// The plan's allocations, cached uploads and launches belong to cfg.device; the caller's current
// device may be another one (multi-GPU hosts). Switch for the duration of the call only.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess || prev == device) return e;
    switched = true;
    return cudaSetDevice(device);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
}  // namespace
extern "C" {

static size_t max_sched_bytes_of(const recom_plan* p, int64_t batch);

recom_status recom_plan_num_slots(const recom_plan* p, int32_t* out) {
  if (!p || !out) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  *out = p->cfg.num_slots;
  return RECOM_OK;
}

recom_status recom_plan_slot_layout(const recom_plan* p, int32_t* ranks, int32_t* elem_bytes) {
  if (!p || !ranks || !elem_bytes) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  for (int s = 0; s < p->cfg.num_slots; ++s) {
    ranks[s] = 1;
    elem_bytes[s] = 1;
  }
  for (const recom_column_desc& d : p->descs) {
    elem_bytes[d.value_slot] = dtype_bytes(d.value_dtype);
    if (d.value_dtype == RECOM_STR && d.bytes_slot >= 0) elem_bytes[d.bytes_slot] = 1;
    if (d.lookup != RECOM_LOOKUP_DENSE && d.seg_slot >= 0) {
      const bool coo = d.seg_format == RECOM_SEG_COO_I64 || d.seg_format == RECOM_SEG_COO_RESHAPE_I64;
      elem_bytes[d.seg_slot] = coo ? 8 : 4;
      ranks[d.seg_slot] = coo ? 2 : 1;  // SparseTensor.indices: [nnz, rank]
      if (d.seg_format == RECOM_SEG_COO_RESHAPE_I64)
        for (int k = 1; k <= ((d.seg_stride & 256) ? 2 : 1); ++k) elem_bytes[d.seg_slot + k] = 8;  // int64 shapes
    }
  }
  return RECOM_OK;
}

recom_status recom_plan_infer_batch(const recom_plan* p, const int64_t* slot_elems, int64_t* batch) {
  if (!p || !slot_elems || !batch) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  for (const recom_column_desc& d : p->descs) {
    if (d.lookup == RECOM_LOOKUP_DENSE) {
      *batch = slot_elems[d.value_slot] - (d.value_dtype == RECOM_STR ? 1 : 0);
      return RECOM_OK;
    }
    if (d.seg_format == RECOM_SEG_ROW_SPLITS_I32) {
      *batch = slot_elems[d.seg_slot] - 1;
      return RECOM_OK;
    }
    if (d.seg_format == RECOM_SEG_SPLIT_STR) {
      *batch = slot_elems[d.value_slot] - 1;
      return RECOM_OK;
    }
    if (d.seg_format == RECOM_SEG_FIXED && d.seg_stride > 0) {
      *batch = (slot_elems[d.value_slot] - (d.value_dtype == RECOM_STR ? 1 : 0)) / d.seg_stride;
      return RECOM_OK;
    }
  }
  return fail(RECOM_ERR_UNSUPPORTED,
              "no column carries the batch size (COO-only plans: pass it through the symbols input)");
}

recom_status recom_plan_output_shapes(const recom_plan* p, int64_t batch, int32_t* shapes) {
  if (!p || !shapes || batch < 0 || batch > 0x7fffffffLL)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  for (size_t i = 0; i < p->descs.size(); ++i) {
    shapes[2 * i] = (int32_t)batch;
    shapes[2 * i + 1] = p->descs[i].dim;
  }
  return RECOM_OK;
}

recom_status recom_plan_scale_rows(const recom_plan* p, const recom_request* req, int64_t row_begin,
                                   int64_t rows, void* out_rows) {
  if (!p || !req || !out_rows || row_begin < 0 || rows < 0 || row_begin + rows > req->batch)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  if (rows == 0) return RECOM_OK;
  if (req->args_offset < 0)
    return fail(RECOM_ERR_UNSUPPORTED, "recom_plan_scale_rows needs a blob with the argument trailer");
  DeviceGuard guard;
  CUDA_TRY(guard.enter(p->cfg.device));
  ReqParams A{};
  A.blob = (const uint8_t*)req->blob;
  A.slot_offsets = (const int64_t*)((const uint8_t*)req->blob + req->args_offset + 16);
  A.batch = req->batch;
  // row_splits built from COO inputs live where recom_plan_run of the same request put them
  size_t off = 0;
  {
    size_t b = 2 * align_up((size_t)p->cfg.num_slots * 8, 16);
    b += align_up((p->descs.size() + 1) * 4, 16);
    b += align_up((size_t)p->cfg.num_tables * 8, 16);
    off = align_up(align_up(b, 64) + max_sched_bytes_of(p, req->batch), 256);
  }
  A.ws_row_splits = req->workspace ? (int32_t*)((uint8_t*)req->workspace + off) : nullptr;
  if ((!p->coos.empty() || !p->splits.empty()) && !req->workspace)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "COO / split inputs: pass the workspace recom_plan_run used");
  launch_scale_rows(p->dev, p->d_out_offsets, A, row_begin, rows, p->cfg.out_width, out_rows,
                    p->elem_bytes == 2, (cudaStream_t)req->stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "scale_rows launch: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

recom_status recom_plan_output_ptrs(const recom_plan* p, void* out, int64_t* ptrs_dev,
                                    recom_stream_t stream) {
  if (!p || !out || !ptrs_dev) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard;
  CUDA_TRY(guard.enter(p->cfg.device));
  launch_output_ptrs((int32_t)p->descs.size(), p->d_out_offsets, p->elem_bytes, out, ptrs_dev,
                     (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "output_ptrs launch: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

// Bytes of the argument data staged per request when it does not travel inside the blob:
// slot offsets, slot element counts, (ids_out hook only) the per-column id prefix, and the table
// pointers when the plan's table-set cache is full.
static size_t arg_block_bytes(const recom_plan* p) {
  size_t b = 2 * align_up((size_t)p->cfg.num_slots * 8, 16);
  b += align_up((p->descs.size() + 1) * 4, 16);
  b += align_up((size_t)p->cfg.num_tables * 8, 16);
  return align_up(b, 64);
}

static size_t max_sched_bytes(const recom_plan* p, int64_t batch);
static size_t max_sched_bytes_of(const recom_plan* p, int64_t batch) { return max_sched_bytes(p, batch); }
static size_t max_sched_bytes(const recom_plan* p, int64_t batch) {
  // only used when the schedule cache is full and the schedule travels with the request
  const size_t ctas = (size_t)p->num_sms * p->ctas_per_sm;
  // column tiles: at least 4 rows (multi-hot) / 32 rows each, and never cut finer than that
  size_t items = 0;
  for (const ColItem& u : p->units) items += (size_t)(batch / (u.seg_slot >= 0 ? 4 : 32)) + 1;
  return items * sizeof(ColItem) + (ctas + p->slabs.size() + 4) * sizeof(Tile) +
         ctas * sizeof(CtaWork) + 256;
}

recom_status recom_plan_workspace_bytes(const recom_plan* p, int64_t batch, int64_t total_ids,
                                        size_t* out) {
  if (!p || !out || batch < 0 || total_ids < 0)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  // [argument block + schedule (upload path only) | row_splits built from COO / split inputs |
  //  hashed tokens of the split inputs: total_ids must cover their payload bytes + batch + 1 each]
  size_t b = align_up(arg_block_bytes(p) + max_sched_bytes(p, batch), 256);
  b += align_up((p->coos.size() + p->splits.size()) * (size_t)(batch + 1) * 4 + 16, 256);
  if (!p->splits.empty()) b += align_up((size_t)total_ids * 4 + 16, 256);
  // the list of deferred long rows: (item, row) pairs, at most one per kDeferIds ids of a unit
  b += align_up(((size_t)total_ids / std::min(kDeferIds, p->defer_ids) + p->units.size() + 2) * 8 + 16, 256);
  *out = b;
  return RECOM_OK;
}

namespace {
struct HostProf {
  bool on = getenv("RECOM_B200_HOST_PROF") != nullptr;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  long calls = 0;
  std::chrono::steady_clock::time_point last;
  void start() { if (on) last = std::chrono::steady_clock::now(); }
  void lap(int i) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::micro>(now - last).count();
    last = now;
  }
  ~HostProf() {
    if (on && calls)
      fprintf(stderr, "[recom_b200 host] %ld calls; us/call: head %.2f, schedule %.2f, validate %.2f, "
              "tables+args %.2f, launch %.2f\n", calls, acc[0] / calls, acc[1] / calls, acc[2] / calls,
              acc[3] / calls, acc[4] / calls);
  }
};
HostProf g_host_prof;


// Device copy of the request's table pointer array: the plan's cached copy when this very set
// of addresses was seen before (every request of a serving process), a new published copy
// otherwise; nullptr when the cache is full (the pointers then travel with the argument block).
const void* const* device_tables(recom_plan* p, const recom_request* req, recom_status* st) {
  *st = RECOM_OK;
  const size_t n = (size_t)p->cfg.num_tables;
  std::lock_guard<std::mutex> lock(p->tables_mu);
  for (const auto& e : p->table_sets)
    if (memcmp(e->host.data(), req->tables, n * sizeof(void*)) == 0) return e->dev;
  for (size_t t = 0; t < n; ++t)
    if (!req->tables[t] || ((uintptr_t)req->tables[t] & 15)) {
      *st = fail(RECOM_ERR_INVALID_ARGUMENT, "table %zu pointer is null or not 16-byte aligned", t);
      return nullptr;
    }
  if (p->table_sets.size() >= (size_t)kTableSets) return nullptr;
  std::unique_ptr<TableSet> e(new TableSet());
  e->host.assign(req->tables, req->tables + n);
  // first sight of a table set (warm-up): a blocking allocation + copy
  if (cudaMalloc((void**)&e->dev, std::max<size_t>(n, 1) * sizeof(void*)) != cudaSuccess ||
      blocking_upload(p, (void*)e->dev, e->host.data(), n * sizeof(void*)) != cudaSuccess) {
    if (e->dev) cudaFree((void*)e->dev);
    *st = fail(RECOM_ERR_CUDA, "table pointer upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  const void* const* dev = e->dev;
  p->table_sets.push_back(std::move(e));
  return dev;
}
}  // namespace

static recom_status run_impl(const recom_plan* cp, const recom_request* req, const recom_shard_out* so);

recom_status recom_plan_run(const recom_plan* cp, const recom_request* req) {
  return run_impl(cp, req, nullptr);
}

recom_status recom_plan_run_sharded(const recom_plan* cp, const recom_request* req,
                                    const recom_shard_out* so) {
  if (!so || !cp || !req) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  if (so->world < 1 || so->world > kMaxPeers || !so->peer_out)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "world must be 1..%d and peer_out non-null", kMaxPeers);
  if (so->rows_per_rank < 32 || so->rows_per_rank > 0x7fffffffLL ||
      req->batch > (int64_t)so->world * so->rows_per_rank)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "rows_per_rank must be >= 32 and cover the batch");
  if (so->col_offset < 0 || (int64_t)so->col_offset + cp->cfg.out_width > so->global_width)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "slice [%d, +%d) outside the global row of %d",
                so->col_offset, cp->cfg.out_width, so->global_width);
  const int es = cp->elem_bytes;
  // the plan's vector width was chosen for its own row pitch and offsets: the global row must
  // keep 16-byte alignment where the plan uses 128-bit chunks
  if (((int64_t)so->global_width * es) % 16 || ((int64_t)so->col_offset * es) % 16)
    return fail(RECOM_ERR_UNSUPPORTED, "global row pitch and slice offset must be 16-byte multiples");
  for (int g = 0; g < so->world; ++g)
    if (!so->peer_out[g] || ((uintptr_t)so->peer_out[g] & 15))
      return fail(RECOM_ERR_INVALID_ARGUMENT, "peer_out[%d] is null or not 16-byte aligned", g);
  if (so->rank < 0 || so->rank >= so->world)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "rank %d outside [0, %d)", so->rank, so->world);
  if (so->peer_flags)
    for (int g = 0; g < so->world; ++g)
      if (!so->peer_flags[g] || ((uintptr_t)so->peer_flags[g] & 3))
        return fail(RECOM_ERR_INVALID_ARGUMENT, "peer_flags[%d] is null or misaligned", g);
  return run_impl(cp, req, so);
}

static recom_status run_impl(const recom_plan* cp, const recom_request* req, const recom_shard_out* so) {
  g_last_launches = 0;
  g_host_prof.start();
  g_host_prof.calls++;
  if (!cp || !req) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  recom_plan* p = const_cast<recom_plan*>(cp);  // only internally locked caches mutate
  DeviceGuard guard;
  CUDA_TRY(guard.enter(p->cfg.device));
  recom_request head = *req;
  if (so) head.out = so->peer_out[0];
  recom_status st = validate_request_head(p, &head);
  if (st != RECOM_OK) return st;
  if (req->batch == 0) return validate_request(p, req);
  const int n = (int)p->descs.size();
  const int64_t B = req->batch;
  const int num_slots = p->cfg.num_slots;
  cudaStream_t stream = (cudaStream_t)req->stream;

  g_host_prof.lap(0);
  const int dens = density_class(p, req);
  const bool flat = prefer_flat(p, B, dens);
  std::shared_ptr<const TileTable> tt = cached_tiles(p, B, flat, dens, &st);
  if (st != RECOM_OK) return st;
  g_host_prof.lap(1);
  const bool sched_in_args = !tt;
  if (sched_in_args) tt = build_tiles(p, B, flat, dens);
  const int grid = tt->grid;

  st = validate_request(p, req);
  if (st != RECOM_OK) return st;
  // longest COO input (grid of the helper kernel)
  int64_t max_coo_nnz = 0;
  for (const DevCoo& c : p->coos)
    max_coo_nnz = std::max<int64_t>(max_coo_nnz, c.kind == 2 ? B : req->slot_elems[c.seg_slot] / c.seg_stride);
  g_host_prof.lap(2);

  // workspace: [argument block + schedule (upload path only) | ws_row_splits]
  size_t off = align_up(arg_block_bytes(p) + max_sched_bytes(p, B), 256);
  const size_t off_rs = off;
  off += align_up((p->coos.size() + p->splits.size()) * (size_t)(B + 1) * 4 + 16, 256);
  const size_t off_split = off;
  if (!p->splits.empty()) {
    if (req->ids_out)
      return fail(RECOM_ERR_UNSUPPORTED, "ids_out with string inputs split on the device (their token count "
                  "is not known on the host)");
    size_t ids = 0;
    for (const DevSplit& sp : p->splits) ids += (size_t)req->slot_elems[sp.bytes_slot] + (size_t)B + 1;
    off += align_up(ids * 4 + 16, 256);
  }
  // long rows of dense requests (density class >= 1) go to long_rows_kernel (flat schedule, plain output)
  const size_t off_long = off;
  int32_t long_cap = 0;
  const bool defer_long = dens >= 1 && flat && !so && !tt->flat.empty();
  if (defer_long) {
    int64_t cap = 0;
    for (const ColItem& u : p->units) {
      if (u.seg_slot < 0 || u.value_slot < 0) continue;
      cap += (req->slot_elems[u.value_slot] - (u.value_dtype == RECOM_STR ? 1 : 0)) / p->defer_ids + 1;
    }
    long_cap = (int32_t)std::min<int64_t>(cap, 0x3fffffff);
    off += align_up(((size_t)long_cap + 1) * 8 + 16, 256);
  }
  if (!req->workspace || req->workspace_bytes < off)
    return fail(RECOM_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes, got %zu", off,
                req->workspace_bytes);
  uint8_t* ws = (uint8_t*)req->workspace;

  const void* const* d_tables = device_tables(p, req, &st);
  if (st != RECOM_OK) return st;

  // Argument data. With the trailer recom_concat_inputs appends to the blob nothing is uploaded
  // here; otherwise (or for the ids_out hook, a full table cache, an uncached schedule) one
  // pinned staging slot -> one cudaMemcpyAsync into the workspace.
  const int64_t* d_slot_offsets = nullptr;
  const int64_t* d_slot_elems = nullptr;
  const int32_t* d_ids_base = nullptr;
  const bool trailer = req->args_offset >= 0;
  if (trailer) {
    if (req->args_offset % 16)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "args_offset must be a multiple of 16");
    const uint8_t* t = (const uint8_t*)req->blob + req->args_offset;
    d_slot_offsets = (const int64_t*)(t + 16);
    d_slot_elems = d_slot_offsets + num_slots;
  }
  Sched sched = sched_in_args ? Sched{} : tt->sched(tt->d_block);
  const ColItem* d_flat = sched_in_args ? nullptr : tt->flat_items(tt->d_block);
  ArgSlot* slot = nullptr;
  std::unique_lock<std::mutex> slot_lock;
  if (!trailer || req->ids_out || !d_tables || sched_in_args) {
    const size_t slots_bytes = align_up((size_t)num_slots * 8, 16);
    const size_t o_elems = slots_bytes, o_ids = 2 * slots_bytes;
    const size_t o_tables = o_ids + align_up(((size_t)n + 1) * 4, 16);
    const size_t args_bytes = arg_block_bytes(p);
    const size_t up_bytes = args_bytes + (sched_in_args ? tt->bytes() : 0);
    slot = &p->ring[p->ring_next.fetch_add(1) % kRing];
    slot_lock = std::unique_lock<std::mutex>(slot->mu);
    if (slot->in_flight) {
      CUDA_TRY(cudaEventSynchronize(slot->done));
      slot->in_flight = false;
    }
    if (slot->cap < up_bytes) {
      if (slot->pinned) cudaFreeHost(slot->pinned);
      slot->pinned = nullptr;
      slot->cap = 0;
      const size_t cap = align_up(up_bytes * 2, 4096);
      CUDA_TRY(cudaHostAlloc((void**)&slot->pinned, cap, cudaHostAllocDefault));
      slot->cap = cap;
    }
    uint8_t* h = slot->pinned;
    memcpy(h, req->slot_offsets, (size_t)num_slots * 8);
    memcpy(h + o_elems, req->slot_elems, (size_t)num_slots * 8);
    int32_t* ids_base = (int32_t*)(h + o_ids);
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
      ids_base[i] = (int32_t)total;
      total += column_num_ids(p, i, req);
      if (total > 0x7fffffffLL) return fail(RECOM_ERR_INVALID_ARGUMENT, "too many ids in one request");
    }
    ids_base[n] = (int32_t)total;
    memcpy(h + o_tables, req->tables, (size_t)p->cfg.num_tables * 8);
    if (sched_in_args) {
      tt->write(h + args_bytes);
      sched = tt->sched(ws + args_bytes);
      d_flat = tt->flat_items(ws + args_bytes);
    }
    CUDA_TRY(cudaMemcpyAsync(ws, h, up_bytes, cudaMemcpyHostToDevice, stream));
    if (!trailer) {
      d_slot_offsets = (const int64_t*)ws;
      d_slot_elems = (const int64_t*)(ws + o_elems);
    }
    d_ids_base = (const int32_t*)(ws + o_ids);
    if (!d_tables) d_tables = (const void* const*)(ws + o_tables);
  }
  g_host_prof.lap(3);

  // Sharded output only: a counter slot (finished warps -> the last one signals the peers) whose
  // previous user, kCounterSlots requests ago, has finished
  uint32_t seq = 0, cslot = 0;
  const bool need_slot = so && so->peer_flags;
  if (need_slot) {
    seq = p->counter_next.fetch_add(1);
    cslot = seq % kCounterSlots;
    if (*(volatile uint32_t*)(p->h_done_seq + cslot) != seq - (uint32_t)kCounterSlots) {
      const auto t0 = std::chrono::steady_clock::now();
      while (*(volatile uint32_t*)(p->h_done_seq + cslot) != seq - (uint32_t)kCounterSlots) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10))
          return fail(RECOM_ERR_CUDA, "no free request slot: %d sharded requests of this plan in flight for 10 s",
                      kCounterSlots);
      }
    }
  }

  ReqParams A{};
  A.blob = (const uint8_t*)req->blob;
  A.out = (uint8_t*)req->out;
  A.ids_out = req->ids_out;
  A.status = req->status_out;
  A.ws_row_splits = (int32_t*)(ws + off_rs);
  A.ws_split_ids = (int32_t*)(ws + off_split);
  A.slot_offsets = d_slot_offsets;
  A.slot_elems = d_slot_elems;
  A.tables = d_tables;
  A.ids_base = d_ids_base;
  A.counters = need_slot ? p->d_counters + (size_t)kCounterInts * cslot : nullptr;
  A.done_seq = need_slot ? p->d_done_seq + cslot : nullptr;
  A.seq = seq;
  {
    static const bool pf = getenv("RECOM_B200_NO_PREFETCH") == nullptr;
    A.blob_bytes = !pf ? 0 : req->args_offset >= 0
                       ? req->args_offset + 16 + 16 * (int64_t)num_slots : blob_extent(p, req);
  }
  A.prof = p->d_prof;
  if (p->d_prof) p->prof_grid = grid;
  A.batch = B;
  A.out_row_bytes = (int64_t)p->cfg.out_width * p->elem_bytes;
  A.world = 0;  // 0: plain output (A.out); >= 1: sharded output through peer_out
  A.rows_per_rank = 0x7fffffff;
  if (so) {
    A.world = so->world;
    A.rows_per_rank = so->rows_per_rank;
    A.out_row_bytes = (int64_t)so->global_width * p->elem_bytes;
    A.col_off_bytes = so->col_offset * p->elem_bytes;
    for (int g = 0; g < so->world; ++g) A.peer_out[g] = (uint8_t*)so->peer_out[g];
    for (int g = 0; g < kMaxPeers; ++g)
      A.peer_flags[g] = so->peer_flags && g < so->world ? (uint32_t*)so->peer_flags[g] : nullptr;
    A.epoch = so->epoch;
    A.rank = so->rank;
    A.wait_in_kernel = (so->flags & RECOM_SHARD_WAIT_IN_KERNEL) && so->peer_flags ? 1u : 0u;
    A.wait_timeout_ms = 10000u;
    A.out = (uint8_t*)so->peer_out[0];
  }

  if (!p->coos.empty()) {
    launch_coo_to_row_splits(p->dev, A, max_coo_nnz, stream);
    g_last_launches++;
  }
  if (!p->splits.empty()) {
    launch_split_hash(p->dev, A, stream);
    g_last_launches++;
  }
  if (defer_long) {
    A.long_list = (int32_t*)(ws + off_long);
    A.long_cap = long_cap;
    A.long_ids = p->defer_ids;
    CUDA_TRY(cudaMemsetAsync(A.long_list, 0, 8, stream));
  }
  // the COO helper was the last launch: what follows may start under it (programmatic dependent
  // launch), its tiles over COO inputs wait before they read the row_splits
  static const bool helper_pdl_on = getenv("RECOM_B200_HELPER_PDL") == nullptr || atoi(getenv("RECOM_B200_HELPER_PDL")) != 0;
  const bool helper_pdl = helper_pdl_on && !p->coos.empty() && p->splits.empty() && !defer_long;
  const bool has_flat = !tt->flat.empty();
  A.total_warps = grid * kWarps + (has_flat ? column_tiles_grid((int)tt->flat.size()) * 4 : 0);
  if (has_flat) {
    // profiling builds: RECOM_B200_PROF_FLAT=1 times the flat kernel's warps instead of the persistent one's
    static const bool prof_flat = getenv("RECOM_B200_PROF_FLAT") != nullptr;
    ReqParams F = A;
    F.prof = nullptr;
    if (p->d_prof && prof_flat) {
      const int cap_rows = p->num_sms * 4 * kWarps;
      F.prof = p->d_prof;
      F.pad0 = std::min<int>(cap_rows, (int)tt->flat.size());
      A.prof = nullptr;
      p->prof_grid = F.pad0 / kWarps;
    }
    launch_column_tiles(p->dev_flat, F, d_flat, (int)tt->flat.size(), p->elem_bytes == 2, stream, helper_pdl);
    g_last_launches++;
  }
  if (grid > 0) {
    // a plain fused request (one launch, no helper in front of it): its prologue may overlap the tail of
    // the previous request on the stream (recom_plan_set_request_overlap / RECOM_B200_REQUEST_PDL=0 turn it off)
    // (a small grid gains nothing: micro100 at batch 256 measured 12.6 -> 13.8 us with it)
    const bool req_pdl = p->request_overlap.load(std::memory_order_relaxed) && !has_flat && p->coos.empty() &&
                         p->splits.empty() && !so && !req->ids_out && !req->status_out && !p->cfg.check_ids &&
                         p->shard_world <= 1 && grid >= p->num_sms;
    // where the column tiles wait (kernels.cu, column_tile): at their entry when the one-id-per-row tiles
    // alone nearly fill the resident warps (model F: 3040 of them over 2960 warps), behind their transform
    // otherwise (model E: 1600)
    A.request_pdl = !req_pdl ? 0 : (int64_t)tt->dense_tiles * 5 > (int64_t)grid * kWarps * 4 ? 2 : 1;
    launch_fused_columns(p->dev, sched, A, grid, p->elem_bytes == 2, stream,
                         (has_flat && p->flat_pdl) || (!has_flat && helper_pdl) || req_pdl);
    g_last_launches++;
  }
  if (defer_long) {
    launch_long_rows(p->dev_flat, A, d_flat, p->num_sms * p->long_ctas_per_sm, p->elem_bytes == 2, stream);
    g_last_launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (slot) {
    cudaEventRecord(slot->done, stream);
    slot->in_flight = true;
  }
  g_host_prof.lap(4);
  if (e != cudaSuccess) {
    if (need_slot) *(volatile uint32_t*)(p->h_done_seq + cslot) = seq;  // the slot was not used
    return fail(RECOM_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  }
  return RECOM_OK;
}

recom_status recom_shard_wait(void* local_flags, int32_t world, uint32_t epoch,
                              uint32_t timeout_ms, void* stream) {
  if (!local_flags || world < 1 || world > kMaxPeers)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "local_flags null or world outside 1..%d", kMaxPeers);
  launch_shard_wait((uint32_t*)local_flags, world, epoch, timeout_ms ? timeout_ms : 10000u,
                    (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "shard wait launch: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

recom_status recom_ipc_alloc(size_t bytes, void** dev_ptr, uint8_t handle[64]) {
  if (!dev_ptr || !handle || bytes == 0) return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUDA_TRY(cudaMalloc(dev_ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) {
    cudaFree(*dev_ptr);
    *dev_ptr = nullptr;
    return fail(RECOM_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(handle, &h, 64);
  return RECOM_OK;
}

recom_status recom_ipc_open(const uint8_t handle[64], void** dev_ptr) {
  if (!dev_ptr || !handle) return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RECOM_OK;
}

recom_status recom_ipc_close(void* dev_ptr) {
  if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return RECOM_OK;
}

recom_status recom_ipc_free(void* dev_ptr) {
  if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
  return RECOM_OK;
}

recom_status recom_plan_algorithmic_bytes(const recom_plan* p, const recom_request* req,
                                          int64_t* bytes_out) {
  if (!p || !req || !bytes_out) return fail(RECOM_ERR_INVALID_ARGUMENT, "null argument");
  const int64_t B = req->batch;
  const int64_t s = p->elem_bytes;
  int64_t total = 0;
  for (size_t i = 0; i < p->descs.size(); ++i) {
    const recom_column_desc& d = p->descs[i];
    const int64_t nnz = column_num_ids(p, (int)i, req);
    int64_t in_bytes;
    if (d.seg_format == RECOM_SEG_SPLIT_STR && d.lookup != RECOM_LOOKUP_DENSE)
      in_bytes = 4 * (B + 1) + req->slot_elems[d.bytes_slot];  // (token count unknown here: no table term)
    else if (d.value_dtype == RECOM_STR)
      in_bytes = 4 * (nnz + 1) + req->slot_elems[d.bytes_slot];
    else
      in_bytes = (int64_t)dtype_bytes(d.value_dtype) * nnz;
    const int64_t seg = d.lookup == RECOM_LOOKUP_DENSE ? 0 : 4 * (B + 1);
    const int64_t tbl = (int64_t)d.dim * s * std::min<int64_t>(nnz, d.table_rows);
    const int64_t outb = (int64_t)d.dim * s * B;
    total += in_bytes + seg + tbl + outb;
  }
  *bytes_out = total;
  return RECOM_OK;
}

recom_status recom_plan_set_request_overlap(recom_plan* p, int32_t on) {
  if (!p) return fail(RECOM_ERR_INVALID_ARGUMENT, "null plan");
  p->request_overlap.store(on != 0, std::memory_order_relaxed);
  return RECOM_OK;
}

recom_status recom_dense_relu(const void* a_bf16, int64_t m, int32_t k, int64_t lda, const void* w_nk_bf16,
                              int32_t n, const float* bias, int32_t relu, float* out, recom_stream_t stream) {
  const char* why = "";
  const int rc = launch_dense_relu(a_bf16, m, k, lda, w_nk_bf16, n, bias, relu, out, (cudaStream_t)stream, &why);
  if (rc == 0) {
    g_last_launches = 1;
    return RECOM_OK;
  }
  return fail(rc == 1 ? RECOM_ERR_INVALID_ARGUMENT : RECOM_ERR_CUDA, "recom_dense_relu: %s", why);
}

recom_status recom_scatter_columns(const void* const* cols, const int32_t* col_dims,
                                   const int32_t* col_offsets, int32_t num_cols, int64_t batch,
                                   int32_t out_width, int32_t elem_bytes, void* out,
                                   recom_stream_t stream_) {
  if (num_cols == 0 || batch == 0) return RECOM_OK;
  if (!cols || !col_dims || !col_offsets || !out || num_cols < 0 || batch < 0 ||
      (elem_bytes != 4 && elem_bytes != 2))
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  for (int i = 0; i < num_cols; ++i)
    if (!cols[i] || col_dims[i] <= 0 || col_offsets[i] < 0 || col_offsets[i] + col_dims[i] > out_width)
      return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: bad pointer/slice", i);
  cudaStream_t stream = (cudaStream_t)stream_;
  // descriptors travel through a small stream-ordered device allocation
  const size_t bytes = (size_t)num_cols * (8 + 4 + 4);
  uint8_t* d = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&d, bytes, stream));
  std::vector<uint8_t> h(bytes);
  memcpy(h.data(), cols, (size_t)num_cols * 8);
  memcpy(h.data() + (size_t)num_cols * 8, col_dims, (size_t)num_cols * 4);
  memcpy(h.data() + (size_t)num_cols * 12, col_offsets, (size_t)num_cols * 4);
  CUDA_TRY(cudaMemcpyAsync(d, h.data(), bytes, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));  // h is pageable and about to die
  launch_scatter_columns((const void* const*)d, (const int32_t*)(d + (size_t)num_cols * 8),
                         (const int32_t*)(d + (size_t)num_cols * 12), num_cols, batch, out_width,
                         elem_bytes, out, stream);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d, stream);
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "scatter launch failed: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

// Addons>ConcatOutputs' host_inputs (concat_outputs_op_gpu.cu.cc:186-232): the columns live in
// host memory; they are uploaded into `d_scratch` (their bytes + 24 per column) and scattered.
recom_status recom_scatter_host_columns(const void* const* host_cols, const int32_t* col_dims,
                                        const int32_t* col_offsets, int32_t num_cols, int64_t batch,
                                        int32_t out_width, int32_t elem_bytes, void* out,
                                        void* d_scratch, recom_stream_t stream_) {
  if (num_cols == 0 || batch == 0) return RECOM_OK;
  if (!host_cols || !col_dims || !col_offsets || !out || !d_scratch || num_cols < 0 || batch < 0 ||
      (elem_bytes != 4 && elem_bytes != 2) || ((uintptr_t)d_scratch & 15))
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  std::vector<const void*> dev(num_cols);
  uint8_t* d = (uint8_t*)d_scratch;
  for (int i = 0; i < num_cols; ++i) {
    if (!host_cols[i] || col_dims[i] <= 0) return fail(RECOM_ERR_INVALID_ARGUMENT, "column %d: bad pointer/dim", i);
    const size_t bytes = (size_t)batch * col_dims[i] * elem_bytes;
    CUDA_TRY(cudaMemcpyAsync(d, host_cols[i], bytes, cudaMemcpyHostToDevice, stream));
    dev[i] = d;
    d += align_up(bytes, 16);
  }
  // (pageable sources: the copies above have been staged when cudaMemcpyAsync returns)
  return recom_scatter_columns(dev.data(), col_dims, col_offsets, num_cols, batch, out_width,
                               elem_bytes, out, stream_);
}

recom_status recom_segment_ids_to_row_splits(const int64_t* coo, int64_t nnz, int32_t stride,
                                             int64_t batch, int32_t* row_splits,
                                             recom_stream_t stream) {
  if (nnz < 0 || batch < 0 || stride < 1 || !row_splits || (nnz > 0 && !coo) || nnz > 0x7fffffffLL)
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  launch_segment_ids_to_row_splits(coo, nnz, stride, batch, row_splits, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

recom_status recom_string_to_hash_bucket_fast(const uint8_t* bytes, const int32_t* offsets,
                                              int64_t n, int64_t num_buckets, int64_t* out,
                                              recom_stream_t stream) {
  if (n == 0) return RECOM_OK;
  if (n < 0 || !offsets || !out || num_buckets < 1)  // op.cc:22-23 requires num_buckets >= 1
    return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  launch_fingerprint64(bytes, offsets, n, num_buckets, (uint64_t*)out, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

recom_status recom_fingerprint64(const uint8_t* bytes, const int32_t* offsets, int64_t n,
                                 uint64_t* out, recom_stream_t stream) {
  if (n == 0) return RECOM_OK;
  if (n < 0 || !offsets || !out) return fail(RECOM_ERR_INVALID_ARGUMENT, "bad argument");
  launch_fingerprint64(bytes, offsets, n, 0, out, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(RECOM_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
  return RECOM_OK;
}

}  // extern "C"
