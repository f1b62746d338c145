// plan.h — internal layout shared by the host planner and the device kernels.
//
// A plan is the B200-native counterpart of the CUDA file the reference generates
// per graph (tensorflow_addons/graph_optimizers/cuda_emitter.cc:2057-2137): the
// same information (per-column transform, lookup kind, table shape, concat
// offset), but as data consumed by ONE precompiled persistent kernel instead of
// as generated code with one 64-thread block per column.
#pragma once
#include <stdint.h>

#include "../../include/recom_b200.h"

namespace recom {

#ifndef RECOM_THREADS
#define RECOM_THREADS 320
#endif
constexpr int kThreads = RECOM_THREADS;  // threads per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kMaxChunks = 32;         // a slab is at most one warp wide
constexpr int kMaxPeers = 8;           // ranks of a fused Mode-A exchange (one NVSwitch box)
constexpr int kDenseTileRows = 32;     // rows of a DENSE warp tile (lane = row while ids are made)
constexpr int kFastTileRows = 16;      // ... of a slab of fast groups only: lane = (group parity, row)
constexpr int kHashTileRows = 8;       // ... of a slab with hashed columns: lane = (column group, row)
constexpr int kIdPitch = kDenseTileRows + 1;  // smem pitch of a column's ids (bank-conflict free)
constexpr int kMinStageRows = 96;      // a DENSE run shorter than this gathers from global memory
constexpr int kSparseRowsPerSub = 4;   // SPARSE warp tile: rows per lane group (32 / W groups)
constexpr int kSparseRowsPerSubHash = 2;  // ... when the ids are hashed on the fly
constexpr int kSparseIdsPerRow = 8;    // smem budget of a SPARSE warp tile, ids per (column, row)
constexpr int kWarpScratchBytes = 4 << 10;  // per warp: row_splits + row ids of a column tile
constexpr int kDeferIds = 256;              // rows longer than this go to long_rows_kernel (when enabled)
constexpr int kLongRowWarps = 8;            // warps of a long_rows_kernel CTA: a row's ids are dealt to them in 8 contiguous parts
constexpr int kFlatScratchBytes = 3 << 10;  // ... of the flat kernel: row_splits + 472 row ids + 1 KB
constexpr int kColTileRowsDense = 128;     // rows of a column tile with one id per row
constexpr int kColTileRowsSparse = 32;     // rows of a multi-hot column tile
constexpr int64_t kHbmTableBytes = 32ll << 20;  // a table beyond this is "HBM resident"
// shared-memory staging limits (per CTA)
constexpr int kStageTableRows = 256;       // only tables with few rows are staged
constexpr int kStageTableBytes = 16 << 10; // ... and at most this large
constexpr int kStageSlabBytes = 56 << 10;  // staged tables of one slab
constexpr int kStageBndFloats = 4096;      // staged boundary sets (with their grids) of one slab

enum SlabClass : int32_t {
  kSlabDenseStaged = 0,  // DENSE lookups, every table staged in shared memory
  kSlabDense = 1,        // DENSE lookups, rows gathered from global memory
  kSlabSparse = 2        // GATHER_SCATTER / SEGMENT lookups
};

// Per-column constants, device resident for the life of the plan.
struct DevColumn {
  int64_t table_rows;
  int64_t num_buckets;
  uint64_t bucket_magic;  // floor(2^64 / num_buckets): multiply-high modulo
  int32_t table;          // index into the request's table pointer array
  int32_t dim;
  int32_t row_bytes;      // dim * sizeof(table element)
  int32_t transform;
  int32_t value_dtype;
  int32_t value_slot;
  int32_t bytes_slot;
  int32_t seg_slot;
  int32_t seg_format;
  int32_t seg_stride;
  int32_t combiner;
  int32_t boundaries_offset;  // into the (deduplicated, 16-byte aligned) boundary pool
  int32_t num_boundaries;
  int32_t sparse_index;   // rank among non-DENSE columns, -1 for DENSE
  int32_t coo_index;      // rank among COO columns, -1 otherwise
  int32_t lookup;
};

// One 16-byte (or element-sized) piece of a slab's output span, owned by a lane.
struct DevChunk {
  int32_t col;         // global column index
  int32_t col_local;   // index of that column inside the slab
  int32_t table_byte;  // byte offset of this chunk inside a table row
  int32_t out_byte;    // byte offset of this chunk inside an output row
};

// A column's membership in a slab, with everything static the kernel needs about the
// column, so that a CTA entering a slab resolves a column with one read of this record
// plus one read of the request's slot / table arrays.
struct DevSlabCol {
  int32_t col;         // global column index
  int32_t stage_off;   // byte offset of its table in the smem table region, -1: global
  int32_t stage_copy;  // 1 if this entry issues the bulk copy (shared tables copy once)
  int32_t bnd_off;     // float offset of its boundaries in the smem region, -1: global
  int32_t transform;
  int32_t value_dtype;
  int32_t value_slot;
  int32_t bytes_slot;
  int32_t table;       // index into the request's table pointer array
  int32_t row_bytes;
  int32_t table_rows;  // < 2^31 (validated)
  int32_t num_boundaries;
  int32_t boundaries_offset;  // float offset in the boundary pool
  int32_t num_buckets;        // < 2^31 (validated)
  uint64_t bucket_magic;
  int32_t seg_slot;     // non-DENSE only
  int32_t coo_index;    // rank among COO columns, -1 otherwise
  int32_t sparse_index; // rank among non-DENSE columns, -1 for DENSE
  int32_t combiner;
  int32_t pad0, pad1;
};

// A column reading an id stream.
struct ColUser {
  int32_t table;       // index into the request's table pointer array
  int32_t row_bytes;
  int32_t table_rows;
  int32_t out_byte;    // byte offset of the column's slice inside an output row
  int32_t combiner;
  int32_t col;         // global column index (ids_out hook)
  int32_t vec_bytes;   // 16, or the element size when the slice is not 16-byte aligned
  int32_t hbm;         // 1: the table is far larger than L2: its rows are prefetched into L2 as
                       // soon as the tile's ids are known (one round trip instead of one per pass)
};
static_assert(sizeof(ColUser) == 32, "ColUser layout");

// An id the stream's filter dropped (recom_column_filter): not gathered, not counted.
constexpr int32_t kDroppedId = (int32_t)0x80000000;

// A stream's pre-lookup filter (recom_column_filter), device resident.
struct DevFilter {
  int32_t kind;
  int32_t num_intervals;
  int64_t lo[RECOM_MAX_INTERVALS];
  int64_t hi[RECOM_MAX_INTERVALS];
  int64_t substitute;
  int32_t not_equal_len;
  uint8_t not_equal[16];
  int32_t pad;
};
static_assert(sizeof(DevFilter) == sizeof(recom_column_filter), "DevFilter layout");

// A column tile of the queue: rows [begin, begin + count) of ONE id stream - a distinct (input,
// transform) pair - worked on by one warp: raw values -> row ids (hash / bucketize / none; every
// lane runs the same transform on consecutive ids: coalesced, no divergence) -> for every column
// that reads the stream (`users`: duplicates of an (input, transform) pair cost one transform,
// pre_lookup_optimizer.cc:596-654 removes the same redundancy on the graph): 128-bit row
// gathers, combine, concat row. The record is self-contained (the stream's static description
// travels with it): one load, then one read of the request's slot offsets.
struct ColItem {
  uint64_t bucket_magic;      // floor(2^64 / num_buckets)
  int32_t num_buckets;
  int32_t transform;
  int32_t value_dtype;
  int32_t value_slot;
  int32_t bytes_slot;         // STR payload, else -1
  int32_t seg_slot;           // row_splits / COO slot; -1: DENSE (one id per row)
  int32_t coo_index;          // rank among COO inputs (row_splits built in the workspace), -1
  int32_t boundaries_offset;  // float offset in the boundary pool
  int32_t num_boundaries;
  int32_t begin;              // first row
  int32_t count;              // rows
  int32_t first_user;         // into the ColUser array
  int32_t num_users;
  int32_t filter;             // index into the plan's filters, -1: none
  ColUser u0;                 // = users[first_user]: the common case of one column per stream
                              // needs no second record
};
static_assert(sizeof(ColItem) == 96, "ColItem layout");


// A COO input (SparseTensor.indices) whose row ids are turned into row_splits in the workspace.
// kind 0: row id = indices[e * seg_stride]; 1 (RECOM_SEG_COO_RESHAPE_I64): the indices of rank
// seg_stride are linearised with the dense_shape in slot seg_slot + 1 and divided by the trailing
// dimensions of the new shape (slot seg_slot + 2, or [prod(shape[:-1]), shape[-1]]); 2
// (RECOM_SEG_FIXED): no indices - row r owns ids [r * T, (r + 1) * T), T = seg_stride or
// elements of `value_slot` (minus one for strings) / batch.
struct DevCoo {
  int32_t seg_slot;
  int32_t seg_stride;
  int32_t kind;
  int32_t has_new_shape;
  int32_t value_slot;
  int32_t value_str;
};

// A string input that is split into tokens on the device (RECOM_SEG_SPLIT_STR): a helper kernel
// in front of the fused one writes its row_splits (ws_row_splits, behind the COO inputs') and the
// hashed tokens (int32 bucket ids, ws_split_ids) into the workspace.
struct DevSplit {
  uint64_t bucket_magic;
  int32_t num_buckets;
  int32_t value_slot;   // int32 offsets[batch + 1]
  int32_t bytes_slot;
  int32_t spec;         // RECOM_SPLIT_SPEC
};
static_assert(sizeof(DevSplit) == 24, "DevSplit layout");

// The same column resolved against one request; lives in shared memory while a CTA (DENSE
// run) or a warp (SPARSE tile) works on the slab.
struct ColCtx {
  const uint8_t* values;      // blob + slot_offsets[value_slot]
  const uint8_t* bytes;       // STR payload, else null
  const uint8_t* table;       // request table pointer
  const int32_t* row_splits;  // non-DENSE only
  uint64_t bucket_magic;
  int32_t num_buckets;
  int32_t table_rows;
  int32_t row_bytes;
  int32_t transform;
  int32_t value_dtype;
  int32_t num_boundaries;
  int32_t bnd_off;            // smem float offset, -1: read the pool in global memory
  int32_t bnd_pool_off;
  int32_t stage_off;          // smem byte offset of the staged table, -1: global
  int32_t combiner;
  int32_t col;                // global column index (ids_out hook)
  int32_t pad;
};
static_assert(sizeof(ColCtx) == 88, "ColCtx layout");

// A distinct boundary set staged by a slab.
struct DevSlabBnd {
  int32_t pool_off;    // float offset in the boundary pool (multiple of 4)
  int32_t floats;      // padded to a multiple of 4
  int32_t smem_off;    // float offset in the smem boundary region
  int32_t pad;
};

// A slab: a group of columns whose output slices are adjacent in the concat row
// and together at most kMaxChunks chunks wide. Work is tiled as (slab, row block).
struct DevSlab {
  int32_t cls;          // SlabClass
  int32_t num_chunks;   // W <= 32
  int32_t first_chunk;  // into the chunk array
  int32_t num_cols;
  int32_t first_col;    // into the slab_cols array
  int32_t vec_bytes;    // 16, or the element size when unaligned
  int32_t tile_rows;    // rows of a warp tile
  int32_t stage_bytes;  // total bytes of staged tables
  int32_t first_bnd;    // into the slab_bnds array
  int32_t num_bnds;
  int32_t bnd_bytes;    // total bytes of staged boundaries
  int32_t first_bnd_pool;  // pool offset of the first staged boundary set (-1: none)
  uint32_t fast8_mask;  // bit g: columns [8g, 8g+8) are float32 bucketize with staged boundaries
  uint32_t uni8_mask;   // ... and share one boundary set
  int16_t all_fast;     // every column belongs to a fast group
  // DENSE run only (filled per run by the scheduler)
  int16_t run_tile_rows;  // rows of this run's warp tiles (<= tile_rows)
  int32_t run_tiles;      // number of warp tiles of this run

};
static_assert(sizeof(DevSlab) == 64, "DevSlab layout");

struct DevPlan {
  const DevColumn* cols;
  const DevSlab* slabs;
  const DevChunk* chunks;
  const DevSlabCol* slab_cols;
  const DevSlabBnd* slab_bnds;
  const float* boundaries;
  const ColUser* users;
  const DevFilter* filters;
  const DevCoo* coos;
  const DevSplit* splits;
  int32_t num_coo;
  int32_t num_splits;
  int32_t num_cols;
  int32_t num_slabs;
  int32_t elem_bytes;   // 4 (fp32) or 2 (bf16): table and output element size
  int32_t check_ids;
  int32_t shard_world;  // Mode B: table rows sharded over this many ranks (1: whole tables)
  int32_t shard_rank;
  // dynamic shared memory layout (bytes from the start):
  // [staged tables | staged boundaries | FAST column contexts | kWarps x warp scratch]
  int32_t smem_off_bnd;
  int32_t smem_off_ctx;
  int32_t smem_off_scratch;
  int32_t scratch_bytes;  // per warp: row ids of a FAST tile, or row_splits + row ids of a column tile
  int32_t smem_bytes;
};

// A unit of work: a row range of a slab. For a DENSE run (a CTA's dense warps walk it in
// 32-row warp tiles) and for a SPARSE warp tile alike. The slab record travels with it so
// that no dependent lookup is needed.
struct Tile {
  int32_t slab_id;      // -1: none
  int32_t row_begin;
  int32_t rows;
  int32_t stage;        // DENSE run: 1 = stage the slab's small tables in shared memory
  DevSlab slab;
};
static_assert(sizeof(Tile) == 80, "Tile layout");

// The work of a CTA: FAST runs runs[first, first + count), walked in order, and column tiles
// queue[q_first, q_first + q_count), most expensive first.
struct CtaWork {
  int32_t first;
  int32_t count;
  int32_t q_first;
  int32_t q_count;
};

// The schedule of a batch size (cached on the device by the plan); a kernel parameter. Static
// per CTA: no global claim counters, no atomics on global memory, no fences inside a launch
// (measured: an atomic claim costs a warp 1-2 us, a gpu-scope release 2-6 us under load - as much
// as the work of a tile).
//   * Column tiles (the queue): everything that is not a FAST slab - hashed, multi-hot, unstaged.
//     The cost-sorted tiles are dealt to the CTAs round robin; a CTA's tiles are contiguous in
//     `queue` (CtaWork::q_first) and taken by its warps from a counter in shared memory.
//   * FAST runs (slabs made of fast bucketize groups, tables staged in shared memory by TMA): every
//     CTA owns the runs cta[c] (usually one: a row range of one slab); its warps take the run's
//     16-row tiles from another shared-memory counter.
//   * Warps [0, fast_warps) of a CTA start with the FAST tiles and go on with column tiles, the
//     others the other way round: the bandwidth-bound store stream of the FAST tiles and the
//     latency-bound chains of the column tiles overlap on every SM from the first microsecond,
//     and whoever runs out of its own kind of work takes the other.
constexpr int kMaxParamCtas = 320;  // CTAs whose work record travels in the kernel parameters

struct Sched {
  const Tile* runs;          // FAST runs, grouped by CTA
  const CtaWork* cta;        // [grid] (device copy: grids beyond kMaxParamCtas)
  const ColItem* queue;      // [num_queue] column tiles, grouped by CTA
  int32_t num_queue;
  int32_t fast_warps;
  int32_t fast_ctas;         // >= 0: exclusive roles - CTAs [0, fast_ctas) own all FAST runs and no
                             // column tiles, the others the column tiles and no FAST run; -1: mixed
  int32_t pad;
  // the CTAs' work records, in the parameter space (constant bank): a warp requests the record
  // of its first column tile with its very first global load, no dependent load in front of it
  CtaWork cta_param[kMaxParamCtas];
};

// Per-request values known on the host at launch; a kernel parameter (no dependent loads).
struct ReqParams {
  const uint8_t* blob;
  uint8_t* out;
  int32_t* ids_out;        // may be null
  int32_t* status;         // may be null
  int32_t* ws_row_splits;  // workspace: row_splits built from COO inputs and split string inputs,
                           // [num_coo + num_splits][batch + 1]
  int32_t* ws_split_ids;   // workspace: hashed tokens of the split inputs, input k at split_ids_offset(k)
  // arrays of the request's argument block (device addresses)
  const int64_t* slot_offsets;  // [num_slots]
  const int64_t* slot_elems;    // [num_slots]
  const void* const* tables;    // [num_tables]
  const int32_t* ids_base;      // [num_cols + 1] prefix of ids per column (ids_out hook only)
  int32_t* counters;            // sharded output only: [1] plan-owned slot counting finished warps
                                // (the last one signals the peers and resets it)
  uint32_t* done_seq;           // host-mapped word of this slot: receives `seq` when the launch ends
  uint32_t seq;
  int32_t pad0;                 // (profiling builds: rows of the flat kernel's timer buffer)
  int32_t total_warps;          // sharded output only: warps of all launches of this request (what
                                // `counters` counts up to)
  int32_t pad1;
  // Long rows (requests of density class >= 1, flat schedule, plain output): a multi-hot row with
  // more than kDeferIds ids is not summed by its tile's one warp but appended to this list -
  // [0] = count, then (item index, row) pairs - and summed by a whole CTA of long_rows_kernel.
  int32_t* long_list;           // null: rows are never deferred
  int32_t long_cap;             // entries the list holds
  int32_t long_ids;             // a row with more ids than this is deferred (kDeferIds)
  int32_t request_pdl;          // fused kernel only: overlap this grid's prologue with the tail of the launch
                                // in front of it (the previous request), see fused_columns_kernel
  int32_t pad3;
  int64_t blob_bytes;           // extent of the blob (L2 prefetch at kernel entry), 0: unknown
  // Sharded output (Mode A, fused exchange): rows [g * rows_per_rank, (g + 1) * rows_per_rank)
  // are stored into peer_out[g] (rank g's [rows_per_rank, global width] tensor, peer mapped);
  // world == 0: everything goes to `out`.
  uint8_t* peer_out[kMaxPeers];
  uint32_t* peer_flags[kMaxPeers];  // completion flags of every rank (null: no signal)
  uint32_t epoch;               // value stored into word `rank` of every peer's flags
  int32_t rank;
  uint32_t wait_in_kernel;      // the last warp also waits for all ranks' flags of this epoch
  uint32_t wait_timeout_ms;
  int64_t rows_per_rank;
  int32_t world;
  int32_t col_off_bytes;        // byte offset of this plan's slice inside a global row
  uint64_t* prof;               // profiling builds only, else null
  int64_t batch;
  int64_t out_row_bytes;
};

// Per-request argument data on the device: int64 slot_offsets[num_slots] then int64
// slot_elems[num_slots] (what Addons>ConcatInputs outputs as `offsets` / `shapes`,
// concat_inputs_ops.cc:42-77). It either travels inside the blob - recom_concat_inputs appends it
// as a trailer, so the one H2D copy of the blob brings it along and recom_plan_run uploads
// nothing (the reference uploads KnlArgs + SymbolResults on every call,
// cuda_emitter.cc:2182-2232) - or, for blobs packed by someone else, is staged through the
// plan's pinned ring into the caller's workspace. The table pointer array lives in a plan-owned
// device copy keyed by content (tables are graph variables: their addresses do not change).
constexpr uint64_t kArgsMagic = 0x5247524132304252ull;  // trailer: {magic, num_slots, offsets[], elems[]}

// kernels.cu
void launch_fused_columns(const DevPlan& plan, const Sched& sched, const ReqParams& req, int grid,
                          bool bf16, cudaStream_t stream, bool pdl);
// the flat kernel of the column tiles (one warp per tile); `plan.scratch_bytes` = its warp scratch
void launch_column_tiles(const DevPlan& plan, const ReqParams& req, const ColItem* items, int num_items,
                         bool bf16, cudaStream_t stream, bool pdl);
int column_tiles_configure(bool bf16, int smem_bytes);
int column_tiles_grid(int num_items);
// the rows the flat kernel deferred (ReqParams::long_list), one CTA per row at a time
void launch_long_rows(const DevPlan& plan, const ReqParams& req, const ColItem* items, int ctas, bool bf16,
                      cudaStream_t stream);
int long_rows_configure(bool bf16, int smem_bytes);  // CTAs (of 4 warps) the flat kernel launches for `num_items` tiles
void launch_shard_wait(uint32_t* local_flags, int world, uint32_t epoch, uint32_t timeout_ms,
                       cudaStream_t stream);
void launch_split_hash(const DevPlan& plan, const ReqParams& req, cudaStream_t stream);
void launch_coo_to_row_splits(const DevPlan& plan, const ReqParams& req, int64_t max_nnz,
                              cudaStream_t stream);
void launch_scatter_columns(const void* const* d_cols, const int32_t* d_dims,
                            const int32_t* d_offsets, int32_t num_cols, int64_t batch,
                            int32_t out_width, int32_t elem_bytes, void* out,
                            cudaStream_t stream);
void launch_segment_ids_to_row_splits(const int64_t* coo, int64_t nnz, int32_t stride,
                                      int64_t batch, int32_t* row_splits,
                                      cudaStream_t stream);
void launch_scale_rows(const DevPlan& plan, const int32_t* d_out_offsets, const ReqParams& req,
                       int64_t row_begin, int64_t rows, int32_t out_width, void* out, bool bf16,
                       cudaStream_t stream);
void launch_output_ptrs(int32_t num_cols, const int32_t* d_out_offsets, int32_t elem_bytes, void* out,
                        int64_t* ptrs_dev, cudaStream_t stream);
void launch_fingerprint64(const uint8_t* bytes, const int32_t* offsets, int64_t n,
                          int64_t num_buckets, uint64_t* out, cudaStream_t stream);
// dense_layer.cu: out = relu(A W^T + bias) on the tcgen05 tensor cores
int launch_dense_relu(const void* a, int64_t M, int32_t K, int64_t lda, const void* w, int32_t N,
                      const float* bias, int relu, float* out, cudaStream_t stream, const char** why);
// configures the dynamic shared memory of the fused kernel; returns resident CTAs per SM
int fused_kernel_configure(bool bf16, int smem_bytes);

}  // namespace recom
