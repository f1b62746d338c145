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
constexpr int kMaxWarpScratch = 6 << 10;
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
};

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
  int32_t sparse_index;       // rank among non-DENSE columns (workspace ids), -1 for DENSE
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
  int32_t num_cols;
  int32_t num_slabs;
  int32_t elem_bytes;   // 4 (fp32) or 2 (bf16): table and output element size
  int32_t check_ids;
  // dynamic shared memory layout (bytes from the start):
  // [staged tables | staged boundaries | DENSE column contexts | kWarps x warp scratch]
  int32_t smem_off_bnd;
  int32_t smem_off_ctx;
  int32_t smem_off_scratch;
  int32_t scratch_bytes;  // per warp: ids of a DENSE warp tile, or ctx + splits + ids of a SPARSE one
  int32_t smem_bytes;
  int32_t pad;
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

// The DENSE runs of a CTA: runs[first, first + count), walked in order.
struct CtaWork {
  int32_t first;
  int32_t count;
};

// The schedule of a batch size (cached on the device by the plan); a kernel parameter.
//   * FAST runs (slabs made of fast bucketize groups with staged tables) are static: every CTA
//     owns the runs cta[c], walked by its dense warps [0, kWarps - queue_warps).
//   * Everything else - multi-hot tiles, tiles of hashed / unstaged DENSE slabs - is a queue of
//     warp tiles in descending cost order. The queue warps of every CTA start there at once
//     (queue warp s of CTA c with tile s * grid + c, no claim needed), the dense warps join when
//     their run is done; further tiles are claimed with one atomicAdd each, one tile ahead.
struct Sched {
  const Tile* runs;          // FAST runs, grouped by CTA
  const CtaWork* cta;        // [grid]
  const Tile* queue;         // [num_queue]
  int32_t num_queue;
  int32_t queue_warps;
  int32_t claim_ahead;       // 0: claim when free, 1: always one tile ahead, 2: ahead except near the end
};

// Per-request values known on the host at launch; a kernel parameter (no dependent loads).
struct ReqParams {
  const uint8_t* blob;
  uint8_t* out;
  int32_t* ids_out;        // may be null
  int32_t* status;         // may be null
  int32_t* ws_ids;         // workspace: transformed ids of non-DENSE columns (overflow only)
  int32_t* ws_row_splits;  // workspace: row_splits built from COO columns
  // arrays of the request's argument block (device addresses; plan-cached or in the workspace)
  const int64_t* slot_offsets;  // [num_slots]
  const void* const* tables;    // [num_tables]
  const int32_t* ids_base;      // [num_cols + 1] prefix of ids per column
  const int32_t* sparse_base;   // [num_sparse + 1] prefix of ids, non-DENSE only
  int32_t* counters;            // [2] plan-owned slot: queue claims, finished warps (self-resetting)
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

// Per-request argument block: a header followed by the arrays ReqParams points to. It holds
// nothing that differs between requests of the same shape (no blob / output / workspace
// pointers), so the plan keeps device copies keyed by content and a steady stream of
// same-shaped requests uploads nothing (replaces the reference's per-call KnlArgs +
// SymbolResults uploads, cuda_emitter.cc:2182-2232).
struct ReqHeader {
  int64_t batch;
  int32_t num_coo;            // number of COO columns
  int32_t pad0;
  // byte offsets from the start of this block
  uint32_t off_slot_offsets;  // int64[num_slots]
  uint32_t off_tables;        // const void*[num_tables]
  uint32_t off_ids_base;      // int32[num_cols + 1]
  uint32_t off_sparse_base;   // int32[num_sparse + 1]
  uint32_t off_coo_cols;      // int32[num_coo] column index of each COO column
  uint32_t total_bytes;
  int64_t pad1;
};

__host__ __device__ inline const int64_t* req_slot_offsets(const ReqHeader* r) {
  return reinterpret_cast<const int64_t*>(reinterpret_cast<const uint8_t*>(r) +
                                          r->off_slot_offsets);
}
__host__ __device__ inline const void* const* req_tables(const ReqHeader* r) {
  return reinterpret_cast<const void* const*>(reinterpret_cast<const uint8_t*>(r) +
                                              r->off_tables);
}
__host__ __device__ inline const int32_t* req_ids_base(const ReqHeader* r) {
  return reinterpret_cast<const int32_t*>(reinterpret_cast<const uint8_t*>(r) +
                                          r->off_ids_base);
}
__host__ __device__ inline const int32_t* req_sparse_base(const ReqHeader* r) {
  return reinterpret_cast<const int32_t*>(reinterpret_cast<const uint8_t*>(r) +
                                          r->off_sparse_base);
}
__host__ __device__ inline const int32_t* req_coo_cols(const ReqHeader* r) {
  return reinterpret_cast<const int32_t*>(reinterpret_cast<const uint8_t*>(r) +
                                          r->off_coo_cols);
}

// kernels.cu
void launch_fused_columns(const DevPlan& plan, const Sched& sched, const ReqParams& req, int grid,
                          bool bf16, cudaStream_t stream);
void launch_shard_wait(uint32_t* local_flags, int world, uint32_t epoch, uint32_t timeout_ms,
                       cudaStream_t stream);
void launch_coo_to_row_splits(const DevPlan& plan, const ReqHeader* d_req, const uint8_t* blob,
                              int32_t* ws_row_splits, int num_coo, int64_t max_nnz,
                              cudaStream_t stream);
void launch_scatter_columns(const void* const* d_cols, const int32_t* d_dims,
                            const int32_t* d_offsets, int32_t num_cols, int64_t batch,
                            int32_t out_width, int32_t elem_bytes, void* out,
                            cudaStream_t stream);
void launch_segment_ids_to_row_splits(const int64_t* coo, int64_t nnz, int32_t stride,
                                      int64_t batch, int32_t* row_splits,
                                      cudaStream_t stream);
void launch_fingerprint64(const uint8_t* bytes, const int32_t* offsets, int64_t n,
                          int64_t num_buckets, uint64_t* out, cudaStream_t stream);
// configures the dynamic shared memory of the fused kernel; returns resident CTAs per SM
int fused_kernel_configure(bool bf16, int smem_bytes);

}  // namespace recom
