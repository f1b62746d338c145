// kernels.cu — hand-written sm_100a kernels of the fused embedding-column stage.
//
// fused_columns_kernel is the hot kernel. It replaces, in one persistent launch,
//   * the reference's FusedKnl (one 64-thread block per column,
//     tensorflow_addons/graph_optimizers/cuda_emitter.cc:2097-2134) with its
//     device templates Bucketize (:233-247), GatherRowsToGlbMem (:250-293),
//     GatherScatterRows (:296-345), SparseSegmentSum/Mean (:348-661) and
//     experiment::SparseSegmentReduce (:820-962),
//   * ConcatOutputsKnl / ScatterBlock
//     (tensorflow_addons/custom_ops/concat_outputs/concat_outputs_op_gpu.cu.cc:85-131):
//     every result is written once, straight into its slice of the [B, sum D] row,
//   * the TF-CPU ops the reference leaves upstream: StringToHashBucketFast and
//     AsString.
//
// Work decomposition: the host builds a per-request table of tiles
// (slab, row block). A slab is a run of columns whose output slices are adjacent
// in the concat row and at most 32 16-byte chunks wide, so one warp store covers
// up to 512 contiguous bytes of one output row. CTAs (256 threads, several per
// SM) pull tiles through an atomic counter, heavy multi-hot tiles first.
//
// Per tile:
//   phase 1  transform raw inputs to row ids (bucketize / hash / cast), one warp
//            per column with lanes along the batch so input reads are coalesced;
//            ids go to shared memory (DENSE) or to the workspace (multi-hot).
//   phase 2  lane = chunk of the slab; per output row: read id(s), 128-bit
//            read-only gathers of the table row chunk, fp32 accumulation,
//            combiner scaling, one 128-bit streaming store into the concat row.
//
// The path is HBM / L2 bound integer+copy work: no tensor cores by design.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "farmhash.cuh"
#include "plan.h"

namespace recom {

// ---------------------------------------------------------------------------
// small load/store helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_cs_v4(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// A chunk of VEC bytes of a table / output row.
template <int VEC>
struct Chunk;
template <>
struct Chunk<16> {
  uint4 v;
  __device__ __forceinline__ void load(const void* p) { v = ldg_nc_v4(p); }
  __device__ __forceinline__ void store(void* p) const { st_cs_v4(p, v); }
  __device__ __forceinline__ void zero() { v = make_uint4(0, 0, 0, 0); }
};
template <>
struct Chunk<4> {
  uint32_t v;
  __device__ __forceinline__ void load(const void* p) { v = __ldg((const uint32_t*)p); }
  __device__ __forceinline__ void store(void* p) const { *(uint32_t*)p = v; }
  __device__ __forceinline__ void zero() { v = 0; }
};
template <>
struct Chunk<2> {
  uint16_t v;
  __device__ __forceinline__ void load(const void* p) { v = __ldg((const uint16_t*)p); }
  __device__ __forceinline__ void store(void* p) const { *(uint16_t*)p = v; }
  __device__ __forceinline__ void zero() { v = 0; }
};

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<const uint32_t*>(&p);
}

// fp32 accumulator of one chunk. NE = elements per chunk.
template <int VEC, bool BF16>
struct Acc;
template <>
struct Acc<16, false> {
  float a[4];
  __device__ __forceinline__ void clear() { a[0] = a[1] = a[2] = a[3] = 0.f; }
  __device__ __forceinline__ void add(const Chunk<16>& c) {
    a[0] += __uint_as_float(c.v.x);
    a[1] += __uint_as_float(c.v.y);
    a[2] += __uint_as_float(c.v.z);
    a[3] += __uint_as_float(c.v.w);
  }
  __device__ __forceinline__ void mul(float s) {
    a[0] *= s; a[1] *= s; a[2] *= s; a[3] *= s;
  }
  __device__ __forceinline__ void div(float s) {
    a[0] = __fdiv_rn(a[0], s); a[1] = __fdiv_rn(a[1], s);
    a[2] = __fdiv_rn(a[2], s); a[3] = __fdiv_rn(a[3], s);
  }
  __device__ __forceinline__ Chunk<16> pack() const {
    Chunk<16> c;
    c.v = make_uint4(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]),
                     __float_as_uint(a[3]));
    return c;
  }
};
template <>
struct Acc<16, true> {
  float a[8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
  }
  __device__ __forceinline__ void add(const Chunk<16>& c) {
    a[0] += bf16_lo(c.v.x); a[1] += bf16_hi(c.v.x);
    a[2] += bf16_lo(c.v.y); a[3] += bf16_hi(c.v.y);
    a[4] += bf16_lo(c.v.z); a[5] += bf16_hi(c.v.z);
    a[6] += bf16_lo(c.v.w); a[7] += bf16_hi(c.v.w);
  }
  __device__ __forceinline__ void mul(float s) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= s;
  }
  __device__ __forceinline__ void div(float s) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __fdiv_rn(a[i], s);
  }
  __device__ __forceinline__ Chunk<16> pack() const {
    Chunk<16> c;
    c.v = make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]),
                     pack_bf16(a[6], a[7]));
    return c;
  }
};
template <>
struct Acc<4, false> {
  float a;
  __device__ __forceinline__ void clear() { a = 0.f; }
  __device__ __forceinline__ void add(const Chunk<4>& c) { a += __uint_as_float(c.v); }
  __device__ __forceinline__ void mul(float s) { a *= s; }
  __device__ __forceinline__ void div(float s) { a = __fdiv_rn(a, s); }
  __device__ __forceinline__ Chunk<4> pack() const {
    Chunk<4> c;
    c.v = __float_as_uint(a);
    return c;
  }
};
template <>
struct Acc<2, true> {
  float a;
  __device__ __forceinline__ void clear() { a = 0.f; }
  __device__ __forceinline__ void add(const Chunk<2>& c) { a += bf16_lo((uint32_t)c.v); }
  __device__ __forceinline__ void mul(float s) { a *= s; }
  __device__ __forceinline__ void div(float s) { a = __fdiv_rn(a, s); }
  __device__ __forceinline__ Chunk<2> pack() const {
    Chunk<2> c;
    c.v = (uint16_t)(pack_bf16(a, 0.f) & 0xffffu);
    return c;
  }
};

// ---------------------------------------------------------------------------
// phase 1: raw value -> table row id
// ---------------------------------------------------------------------------
// TF Bucketize == std::upper_bound over float boundaries with `value < boundary`
// (bucketize_op.cc:42-44; the reference's device Bucketize, cuda_emitter.cc:233-247).
template <typename T>
__device__ __forceinline__ int32_t bucketize(const float* __restrict__ bnd, int32_t nb, T v) {
  int32_t lo = 0, n = nb;
  while (n > 0) {
    const int32_t half = n >> 1;
    const int32_t mid = lo + half;
    const bool right = !(v < __ldg(bnd + mid));
    lo = right ? mid + 1 : lo;
    n = right ? n - half - 1 : half;
  }
  return lo;
}

__device__ __forceinline__ int32_t column_id(const DevColumn& c, const uint8_t* __restrict__ blob,
                                             const int64_t* __restrict__ slot_off,
                                             const float* __restrict__ boundaries, int64_t i) {
  const uint8_t* vals = blob + slot_off[c.value_slot];
  switch (c.transform) {
    case RECOM_XFORM_NONE:
      if (c.value_dtype == RECOM_I32) return __ldg((const int32_t*)vals + i);
      return (int32_t)__ldg((const long long*)vals + i);
    case RECOM_XFORM_BUCKETIZE: {
      const float* b = boundaries + c.boundaries_offset;
      switch (c.value_dtype) {
        case RECOM_F32: return bucketize<float>(b, c.num_boundaries, __ldg((const float*)vals + i));
        case RECOM_I32: return bucketize<int32_t>(b, c.num_boundaries, __ldg((const int32_t*)vals + i));
        case RECOM_I64:
          return bucketize<long long>(b, c.num_boundaries, __ldg((const long long*)vals + i));
        default: return bucketize<double>(b, c.num_boundaries, __ldg((const double*)vals + i));
      }
    }
    case RECOM_XFORM_HASH_STR: {
      const int32_t* off = (const int32_t*)vals;
      const int32_t o0 = __ldg(off + i), o1 = __ldg(off + i + 1);
      const uint64_t h =
          fh::fingerprint64(blob + slot_off[c.bytes_slot] + o0, (uint32_t)(o1 - o0));
      return (int32_t)(h % (uint64_t)c.num_buckets);
    }
    default: {  // RECOM_XFORM_HASH_INT
      uint64_t h;
      if (c.value_dtype == RECOM_I32) {
        const int32_t v = __ldg((const int32_t*)vals + i);
        h = (v >= 0 && v < 100000000) ? fh::fingerprint64_of_decimal_u32small((uint32_t)v)
                                      : fh::fingerprint64_of_decimal((int64_t)v);
      } else {
        h = fh::fingerprint64_of_decimal((int64_t)__ldg((const long long*)vals + i));
      }
      return (int32_t)(h % (uint64_t)c.num_buckets);
    }
  }
}

__device__ __forceinline__ int32_t checked_id(int32_t id, int64_t rows, int32_t* status) {
  if ((uint64_t)(int64_t)id >= (uint64_t)rows) {
    if (status) atomicExch(status, RECOM_ERR_ID_OUT_OF_RANGE);
    return 0;
  }
  return id;
}

struct __align__(16) TileSmem {
  int32_t ids[kMaxChunks * kIdPitch];
  int32_t next_tile;
};

// ---------------------------------------------------------------------------
// DENSE slab: exactly one id per (column, row)
// ---------------------------------------------------------------------------
template <int VEC, bool CHECK>
__device__ __forceinline__ void dense_tile(const DevPlan& P, const ReqHeader* __restrict__ R,
                                           const DevSlab& slab, const Tile& tile,
                                           TileSmem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* slot_off = req_slot_offsets(R);
  const int32_t* ids_base = req_ids_base(R);
  const int rows = tile.rows;
  const int64_t b0 = tile.row_begin;

  // phase 1: one warp per column, lanes along the batch
  for (int cl = warp; cl < slab.num_cols; cl += kWarps) {
    const int ci = P.slab_cols[slab.first_col + cl];
    const DevColumn c = P.cols[ci];
    for (int r = lane; r < rows; r += 32) {
      int32_t id = column_id(c, R->blob, slot_off, P.boundaries, b0 + r);
      if (CHECK) id = checked_id(id, c.table_rows, R->status);
      sm.ids[cl * kIdPitch + r] = id;
      if (R->ids_out) R->ids_out[ids_base[ci] + b0 + r] = id;
    }
  }
  __syncthreads();

  // phase 2: lane = chunk; rows_per_pass rows per warp pass
  const int W = slab.num_chunks;
  const int rpp = 32 / W;  // rows handled by one warp per pass
  const int sub = lane / W;
  const int chunk = lane - sub * W;
  if (sub < rpp) {
    const DevChunk ch = P.chunks[slab.first_chunk + chunk];
    const DevColumn& c = P.cols[ch.col];
    const uint8_t* tbl = (const uint8_t*)req_tables(R)[c.table] + ch.table_byte;
    const int64_t row_bytes = c.row_bytes;
    const int32_t* ids = sm.ids + ch.col_local * kIdPitch;
    uint8_t* outp = R->out + b0 * R->out_row_bytes + ch.out_byte;
    const int64_t orb = R->out_row_bytes;
    const int step = kWarps * rpp;
    int r = warp * rpp + sub;
    // 4 independent gathers in flight per lane
    for (; r + 3 * step < rows; r += 4 * step) {
      Chunk<VEC> v0, v1, v2, v3;
      v0.load(tbl + (int64_t)ids[r] * row_bytes);
      v1.load(tbl + (int64_t)ids[r + step] * row_bytes);
      v2.load(tbl + (int64_t)ids[r + 2 * step] * row_bytes);
      v3.load(tbl + (int64_t)ids[r + 3 * step] * row_bytes);
      v0.store(outp + (int64_t)r * orb);
      v1.store(outp + (int64_t)(r + step) * orb);
      v2.store(outp + (int64_t)(r + 2 * step) * orb);
      v3.store(outp + (int64_t)(r + 3 * step) * orb);
    }
    for (; r < rows; r += step) {
      Chunk<VEC> v;
      v.load(tbl + (int64_t)ids[r] * row_bytes);
      v.store(outp + (int64_t)r * orb);
    }
  }
  __syncthreads();  // ids are reused by the next tile
}

// ---------------------------------------------------------------------------
// SPARSE slab: row b owns ids [row_splits[b], row_splits[b+1])
// ---------------------------------------------------------------------------
__device__ __forceinline__ const int32_t* column_row_splits(const DevColumn& c,
                                                            const ReqHeader* R,
                                                            const int64_t* slot_off) {
  if (c.coo_index >= 0) return R->ws_row_splits + (int64_t)c.coo_index * (R->batch + 1);
  return (const int32_t*)(R->blob + slot_off[c.seg_slot]);
}

template <int VEC, bool BF16, bool CHECK>
__device__ __forceinline__ void sparse_tile(const DevPlan& P, const ReqHeader* __restrict__ R,
                                            const DevSlab& slab, const Tile& tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* slot_off = req_slot_offsets(R);
  const int32_t* ids_base = req_ids_base(R);
  const int32_t* sparse_base = req_sparse_base(R);
  const int rows = tile.rows;
  const int64_t b0 = tile.row_begin;

  // phase 1: materialise row ids of transformed columns for this row block
  bool wrote = false;
  for (int cl = 0; cl < slab.num_cols; ++cl) {
    const int ci = P.slab_cols[slab.first_col + cl];
    const DevColumn c = P.cols[ci];
    const bool xform = c.transform != RECOM_XFORM_NONE;
    if (!xform && !R->ids_out && !CHECK) continue;
    const int32_t* rs = column_row_splits(c, R, slot_off);
    const int32_t e0 = rs[b0], e1 = rs[b0 + rows];
    int32_t* ws = R->ws_ids + sparse_base[c.sparse_index];
    for (int32_t e = e0 + (int)threadIdx.x; e < e1; e += kThreads) {
      int32_t id = column_id(c, R->blob, slot_off, P.boundaries, e);
      if (CHECK) id = checked_id(id, c.table_rows, R->status);
      if (xform || CHECK) ws[e] = id;
      if (R->ids_out) R->ids_out[ids_base[ci] + e] = id;
    }
    wrote |= (xform || CHECK);
  }
  if (wrote) __syncthreads();  // block-scope visibility of the global writes above

  // phase 2
  const int W = slab.num_chunks;
  const int rpp = 32 / W;
  const int sub = lane / W;
  const int chunk = lane - sub * W;
  if (sub >= rpp) return;
  const DevChunk ch = P.chunks[slab.first_chunk + chunk];
  const DevColumn& c = P.cols[ch.col];
  const uint8_t* tbl = (const uint8_t*)req_tables(R)[c.table] + ch.table_byte;
  const int64_t row_bytes = c.row_bytes;
  const int32_t* rs = column_row_splits(c, R, slot_off) + b0;
  const bool from_ws = (c.transform != RECOM_XFORM_NONE) || CHECK;
  const int32_t* ids32 = from_ws ? R->ws_ids + sparse_base[c.sparse_index]
                                 : (const int32_t*)(R->blob + slot_off[c.value_slot]);
  const bool ids64 = !from_ws && c.value_dtype == RECOM_I64;
  const long long* idsll = (const long long*)ids32;
  uint8_t* outp = R->out + b0 * R->out_row_bytes + ch.out_byte;
  const int combiner = c.combiner;

  for (int r = warp * rpp + sub; r < rows; r += kWarps * rpp) {
    const int32_t beg = rs[r], end = rs[r + 1];
    Acc<VEC, BF16> acc;
    acc.clear();
    int32_t e = beg;
    for (; e + 3 < end; e += 4) {  // 4 independent row gathers in flight
      int64_t i0, i1, i2, i3;
      if (ids64) {
        i0 = idsll[e]; i1 = idsll[e + 1]; i2 = idsll[e + 2]; i3 = idsll[e + 3];
      } else {
        i0 = ids32[e]; i1 = ids32[e + 1]; i2 = ids32[e + 2]; i3 = ids32[e + 3];
      }
      Chunk<VEC> v0, v1, v2, v3;
      v0.load(tbl + i0 * row_bytes);
      v1.load(tbl + i1 * row_bytes);
      v2.load(tbl + i2 * row_bytes);
      v3.load(tbl + i3 * row_bytes);
      acc.add(v0); acc.add(v1); acc.add(v2); acc.add(v3);
    }
    for (; e < end; ++e) {
      const int64_t i0 = ids64 ? (int64_t)idsll[e] : (int64_t)ids32[e];
      Chunk<VEC> v;
      v.load(tbl + i0 * row_bytes);
      acc.add(v);
    }
    const int32_t n = end - beg;
    if (n > 1 && combiner != RECOM_COMBINE_SUM) {
      // TF-CPU scaling (segment_reduction_ops_impl.h:619-629,765-770):
      // n < 10 multiplies by 1/m, n >= 10 divides by m; m = n or sqrt(n).
      const float m = combiner == RECOM_COMBINE_MEAN ? (float)n
                                                     : (float)__dsqrt_rn((double)n);
      if (n < 10) acc.mul(__fdiv_rn(1.0f, m));
      else acc.div(m);
    }
    acc.pack().store(outp + (int64_t)r * R->out_row_bytes);
  }
}

// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
template <bool BF16, bool CHECK>
__global__ void __launch_bounds__(kThreads, 4)
fused_columns_kernel(const DevPlan P, ReqHeader* __restrict__ R) {
  __shared__ TileSmem sm;
  constexpr int ELEM = BF16 ? 2 : 4;
  const Tile* tiles = req_tiles(R);
  int t = blockIdx.x;
  while (t < R->num_tiles) {
    const Tile tile = tiles[t];
    const DevSlab slab = P.slabs[tile.slab];
    if (threadIdx.x == 0) sm.next_tile = atomicAdd(&R->tile_counter, 1);
    if (slab.cls == kSlabDense) {
      if (slab.vec_bytes == 16) dense_tile<16, CHECK>(P, R, slab, tile, sm);
      else dense_tile<ELEM, CHECK>(P, R, slab, tile, sm);
    } else {
      if (slab.vec_bytes == 16) sparse_tile<16, BF16, CHECK>(P, R, slab, tile);
      else sparse_tile<ELEM, BF16, CHECK>(P, R, slab, tile);
      __syncthreads();
    }
    t = sm.next_tile;
    __syncthreads();
  }
}

int fused_kernel_max_ctas_per_sm(bool bf16) {
  int n = 0;
  if (bf16)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_columns_kernel<true, false>,
                                                  kThreads, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_columns_kernel<false, false>,
                                                  kThreads, 0);
  return n;
}

void launch_fused_columns(const DevPlan& plan, ReqHeader* d_req, int grid, bool bf16,
                          cudaStream_t stream) {
  const bool check = plan.check_ids != 0;
  if (bf16) {
    if (check) fused_columns_kernel<true, true><<<grid, kThreads, 0, stream>>>(plan, d_req);
    else fused_columns_kernel<true, false><<<grid, kThreads, 0, stream>>>(plan, d_req);
  } else {
    if (check) fused_columns_kernel<false, true><<<grid, kThreads, 0, stream>>>(plan, d_req);
    else fused_columns_kernel<false, false><<<grid, kThreads, 0, stream>>>(plan, d_req);
  }
}

// ---------------------------------------------------------------------------
// COO row ids -> CSR row_splits (replaces experiment::ComputeSegmentOffsets,
// cuda_emitter.cc:768-818, which walks tiles serially in one block).
// Every boundary i in [0, nnz] fills row_splits for the rows in (row[i-1], row[i]].
// ---------------------------------------------------------------------------
__device__ __forceinline__ void fill_row_splits(const long long* __restrict__ coo, int64_t nnz,
                                                int32_t stride, int64_t batch,
                                                int32_t* __restrict__ rs, int64_t i) {
  const int64_t prev = i == 0 ? -1 : (int64_t)__ldg(coo + (i - 1) * stride);
  int64_t cur = i == nnz ? batch : (int64_t)__ldg(coo + i * stride);
  if (cur > batch) cur = batch;
  for (int64_t r = prev + 1; r <= cur; ++r) rs[r] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
coo_to_row_splits_kernel(const DevPlan P, const ReqHeader* __restrict__ R) {
  const int k = blockIdx.y;
  const int ci = req_coo_cols(R)[k];
  const DevColumn c = P.cols[ci];
  const int32_t* ids_base = req_ids_base(R);
  const int64_t nnz = ids_base[ci + 1] - ids_base[ci];
  const long long* coo = (const long long*)(R->blob + req_slot_offsets(R)[c.seg_slot]);
  int32_t* rs = R->ws_row_splits + (int64_t)k * (R->batch + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz;
       i += (int64_t)gridDim.x * blockDim.x)
    fill_row_splits(coo, nnz, c.seg_stride, R->batch, rs, i);
}

void launch_coo_to_row_splits(const DevPlan& plan, const ReqHeader* d_req, int num_coo,
                              int64_t max_nnz, cudaStream_t stream) {
  int64_t bx = (max_nnz + 1 + 255) / 256;
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)num_coo);
  coo_to_row_splits_kernel<<<grid, 256, 0, stream>>>(plan, d_req);
}

__global__ void __launch_bounds__(256)
segment_ids_to_row_splits_kernel(const long long* __restrict__ coo, int64_t nnz, int32_t stride,
                                 int64_t batch, int32_t* __restrict__ rs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz;
       i += (int64_t)gridDim.x * blockDim.x)
    fill_row_splits(coo, nnz, stride, batch, rs, i);
}

void launch_segment_ids_to_row_splits(const int64_t* coo, int64_t nnz, int32_t stride,
                                      int64_t batch, int32_t* row_splits,
                                      cudaStream_t stream) {
  int64_t bx = (nnz + 1 + 255) / 256;
  if (bx > 148 * 8) bx = 148 * 8;
  segment_ids_to_row_splits_kernel<<<(unsigned)bx, 256, 0, stream>>>(
      (const long long*)coo, nnz, stride, batch, row_splits);
}

// ---------------------------------------------------------------------------
// ConcatOutputs' host_inputs capability: scatter externally computed columns.
// grid.y = column, threads along (row, element) of that column.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_columns_kernel(const void* const* __restrict__ cols, const int32_t* __restrict__ dims,
                       const int32_t* __restrict__ offsets, int64_t batch, int32_t out_width,
                       int32_t elem_bytes, uint8_t* __restrict__ out) {
  const int c = blockIdx.y;
  const int32_t d = dims[c];
  const int64_t total = batch * d;
  const uint8_t* src = (const uint8_t*)cols[c];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / d;
    const int32_t e = (int32_t)(i - p * d);
    const int64_t o = (p * out_width + offsets[c] + e) * elem_bytes;
    if (elem_bytes == 4) *(uint32_t*)(out + o) = *(const uint32_t*)(src + i * 4);
    else *(uint16_t*)(out + o) = *(const uint16_t*)(src + i * 2);
  }
}

void launch_scatter_columns(const void* const* d_cols, const int32_t* d_dims,
                            const int32_t* d_offsets, int32_t num_cols, int64_t batch,
                            int32_t out_width, int32_t elem_bytes, void* out,
                            cudaStream_t stream) {
  dim3 grid(64, (unsigned)num_cols);
  scatter_columns_kernel<<<grid, 256, 0, stream>>>(d_cols, d_dims, d_offsets, batch, out_width,
                                                  elem_bytes, (uint8_t*)out);
}

// ---------------------------------------------------------------------------
// Stand-alone device StringToHashBucketFast / Fingerprint64 (parity hook).
// num_buckets == 0 writes the raw 64-bit fingerprint.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fingerprint64_kernel(const uint8_t* __restrict__ bytes, const int32_t* __restrict__ offsets,
                     int64_t n, uint64_t num_buckets, uint64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t o0 = offsets[i], o1 = offsets[i + 1];
    const uint64_t h = fh::fingerprint64(bytes + o0, (uint32_t)(o1 - o0));
    out[i] = num_buckets ? h % num_buckets : h;
  }
}

void launch_fingerprint64(const uint8_t* bytes, const int32_t* offsets, int64_t n,
                          int64_t num_buckets, uint64_t* out, cudaStream_t stream) {
  int64_t bx = (n + 255) / 256;
  if (bx > 148 * 8) bx = 148 * 8;
  if (bx < 1) bx = 1;
  fingerprint64_kernel<<<(unsigned)bx, 256, 0, stream>>>(bytes, offsets, n,
                                                        (uint64_t)num_buckets, out);
}

}  // namespace recom
