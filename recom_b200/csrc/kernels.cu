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
// Work decomposition: a table of tiles (slab, row block) per batch size, cached on
// the device. A slab is a run of columns whose output slices are adjacent in the
// concat row and at most 32 16-byte chunks wide, so one warp store covers up to
// 512 contiguous bytes of one output row. Persistent CTAs (256 threads, 3 per SM)
// first drain a shared pool of multi-hot tiles, then walk a static run of dense
// tiles of ONE slab (so what they staged in shared memory is reused), then drain
// a shared pool of left-over dense tiles (load balance).
//
// Per tile:
//   stage    (dense, on slab change) the slab's small tables and bucket boundaries
//            are copied into shared memory with TMA bulk copies (cp.async.bulk,
//            mbarrier completion), overlapping phase 1.
//   phase 1  transform raw inputs to row ids (bucketize / hash / cast), one warp
//            per column with lanes along the batch so input reads are coalesced;
//            ids go to shared memory (DENSE) or to the workspace (multi-hot).
//   phase 2  lane = chunk of the slab; per output row: read id(s), 128-bit gathers
//            of the table row chunk (from shared memory when staged, else
//            ld.global.nc), fp32 accumulation, combiner scaling, one 128-bit
//            streaming store into the concat row.
// One __syncthreads per dense tile (ids and the next tile index are double
// buffered), so fast warps start the next tile's phase 1 while others still store.
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
// shared-memory / TMA helpers (sm_100a)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waiting warps back off with nanosleep: a warp spinning on try_wait competes for issue slots
// with the warps (often a single one doing serial work) that everybody is waiting for.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}
// TMA bulk copy global -> shared (UBLKCP), completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// A chunk read from shared memory.
template <int VEC>
__device__ __forceinline__ Chunk<VEC> lds_chunk(const uint8_t* p);
template <>
__device__ __forceinline__ Chunk<16> lds_chunk<16>(const uint8_t* p) {
  Chunk<16> c;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(c.v.x), "=r"(c.v.y), "=r"(c.v.z), "=r"(c.v.w)
               : "r"(smem_u32(p)));
  return c;
}
template <>
__device__ __forceinline__ Chunk<4> lds_chunk<4>(const uint8_t* p) {
  Chunk<4> c;
  c.v = *reinterpret_cast<const uint32_t*>(p);
  return c;
}
template <>
__device__ __forceinline__ Chunk<2> lds_chunk<2>(const uint8_t* p) {
  Chunk<2> c;
  c.v = *reinterpret_cast<const uint16_t*>(p);
  return c;
}

// ---------------------------------------------------------------------------
// phase 1: raw value -> table row id
// ---------------------------------------------------------------------------
// TF Bucketize == std::upper_bound over float boundaries with `value < boundary`
// (bucketize_op.cc:42-44; the reference's device Bucketize, cuda_emitter.cc:233-247).
// N independent searches run in lock step (same boundary array) for ILP. `bnd` may
// point to shared or global memory.
template <typename T, int N>
__device__ __forceinline__ void bucketize_n(const float* __restrict__ bnd, int32_t nb,
                                            const T (&v)[N], int32_t (&out)[N]) {
  int32_t lo[N];
#pragma unroll
  for (int k = 0; k < N; ++k) lo[k] = 0;
  int32_t n = nb;
  // Every search sees the same sequence of range lengths: with len = n, the left part has
  // `half` elements and the right part n - half - 1 <= half; using `half` for both keeps the
  // searches in lock step and is still a correct upper_bound (the probe is clamped).
  while (n > 0) {
    const int32_t half = n >> 1;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const int32_t mid = min(lo[k] + half, nb - 1);
      const bool right = (lo[k] + half < nb) && !(v[k] < bnd[mid]);
      lo[k] = right ? lo[k] + half + 1 : lo[k];
    }
    n = half;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) out[k] = min(lo[k], nb);
}

template <typename T>
__device__ __forceinline__ int32_t bucketize(const float* __restrict__ bnd, int32_t nb, T v) {
  const T vv[1] = {v};
  int32_t o[1];
  bucketize_n<T, 1>(bnd, nb, vv, o);
  return o[0];
}

// Any transform, one element. `bnd` = this column's boundaries (shared or global).
__device__ __forceinline__ int32_t column_id(const ColCtx& c, const float* __restrict__ bnd,
                                             int64_t i) {
  const uint8_t* vals = c.values;
  switch (c.transform) {
    case RECOM_XFORM_NONE:
      if (c.value_dtype == RECOM_I32) return __ldg((const int32_t*)vals + i);
      return (int32_t)__ldg((const long long*)vals + i);
    case RECOM_XFORM_BUCKETIZE: {
      switch (c.value_dtype) {
        case RECOM_F32: return bucketize<float>(bnd, c.num_boundaries, __ldg((const float*)vals + i));
        case RECOM_I32: return bucketize<int32_t>(bnd, c.num_boundaries, __ldg((const int32_t*)vals + i));
        case RECOM_I64:
          return bucketize<long long>(bnd, c.num_boundaries, __ldg((const long long*)vals + i));
        default: return bucketize<double>(bnd, c.num_boundaries, __ldg((const double*)vals + i));
      }
    }
    case RECOM_XFORM_HASH_STR: {
      const int32_t* off = (const int32_t*)vals;
      const int32_t o0 = __ldg(off + i), o1 = __ldg(off + i + 1);
      const uint64_t h = fh::fingerprint64(c.bytes + o0, (uint32_t)(o1 - o0));
      return (int32_t)fh::mod_bucket(h, (uint64_t)c.num_buckets, c.bucket_magic);
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
      return (int32_t)fh::mod_bucket(h, (uint64_t)c.num_buckets, c.bucket_magic);
    }
  }
}

__device__ __forceinline__ int32_t checked_id(int32_t id, int32_t rows, int32_t* status) {
  if ((uint32_t)id >= (uint32_t)rows) {
    if (status) atomicExch(status, RECOM_ERR_ID_OUT_OF_RANGE);
    return 0;
  }
  return id;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// Static shared state of a CTA. The dynamic part holds, for a DENSE slab,
// [staged tables | staged boundaries | ids of the tile]; a SPARSE tile reuses all of it for
// the tile's row_splits and row ids.
struct __align__(16) CtaShared {
  Tile next[2];            // the tile two steps ahead (slab < 0: none), double buffered
  uint64_t mbar_bnd;       // boundaries of the staged slab have landed
  uint64_t mbar_tbl;       // tables of the staged slab have landed
  uint64_t full[2];        // ids ring: slot filled by the producer warps
  uint64_t empty[2];       // ids ring: slot released by the consumer warps
  ColCtx ctx[kMaxChunks];  // the columns of the slab the CTA is working on
  int32_t sp_cnt[kMaxChunks];  // SPARSE tile: ids of the tile per column, -1: not staged
};

// Request-wide values every tile needs, read once per CTA.
struct ReqRegs {
  uint8_t* out;
  int64_t out_row_bytes;
  int32_t* status;
};

// ---------------------------------------------------------------------------
// entering a slab: resolve its columns against the request (one thread per column), start
// the TMA bulk copies of its small tables and bucket boundaries into shared memory
// ---------------------------------------------------------------------------
__device__ __forceinline__ void enter_slab(const DevPlan& P, const ReqParams& A,
                                           const DevSlab& slab, CtaShared& cs, uint8_t* s_tables,
                                           float* s_bnd, bool keep_bnd) {
  const int tid = threadIdx.x;
  if (tid < slab.num_cols) {
    const DevSlabCol sc = P.slab_cols[slab.first_col + tid];
    ColCtx x;
    x.values = A.blob + A.slot_offsets[sc.value_slot];
    x.bytes = sc.bytes_slot >= 0 ? A.blob + A.slot_offsets[sc.bytes_slot] : nullptr;
    x.table = (const uint8_t*)A.tables[sc.table];
    x.row_splits = nullptr;
    x.ws_ids = nullptr;
    if (sc.sparse_index >= 0) {
      x.row_splits = sc.coo_index >= 0
                         ? A.ws_row_splits + (int64_t)sc.coo_index * (A.batch + 1)
                         : (const int32_t*)(A.blob + A.slot_offsets[sc.seg_slot]);
      x.ws_ids = A.ws_ids + A.sparse_base[sc.sparse_index];
    }
    x.ids_out = A.ids_out ? A.ids_out + A.ids_base[sc.col] : nullptr;
    x.bucket_magic = sc.bucket_magic;
    x.num_buckets = sc.num_buckets;
    x.table_rows = sc.table_rows;
    x.row_bytes = sc.row_bytes;
    x.transform = sc.transform;
    x.value_dtype = sc.value_dtype;
    x.num_boundaries = sc.num_boundaries;
    x.bnd_off = sc.bnd_off;
    x.bnd_pool_off = sc.boundaries_offset;
    x.stage_off = sc.stage_off;
    x.combiner = sc.combiner;
    cs.ctx[tid] = x;
    if (sc.stage_off >= 0 && sc.stage_copy) {
      fence_proxy_async();  // the region was read through the generic proxy by the last slab
      bulk_g2s(s_tables + sc.stage_off, x.table, (uint32_t)sc.table_rows * (uint32_t)sc.row_bytes,
               &cs.mbar_tbl);
    }
  } else if (!keep_bnd && tid >= 32 && tid - 32 < slab.num_bnds) {
    const DevSlabBnd b = P.slab_bnds[slab.first_bnd + tid - 32];
    fence_proxy_async();
    bulk_g2s(s_bnd + b.smem_off, P.boundaries + b.pool_off, (uint32_t)b.floats * 4, &cs.mbar_bnd);
  }
  const int bnd_bytes = keep_bnd ? 0 : slab.bnd_bytes;
  if (tid == 0 && slab.stage_bytes + bnd_bytes > 0) {
    // complete_tx of a copy issued above may precede this expect_tx: the phase cannot
    // complete before this (the only) arrival
    mbar_expect_tx(&cs.mbar_bnd, (uint32_t)bnd_bytes);
    mbar_expect_tx(&cs.mbar_tbl, (uint32_t)slab.stage_bytes);
  }
}

// ---------------------------------------------------------------------------
// DENSE slab: exactly one id per (column, row)
// ---------------------------------------------------------------------------
// Is this column on the fast path: float32 input bucketized against boundaries staged in smem?
__device__ __forceinline__ bool fast_bucketize(const ColCtx& c) {
  return c.transform == RECOM_XFORM_BUCKETIZE && c.value_dtype == RECOM_F32 && c.bnd_off >= 0;
}

// 8 upper_bound searches in lock step over boundaries in shared memory; search k uses
// boundary set k / 2 (see bucketize_n for why `half` serves both sides of a probe).
__device__ __forceinline__ void search_smem8(const uint32_t (&bp)[4], const int32_t (&nb)[4],
                                             const float (&v)[8], int32_t (&lo)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) lo[k] = 0;
  for (int32_t n = max(max(nb[0], nb[1]), max(nb[2], nb[3])); n > 0; n >>= 1) {
    const int32_t half = n >> 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int32_t idx = lo[k] + half;
      const float b = lds_f32(bp[k >> 1] + 4u * (uint32_t)min(idx, nb[k >> 1] - 1));
      const bool right = idx < nb[k >> 1] && !(v[k] < b);
      lo[k] = right ? idx + 1 : lo[k];
    }
  }
}

// DENSE tiles are processed by two groups of warps that hand over 64-row sub-tiles through a
// two-slot ring of row ids in shared memory (mbarrier full / empty per slot):
//   producers (warps 0-3): inputs -> row ids (bucketize / hash), lanes along the batch;
//   consumers (warps 4-7): lane = chunk; gather the table rows, store into the concat rows.
// The stores, which bound this path, therefore never wait for id computation or input latency.
//
// Producer side of one sub-tile. A producer warp owns the columns pw, pw + 4, pw + 8, ... and
// takes them four at a time: 8 coalesced loads in flight, then 8 searches in lock step. The
// id slot receives the byte offset of the row inside the smem table region when the table is
// staged, else the raw row id.
template <bool CHECK>
__device__ __forceinline__ void produce_ids(const DevPlan& P, const ReqRegs& Q, const CtaShared& cs,
                                            int num_cols, int64_t b0, int rows, const float* s_bnd,
                                            int32_t* s_ids, int pw, int lane) {
  for (int cbase = pw; cbase < num_cols; cbase += 4 * kProducerWarps) {
    const int nq = min(4, (num_cols - cbase + kProducerWarps - 1) / kProducerWarps);
    bool fast = nq == 4;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < nq) fast = fast && fast_bucketize(cs.ctx[cbase + q * kProducerWarps]);
    if (fast) {
      float v[8];
      uint32_t bp[4];
      int32_t nb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const ColCtx& c = cs.ctx[cbase + q * kProducerWarps];
        const float* in = (const float*)c.values + b0;
        v[2 * q] = lane < rows ? __ldg(in + lane) : 0.f;
        v[2 * q + 1] = lane + 32 < rows ? __ldg(in + lane + 32) : 0.f;
        bp[q] = smem_u32(s_bnd + c.bnd_off);
        nb[q] = c.num_boundaries;
      }
      int32_t lo[8];
      search_smem8(bp, nb, v, lo);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int cl = cbase + (k >> 1) * kProducerWarps;
        const ColCtx& c = cs.ctx[cl];
        const int r = lane + 32 * (k & 1);
        if (r < rows) {
          int32_t id = lo[k];
          if (CHECK) id = checked_id(id, c.table_rows, Q.status);
          s_ids[cl * kIdPitch + r] = c.stage_off >= 0 ? c.stage_off + id * c.row_bytes : id;
          if (c.ids_out) c.ids_out[b0 + r] = id;
        }
      }
      continue;
    }
    for (int q = 0; q < nq; ++q) {
      const int cl = cbase + q * kProducerWarps;
      const ColCtx& c = cs.ctx[cl];
      const float* bnd = c.bnd_off >= 0 ? s_bnd + c.bnd_off : P.boundaries + c.bnd_pool_off;
      for (int r = lane; r < rows; r += 32) {
        int32_t id = column_id(c, bnd, b0 + r);
        if (CHECK) id = checked_id(id, c.table_rows, Q.status);
        s_ids[cl * kIdPitch + r] = c.stage_off >= 0 ? c.stage_off + id * c.row_bytes : id;
        if (c.ids_out) c.ids_out[b0 + r] = id;
      }
    }
  }
}

// Consumer side of one sub-tile: lane = chunk; 32 / W rows per warp pass; UNROLL independent
// rows in flight.
template <int VEC, bool STAGED>
__device__ __forceinline__ void consume_ids(const ReqRegs& Q, const CtaShared& cs, int W,
                                            const DevChunk& ch, int64_t b0, int rows,
                                            const uint8_t* s_tables, const int32_t* s_ids, int cw,
                                            int lane) {
  const int rpp = 32 / W;  // rows handled by one warp per pass
  const int sub = lane / W;
  if (sub >= rpp) return;
  const uint8_t* tbl;
  int64_t row_bytes = 0;
  if (STAGED) {
    tbl = s_tables + ch.table_byte;
  } else {
    const ColCtx& c = cs.ctx[ch.col_local];
    tbl = c.table + ch.table_byte;
    row_bytes = c.row_bytes;
  }
  const int32_t* ids = s_ids + ch.col_local * kIdPitch;
  const int64_t orb = Q.out_row_bytes;
  const int step = kConsumerWarps * rpp;
  int r = cw * rpp + sub;
  uint8_t* outp = Q.out + (b0 + r) * orb + ch.out_byte;
  const int64_t ostep = (int64_t)step * orb;
  constexpr int UNROLL = 8;
  for (; r + (UNROLL - 1) * step < rows; r += UNROLL * step) {
    Chunk<VEC> v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int32_t o = ids[r + k * step];
      if (STAGED) v[k] = lds_chunk<VEC>(tbl + o);
      else v[k].load(tbl + (int64_t)o * row_bytes);
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      v[k].store(outp);
      outp += ostep;
    }
  }
  for (; r < rows; r += step) {
    const int32_t o = ids[r];
    Chunk<VEC> v;
    if (STAGED) v = lds_chunk<VEC>(tbl + o);
    else v.load(tbl + (int64_t)o * row_bytes);
    v.store(outp);
    outp += ostep;
  }
}

// One DENSE tile (any number of rows), producers and consumers running concurrently.
// `ring_it` counts the sub-tiles handed over since the kernel started (same value in every
// thread); slot = ring_it & 1, and use u of a slot completes phase u of its barriers.
template <int ELEM, bool CHECK>
__device__ __forceinline__ void dense_tile(const DevPlan& P, const ReqRegs& Q, CtaShared& cs,
                                           const DevSlab& slab, const DevChunk& ch, const Tile& tile,
                                           const uint8_t* s_tables, const float* s_bnd,
                                           int32_t* s_ids, int ids_slot, uint32_t stage_parity,
                                           uint32_t& ring_it) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsub = (tile.rows + kSubRows - 1) / kSubRows;
  uint32_t it = ring_it;
  if (warp < kProducerWarps) {
    if (slab.bnd_bytes > 0) mbar_wait(&cs.mbar_bnd, stage_parity);
    for (int j = 0; j < nsub; ++j, ++it) {
      const int slot = it & 1;
      if (it >= 2) mbar_wait(&cs.empty[slot], ((it >> 1) - 1) & 1);
      const int rows = min(kSubRows, tile.rows - j * kSubRows);
      produce_ids<CHECK>(P, Q, cs, slab.num_cols, (int64_t)tile.row_begin + j * kSubRows, rows, s_bnd,
                         s_ids + slot * ids_slot, warp, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&cs.full[slot]);
    }
  } else {
    if (slab.cls == kSlabDenseStaged) mbar_wait(&cs.mbar_tbl, stage_parity);
    const int cw = warp - kProducerWarps;
    for (int j = 0; j < nsub; ++j, ++it) {
      const int slot = it & 1;
      mbar_wait(&cs.full[slot], (it >> 1) & 1);
      const int rows = min(kSubRows, tile.rows - j * kSubRows);
      const int64_t b0 = (int64_t)tile.row_begin + j * kSubRows;
      const int32_t* ids = s_ids + slot * ids_slot;
      if (slab.cls == kSlabDenseStaged) {
        if (slab.vec_bytes == 16) consume_ids<16, true>(Q, cs, slab.num_chunks, ch, b0, rows, s_tables, ids, cw, lane);
        else consume_ids<ELEM, true>(Q, cs, slab.num_chunks, ch, b0, rows, s_tables, ids, cw, lane);
      } else {
        if (slab.vec_bytes == 16) consume_ids<16, false>(Q, cs, slab.num_chunks, ch, b0, rows, s_tables, ids, cw, lane);
        else consume_ids<ELEM, false>(Q, cs, slab.num_chunks, ch, b0, rows, s_tables, ids, cw, lane);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&cs.empty[slot]);
    }
  }
  ring_it += (uint32_t)nsub;
}

// ---------------------------------------------------------------------------
// SPARSE slab: row b owns ids [row_splits[b], row_splits[b+1])
// ---------------------------------------------------------------------------
// The tile's row_splits and row ids are staged in shared memory (`sp`, `stride` ints per
// column: rows + 1 splits, then the ids), so that phase 2's only global accesses are the
// table rows. A column whose ids do not fit (very long rows) falls back to the workspace.
//
// phase 1: one warp per column, so the dependent chains (row_splits -> offsets -> bytes ->
// hash) of different columns overlap.
template <bool CHECK>
__device__ __forceinline__ void sparse_phase1(const DevPlan& P, const ReqRegs& Q, CtaShared& cs,
                                              int num_cols, const Tile& tile, int32_t* sp,
                                              int stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = tile.rows;
  const int64_t b0 = tile.row_begin;
  const int cap = stride - (rows + 1);
  for (int cl = warp; cl < num_cols; cl += kWarps) {
    const ColCtx& c = cs.ctx[cl];
    int32_t* rs_s = sp + cl * stride;
    int32_t* ids_s = rs_s + rows + 1;
    for (int r = lane; r <= rows; r += 32) rs_s[r] = c.row_splits[b0 + r];
    __syncwarp();
    const int32_t e0 = rs_s[0], e1 = rs_s[rows];
    const bool staged = e1 - e0 <= cap;
    if (lane == 0) cs.sp_cnt[cl] = staged ? e1 - e0 : -1;
    const bool xform = c.transform != RECOM_XFORM_NONE;
    if (!staged && !xform && !c.ids_out && !CHECK) continue;
    const float* bnd = P.boundaries + c.bnd_pool_off;
    for (int32_t e = e0 + lane; e < e1; e += 32) {
      int32_t id = column_id(c, bnd, e);
      if (CHECK) id = checked_id(id, c.table_rows, Q.status);
      if (staged) ids_s[e - e0] = id;
      else if (xform || CHECK) c.ws_ids[e] = id;
      if (c.ids_out) c.ids_out[e] = id;
    }
  }
}

template <int VEC, bool BF16, bool CHECK>
__device__ __forceinline__ void sparse_phase2(const ReqRegs& Q, const CtaShared& cs, int W,
                                              const DevChunk& ch, const Tile& tile,
                                              const int32_t* sp, int stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = tile.rows;
  const int64_t b0 = tile.row_begin;
  const int rpp = 32 / W;
  const int sub = lane / W;
  if (sub >= rpp) return;
  const ColCtx& c = cs.ctx[ch.col_local];
  const uint8_t* tbl = c.table + ch.table_byte;
  const int64_t row_bytes = c.row_bytes;
  const int32_t* rs = sp + ch.col_local * stride;
  const bool staged = cs.sp_cnt[ch.col_local] >= 0;
  // id source: smem (relative to the tile's first id), the workspace, or the raw input slot
  const bool from_ws = (c.transform != RECOM_XFORM_NONE) || CHECK;
  const int32_t* ids32 = staged ? rs + rows + 1 - rs[0]
                                : (from_ws ? c.ws_ids : (const int32_t*)c.values);
  const bool ids64 = !staged && !from_ws && c.value_dtype == RECOM_I64;
  const long long* idsll = (const long long*)c.values;
  uint8_t* outp = Q.out + b0 * Q.out_row_bytes + ch.out_byte;
  const int combiner = c.combiner;
  auto id_at = [&](int32_t e) -> int64_t {
    return ids64 ? (int64_t)idsll[e] : (int64_t)ids32[e];
  };
  constexpr int U = BF16 ? 4 : 8;  // independent row gathers in flight per lane

  for (int r = warp * rpp + sub; r < rows; r += kWarps * rpp) {
    const int32_t beg = rs[r], end = rs[r + 1];
    Acc<VEC, BF16> acc;
    acc.clear();
    int32_t e = beg;
    for (; e + U - 1 < end; e += U) {
      Chunk<VEC> v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) v[k].load(tbl + id_at(e + k) * row_bytes);
#pragma unroll
      for (int k = 0; k < U; ++k) acc.add(v[k]);
    }
    if (U == 8 && e + 3 < end) {
      Chunk<VEC> v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k].load(tbl + id_at(e + k) * row_bytes);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc.add(v[k]);
      e += 4;
    }
    {
      // up to 3 left: issue them together as well (named registers: a conditionally indexed
      // array would live in local memory)
      const int left = end - e;
      Chunk<VEC> v0, v1, v2;
      v0.zero();
      v1.zero();
      v2.zero();
      if (left > 0) v0.load(tbl + id_at(e) * row_bytes);
      if (left > 1) v1.load(tbl + id_at(e + 1) * row_bytes);
      if (left > 2) v2.load(tbl + id_at(e + 2) * row_bytes);
      acc.add(v0);
      acc.add(v1);
      acc.add(v2);
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
    acc.pack().store(outp + (int64_t)r * Q.out_row_bytes);
  }
}

// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
// Tile sources of a CTA: ST = its own static DENSE macro tile (Sched::start), SP = the shared
// pool of SPARSE tiles (long DRAM-latency chains), DP = the shared pool of small left-over
// DENSE tiles (load balance). "Sparse first" CTAs walk SP, DP; the others ST, SP, DP, so that
// latency-bound gathers and bandwidth-bound dense stores overlap on every SM.
// Pool claims are issued one tile before they are looked at, so the atomic's latency is not
// waited for (except when a pool runs dry). Lives in thread 0.
struct TileCursor {
  enum Kind : int { SP = 0, DP = 1, DONE = 2 };
  struct Claim {
    int kind;
    int raw;  // the atomic's result, untouched until resolve()
  };
  int kind;
  int num_sparse, num_tiles;

  __device__ __forceinline__ void init(const Sched& S) {
    kind = SP;
    num_sparse = S.num_sparse_tiles;
    num_tiles = S.num_tiles;
  }
  __device__ __forceinline__ Claim issue(const ReqParams& A) {
    if (kind == SP) {
      if (num_sparse > 0) return Claim{SP, atomicAdd(A.counters, 1)};
      kind = DP;
    }
    if (kind == DP) {
      if (num_sparse < num_tiles) return Claim{DP, atomicAdd(A.counters + 1, 1)};
      kind = DONE;
    }
    return Claim{DONE, -1};
  }
  // -> tile index, or -1 when every source is exhausted
  __device__ __forceinline__ int resolve(const ReqParams& A, Claim c) {
    for (;;) {
      if (c.kind == DONE) return -1;
      if (c.kind == SP && c.raw < num_sparse) return c.raw;
      if (c.kind == DP && num_sparse + c.raw < num_tiles) return num_sparse + c.raw;
      if (kind == c.kind) ++kind;  // that pool is dry
      c = issue(A);
    }
  }
};

// Optional per-CTA phase timing (thread 0's view), compiled in only with -DRECOM_PROFILE
// (make prof): slots are summed over the CTA's tiles and written to R->prof.
#ifdef RECOM_PROFILE
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// phase slots count SM cycles (clock64: cheap, exact); slots 14 / 15 hold the CTA's start /
// end in globaltimer ns
#define PROF_DECL                                                       \
  uint32_t prof_acc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   \
  const uint64_t prof_t0 = gtimer();                                    \
  uint32_t prof_last = (uint32_t)clock64();
#define PROF_T(i)                                   \
  if (threadIdx.x == 0) {                           \
    const uint32_t now_ = (uint32_t)clock64();      \
    prof_acc[i] += now_ - prof_last;                \
    prof_last = now_;                               \
  }
#define PROF_COUNT(i) if (threadIdx.x == 0) prof_acc[i] += 1;
#define PROF_END                                                        \
  if (threadIdx.x == 0 && A.prof) {                                     \
    uint64_t* o_ = A.prof + blockIdx.x * 16;                            \
    _Pragma("unroll") for (int i_ = 0; i_ < 14; ++i_) o_[i_] = prof_acc[i_]; \
    o_[14] = prof_t0;                                                   \
    o_[15] = gtimer();                                                  \
  }
#else
#define PROF_DECL
#define PROF_T(i)
#define PROF_COUNT(i)
#define PROF_END
#endif

__device__ __forceinline__ void cp_async_tile(Tile* smem_dst, const Tile* gsrc) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    cp_async_16(reinterpret_cast<uint8_t*>(smem_dst) + 16 * i,
                reinterpret_cast<const uint8_t*>(gsrc) + 16 * i);
}

template <bool BF16, bool CHECK>
__global__ void __launch_bounds__(kThreads, 3)
fused_columns_kernel(const DevPlan P, const Sched S, const ReqParams A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ CtaShared cs;
  constexpr int ELEM = BF16 ? 2 : 4;
  uint8_t* s_tables = smem;
  float* s_bnd = reinterpret_cast<float*>(smem + P.smem_off_bnd);
  int32_t* s_ids = reinterpret_cast<int32_t*>(smem + P.smem_off_ids);
  ReqRegs Q;
  Q.out = A.out;
  Q.out_row_bytes = A.out_row_bytes;
  Q.status = A.status;

  PROF_DECL
  // The first tile: the CTA's static macro tile, read by every thread straight from the
  // schedule (one global round trip, no hand-over), or a claim from the sparse pool.
  Tile tile = S.start[blockIdx.x];
  const bool sparse_first = (int)blockIdx.x < S.sparse_first_ctas || tile.slab_id < 0;
  TileCursor cursor;
  TileCursor::Claim pend{TileCursor::DONE, -1};
  if (threadIdx.x == 0) {
    mbar_init(&cs.mbar_bnd, 1);
    mbar_init(&cs.mbar_tbl, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&cs.full[i], kProducerWarps);
      mbar_init(&cs.empty[i], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cursor.init(S);
    if (sparse_first) {
      const int i0 = cursor.resolve(A, cursor.issue(A));
      if (i0 >= 0) cs.next[1] = S.tiles[i0];
      else cs.next[1].slab_id = -1;
    }
    pend = cursor.issue(A);
  }
  __syncthreads();  // mbarriers initialised (and cs.next[1] written)
  if (sparse_first) {
    tile = cs.next[1];
    __syncthreads();  // read before thread 0 reuses the slot
  }
  PROF_T(13)

  int cur_slab = -1;
  int staged_bnd = -1;        // pool offset of the single boundary set resident in smem, or -1
  uint32_t stage_parity = 1;  // parity of the phase completed by the latest staging
  DevChunk ch;
  ch.col = ch.col_local = ch.table_byte = ch.out_byte = 0;
  int sp_stride = 0;
  uint32_t ring_it = 0;
  const int ids_slot = P.smem_ids_cols * kIdPitch;

  for (int k = 0; tile.slab_id >= 0; ++k) {
    const int buf = k & 1;
    const DevSlab& slab = tile.slab;
    PROF_T(0)
    PROF_COUNT(11)
    if (tile.slab_id != cur_slab) {
      // (every warp has passed the barrier that ends the previous tile)
      cur_slab = tile.slab_id;
      // a slab whose only boundary set is already resident keeps it (model E/F: one set for all)
      const bool keep_bnd = slab.num_bnds == 1 && slab.first_bnd_pool == staged_bnd;
      if (slab.cls == kSlabSparse) staged_bnd = -1;  // a sparse tile overwrites the whole region
      else if (slab.num_bnds == 1) staged_bnd = slab.first_bnd_pool;
      else if (slab.num_bnds > 1) staged_bnd = -1;
      if (slab.stage_bytes + (keep_bnd ? 0 : slab.bnd_bytes) > 0) stage_parity ^= 1;
      enter_slab(P, A, slab, cs, s_tables, s_bnd, keep_bnd);
      {
        const int lane = threadIdx.x & 31;
        ch = P.chunks[slab.first_chunk + lane % slab.num_chunks];
      }
      if (slab.cls == kSlabSparse)
        sp_stride = (P.smem_bytes / 4 / slab.num_cols) & ~3;
      __syncthreads();  // the column contexts are visible
      PROF_COUNT(12)
    }
    PROF_T(1)
    // thread 0: look at the claim made one tile ago, start copying that tile's record into
    // cs.next (read by everyone after the barrier that ends this tile), claim one more.
    // Placed right after a barrier: a __syncthreads waits for outstanding atomics, so the
    // claim must have the whole tile to come back.
    if (threadIdx.x == 0) {
      const int idx = cursor.resolve(A, pend);
      PROF_T(2)
      if (idx >= 0) {
        cp_async_tile(&cs.next[buf], S.tiles + idx);
        pend = cursor.issue(A);
      } else {
        cs.next[buf].slab_id = -1;
        pend = TileCursor::Claim{TileCursor::DONE, -1};
      }
      PROF_T(3)
    }
    if (slab.cls == kSlabSparse) {
      int32_t* sp = reinterpret_cast<int32_t*>(smem);
      sparse_phase1<CHECK>(P, Q, cs, slab.num_cols, tile, sp, sp_stride);
      PROF_T(8)
      __syncthreads();  // staged splits / ids are visible
      PROF_T(9)
      if (slab.vec_bytes == 16)
        sparse_phase2<16, BF16, CHECK>(Q, cs, slab.num_chunks, ch, tile, sp, sp_stride);
      else
        sparse_phase2<ELEM, BF16, CHECK>(Q, cs, slab.num_chunks, ch, tile, sp, sp_stride);
      PROF_T(10)
    } else {
      dense_tile<ELEM, CHECK>(P, Q, cs, slab, ch, tile, s_tables, s_bnd, s_ids, ids_slot, stage_parity,
                              ring_it);
      PROF_T(6)
    }
    if (threadIdx.x == 0) cp_async_wait_all();
    __syncthreads();  // end of tile: the ids ring / staged splits may be reused, cs.next is visible
    PROF_T(7)
    tile = cs.next[buf];
  }
  PROF_END
}

template <bool BF16, bool CHECK>
static int configure_one(int smem_bytes) {
  auto fn = fused_columns_kernel<BF16, CHECK>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) !=
      cudaSuccess)
    return -1;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kThreads, smem_bytes) != cudaSuccess)
    return -1;
  return n;
}

int fused_kernel_configure(bool bf16, int smem_bytes) {
  const int a = bf16 ? configure_one<true, false>(smem_bytes) : configure_one<false, false>(smem_bytes);
  const int b = bf16 ? configure_one<true, true>(smem_bytes) : configure_one<false, true>(smem_bytes);
  return a < b ? a : b;
}

void launch_fused_columns(const DevPlan& plan, const Sched& sched, const ReqParams& req, int grid,
                          bool bf16, cudaStream_t stream) {
  const bool check = plan.check_ids != 0;
  const size_t sm = (size_t)plan.smem_bytes;
  if (bf16) {
    if (check) fused_columns_kernel<true, true><<<grid, kThreads, sm, stream>>>(plan, sched, req);
    else fused_columns_kernel<true, false><<<grid, kThreads, sm, stream>>>(plan, sched, req);
  } else {
    if (check) fused_columns_kernel<false, true><<<grid, kThreads, sm, stream>>>(plan, sched, req);
    else fused_columns_kernel<false, false><<<grid, kThreads, sm, stream>>>(plan, sched, req);
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
    out[i] = num_buckets ? h % num_buckets : h;  // plain 64-bit modulo: checks mod_bucket
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
