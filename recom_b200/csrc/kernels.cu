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
// Work decomposition (v4): a STATIC schedule per batch size, cached on the device. A slab is
// a run of columns whose output slices are adjacent in the concat row and at most 32 16-byte
// chunks wide, so one warp store covers up to 512 contiguous bytes of one output row.
// Persistent CTAs (512 threads, 2 per SM); the warps of a CTA are split into
//   dense warps   - the CTA owns one DENSE run (a row range of one slab; more when there are
//                   more slabs than CTAs). The run's column contexts, bucket boundaries and
//                   small tables are staged in shared memory once (TMA bulk copies, mbarrier
//                   completion); then every warp walks its own 32-row tiles alone:
//                     phase 1  lane = row: coalesced input loads of 8 columns in flight, 8
//                              upper_bound searches in lock step (or hashing, 4 chains at a
//                              time) -> row ids in the warp's shared-memory scratch;
//                     phase 2  lane = 16-byte chunk: 128-bit gathers from the staged tables
//                              (or ld.global.nc), 128-bit streaming stores into the concat row.
//   sparse warps  - a static, interleaved share of the SPARSE (multi-hot) warp tiles: row_splits
//                   and transformed ids of all the slab's columns in the warp's scratch, then
//                   one pipelined gather stream per lane (fp32 accumulation, TF-CPU scaling).
// No block-wide barrier after start-up, no atomics, no per-request upload: latency-bound
// multi-hot gathers and bandwidth-bound dense stores overlap on every SM.
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

// One upper_bound through the set's grid (append_grid in capi.cu; the layout bucket_smem8 reads
// from shared memory): header and cell word narrow the search to the candidates of the value's
// cell - usually one boundary - so a queue tile pays three dependent loads per id instead of
// log2(n) + 1. Exact for every float (the cells were built over the whole key space); NaN
// compares false with everything: the last bucket.
__device__ __forceinline__ int32_t bucketize_grid(const float* __restrict__ bnd, int32_t nb,
                                                  float v) {
  if (nb <= 0) return 0;  // no boundaries, no set in the pool
  const float* hdr = bnd + ((nb + 3) & ~3);
  const float b0 = hdr[0], inv = hdr[1];
  const int32_t cells = __float_as_int(hdr[2]);
  if (cells == 0) return bucketize<float>(bnd, nb, v);  // more than 65535 boundaries: no grid
  int32_t cell = __float2int_rz((v - b0) * inv);
  cell = min(max(cell, 0), cells - 1);
  const uint32_t w = __float_as_uint(hdr[4 + cell]);
  if (v != v) return nb;
  int32_t lo = (int32_t)(w & 0xffffu);
  const int32_t hi = (int32_t)(w >> 16);
  for (int32_t n = hi - lo; n > 0; n >>= 1) {
    const int32_t idx = lo + (n >> 1);
    const float b = bnd[min(idx, nb - 1)];
    const bool right = idx < hi && !(v < b);
    lo = right ? idx + 1 : lo;
  }
  return lo;
}

// ---------------------------------------------------------------------------
// raw value -> row id, N at a time
// ---------------------------------------------------------------------------
// Every id is the end of a chain of dependent memory round trips (value; for strings:
// offsets -> bytes). ids_n runs N chains side by side: stage 1 loads the N raw values, stage 2
// the words of the N strings, stage 3 computes. A warp never waits for one chain alone.

// I32 / F32: the low 32 bits; I64 / F64: 64 bits; STR: begin offset | end offset << 32.
__device__ __forceinline__ uint64_t load_raw(const ColCtx& c, int64_t i) {
  const uint8_t* vals = c.values;
  switch (c.value_dtype) {
    case RECOM_STR: {
      const uint32_t* off = (const uint32_t*)vals;
      const uint32_t o0 = __ldg(off + i), o1 = __ldg(off + i + 1);
      return (uint64_t)o0 | ((uint64_t)o1 << 32);
    }
    case RECOM_I64:
    case RECOM_F64: return (uint64_t)__ldg((const unsigned long long*)vals + i);
    default: return (uint64_t)__ldg((const uint32_t*)vals + i);
  }
}

// The (at most three) aligned 8-byte words that cover a string of 1..16 bytes. Every word
// read holds at least one byte of the string (see fh::MemBytes on slot padding).
struct StrWords {
  uint64_t w0, w1, w2;
};
__device__ __forceinline__ StrWords load_short_words(const uint8_t* p, uint32_t len) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a & ~uintptr_t(7));
  const uint32_t words = ((uint32_t)(a & 7) + len + 7) >> 3;
  StrWords w;
  w.w0 = __ldg(q);
  w.w1 = words > 1 ? __ldg(q + 1) : 0ull;
  w.w2 = words > 2 ? __ldg(q + 2) : 0ull;
  return w;
}
__device__ __forceinline__ fh::RegBytes short_bytes(const StrWords& w, const uint8_t* p) {
  const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 7) * 8;
  fh::RegBytes s;
  if (sh == 0) {
    s.w0 = w.w0;
    s.w1 = w.w1;
  } else {
    s.w0 = (w.w0 >> sh) | (w.w1 << (64 - sh));
    s.w1 = (w.w1 >> sh) | (w.w2 << (64 - sh));
  }
  s.w2 = 0;
  return s;
}

// Stage 3. `bnd` = this column's boundaries (shared or global memory).
__device__ __noinline__ int32_t finish_id(const ColCtx& c, const float* __restrict__ bnd,
                                          uint64_t raw, const StrWords sw) {
  switch (c.transform) {
    case RECOM_XFORM_NONE:
      return c.value_dtype == RECOM_I32 ? (int32_t)(uint32_t)raw : (int32_t)(long long)raw;
    case RECOM_XFORM_BUCKETIZE: {
      switch (c.value_dtype) {
#ifndef RECOM_NO_GRID_BUCKETIZE
        // integers compare as floats (`value < boundary` converts the integer, like TF's
        // upper_bound over float boundaries does): the same search on the converted value
        case RECOM_F32: return bucketize_grid(bnd, c.num_boundaries, __uint_as_float((uint32_t)raw));
        case RECOM_I32: return bucketize_grid(bnd, c.num_boundaries, (float)(int32_t)(uint32_t)raw);
        case RECOM_I64: return bucketize_grid(bnd, c.num_boundaries, (float)(long long)raw);
#else
        case RECOM_F32: return bucketize<float>(bnd, c.num_boundaries, __uint_as_float((uint32_t)raw));
        case RECOM_I32: return bucketize<int32_t>(bnd, c.num_boundaries, (int32_t)(uint32_t)raw);
        case RECOM_I64: return bucketize<long long>(bnd, c.num_boundaries, (long long)raw);
#endif
        default: return bucketize<double>(bnd, c.num_boundaries, __longlong_as_double((long long)raw));
      }
    }
    case RECOM_XFORM_HASH_STR: {
      const uint32_t o0 = (uint32_t)raw, len = (uint32_t)(raw >> 32) - o0;
      const uint8_t* p = c.bytes + o0;
      uint64_t h;
      if (len - 1u < 16u) h = fh::hash_len_0_16(short_bytes(sw, p), len);
      else h = fh::fingerprint64(p, len);  // empty, or longer than 16 bytes
      return (int32_t)fh::mod_bucket(h, (uint64_t)c.num_buckets, c.bucket_magic);
    }
    default: {  // RECOM_XFORM_HASH_INT
      uint64_t h;
      if (c.value_dtype == RECOM_I32) {
        const int32_t v = (int32_t)(uint32_t)raw;
        h = (v >= 0 && v < 100000000) ? fh::fingerprint64_of_decimal_u32small((uint32_t)v)
                                      : fh::fingerprint64_of_decimal((int64_t)v);
      } else {
        h = fh::fingerprint64_of_decimal((int64_t)raw);
      }
      return (int32_t)fh::mod_bucket(h, (uint64_t)c.num_buckets, c.bucket_magic);
    }
  }
}

// Four chains at a time (stage 3, the bulk of the code, is the out-of-line finish_id).
struct Ids4 {
  int32_t id[4];
};
#ifdef RECOM_IDS_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
Ids4 ids4(const ColCtx* c0, const ColCtx* c1, const ColCtx* c2,
                                  const ColCtx* c3, const float* bnd_smem, const float* bnd_pool,
                                  int64_t i0, int64_t i1, int64_t i2, int64_t i3, uint32_t okmask) {
  const ColCtx* c[4] = {c0, c1, c2, c3};
  const int64_t i[4] = {i0, i1, i2, i3};
  uint64_t raw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) raw[k] = ((okmask >> k) & 1u) ? load_raw(*c[k], i[k]) : 0ull;
  StrWords sw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sw[k].w0 = sw[k].w1 = sw[k].w2 = 0ull;
    if (((okmask >> k) & 1u) && c[k]->value_dtype == RECOM_STR) {
      const uint32_t o0 = (uint32_t)raw[k], len = (uint32_t)(raw[k] >> 32) - o0;
      if (len - 1u < 16u) sw[k] = load_short_words(c[k]->bytes + o0, len);
    }
  }
  Ids4 r;
  // Common case, kept in straight-line code so that the four dependent multiply / xor chains
  // interleave in the pipeline: every live chain of every lane hashes a string of 8..16 bytes.
  bool str816 = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t len = (uint32_t)(raw[k] >> 32) - (uint32_t)raw[k];
    str816 = str816 && (!((okmask >> k) & 1u) ||
                        (c[k]->transform == RECOM_XFORM_HASH_STR && len - 8u <= 8u));
  }
  if (__all_sync(0xffffffffu, str816)) {
    uint64_t h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t o0 = (uint32_t)raw[k], len = (uint32_t)(raw[k] >> 32) - o0;
      const fh::RegBytes sb = short_bytes(sw[k], c[k]->bytes + o0);
      h[k] = fh::hash_len_8_16(sb.w0, sb.fetch64(len - 8), len);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      r.id[k] = ((okmask >> k) & 1u)
                    ? (int32_t)fh::mod_bucket(h[k], (uint64_t)c[k]->num_buckets, c[k]->bucket_magic)
                    : 0;
    return r;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r.id[k] = 0;
    if ((okmask >> k) & 1u) {
      const float* bnd = c[k]->bnd_off >= 0 ? bnd_smem + c[k]->bnd_off : bnd_pool + c[k]->bnd_pool_off;
      r.id[k] = finish_id(*c[k], bnd, raw[k], sw[k]);
    }
  }
  return r;
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
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Spin until word g of a completion-flag array has reached `epoch` (wrap-around safe); after
// timeout_ms the last word of the array is set to 1 and the wait gives up.
__device__ __forceinline__ void wait_flag(uint32_t* flags, int g, uint32_t epoch,
                                          uint32_t timeout_ms) {
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + g) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > (uint64_t)timeout_ms * 1000000ull) {
      flags[RECOM_SHARD_FLAG_WORDS - 1] = 1;
      break;
    }
    __nanosleep(64);
  }
}

// Request-wide values every tile needs.
struct ReqRegs {
  uint8_t* out;
  int64_t out_row_bytes;
  int32_t* status;
  int32_t* ids_out;
  const int32_t* ids_base;
  bool check;
};

template <bool SH>
struct ReqRegsT;
template <>
struct ReqRegsT<false> : ReqRegs {
  static constexpr bool kSharded = false;
  __device__ __forceinline__ void bind(const ReqParams&) {}
};
template <>
struct ReqRegsT<true> : ReqRegs {
  static constexpr bool kSharded = true;
  const ReqParams* A;
  __device__ __forceinline__ void bind(const ReqParams& a) { A = &a; }
};

// Where the rows a lane stores go: a cursor over the output rows b0 + r0, + step_rows, ...
// Single GPU: one pointer bumped by a constant. Sharded output (Mode A, fused exchange): the
// rows of a tile (at most 32, never more than rows_per_rank) belong to at most two ranks; the
// cursor runs in the first owner's address space (base pre-biased by the owner's first row) and
// rows at or past the boundary are displaced by the distance between the two owners' bases.
template <bool SH>
struct RowCursor;
template <>
struct RowCursor<false> {
  uint8_t* p;
  int64_t step;
  __device__ __forceinline__ RowCursor(const ReqRegs& Q, int64_t b0, int r0, int step_rows,
                                       int32_t out_byte) {
    p = Q.out + (b0 + r0) * Q.out_row_bytes + out_byte;
    step = (int64_t)step_rows * Q.out_row_bytes;
  }
  __device__ __forceinline__ uint8_t* next() {
    uint8_t* q = p;
    p += step;
    return q;
  }
};
template <>
struct RowCursor<true> {
  uint8_t* p;
  int64_t step;
  int64_t delta;   // second owner's biased base - first owner's
  int32_t left;    // rows until the boundary (<= 0: past it)
  int32_t step_rows;
  __device__ __forceinline__ RowCursor(const ReqRegsT<true>& Q, int64_t b0, int r0,
                                       int step_rows_, int32_t out_byte) {
    const ReqParams& A = *Q.A;
    const int32_t rpr = (int32_t)A.rows_per_rank;
    const int32_t g = (int32_t)b0 / rpr;
    const int32_t g1 = min(g + 1, A.world - 1);
    const int64_t orb = Q.out_row_bytes;
    uint8_t* base0 = A.peer_out[g] - (int64_t)g * rpr * orb;
    uint8_t* base1 = A.peer_out[g1] - (int64_t)g1 * rpr * orb;
    delta = base1 - base0;
    left = (g + 1) * rpr - ((int32_t)b0 + r0);
    step_rows = step_rows_;
    p = base0 + (b0 + r0) * orb + A.col_off_bytes + out_byte;
    step = (int64_t)step_rows * orb;
  }
  __device__ __forceinline__ uint8_t* next() {
    uint8_t* q = left > 0 ? p : p + delta;
    p += step;
    left -= step_rows;
    return q;
  }
};

// Resolve a column against the request (one thread per column): one read of the static
// record plus one read each of the request's slot / table arrays.
__device__ __forceinline__ void build_ctx(const ReqParams& A, const DevSlabCol& sc, bool stage_tbl,
                                          bool stage_bnd, ColCtx& x) {
  x.values = A.blob + A.slot_offsets[sc.value_slot];
  x.bytes = sc.bytes_slot >= 0 ? A.blob + A.slot_offsets[sc.bytes_slot] : nullptr;
  x.table = (const uint8_t*)A.tables[sc.table];
  x.row_splits = nullptr;
  if (sc.sparse_index >= 0)
    x.row_splits = sc.coo_index >= 0 ? A.ws_row_splits + (int64_t)sc.coo_index * (A.batch + 1)
                                     : (const int32_t*)(A.blob + A.slot_offsets[sc.seg_slot]);
  x.bucket_magic = sc.bucket_magic;
  x.num_buckets = sc.num_buckets;
  x.table_rows = sc.table_rows;
  x.row_bytes = sc.row_bytes;
  x.transform = sc.transform;
  x.value_dtype = sc.value_dtype;
  x.num_boundaries = sc.num_boundaries;
  x.bnd_off = stage_bnd ? sc.bnd_off : -1;
  x.bnd_pool_off = sc.boundaries_offset;
  x.stage_off = stage_tbl ? sc.stage_off : -1;
  x.combiner = sc.combiner;
  x.col = sc.col;
  x.sparse_index = sc.sparse_index;
}

// ---------------------------------------------------------------------------
// DENSE run: exactly one id per (column, row). The run's dense warps share the column
// contexts, the bucket boundaries and (run.stage) the slab's small tables in shared memory;
// everything else is private to a warp, which walks its 32-row tiles without any block-wide
// synchronisation.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void enter_run(const DevPlan& P, const ReqParams& A, const Tile& run,
                                          ColCtx* ctx, uint8_t* s_tables, float* s_bnd,
                                          bool keep_bnd, uint64_t* mbar_tbl, uint64_t* mbar_bnd,
                                          int dense_warps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DevSlab& slab = run.slab;
  const bool stage = run.stage != 0 && slab.stage_bytes > 0;
  if (warp == 0) {
    if (lane < slab.num_cols) {
      const DevSlabCol sc = P.slab_cols[slab.first_col + lane];
      ColCtx x;
      build_ctx(A, sc, stage, true, x);
      ctx[lane] = x;
      if (stage && sc.stage_off >= 0 && sc.stage_copy) {
        fence_proxy_async();  // the region was read through the generic proxy by the last run
        bulk_g2s(s_tables + sc.stage_off, x.table, (uint32_t)sc.table_rows * (uint32_t)sc.row_bytes,
                 mbar_tbl);
      }
    }
    // complete_tx of a copy issued above may precede this expect_tx: the phase cannot
    // complete before this (the only) arrival
    if (lane == 0 && stage) mbar_expect_tx(mbar_tbl, (uint32_t)slab.stage_bytes);
  }
  if (warp == (dense_warps > 1 ? 1 : 0) && !keep_bnd && slab.bnd_bytes > 0) {
    if (lane < slab.num_bnds) {
      const DevSlabBnd b = P.slab_bnds[slab.first_bnd + lane];
      fence_proxy_async();
      bulk_g2s(s_bnd + b.smem_off, P.boundaries + b.pool_off, (uint32_t)b.floats * 4, mbar_bnd);
    }
    if (lane == 0) mbar_expect_tx(mbar_bnd, (uint32_t)slab.bnd_bytes);
  }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// 8 upper_bound searches in lock step, one per column, over boundary sets in shared memory.
// The set's grid (see append_grid in capi.cu) narrows every search to the candidates [lo, hi]
// of the value's cell first - usually one boundary - and the same `value < boundary`
// comparisons as TF finish it (see bucketize_n for why `half` serves both sides of a probe).
__device__ __forceinline__ void bucket_smem8(const uint32_t (&bp)[8], const int32_t (&nb)[8],
                                             const float (&v)[8], int32_t (&lo)[8]) {
  int32_t hi[8];
  int32_t nmax = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t hdr = bp[k] + 4u * (uint32_t)((nb[k] + 3) & ~3);
    const float b0 = lds_f32(hdr), inv = lds_f32(hdr + 4);
    const int32_t cells = (int32_t)lds_u32(hdr + 8);
    int32_t cell = __float2int_rz((v[k] - b0) * inv);
    cell = min(max(cell, 0), cells - 1);
    const uint32_t w = lds_u32(hdr + 16 + 4u * (uint32_t)cell);
    const bool nan = v[k] != v[k];  // every comparison is false: the last bucket
    lo[k] = nan ? nb[k] : (int32_t)(w & 0xffffu);
    hi[k] = nan ? nb[k] : (int32_t)(w >> 16);
    nmax = max(nmax, hi[k] - lo[k]);
  }
  for (int32_t n = nmax; n > 0; n >>= 1) {
    const int32_t half = n >> 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int32_t idx = lo[k] + half;
      const float b = lds_f32(bp[k] + 4u * (uint32_t)min(idx, nb[k] - 1));
      const bool right = idx < hi[k] && !(v[k] < b);
      lo[k] = right ? idx + 1 : lo[k];
    }
  }
}

// The same for 8 columns that share ONE boundary set (the common case: a model bucketizes many
// features over the same boundaries): header and set address are read once.
__device__ __forceinline__ void bucket_smem8_uniform(uint32_t bp, int32_t nb, const float (&v)[8],
                                                     int32_t (&lo)[8]) {
  const uint32_t hdr = bp + 4u * (uint32_t)((nb + 3) & ~3);
  const float b0 = lds_f32(hdr), inv = lds_f32(hdr + 4);
  const int32_t cells = (int32_t)lds_u32(hdr + 8);
  int32_t hi[8];
  int32_t nmax = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int32_t cell = __float2int_rz((v[k] - b0) * inv);
    cell = min(max(cell, 0), cells - 1);
    const uint32_t w = lds_u32(hdr + 16 + 4u * (uint32_t)cell);
    const bool nan = v[k] != v[k];
    lo[k] = nan ? nb : (int32_t)(w & 0xffffu);
    hi[k] = nan ? nb : (int32_t)(w >> 16);
    nmax = max(nmax, hi[k] - lo[k]);
  }
  for (int32_t n = nmax; n > 0; n >>= 1) {
    const int32_t half = n >> 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int32_t idx = lo[k] + half;
      const float b = lds_f32(bp + 4u * (uint32_t)min(idx, nb - 1));
      const bool right = idx < hi[k] && !(v[k] < b);
      lo[k] = right ? idx + 1 : lo[k];
    }
  }
}

// phase 1 of a DENSE warp tile. A tile has at most slab.tile_rows rows: 32; 16 for slabs made
// of fast groups only (two lane groups take alternate groups of 8 columns, so a 16-column slab
// needs one pass and its stores start after half the time); 8 for slabs with hashed columns,
// whose ids cost hundreds of instructions each (four lane groups of 4 columns). lane = (column
// group, row). The id slot receives the byte offset of the row inside the smem table region
// when the table is staged, else the raw row id.
// The id slot holds the raw row id; phase 2, whose lanes own one column each, turns it into an
// address. Bounds check and the ids_out parity hook are off the fast path.
__device__ __forceinline__ void put_id(const ReqRegs& Q, const ColCtx& c, int32_t* s_ids, int cl,
                                       int row, int64_t i, int32_t id) {
  if (Q.check | (Q.ids_out != nullptr)) {
    if (Q.check) id = checked_id(id, c.table_rows, Q.status);
    if (Q.ids_out) Q.ids_out[Q.ids_base[c.col] + i] = id;
  }
  s_ids[cl * kIdPitch + row] = id;
}

// FAST run (every column belongs to a fast group): lane group cg takes the groups cg,
// cg + ncg, ...; 8 coalesced loads in flight, 8 searches in lock step.
__device__ __forceinline__ void dense_phase1_fast(const ReqRegs& Q, const ColCtx* ctx,
                                                  const DevSlab& slab, int64_t b0, int rows,
                                                  const float* s_bnd, int32_t* s_ids, int lane) {
  const int nc = slab.num_cols;
  const int TR = slab.tile_rows;
  const int row = lane & (TR - 1), cg = lane / TR, ncg = 32 / TR;
  const bool valid = row < rows;
  const int64_t i = b0 + row;
  for (int cbase = cg * 8; cbase < nc; cbase += ncg * 8) {
    float v[8];
    int32_t lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = valid ? __ldg((const float*)ctx[cbase + k].values + i) : 0.f;
    if ((slab.uni8_mask >> (cbase >> 3)) & 1u) {
      const ColCtx& c0 = ctx[cbase];
      bucket_smem8_uniform(smem_u32(s_bnd + c0.bnd_off), c0.num_boundaries, v, lo);
    } else {
      uint32_t bp[8];
      int32_t nb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const ColCtx& c = ctx[cbase + k];
        bp[k] = smem_u32(s_bnd + c.bnd_off);
        nb[k] = c.num_boundaries;
      }
      bucket_smem8(bp, nb, v, lo);
    }
    if (valid) {
#pragma unroll
      for (int k = 0; k < 8; ++k) put_id(Q, ctx[cbase + k], s_ids, cbase + k, row, i, lo[k]);
    }
  }
}

// Queue tile, any transform: 4 chains per lane; the lane groups take 4 columns each.
__device__ __forceinline__ void dense_phase1_generic(const DevPlan& P, const ReqRegs& Q,
                                                     const ColCtx* ctx, const DevSlab& slab,
                                                     int64_t b0, int rows, int32_t* s_ids,
                                                     int lane) {
  const int nc = slab.num_cols;
  const int TR = slab.tile_rows;
  const int row = lane & (TR - 1), cg = lane / TR, ncg = 32 / TR;
  const bool valid = row < rows;
  const int64_t i = b0 + row;
  for (int cbase = 0; cbase < nc; cbase += 4 * ncg) {
    const int mine = cbase + cg * 4;
    uint32_t okmask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) okmask |= (valid && mine + k < nc) ? (1u << k) : 0u;
    auto cx = [&](int k) { return ctx + (mine + k < nc ? mine + k : 0); };
    const Ids4 r = ids4(cx(0), cx(1), cx(2), cx(3), nullptr, P.boundaries, i, i, i, i, okmask);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((okmask >> k) & 1u) put_id(Q, ctx[mine + k], s_ids, mine + k, row, i, r.id[k]);
  }
}

// phase 2 of a DENSE warp tile: lane = chunk; 32 / W rows per pass; UNROLL independent rows in
// flight; 128-bit gathers from the staged tables (or ld.global.nc), 128-bit streaming stores
// straight into the concat rows.
template <int VEC, bool STAGED, class QT>
__device__ __forceinline__ void dense_phase2_impl(const QT& Q, const ColCtx* ctx, int W,
                                             const DevChunk& ch, int64_t b0, int rows,
                                             const uint8_t* s_tables, const int32_t* s_ids,
                                             int lane) {
#ifdef RECOM_TEST_NODP2
  return;
#endif
  const int rpp = 32 / W;  // rows handled per pass
  const int sub = lane / W;
  if (sub >= rpp) return;
  const ColCtx& c = ctx[ch.col_local];
  const uint8_t* tbl = (STAGED ? s_tables + c.stage_off : c.table) + ch.table_byte;
  const int32_t row_bytes = c.row_bytes;
  const int32_t* ids = s_ids + ch.col_local * kIdPitch;
  int r = sub;
  RowCursor<QT::kSharded> outp(Q, b0, r, rpp, ch.out_byte);
  #ifndef RECOM_DENSE_U
#define RECOM_DENSE_U 8
#endif
  constexpr int UNROLL = RECOM_DENSE_U;
  for (; r + (UNROLL - 1) * rpp < rows; r += UNROLL * rpp) {
    Chunk<VEC> v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int32_t o = ids[r + k * rpp];
      if (STAGED) v[k] = lds_chunk<VEC>(tbl + o * row_bytes);
      else v[k].load(tbl + (int64_t)o * row_bytes);
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      v[k].store(outp.next());
    }
  }
  for (; r < rows; r += rpp) {
    const int32_t o = ids[r];
    Chunk<VEC> v;
    if (STAGED) v = lds_chunk<VEC>(tbl + o * row_bytes);
    else v.load(tbl + (int64_t)o * row_bytes);
    v.store(outp.next());
  }
}

// The 16-byte variants are the hot ones and inlined; element-wise slabs (unaligned slices) are
// rare and kept out of line to keep the hot code small.
template <int VEC, bool STAGED, class QT>
__device__ __noinline__ void dense_phase2_cold(const QT& Q, const ColCtx* ctx, int W,
                                               const DevChunk& ch, int64_t b0, int rows,
                                               const uint8_t* s_tables, const int32_t* s_ids,
                                               int lane) {
  dense_phase2_impl<VEC, STAGED>(Q, ctx, W, ch, b0, rows, s_tables, s_ids, lane);
}
template <int VEC, bool STAGED, class QT>
__device__ __forceinline__ void dense_phase2(const QT& Q, const ColCtx* ctx, int W,
                                             const DevChunk& ch, int64_t b0, int rows,
                                             const uint8_t* s_tables, const int32_t* s_ids,
                                             int lane) {
  if (VEC == 16) dense_phase2_impl<VEC, STAGED>(Q, ctx, W, ch, b0, rows, s_tables, s_ids, lane);
  else dense_phase2_cold<VEC, STAGED>(Q, ctx, W, ch, b0, rows, s_tables, s_ids, lane);
}

// ---------------------------------------------------------------------------
// SPARSE warp tile: row b owns ids [row_splits[b], row_splits[b+1]) of every column of the
// slab. One warp, no block-wide synchronisation. Warp scratch (shared memory):
//   [ColCtx x cols | prefix (cols + 1) | row_splits (cols x (tile_rows + 1)) | ids]
// phase 1 flattens the ids of all columns of the tile into one index space, so that every lane
// has work whatever the row lengths are, and transforms them four chains at a time;
// phase 2 (lane = chunk) walks the ids of its rows as ONE stream with up to U independent
// 128-bit row gathers in flight across row boundaries, fp32 accumulation, TF-CPU scaling, one
// 128-bit store per row. A tile whose ids do not fit the scratch uses the workspace instead.
// ---------------------------------------------------------------------------
// End of a row: combiner scaling and the row's one store. Out of line: the gather stream below
// reaches it from every unrolled slot, and inlining it there would multiply the code size.
template <int VEC, bool BF16>
__device__ __noinline__ void flush_row(Acc<VEC, BF16> acc, int32_t n, int combiner, uint8_t* outp) {
  if (n > 1 && combiner != RECOM_COMBINE_SUM) {
    // TF-CPU scaling (segment_reduction_ops_impl.h:619-629,765-770):
    // n < 10 multiplies by 1/m, n >= 10 divides by m; m = n or sqrt(n).
    const float m = combiner == RECOM_COMBINE_MEAN ? (float)n : (float)__dsqrt_rn((double)n);
    if (n < 10) acc.mul(__fdiv_rn(1.0f, m));
    else acc.div(m);
  }
  acc.pack().store(outp);
}

template <int VEC, bool BF16, class QT>
__device__ __forceinline__ void sparse_phase2_impl(const QT& Q, const ColCtx& c, int W,
                                              const DevChunk& ch, int64_t b0, int rows,
                                              const int32_t* sp, const int32_t* ids, int lane) {
#ifdef RECOM_TEST_NOSP2
  return;
#endif
  const int rpp = 32 / W;
  const int sub = lane / W;
  if (sub >= rpp) return;
  const int rps = (rows + rpp - 1) / rpp;  // rows of this lane group
  int r = sub * rps;
  const int r_end = min(rows, r + rps);
  if (r >= r_end) return;
  const uint8_t* tbl = c.table + ch.table_byte;
  const int64_t row_bytes = c.row_bytes;
  const int combiner = c.combiner;
  RowCursor<QT::kSharded> outp(Q, b0, r, 1, ch.out_byte);
  #ifndef RECOM_SPARSE_U
#define RECOM_SPARSE_U 8
#endif
  constexpr int U = BF16 ? 4 : RECOM_SPARSE_U;
  int32_t e = sp[r];
  const int32_t e_end = sp[r_end];
  int32_t row_beg = e, row_end = sp[r + 1];
  Acc<VEC, BF16> acc;
  acc.clear();
  auto flush = [&]() {
    flush_row<VEC, BF16>(acc, row_end - row_beg, combiner, outp.next());
  };
  while (e < e_end) {
    Chunk<VEC> v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (e + k < e_end) v[k].load(tbl + (int64_t)ids[e + k] * row_bytes);
      else v[k].zero();
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (e + k < e_end) {
        while (e + k >= row_end) {  // the stream enters the next (possibly empty) row
          flush();
          ++r;
          row_beg = row_end;
          row_end = sp[r + 1];
          acc.clear();
        }
        acc.add(v[k]);
      }
    }
    e += U;
  }
  for (;;) {  // the row being accumulated, then the empty rows after the last id
    flush();
    if (++r >= r_end) break;
    row_beg = row_end;
    row_end = sp[r + 1];
    acc.clear();
  }
}

template <int VEC, bool BF16, class QT>
__device__ __noinline__ void sparse_phase2_cold(const QT& Q, const ColCtx& c, int W,
                                                const DevChunk& ch, int64_t b0, int rows,
                                                const int32_t* sp, const int32_t* ids, int lane) {
  sparse_phase2_impl<VEC, BF16>(Q, c, W, ch, b0, rows, sp, ids, lane);
}
template <int VEC, bool BF16, class QT>
__device__ __forceinline__ void sparse_phase2(const QT& Q, const ColCtx& c, int W,
                                              const DevChunk& ch, int64_t b0, int rows,
                                              const int32_t* sp, const int32_t* ids, int lane) {
  if (VEC == 16) sparse_phase2_impl<VEC, BF16>(Q, c, W, ch, b0, rows, sp, ids, lane);
  else sparse_phase2_cold<VEC, BF16>(Q, c, W, ch, b0, rows, sp, ids, lane);
}

// A DENSE tile from the queue (hashed or unstaged slab): like a SPARSE tile, one warp on its
// own - column contexts and row ids in the warp's scratch, rows gathered with ld.global.nc.
// (The tile's slab record has its fast-group masks cleared: boundaries are read from the pool.)
template <bool BF16, class QT>
__device__ __forceinline__ void generic_dense_tile(const DevPlan& P, const ReqParams& A,
                                                   const QT& Q, const Tile& tile,
                                                   bool new_slab, uint8_t* scratch, DevChunk& ch,
                                                   int lane) {
  constexpr int ELEM = BF16 ? 2 : 4;
  const DevSlab& slab = tile.slab;
  const int nc = slab.num_cols;
  ColCtx* ctx = reinterpret_cast<ColCtx*>(scratch);
  int32_t* s_ids = reinterpret_cast<int32_t*>(scratch + ((nc * (int)sizeof(ColCtx) + 15) & ~15));
  if (new_slab) {
    __syncwarp();
    if (lane < nc) {
      ColCtx x;
      build_ctx(A, P.slab_cols[slab.first_col + lane], false, false, x);
      ctx[lane] = x;
    }
    ch = P.chunks[slab.first_chunk + lane % slab.num_chunks];
    __syncwarp();
  }
  dense_phase1_generic(P, Q, ctx, slab, tile.row_begin, tile.rows, s_ids, lane);
  __syncwarp();
  if (slab.vec_bytes == 16)
    dense_phase2<16, false>(Q, ctx, slab.num_chunks, ch, tile.row_begin, tile.rows, nullptr, s_ids, lane);
  else
    dense_phase2<ELEM, false>(Q, ctx, slab.num_chunks, ch, tile.row_begin, tile.rows, nullptr, s_ids, lane);
  __syncwarp();
}

// Optional per-warp timing, compiled in only with -DRECOM_PROFILE (make prof): every warp
// records [start, end] in globaltimer ns plus the cycles it spent per phase.
#ifdef RECOM_PROFILE
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PROF_DECL                                     \
  uint32_t prof_acc[6] = {0, 0, 0, 0, 0, 0};          \
  const uint64_t prof_t0 = gtimer();                  \
  uint32_t prof_last = (uint32_t)clock64();
#define PROF_T(i)                                     \
  {                                                   \
    const uint32_t now_ = (uint32_t)clock64();        \
    prof_acc[i] += now_ - prof_last;                  \
    prof_last = now_;                                 \
  }
#define PROF_END                                                                       \
  if ((threadIdx.x & 31) == 0 && A.prof) {                                             \
    uint64_t* o_ = A.prof + ((size_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * 8;    \
    _Pragma("unroll") for (int i_ = 0; i_ < 6; ++i_) o_[i_] = prof_acc[i_];            \
    o_[6] = prof_t0;                                                                   \
    o_[7] = gtimer();                                                                  \
  }
#else
#define PROF_DECL
#define PROF_T(i)
#define PROF_END
#endif

// Rare paths of a SPARSE tile's phase 1, out of line so that their loads are not hoisted into
// the hot loop: ids of a tile that does not fit the scratch go to the workspace; the parity
// hook receives every id.
__device__ __noinline__ void spill_id(const ReqParams& A, const ReqRegs& Q, const ColCtx& c,
                                      int64_t e, int32_t v, bool in_smem) {
  if (!in_smem) A.ws_ids[A.sparse_base[c.sparse_index] + e] = v;
  if (Q.ids_out) Q.ids_out[Q.ids_base[c.col] + e] = v;
}

template <bool BF16, class QT>
__device__ __forceinline__ void sparse_warp_tile(const DevPlan& P, const ReqParams& A,
                                                 const QT& Q, const Tile& tile, bool new_slab,
                                                 uint8_t* scratch, DevChunk& ch, int lane
#ifdef RECOM_PROFILE
                                                 , uint32_t (&prof_acc)[6], uint32_t& prof_last
#endif
                                                 ) {
  constexpr int ELEM = BF16 ? 2 : 4;
  const DevSlab& slab = tile.slab;
  const int nc = slab.num_cols, rows = tile.rows;
  const int pitch = slab.tile_rows + 1;
  const int64_t b0 = tile.row_begin;
  ColCtx* ctx = reinterpret_cast<ColCtx*>(scratch);
  int32_t* prefix = reinterpret_cast<int32_t*>(scratch + nc * sizeof(ColCtx));
  int32_t* splits = prefix + nc + 1;
  int32_t* ids_s = splits + nc * pitch;
  const int cap = (P.scratch_bytes - (int)(reinterpret_cast<uint8_t*>(ids_s) - scratch)) / 4;
  if (new_slab) {
    __syncwarp();
    if (lane < nc) {
      ColCtx x;
      build_ctx(A, P.slab_cols[slab.first_col + lane], false, false, x);
      ctx[lane] = x;
    }
    ch = P.chunks[slab.first_chunk + lane % slab.num_chunks];
    __syncwarp();
  }
  PROF_T(3)
  // the tile's row_splits, every column at once
  for (int idx = lane; idx < nc * (rows + 1); idx += 32) {
    const int cl = idx / (rows + 1), r = idx - cl * (rows + 1);
    splits[cl * pitch + r] = __ldg(ctx[cl].row_splits + b0 + r);
  }
  __syncwarp();
  // ids per column -> exclusive prefix
  int32_t incl = lane < nc ? splits[lane * pitch + rows] - splits[lane * pitch] : 0;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane < nc) prefix[lane + 1] = incl;
  if (lane == 0) prefix[0] = 0;
  const int32_t total = __shfl_sync(0xffffffffu, incl, 31);
  __syncwarp();
  const bool in_smem = total <= cap;
  PROF_T(4)

  // phase 1: flat index f -> (column, element), four chains per lane
  for (int32_t f0 = 0; f0 < total; f0 += 128) {
    const ColCtx* c[4];
    int64_t ee[4];
    uint32_t okmask = 0;
    int32_t f[4], pos[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[k] = f0 + 32 * k + lane;
      okmask |= f[k] < total ? (1u << k) : 0u;
      pos[k] = 0;
    }
    // largest column with prefix[column] <= f: four searches side by side, five fixed steps
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cand = pos[k] + step;
        const int32_t pv = prefix[min(cand, nc - 1)];
        pos[k] = (cand < nc && pv <= f[k]) ? cand : pos[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cl = ((okmask >> k) & 1u) ? pos[k] : 0;
      c[k] = ctx + cl;
      ee[k] = (int64_t)splits[cl * pitch] + (f[k] - prefix[cl]);
    }
    const Ids4 r = ids4(c[0], c[1], c[2], c[3], nullptr, P.boundaries, ee[0], ee[1], ee[2], ee[3], okmask);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!((okmask >> k) & 1u)) continue;
      int32_t v = r.id[k];
      if (Q.check) v = checked_id(v, c[k]->table_rows, Q.status);
      if (in_smem) ids_s[f0 + 32 * k + lane] = v;
      if (!in_smem | (Q.ids_out != nullptr)) spill_id(A, Q, *c[k], ee[k], v, in_smem);
    }
  }
  __syncwarp();
  PROF_T(5)

  // phase 2
  const int cl = ch.col_local;
  const ColCtx& cc = ctx[cl];
  const int32_t* sp = splits + cl * pitch;
  // ids indexed by the absolute element number e
  const int32_t* ids = in_smem ? ids_s + prefix[cl] - sp[0]
                               : A.ws_ids + A.sparse_base[cc.sparse_index];
  if (slab.vec_bytes == 16)
    sparse_phase2<16, BF16>(Q, cc, slab.num_chunks, ch, b0, rows, sp, ids, lane);
  else
    sparse_phase2<ELEM, BF16>(Q, cc, slab.num_chunks, ch, b0, rows, sp, ids, lane);
  __syncwarp();
}

// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
template <bool BF16, bool SH>
#ifndef RECOM_MIN_BLOCKS
#define RECOM_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kThreads, RECOM_MIN_BLOCKS)
fused_columns_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ Sched S,
                     const __grid_constant__ ReqParams A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t mbar_tbl;  // tables of the staged slab have landed
  __shared__ uint64_t mbar_bnd;  // boundaries of the staged slab have landed
  constexpr int ELEM = BF16 ? 2 : 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* s_tables = smem;
  float* s_bnd = reinterpret_cast<float*>(smem + P.smem_off_bnd);
  ColCtx* s_ctx = reinterpret_cast<ColCtx*>(smem + P.smem_off_ctx);
  uint8_t* scratch = smem + P.smem_off_scratch + warp * P.scratch_bytes;
  ReqRegsT<SH> Q;
  Q.out = A.out;
  Q.bind(A);
  Q.out_row_bytes = A.out_row_bytes;
  Q.status = A.status;
  Q.ids_out = A.ids_out;
  Q.ids_base = A.ids_base;
  Q.check = P.check_ids != 0;
  PROF_DECL

  if (threadIdx.x == 0) {
    mbar_init(&mbar_tbl, 1);
    mbar_init(&mbar_bnd, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int D = kWarps - S.queue_warps;  // dense warps
  DevChunk ch;
  ch.col = ch.col_local = ch.table_byte = ch.out_byte = 0;

  if (warp < D) {
    // ---- dense warps: the CTA's FAST runs, in order ----
    const CtaWork cw = S.cta[blockIdx.x];
    named_bar_sync(1, D * 32);  // the mbarriers are initialised (thread 0 is a dense thread)
    uint32_t tbl_par = 0, bnd_par = 0;  // parity of the phase the next staging completes
    int staged_bnd = -1;  // pool offset of the single boundary set resident in smem, or -1
    int32_t* s_ids = reinterpret_cast<int32_t*>(scratch);
    for (int ri = 0; ri < cw.count; ++ri) {
      const Tile run = S.runs[cw.first + ri];
      const DevSlab& slab = run.slab;
      if (ri > 0) named_bar_sync(1, D * 32);  // every dense warp is done with the previous run
      // a run whose only boundary set is already resident keeps it (model E/F: one set for all)
      const bool keep_bnd = slab.num_bnds == 1 && slab.first_bnd_pool == staged_bnd;
      if (slab.num_bnds == 1) staged_bnd = slab.first_bnd_pool;
      else if (slab.num_bnds > 1) staged_bnd = -1;
      const bool wait_bnd = !keep_bnd && slab.bnd_bytes > 0;
      const bool staged = run.stage != 0 && slab.stage_bytes > 0;
      enter_run(P, A, run, s_ctx, s_tables, s_bnd, keep_bnd, &mbar_tbl, &mbar_bnd, D);
      ch = P.chunks[slab.first_chunk + lane % slab.num_chunks];
      named_bar_sync(1, D * 32);  // the column contexts are visible
      PROF_T(0)
      if (wait_bnd) {
        mbar_wait(&mbar_bnd, bnd_par);
        bnd_par ^= 1;
      }
      bool tbl_ready = !staged;
      // the run's rows are split evenly into run_tiles tiles of run_tile_rows (<= tile_rows)
      // rows, so that every dense warp of the CTA gets the same number of (smaller) tiles
      const int rpt = slab.run_tile_rows;
      for (int t = warp; t < slab.run_tiles; t += D) {
        const int64_t b0 = (int64_t)run.row_begin + (int64_t)t * rpt;
        const int rows = min(rpt, run.rows - t * rpt);
#ifdef RECOM_TEST_NOP1
        for (int q = lane; q < slab.num_cols * kIdPitch; q += 32) s_ids[q] = (q & 63) * 32;
#else
        dense_phase1_fast(Q, s_ctx, slab, b0, rows, s_bnd, s_ids, lane);
#endif
        __syncwarp();
        PROF_T(1)
        if (!tbl_ready) {
          mbar_wait(&mbar_tbl, tbl_par);
          tbl_ready = true;
        }
        if (staged) {
          if (slab.vec_bytes == 16) dense_phase2<16, true>(Q, s_ctx, slab.num_chunks, ch, b0, rows, s_tables, s_ids, lane);
          else dense_phase2<ELEM, true>(Q, s_ctx, slab.num_chunks, ch, b0, rows, s_tables, s_ids, lane);
        } else {
          if (slab.vec_bytes == 16) dense_phase2<16, false>(Q, s_ctx, slab.num_chunks, ch, b0, rows, s_tables, s_ids, lane);
          else dense_phase2<ELEM, false>(Q, s_ctx, slab.num_chunks, ch, b0, rows, s_tables, s_ids, lane);
        }
        __syncwarp();
        PROF_T(2)
      }
      if (staged) tbl_par ^= 1;
    }
  }

  // ---- every warp: the queue. Queue warps start with a static tile; a claim is always made
  // one tile ahead, so its round trip is hidden behind the tile being processed. ----
  {
    const int nq = S.num_queue;
    const int first_claim = (int)gridDim.x * S.queue_warps;  // tiles below it are static firsts
    // Queue claims (Sched::claim_ahead). 0, the default: a warp claims its next tile when it is
    // free, so the warps that finish early take what is left (the exposed atomic round trip,
    // ~1 us per tile, costs less than the imbalance: model E 48.5 -> 41.8 us, model F 47.9 ->
    // 40.4). 1: always one tile ahead (round trip hidden behind the tile, but every warp holds a
    // tile in reserve that nobody else can take). 2: ahead except within the last grid * warps
    // tiles of the queue.
    const int horizon = S.claim_ahead == 1 ? 0x7fffffff
                        : S.claim_ahead == 0 ? 0 : nq - (int)gridDim.x * kWarps;
    int t = warp >= D ? (warp - D) * (int)gridDim.x + (int)blockIdx.x : nq;
    int ahead = nq;
    if (nq > 0 && warp < D) {  // no static first tile: the claim is the first tile
      if (lane == 0) ahead = first_claim + atomicAdd(A.counters, 1);
      t = __shfl_sync(0xffffffffu, ahead, 0);
    }
    int cur_slab = -1;
    while (t < nq) {
      const bool pre = t < horizon;
      if (pre && lane == 0) ahead = first_claim + atomicAdd(A.counters, 1);
      const Tile tile = S.queue[t];
      const bool new_slab = tile.slab_id != cur_slab;
      cur_slab = tile.slab_id;
      PROF_T(0)
      if (tile.slab.cls == kSlabSparse) {
#ifdef RECOM_PROFILE
        sparse_warp_tile<BF16>(P, A, Q, tile, new_slab, scratch, ch, lane, prof_acc, prof_last);
#else
        sparse_warp_tile<BF16>(P, A, Q, tile, new_slab, scratch, ch, lane);
#endif
      } else {
        generic_dense_tile<BF16>(P, A, Q, tile, new_slab, scratch, ch, lane);
      }
      PROF_T(2)
      if (!pre && lane == 0) ahead = first_claim + atomicAdd(A.counters, 1);
      t = __shfl_sync(0xffffffffu, ahead, 0);
    }
    // the last warp of the grid to get here leaves the slot clean for the next launch and,
    // with a sharded output, tells every peer that this rank's slices of its rows are in place
    const bool signal = SH && A.peer_flags[0] != nullptr;
    if (signal) {
      __syncwarp();  // the lanes' row stores are ordered before lane 0's fence
      if (lane == 0) __threadfence_system();
    }
    if (nq > 0 || signal) {
      bool last = false;
      if (lane == 0) {
        last = atomicAdd(A.counters + 1, 1) == (int)gridDim.x * kWarps - 1;
        if (last) {
          A.counters[0] = 0;
          A.counters[1] = 0;
        }
      }
      if (signal && __shfl_sync(0xffffffffu, (int)last, 0)) {
        __threadfence_system();
        if (lane < A.world)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.peer_flags[lane] + A.rank),
                       "r"(A.epoch)
                       : "memory");
        // the barrier inside the kernel: this launch ends when every rank's slices of this
        // rank's rows have landed (lane g watches rank g's word of the local flag array)
        if (A.wait_in_kernel && lane < A.world)
          wait_flag(A.peer_flags[A.rank], lane, A.epoch, A.wait_timeout_ms);
      }
    }
  }
  PROF_END
}

template <bool BF16>
static int configure_one(int smem_bytes) {
  auto fn = fused_columns_kernel<BF16, false>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) !=
          cudaSuccess ||
      cudaFuncSetAttribute(fused_columns_kernel<BF16, true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return -1;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kThreads, smem_bytes) != cudaSuccess)
    return -1;
  return n;
}

int fused_kernel_configure(bool bf16, int smem_bytes) {
  return bf16 ? configure_one<true>(smem_bytes) : configure_one<false>(smem_bytes);
}

void launch_fused_columns(const DevPlan& plan, const Sched& sched, const ReqParams& req, int grid,
                          bool bf16, cudaStream_t stream) {
  const size_t sm = (size_t)plan.smem_bytes;
#ifdef RECOM_TEST_NOLAUNCH
  (void)sm; (void)sched; (void)req; (void)grid; (void)stream; (void)bf16;
  return;
#endif
  const bool sh = req.world > 0;  // sharded output: rows stored into their owners' tensors
  if (bf16) {
    if (sh) fused_columns_kernel<true, true><<<grid, kThreads, sm, stream>>>(plan, sched, req);
    else fused_columns_kernel<true, false><<<grid, kThreads, sm, stream>>>(plan, sched, req);
  } else {
    if (sh) fused_columns_kernel<false, true><<<grid, kThreads, sm, stream>>>(plan, sched, req);
    else fused_columns_kernel<false, false><<<grid, kThreads, sm, stream>>>(plan, sched, req);
  }
}

// Stream-side end of the fused exchange: lane g waits for rank g's completion signal.
__global__ void __launch_bounds__(32)
shard_wait_kernel(uint32_t* flags, int world, uint32_t epoch, uint32_t timeout_ms) {
  const int g = threadIdx.x;
  if (g < world) wait_flag(flags, g, epoch, timeout_ms);
}

void launch_shard_wait(uint32_t* local_flags, int world, uint32_t epoch, uint32_t timeout_ms,
                       cudaStream_t stream) {
  shard_wait_kernel<<<1, 32, 0, stream>>>(local_flags, world, epoch, timeout_ms);
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
coo_to_row_splits_kernel(const DevPlan P, const ReqHeader* __restrict__ R,
                         const uint8_t* __restrict__ blob, int32_t* __restrict__ ws_row_splits) {
  const int k = blockIdx.y;
  const int ci = req_coo_cols(R)[k];
  const DevColumn c = P.cols[ci];
  const int32_t* ids_base = req_ids_base(R);
  const int64_t nnz = ids_base[ci + 1] - ids_base[ci];
  const long long* coo = (const long long*)(blob + req_slot_offsets(R)[c.seg_slot]);
  int32_t* rs = ws_row_splits + (int64_t)k * (R->batch + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz;
       i += (int64_t)gridDim.x * blockDim.x)
    fill_row_splits(coo, nnz, c.seg_stride, R->batch, rs, i);
}

void launch_coo_to_row_splits(const DevPlan& plan, const ReqHeader* d_req, const uint8_t* blob,
                              int32_t* ws_row_splits, int num_coo, int64_t max_nnz,
                              cudaStream_t stream) {
  int64_t bx = (max_nnz + 1 + 255) / 256;
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)num_coo);
  coo_to_row_splits_kernel<<<grid, 256, 0, stream>>>(plan, d_req, blob, ws_row_splits);
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
