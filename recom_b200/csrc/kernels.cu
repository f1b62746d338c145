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
// Work decomposition (v7): a STATIC schedule per (batch size, schedule kind,
// density class), cached on the device. A slab is a run of columns whose output slices are adjacent
// in the concat row and at most 32 16-byte chunks wide, so one warp store covers up to 512 contiguous
// bytes of one output row.
//   fused_columns_kernel  persistent CTAs (320 threads, 2 per SM). Every CTA owns
//     FAST runs    - a row range of a slab of float32-bucketize columns over small tables: column
//                    contexts, boundary sets (with their grid index) and tables staged in shared
//                    memory once per CTA (TMA bulk copies, mbarrier completion); the warps take 16-row
//                    tiles from a shared-memory counter: phase 1 lane = (column group, row): values by
//                    cp.async, 8 exact upper_bound searches in lock step -> row ids in the warp's
//                    scratch; phase 2 lane = 16-byte chunk: 128-bit loads from the staged tables,
//                    128-bit streaming stores into the concat row;
//     column tiles - everything else (hashed, multi-hot, unstaged, filtered): one id stream x rows per
//                    warp (column_tile): transform (FarmHash / decimal rendering / bucketize / filters)
//                    -> row ids in the scratch -> 128-bit row gathers, fp32 accumulation in id order,
//                    TF-CPU scaling, one store per row - once per column that reads the stream.
//     No global atomics, no fences, no per-request upload.
//   column_tiles_kernel   the flat schedule: the same column_tile code, one warp per tile, 4-warp CTAs,
//                    in front of the persistent kernel (which then keeps the FAST runs) when the column
//                    tiles outweigh the FAST slabs (capi.cu prefer_flat).
//   long_rows_kernel      rows of hundreds to thousands of ids that the flat kernel defers: one 8-warp
//                    CTA per row, partial sums added in warp order.
// Launches are chained with programmatic dependent launch where one does not read what the other
// writes: the persistent kernel behind the flat one, either behind the COO helper (tiles over COO
// inputs wait), and consecutive plain requests on a stream (reads run ahead, every tile waits in
// front of its first store: REQUEST_PDL_WAIT).
//
// The path is HBM / L2 bound integer+copy work: no tensor cores by design (the dense layer behind the
// concat, dense_layer.cu, is the tensor-core kernel of this library).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

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
  static constexpr int kFloats = 4;
  float a[4];
  __device__ __forceinline__ float& at(int i) { return a[i]; }
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
  static constexpr int kFloats = 8;
  float a[8];
  __device__ __forceinline__ float& at(int i) { return a[i]; }
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
  static constexpr int kFloats = 1;
  float a;
  __device__ __forceinline__ float& at(int) { return a; }
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
  static constexpr int kFloats = 1;
  float a;
  __device__ __forceinline__ float& at(int) { return a; }
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

// A chunk read from shared memory (32-bit shared address).
template <int VEC>
__device__ __forceinline__ Chunk<VEC> lds_chunk(uint32_t addr);
template <>
__device__ __forceinline__ Chunk<16> lds_chunk<16>(uint32_t addr) {
  Chunk<16> c;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(c.v.x), "=r"(c.v.y), "=r"(c.v.z), "=r"(c.v.w)
               : "r"(addr));
  return c;
}
template <>
__device__ __forceinline__ Chunk<4> lds_chunk<4>(uint32_t addr) {
  Chunk<4> c;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.v) : "r"(addr));
  return c;
}
template <>
__device__ __forceinline__ Chunk<2> lds_chunk<2>(uint32_t addr) {
  Chunk<2> c;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c.v) : "r"(addr));
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
// transforms: the raw values of an id stream -> row ids
// ---------------------------------------------------------------------------
// A column tile covers a contiguous range [e0, e1) of ONE stream, so a warp runs one transform
// without divergence and its loads are coalesced: lane = consecutive ids, 4 x 32 ids per pass,
// four independent chains per lane (value; for strings: offsets -> words -> hash). The ids go
// to the warp's scratch in shared memory (s_ids[e - e0]) for the gather that follows.

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

// The hashes themselves live out of line, ONE copy each: inlined into the four unrolled chains
// of a transform they were 13 KB of straight-line code that every warp runs once per tile - a
// quarter of the kernel's stall samples were instruction-cache misses (no_instruction) there.
// The loads in front of them stay batched; 20 warps per SM interleave the multiply chains.
__device__ __noinline__ uint64_t hash_8_16_ool(uint64_t first, uint64_t last, uint32_t len) {
  return fh::hash_len_8_16(first, last, len);
}
__device__ __noinline__ uint64_t hash_decimal_small_ool(uint32_t v) {
  return fh::fingerprint64_of_decimal_u32small(v);
}
__device__ __noinline__ uint64_t hash_decimal_ool(long long v) { return fh::fingerprint64_of_decimal((int64_t)v); }

// Any length (rare path of the string transform, out of line).
__device__ __noinline__ uint64_t hash_any(const uint8_t* p, uint32_t len, const StrWords sw) {
  if (len - 1u < 16u) return fh::hash_len_0_16(short_bytes(sw, p), len);
  return fh::fingerprint64(p, len);  // empty, or longer than 16 bytes
}

__device__ __forceinline__ void sts_i32(uint32_t addr, int32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Pre-lookup filters on raw integer values (recom_column_filter; the reference's
// Addons>GatherIndiceValue / Addons>SelectValue, gather_indice_value_ops.cc:33-70,
// select_value_ops.cc:33-56). -> false: the id is dropped. `f` is warp uniform (one stream per
// tile), null for no filter.
__device__ __forceinline__ bool in_intervals(const DevFilter* f, long long v) {
  bool in = false;
  for (int i = 0; i < f->num_intervals; ++i) in = in || (v >= f->lo[i] && v <= f->hi[i]);
  return in;
}
template <typename T>
__device__ __forceinline__ bool filter_value(const DevFilter* f, T& v) {
  if (!f) return true;
  if (f->kind == RECOM_FILTER_KEEP) return in_intervals(f, (long long)v);
  if (f->kind == RECOM_FILTER_SELECT && !in_intervals(f, (long long)v)) v = (T)f->substitute;
  return true;
}
// Addons>GatherStrValueGenIndice (gather_str_value_gen_indice_ops.cc:31-58): drop the strings equal
// to the filter's constant.
__device__ __noinline__ bool string_equals(const DevFilter* f, const uint8_t* p, uint32_t len) {
  if ((int32_t)len != f->not_equal_len) return false;
  for (uint32_t i = 0; i < len; ++i)
    if (__ldg(p + i) != f->not_equal[i]) return false;
  return true;
}

// TF StringToHashBucketFast over ids [e0, e1) of a packed string slot. s_ids: shared-memory
// address of the id of e0.
__device__ __forceinline__ void xform_hash_str(const uint32_t* __restrict__ off,
                                               const uint8_t* __restrict__ bytes, uint32_t nbk,
                                               uint64_t magic, const DevFilter* flt, int64_t e0,
                                               int64_t e1, uint32_t s_ids, int lane) {
  for (int64_t base = e0 + lane; base < e1 + lane; base += 128) {
    uint32_t o0[4], len[4];
    uint32_t okmask = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t e = base + 32 * k;
      const bool ok = e < e1;
      okmask |= ok ? (1u << k) : 0u;
      o0[k] = ok ? __ldg(off + e) : 0u;
      len[k] = (ok ? __ldg(off + e + 1) : 0u) - o0[k];
    }
    StrWords sw[4];
    bool all816 = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sw[k].w0 = sw[k].w1 = sw[k].w2 = 0ull;
      const bool ok = (okmask >> k) & 1u;
      if (ok && len[k] - 1u < 16u) sw[k] = load_short_words(bytes + o0[k], len[k]);
      all816 = all816 && (!ok || len[k] - 8u <= 8u);
    }
    uint64_t h[4];
    // Common case, kept in straight-line code so that the four dependent multiply / xor chains
    // interleave in the pipeline: every live string of every lane has 8..16 bytes.
    if (__all_sync(0xffffffffu, all816)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        h[k] = 0ull;
        if (base - lane + 32 * k >= e1) continue;  // (warp uniform) a short tile: no live string in this chain
        const fh::RegBytes sb = short_bytes(sw[k], bytes + o0[k]);
        h[k] = hash_8_16_ool(sb.w0, sb.fetch64(len[k] - 8), len[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        h[k] = ((okmask >> k) & 1u) ? hash_any(bytes + o0[k], len[k], sw[k]) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((okmask >> k) & 1u) {
        int32_t id = (int32_t)fh::mod_bucket(h[k], nbk, magic);
        if (flt && flt->kind == RECOM_FILTER_STR_NE && string_equals(flt, bytes + o0[k], len[k])) id = kDroppedId;
        sts_i32(s_ids + 4u * (uint32_t)(base + 32 * k - e0), id);
      }
  }
}

// AsString(int) + StringToHashBucketFast (feature_column_v2.py:3414-3419).
template <typename T>
__device__ __forceinline__ void xform_hash_int(const T* __restrict__ vals, uint32_t nbk,
                                               uint64_t magic, const DevFilter* flt, int64_t e0,
                                               int64_t e1, uint32_t s_ids, int lane) {
  for (int64_t base = e0 + lane; base < e1 + lane; base += 128) {
    T v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + 32 * k < e1 ? __ldg(vals + base + 32 * k) : (T)0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (base + 32 * k >= e1) continue;
      if (!filter_value(flt, v[k])) {
        sts_i32(s_ids + 4u * (uint32_t)(base + 32 * k - e0), kDroppedId);
        continue;
      }
      const uint64_t h = (v[k] >= 0 && v[k] < 100000000) ? hash_decimal_small_ool((uint32_t)v[k])
                                                         : hash_decimal_ool((long long)v[k]);
      sts_i32(s_ids + 4u * (uint32_t)(base + 32 * k - e0), (int32_t)fh::mod_bucket(h, nbk, magic));
    }
  }
}

// Four upper_bound searches over one boundary set in lock step, through the set's grid (see
// bucketize_grid): the header is read once per tile.
__device__ __forceinline__ void bucket_grid4(const float* __restrict__ bnd, int32_t nb, float b0,
                                             float inv, int32_t cells, const float* __restrict__ hdr,
                                             const float (&v)[4], int32_t (&lo)[4]) {
  int32_t hi[4];
  int32_t nmax = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int32_t cell = __float2int_rz((v[k] - b0) * inv);
    cell = min(max(cell, 0), cells - 1);
    const uint32_t w = __float_as_uint(__ldg(hdr + 4 + cell));
    const bool nan = v[k] != v[k];  // every comparison is false: the last bucket
    lo[k] = nan ? nb : (int32_t)(w & 0xffffu);
    hi[k] = nan ? nb : (int32_t)(w >> 16);
    nmax = max(nmax, hi[k] - lo[k]);
  }
  for (int32_t n = nmax; n > 0; n >>= 1) {
    const int32_t half = n >> 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int32_t idx = lo[k] + half;
      const float b = __ldg(bnd + min(idx, nb - 1));
      const bool right = idx < hi[k] && !(v[k] < b);
      lo[k] = right ? idx + 1 : lo[k];
    }
  }
}

// TF Bucketize. Integers compare as floats (`value < boundary` converts the integer, like
// std::upper_bound over float boundaries does in bucketize_op.cc:42-44); doubles compare as
// doubles against the widened boundary.
template <typename T>
__device__ __forceinline__ void xform_bucketize(const T* __restrict__ vals,
                                                const float* __restrict__ bnd, int32_t nb,
                                                const DevFilter* flt, int64_t e0, int64_t e1,
                                                uint32_t s_ids, int lane) {
  const float* hdr = bnd + ((nb + 3) & ~3);
  float b0 = 0.f, inv = 0.f;
  int32_t cells = 0;
  if (nb > 0) {
    b0 = __ldg(hdr);
    inv = __ldg(hdr + 1);
    cells = __float_as_int(__ldg(hdr + 2));
  }
  for (int64_t base = e0 + lane; base < e1 + lane; base += 128) {
    T v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + 32 * k < e1 ? __ldg(vals + base + 32 * k) : (T)0;
    bool keep[4] = {true, true, true, true};
    if (flt && (std::is_same<T, int32_t>::value || std::is_same<T, long long>::value)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) keep[k] = filter_value(flt, v[k]);
    }
    int32_t id[4];
    if (std::is_same<T, double>::value) {
#pragma unroll
      for (int k = 0; k < 4; ++k) id[k] = bucketize<T>(bnd, nb, v[k]);
    } else if (cells == 0) {  // no boundaries, or more than 65535 of them: plain search
#pragma unroll
      for (int k = 0; k < 4; ++k) id[k] = nb > 0 ? bucketize<float>(bnd, nb, (float)v[k]) : 0;
    } else {
      float f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = (float)v[k];
      bucket_grid4(bnd, nb, b0, inv, cells, hdr, f, id);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (base + 32 * k < e1) sts_i32(s_ids + 4u * (uint32_t)(base + 32 * k - e0), keep[k] ? id[k] : kDroppedId);
  }
}

// Transform NONE: the values are the ids.
template <typename T>
__device__ __forceinline__ void xform_copy(const T* __restrict__ vals, const DevFilter* flt, int64_t e0,
                                           int64_t e1, uint32_t s_ids, int lane) {
  for (int64_t base = e0 + lane; base < e1 + lane; base += 128) {
    T v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + 32 * k < e1 ? __ldg(vals + base + 32 * k) : (T)0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (base + 32 * k < e1) {
        const bool keep = filter_value(flt, v[k]);
        sts_i32(s_ids + 4u * (uint32_t)(base + 32 * k - e0), keep ? (int32_t)v[k] : kDroppedId);
      }
  }
}

// ids [e0, e1) of the tile's stream -> s_ids[0, e1 - e0). The kinds are out of line: a warp
// runs exactly one of them per tile, and the hot code stays small.
__device__ __noinline__ void run_hash_str(const uint8_t* values, const uint8_t* bytes, uint32_t nbk,
                                          uint64_t magic, const DevFilter* flt, int64_t e0, int64_t e1,
                                          uint32_t s_ids, int lane) {
  xform_hash_str((const uint32_t*)values, bytes, nbk, magic, flt, e0, e1, s_ids, lane);
}
__device__ __noinline__ void run_hash_int(const uint8_t* values, bool i64, uint32_t nbk,
                                          uint64_t magic, const DevFilter* flt, int64_t e0, int64_t e1,
                                          uint32_t s_ids, int lane) {
  if (i64) xform_hash_int((const long long*)values, nbk, magic, flt, e0, e1, s_ids, lane);
  else xform_hash_int((const int32_t*)values, nbk, magic, flt, e0, e1, s_ids, lane);
}
__device__ __noinline__ void run_bucketize(const uint8_t* values, int32_t dtype, const float* bnd,
                                           int32_t nb, const DevFilter* flt, int64_t e0, int64_t e1,
                                           uint32_t s_ids, int lane) {
  switch (dtype) {
    case RECOM_F32: xform_bucketize((const float*)values, bnd, nb, flt, e0, e1, s_ids, lane); break;
    case RECOM_I32: xform_bucketize((const int32_t*)values, bnd, nb, flt, e0, e1, s_ids, lane); break;
    case RECOM_I64: xform_bucketize((const long long*)values, bnd, nb, flt, e0, e1, s_ids, lane); break;
    default: xform_bucketize((const double*)values, bnd, nb, flt, e0, e1, s_ids, lane); break;
  }
}
__device__ __noinline__ void run_copy(const uint8_t* values, bool i64, const DevFilter* flt, int64_t e0,
                                      int64_t e1, uint32_t s_ids, int lane) {
  if (i64) xform_copy((const long long*)values, flt, e0, e1, s_ids, lane);
  else xform_copy((const int32_t*)values, flt, e0, e1, s_ids, lane);
}

__device__ __forceinline__ void transform_ids(const DevPlan& P, const ColItem& it,
                                              const uint8_t* values, const uint8_t* bytes,
                                              int64_t e0, int64_t e1, uint32_t s_ids, int lane) {
  const DevFilter* flt = it.filter >= 0 ? P.filters + it.filter : nullptr;
  switch (it.transform) {
    case RECOM_XFORM_HASH_STR:
      run_hash_str(values, bytes, (uint32_t)it.num_buckets, it.bucket_magic, flt, e0, e1, s_ids, lane);
      break;
    case RECOM_XFORM_HASH_INT:
      run_hash_int(values, it.value_dtype == RECOM_I64, (uint32_t)it.num_buckets, it.bucket_magic,
                   flt, e0, e1, s_ids, lane);
      break;
    case RECOM_XFORM_BUCKETIZE:
      run_bucketize(values, it.value_dtype, P.boundaries + it.boundaries_offset,
                    it.num_boundaries, flt, e0, e1, s_ids, lane);
      break;
    default:
      run_copy(values, it.value_dtype == RECOM_I64, flt, e0, e1, s_ids, lane);
      break;
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

// See fused_columns_kernel: with A.request_pdl this grid started under the previous request's tail.
#define REQUEST_PDL_WAIT(A) \
  if ((A).request_pdl) asm volatile("griddepcontrol.wait;" ::: "memory")

// Request-wide values every tile needs.
struct ReqRegs {
  uint8_t* out;
  int64_t out_row_bytes;
  int32_t* status;
  int32_t* ids_out;
  const int32_t* ids_base;
  bool check;
  // Mode B (table rows sharded over `shard_world` ranks, row r on rank r % world at local row
  // r / world): ids of other ranks' rows are skipped, the plan's combiners are SUM and the
  // caller scales after its cross-rank reduction. 1: whole tables.
  int32_t shard_world, shard_rank;
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


// Resolve a column of a FAST slab against the request (one thread per column): one read of the
// static record plus one read each of the request's slot / table arrays.
__device__ __forceinline__ void build_ctx(const ReqParams& A, const DevSlabCol& sc, bool stage_tbl,
                                          bool stage_bnd, ColCtx& x) {
  x.values = A.blob + A.slot_offsets[sc.value_slot];
  x.bytes = sc.bytes_slot >= 0 ? A.blob + A.slot_offsets[sc.bytes_slot] : nullptr;
  x.table = (const uint8_t*)A.tables[sc.table];
  x.row_splits = nullptr;
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
  x.pad = 0;
}

// ---------------------------------------------------------------------------
// FAST run: a row range of a slab whose columns are all float32 bucketize over boundaries and
// tables small enough for shared memory. The CTA's warps share the column contexts, the bucket
// boundaries and the slab's tables (staged by TMA bulk copies, mbarrier completion); everything
// else is private to a warp, which works on its 16-row tiles without any block-wide
// synchronisation. Called by warp 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void enter_run(const DevPlan& P, const ReqParams& A, const Tile& run,
                                          ColCtx* ctx, uint8_t* s_tables, float* s_bnd,
                                          bool keep_bnd, uint64_t* mbar_tbl, uint64_t* mbar_bnd) {
  const int lane = threadIdx.x & 31;
  const DevSlab& slab = run.slab;
  const bool stage = run.stage != 0 && slab.stage_bytes > 0;
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
  if (!keep_bnd && slab.bnd_bytes > 0) {
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

// ---------------------------------------------------------------------------
// FAST tiles. Everything a warp touches between its global loads (the raw values, prefetched one
// tile ahead) and its global stores lives in shared memory and is addressed with 32-bit shared
// addresses.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t lds_u64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
#define CTX_OFF(field) ((uint32_t)offsetof(ColCtx, field))

// Bounds check and the ids_out parity hook of a FAST tile, off the fast path.
__device__ __noinline__ int32_t fast_note_id(const ReqRegs Q, uint32_t ctx, int64_t i, int32_t id) {
  if (Q.check) id = checked_id(id, (int32_t)lds_u32(ctx + CTX_OFF(table_rows)), Q.status);
  if (Q.ids_out) Q.ids_out[Q.ids_base[lds_u32(ctx + CTX_OFF(col))] + i] = id;
  return id;
}

// The raw values of a tile, requested one tile ahead: asynchronous 4-byte copies (LDGSTS)
// straight into the warp's value buffer in shared memory, slot (column, row) - no registers
// are held while the previous tile is worked on, and nothing waits before phase 1 does. Every
// lane copies (and later reads) the values of its own (column groups, row).
__device__ __forceinline__ void fast_prefetch(uint32_t s_ctx, const DevSlab& slab, int64_t b0,
                                              int rows, uint32_t s_val, int lane) {
  const int nc = slab.num_cols;
  const int TR = slab.tile_rows;
  const int row = lane & (TR - 1), cg = lane / TR, ncg = 32 / TR;
  if (row < rows) {
    for (int cbase = cg * 8; cbase < nc; cbase += ncg * 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float* p = reinterpret_cast<const float*>(
            lds_u64(s_ctx + (uint32_t)(cbase + k) * (uint32_t)sizeof(ColCtx) + CTX_OFF(values)));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_val + 4u * (uint32_t)((cbase + k) * TR + row)),
                     "l"(p + b0 + row)
                     : "memory");
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// phase 1: lane = (column group of 8, row): 8 upper_bound searches in lock step through the
// boundary set's grid; the id slot of (column, row) receives the byte offset of the table row
// inside the staged table region (STAGED) or the raw row id. The values come from the warp's
// value buffer (fast_prefetch). Lane group cg takes the groups cg, cg + ncg, ...
template <bool STAGED>
__device__ __forceinline__ void fast_phase1(const ReqRegs& Q, uint32_t s_ctx, const DevSlab& slab,
                                            int64_t b0, int rows, uint32_t s_bnd, uint32_t s_ids,
                                            uint32_t s_val, int lane) {
  const int nc = slab.num_cols;
  const int TR = slab.tile_rows;
  const int row = lane & (TR - 1), cg = lane / TR, ncg = 32 / TR;
  const bool valid = row < rows;
  const int64_t i = b0 + row;
  const bool hook = Q.check | (Q.ids_out != nullptr) | (Q.shard_world > 1);
  for (int cbase = cg * 8; cbase < nc; cbase += ncg * 8) {
    const uint32_t ctx = s_ctx + (uint32_t)cbase * (uint32_t)sizeof(ColCtx);
    float v[8];
    int32_t lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = valid ? lds_f32(s_val + 4u * (uint32_t)((cbase + k) * TR + row)) : 0.f;
    if ((slab.uni8_mask >> (cbase >> 3)) & 1u) {
      bucket_smem8_uniform(s_bnd + 4u * lds_u32(ctx + CTX_OFF(bnd_off)),
                           (int32_t)lds_u32(ctx + CTX_OFF(num_boundaries)), v, lo);
    } else {
      uint32_t bp[8];
      int32_t nb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t c = ctx + k * (uint32_t)sizeof(ColCtx);
        bp[k] = s_bnd + 4u * lds_u32(c + CTX_OFF(bnd_off));
        nb[k] = (int32_t)lds_u32(c + CTX_OFF(num_boundaries));
      }
      bucket_smem8(bp, nb, v, lo);
    }
    if (valid) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t c = ctx + k * (uint32_t)sizeof(ColCtx);
        int32_t id = lo[k];
        if (hook) id = fast_note_id(Q, c, i, id);
        if (STAGED) id = (int32_t)lds_u32(c + CTX_OFF(stage_off)) + id * (int32_t)lds_u32(c + CTX_OFF(row_bytes));
        sts_i32(s_ids + 4u * (uint32_t)((cbase + k) * kIdPitch + row), id);
      }
    }
  }
}

// phase 2: lane = chunk of the slab's span; 32 / W rows per pass; U independent rows in
// flight; 128-bit loads from the staged tables (or ld.global.nc), 128-bit streaming stores
// straight into the concat rows.
template <int VEC, bool STAGED, class QT>
__device__ __forceinline__ void fast_phase2_impl(const QT& Q, uint32_t s_ctx, int W,
                                                 const DevChunk& ch, int64_t b0, int rows,
                                                 uint32_t s_tables, uint32_t s_ids, int lane) {
#ifdef RECOM_TEST_NODP2
  return;
#endif
  const int rpp = 32 / W;  // rows handled per pass
  const int sub = lane / W;
  if (sub >= rpp) return;
  const uint32_t ids = s_ids + 4u * (uint32_t)(ch.col_local * kIdPitch);
  RowCursor<QT::kSharded> outp(Q, b0, sub, rpp, ch.out_byte);
  constexpr int U = 8;
  const uint32_t ctx = s_ctx + (uint32_t)ch.col_local * (uint32_t)sizeof(ColCtx);
  const uint32_t tbl_s = s_tables + (uint32_t)ch.table_byte;
  const uint8_t* tbl_g = nullptr;
  int32_t row_bytes = 0;
  if (!STAGED) {
    tbl_g = reinterpret_cast<const uint8_t*>(lds_u64(ctx + CTX_OFF(table))) + ch.table_byte;
    row_bytes = (int32_t)lds_u32(ctx + CTX_OFF(row_bytes));
  }
  int r = sub;
  for (; r + (U - 1) * rpp < rows; r += U * rpp) {  // full passes: no predicates
    Chunk<VEC> v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t o = lds_u32(ids + 4u * (uint32_t)(r + k * rpp));
      if (STAGED) v[k] = lds_chunk<VEC>(tbl_s + o);
      else v[k].load(tbl_g + (int64_t)o * row_bytes);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) v[k].store(outp.next());
  }
  for (; r < rows; r += rpp) {  // ragged tail of the run's last tile
    const uint32_t o = lds_u32(ids + 4u * (uint32_t)r);
    Chunk<VEC> v;
    if (STAGED) v = lds_chunk<VEC>(tbl_s + o);
    else v.load(tbl_g + (int64_t)o * row_bytes);
    v.store(outp.next());
  }
}

// The 16-byte variants are the hot ones and inlined; element-wise slabs (unaligned slices) are
// rare and kept out of line to keep the hot code small.
template <int VEC, bool STAGED, class QT>
__device__ __noinline__ void fast_phase2_cold(const QT Q, uint32_t s_ctx, int W, const DevChunk ch,
                                              int64_t b0, int rows, uint32_t s_tables,
                                              uint32_t s_ids, int lane) {
  fast_phase2_impl<VEC, STAGED>(Q, s_ctx, W, ch, b0, rows, s_tables, s_ids, lane);
}
template <int VEC, bool STAGED, class QT>
__device__ __forceinline__ void fast_phase2(const QT& Q, uint32_t s_ctx, int W, const DevChunk& ch,
                                            int64_t b0, int rows, uint32_t s_tables, uint32_t s_ids,
                                            int lane) {
  if (VEC == 16) fast_phase2_impl<VEC, STAGED>(Q, s_ctx, W, ch, b0, rows, s_tables, s_ids, lane);
  else fast_phase2_cold<VEC, STAGED>(Q, s_ctx, W, ch, b0, rows, s_tables, s_ids, lane);
}

// Optional per-warp timing, compiled in only with -DRECOM_PROFILE (make prof): every warp
// records [start, end] in globaltimer ns plus the cycles it spent per phase:
// 0 claims / records / entry, 1 FAST phase 1, 2 FAST phase 2, 3 column tile: transform,
// 4 column tile: gathers, 5 waits for the FAST staging.
#ifdef RECOM_PROFILE
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct ProfRef {
  uint32_t* acc;
  uint32_t* last;
};
#define PROF_PARAM , ProfRef pr
#define PROF_ARG , pr
#define PROF_T(i)                                     \
  {                                                   \
    const uint32_t now_ = (uint32_t)clock64();        \
    pr.acc[i] += now_ - *pr.last;                     \
    *pr.last = now_;                                  \
  }
#else
#define PROF_PARAM
#define PROF_ARG
#define PROF_T(i)
#endif
// -DRECOM_PROFILE_COL (with -DRECOM_PROFILE): level timers of the column tiles instead - 0 claim +
// record, 1 slot offsets / column record, 2 row_splits, 3 transform, 4 gathers; a lap waits for
// the value `dep` (what the level's loads produce).
#if defined(RECOM_PROFILE) && defined(RECOM_PROFILE_COL)
#define PROF_C(i, dep)                                                                   \
  {                                                                                      \
    uint32_t now_;                                                                       \
    asm volatile("mov.u32 %0, %%clock;" : "=r"(now_) : "r"((uint32_t)(dep)) : "memory"); \
    pr.acc[i] += now_ - *pr.last;                                                        \
    *pr.last = now_;                                                                     \
  }
#undef PROF_T
#define PROF_T(i)
#else
#define PROF_C(i, dep)
#endif

// ---------------------------------------------------------------------------
// column tiles (the queue): one warp, rows [begin, begin + count) of one id stream.
//   transform  lane = consecutive ids -> row ids in the warp's scratch (shared memory)
//   gather     lane = (lane group, 16-byte chunk of the column's row): 128-bit row gathers,
//              fp32 accumulation for multi-hot rows, TF-CPU scaling, 128-bit stores into the
//              column's slice of the concat row; once per column that reads the stream.
// Warp scratch: [row_splits of the tile (kColSplitBytes) | row ids].
// ---------------------------------------------------------------------------
constexpr int kColSplitBytes = 160;  // up to 33 row_splits of a multi-hot tile, 16-byte padded

__device__ __forceinline__ int32_t lds_i32(uint32_t addr) {
  int32_t v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Bounds check and the ids_out parity hook, off the fast path.
__device__ __noinline__ int32_t note_id(const ReqRegs Q, int32_t col, int32_t table_rows, int64_t i,
                                        int32_t id) {
  if (id == kDroppedId) return id;  // dropped by the stream's filter
  if (Q.check) id = checked_id(id, table_rows, Q.status);
  if (Q.ids_out && col >= 0) Q.ids_out[Q.ids_base[col] + i] = id;
  if (Q.shard_world > 1)  // -1: another rank's row (the caller skips the gather)
    id = id % Q.shard_world == Q.shard_rank ? id / Q.shard_world : -1;
  return id;
}

// A column reading the tile's stream, resolved against the request.
struct UserRegs {
  const uint8_t* table;
  int32_t row_bytes, table_rows, out_byte, combiner, col, vec_bytes, hbm;
};
__device__ __forceinline__ UserRegs load_user(const DevPlan& P, const ReqParams& A, int32_t index) {
  const int4* q = reinterpret_cast<const int4*>(P.users + index);
  const int4 a = __ldg(q), b = __ldg(q + 1);
  UserRegs u;
  u.table = (const uint8_t*)A.tables[a.x];
  u.row_bytes = a.y;
  u.table_rows = a.z;
  u.out_byte = a.w;
  u.combiner = b.x;
  u.col = b.y;
  u.vec_bytes = b.z;
  u.hbm = b.w;
  return u;
}

__device__ __forceinline__ UserRegs first_user(const ReqParams& A, const ColItem& it) {
  UserRegs u;
  u.table = (const uint8_t*)A.tables[it.u0.table];
  u.row_bytes = it.u0.row_bytes;
  u.table_rows = it.u0.table_rows;
  u.out_byte = it.u0.out_byte;
  u.combiner = it.u0.combiner;
  u.col = it.u0.col;
  u.vec_bytes = it.u0.vec_bytes;
  u.hbm = it.u0.hbm;
  return u;
}

// The table rows of ids s_ids[0, n) -> L2 (tables far larger than L2: HBM latency and the page
// walks behind it are paid once per tile, all rows in flight together, instead of once per
// pass of the gather).
__device__ __forceinline__ void prefetch_rows(const UserRegs& u, uint32_t s_ids, int n, int lane) {
  for (int j = lane; j < n; j += 32) {
    const int32_t id = lds_i32(s_ids + 4u * (uint32_t)j);
    if (id < 0) continue;  // dropped by the stream's filter
    const uint8_t* row = u.table + (int64_t)id * u.row_bytes;
    for (int b = 0; b < u.row_bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
  }
}

// Where the rows a lane of a column tile stores go. Sharded output: a tile of up to 128 rows
// may span several owners, so the owner is looked up per row.
template <bool SH>
struct ColCursor;
template <>
struct ColCursor<false> : RowCursor<false> {
  __device__ __forceinline__ ColCursor(const ReqRegs& Q, int64_t b0, int r0, int step_rows,
                                       int32_t out_byte)
      : RowCursor<false>(Q, b0, r0, step_rows, out_byte) {}
};
template <>
struct ColCursor<true> {
  const ReqParams* A;
  int64_t orb;
  int32_t row, step_rows, byte;
  __device__ __forceinline__ ColCursor(const ReqRegsT<true>& Q, int64_t b0, int r0, int step_rows_,
                                       int32_t out_byte) {
    A = Q.A;
    orb = Q.out_row_bytes;
    row = (int32_t)b0 + r0;
    step_rows = step_rows_;
    byte = A->col_off_bytes + out_byte;
  }
  __device__ __forceinline__ uint8_t* next() {
    const int32_t rpr = (int32_t)A->rows_per_rank;
    const int32_t g = min(row / rpr, A->world - 1);
    uint8_t* q = A->peer_out[g] + (int64_t)(row - g * rpr) * orb + byte;
    row += step_rows;
    return q;
  }
};

// One id per row (DENSE lookups): rows [b0, b0 + n), ids in s_ids[0, n). lane = (lane group,
// chunk); U rows in flight per lane.
template <int VEC, class QT>
__device__ __forceinline__ void gather_rows(const QT& Q, const UserRegs& u, int64_t b0, int n,
                                            uint32_t s_ids, int lane) {
  const int C = u.row_bytes / VEC;
  const bool hook = Q.check | (Q.ids_out != nullptr) | (Q.shard_world > 1);
  constexpr int U = 8;
  for (int c0 = 0; c0 < C; c0 += 32) {  // rows wider than 32 chunks: 32 chunks at a time
    const int W = min(32, C - c0);
    const int rpp = 32 / W, sub = lane / W;
    if (sub >= rpp) continue;
    const int chb = (c0 + lane - sub * W) * VEC;
    const uint8_t* tbl = u.table + chb;
    const int32_t hook_col = chb == 0 ? u.col : -1;
    ColCursor<QT::kSharded> outp(Q, b0, sub, rpp, u.out_byte + chb);
    for (int r = sub; r < n; r += U * rpp) {
      int32_t id[U];
#pragma unroll
      for (int k = 0; k < U; ++k) id[k] = r + k * rpp < n ? lds_i32(s_ids + 4u * (uint32_t)(r + k * rpp)) : 0;
      if (hook) {
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (r + k * rpp < n) id[k] = note_id(Q, hook_col, u.table_rows, b0 + r + k * rpp, id[k]);
      }
      Chunk<VEC> v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (r + k * rpp < n && id[k] >= 0) v[k].load(tbl + (int64_t)id[k] * u.row_bytes);
        else v[k].zero();  // (row-sharded tables: another rank's row)
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        uint8_t* o = outp.next();
        if (r + k * rpp < n) v[k].store(o);
      }
    }
  }
}
template <int VEC, class QT>
__device__ __noinline__ void gather_rows_cold(const QT Q, const UserRegs u, int64_t b0, int n,
                                              uint32_t s_ids, int lane) {
  gather_rows<VEC>(Q, u, b0, n, s_ids, lane);
}

// End of a row: combiner scaling (TF-CPU: segment_reduction_ops_impl.h:619-629,765-770 - n < 10
// multiplies by 1/m, n >= 10 divides by m; m = n or sqrt(n)) and the row's one store.
template <int VEC, bool BF16>
__device__ __forceinline__ void flush_row(Acc<VEC, BF16>& acc, int32_t n, int combiner, uint8_t* outp) {
  if (n > 1 && combiner != RECOM_COMBINE_SUM) {
    const float m = combiner == RECOM_COMBINE_MEAN ? (float)n : (float)__dsqrt_rn((double)n);
    if (n < 10) acc.mul(__fdiv_rn(1.0f, m));
    else acc.div(m);
  }
  acc.pack().store(outp);
}

// Rows with more ids than this are summed by all lane groups together (coop_row).
constexpr int kCoopRowIds = 64;

// Multi-hot rows [b0, b0 + rows): row r owns ids [sp[r], sp[r + 1]) of the stream, held in
// s_ids[e - e_base]. lane = (lane group, chunk); a lane group takes the rows sub, sub + rpp, ...
// one at a time: up to U independent row gathers in flight, fp32 accumulation in id order (the
// order of the TF-CPU kernel: bit-identical sums), one store per row; empty rows store zeros.
// Rows longer than kCoopRowIds are left to coop_row.
template <int VEC, bool BF16, class QT>
__device__ __forceinline__ void sparse_rows(const QT& Q, const UserRegs& u, int64_t b0, int rows,
                                            uint32_t s_sp, uint32_t s_ids, int32_t e_base, int lane) {
  const int C = u.row_bytes / VEC;
  const bool hook = Q.check | (Q.ids_out != nullptr) | (Q.shard_world > 1);
#ifndef RECOM_SPARSE_U
#define RECOM_SPARSE_U 10  // rows of 9-10 ids (dlrm.py:52: 1..10 per row) in one pass: mixed1000 123.2 -> 120.5, model F 49.9 -> 48.0, E multi-hot 16.2 -> 14.5 us
#endif
  constexpr int U = BF16 ? 4 : RECOM_SPARSE_U;
  const uint32_t ids0 = s_ids - 4u * (uint32_t)e_base;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int W = min(32, C - c0);
    const int rpp = 32 / W, sub = lane / W;
    if (sub >= rpp) continue;
    const int chb = (c0 + lane - sub * W) * VEC;
    const uint8_t* tbl = u.table + chb;
    const int64_t row_bytes = u.row_bytes;
    const int32_t hook_col = chb == 0 ? u.col : -1;
    ColCursor<QT::kSharded> outp(Q, b0, sub, rpp, u.out_byte + chb);
    for (int r = sub; r < rows; r += rpp) {
      const int32_t e0 = lds_i32(s_sp + 4u * (uint32_t)r), e1 = lds_i32(s_sp + 4u * (uint32_t)(r + 1));
      uint8_t* o = outp.next();
      if (e1 - e0 > kCoopRowIds) continue;
      Acc<VEC, BF16> acc;
      acc.clear();
      int32_t kept = 0;  // ids the stream's filter did not drop
      for (int32_t e = e0; e < e1; e += U) {
        int32_t id[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          id[k] = e + k < e1 ? lds_i32(ids0 + 4u * (uint32_t)(e + k)) : kDroppedId;
          kept += id[k] != kDroppedId;
        }
        if (hook) {
#pragma unroll
          for (int k = 0; k < U; ++k)
            if (e + k < e1) id[k] = note_id(Q, hook_col, u.table_rows, e + k, id[k]);
        }
        Chunk<VEC> v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          if (e + k < e1 && id[k] >= 0) v[k].load(tbl + (int64_t)id[k] * row_bytes);
          else v[k].zero();
        }
#pragma unroll
        for (int k = 0; k < U; ++k) acc.add(v[k]);  // + 0.0f for the slots past the row's end
      }
      flush_row<VEC, BF16>(acc, kept, u.combiner, o);
    }
  }
}
template <int VEC, bool BF16, class QT>
__device__ __noinline__ void sparse_rows_cold(const QT Q, const UserRegs u, int64_t b0, int rows,
                                              uint32_t s_sp, uint32_t s_ids, int32_t e_base, int lane) {
  sparse_rows<VEC, BF16>(Q, u, b0, rows, s_sp, s_ids, e_base, lane);
}

// Ids of s_ids[0, n) the stream's filter kept (warp uniform).
__device__ __forceinline__ int32_t count_kept(uint32_t s_ids, int32_t n, int lane) {
  int32_t c = 0;
  for (int32_t j = lane; j < n; j += 32) c += lds_i32(s_ids + 4u * (uint32_t)j) != kDroppedId;
  return __reduce_add_sync(0xffffffffu, c);
}

// The partial sums of a long row, accumulated over the ids s_ids[0, cn) (stream positions
// e_first + j): the lane groups take alternate ids. (Another summation order than the stream's:
// within the fp32 tolerance, not bit-identical to a sequential sum.)
template <int VEC, bool BF16, class QT>
__device__ __forceinline__ void coop_accumulate(const QT& Q, const UserRegs& u, Acc<VEC, BF16>& acc,
                                                const uint8_t* tbl, int32_t hook_col, int rpp, int sub,
                                                uint32_t s_ids, int32_t cn, int32_t e_first) {
  const bool hook = Q.check | (Q.ids_out != nullptr) | (Q.shard_world > 1);
  for (int32_t j = sub; j < cn; j += 4 * rpp) {
    int32_t id[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) id[k] = j + k * rpp < cn ? lds_i32(s_ids + 4u * (uint32_t)(j + k * rpp)) : 0;
    if (hook) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (j + k * rpp < cn) id[k] = note_id(Q, hook_col, u.table_rows, e_first + j + k * rpp, id[k]);
    }
    Chunk<VEC> v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (j + k * rpp < cn && id[k] >= 0) v[k].load(tbl + (int64_t)id[k] * u.row_bytes);
      else v[k].zero();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc.add(v[k]);
  }
}
// ... and their meeting in shared memory (s_tmp: 32 * kFloats floats), scaling and store by
// lane group 0.
template <int VEC, bool BF16, class QT>
__device__ __forceinline__ void coop_finish(const QT& Q, const UserRegs& u, Acc<VEC, BF16>& acc,
                                            int W, int rpp, int sub, int chb, int64_t row, int32_t n,
                                            uint32_t s_tmp, int lane) {
  constexpr int NE = Acc<VEC, BF16>::kFloats;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NE; ++i)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_tmp + 4u * (uint32_t)(lane * NE + i)), "f"(acc.at(i))
                 : "memory");
  __syncwarp();
  if (sub == 0) {
    for (int s = 1; s < rpp; ++s) {
#pragma unroll
      for (int i = 0; i < NE; ++i) acc.at(i) += lds_f32(s_tmp + 4u * (uint32_t)((s * W + lane) * NE + i));
    }
    ColCursor<QT::kSharded> outp(Q, row, 0, 1, u.out_byte + chb);
    flush_row<VEC, BF16>(acc, n, u.combiner, outp.next());
  }
  __syncwarp();
}

// The rows of a tile with more than kCoopRowIds ids (ids in the scratch): all lane groups sum
// one row together. s_tmp: scratch behind the tile's ids.
template <int VEC, bool BF16, class QT>
__device__ __noinline__ void coop_rows(const QT Q, const UserRegs u, int64_t b0, int rows,
                                       uint32_t s_sp, uint32_t s_ids, int32_t e_base, uint32_t s_tmp,
                                       int lane) {
  const int C = u.row_bytes / VEC;
  for (int r = 0; r < rows; ++r) {
    const int32_t e0 = lds_i32(s_sp + 4u * (uint32_t)r), e1 = lds_i32(s_sp + 4u * (uint32_t)(r + 1));
    if (e1 - e0 <= kCoopRowIds) continue;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int W = min(32, C - c0);
      const int rpp = 32 / W, sub = lane / W;
      const int chb = (c0 + lane - sub * W) * VEC;
      Acc<VEC, BF16> acc;
      acc.clear();
      const int32_t kept = count_kept(s_ids + 4u * (uint32_t)(e0 - e_base), e1 - e0, lane);
      if (sub < rpp)
        coop_accumulate<VEC, BF16>(Q, u, acc, u.table + chb, chb == 0 ? u.col : -1, rpp, sub,
                                   s_ids + 4u * (uint32_t)(e0 - e_base), e1 - e0, e0);
      coop_finish<VEC, BF16>(Q, u, acc, W, rpp, sub, chb, b0 + r, kept, s_tmp, lane);
    }
  }
}

// A row with more ids than the scratch holds (long tails of variable-length features): the
// ids are transformed `cap` at a time.
template <int VEC, bool BF16, class QT>
__device__ __noinline__ void long_row(const DevPlan& P, const QT Q, const ColItem it,
                                      const UserRegs u, const uint8_t* values, const uint8_t* bytes,
                                      int64_t row, int32_t e0, int32_t e1, uint32_t s_ids, int capc,
                                      uint32_t s_tmp, int lane) {
  const int C = u.row_bytes / VEC;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int W = min(32, C - c0);
    const int rpp = 32 / W, sub = lane / W;
    const int chb = (c0 + lane - sub * W) * VEC;
    Acc<VEC, BF16> acc;
    acc.clear();
    int32_t kept = 0;
    for (int32_t cb = e0; cb < e1; cb += capc) {
      const int32_t cn = min(e1 - cb, capc);
      transform_ids(P, it, values, bytes, cb, cb + cn, s_ids, lane);
      __syncwarp();
      kept += count_kept(s_ids, cn, lane);
      if (sub < rpp)
        coop_accumulate<VEC, BF16>(Q, u, acc, u.table + chb, chb == 0 ? u.col : -1, rpp, sub, s_ids, cn, cb);
      __syncwarp();
    }
    coop_finish<VEC, BF16>(Q, u, acc, W, rpp, sub, chb, row, kept, s_tmp, lane);
  }
}

// Rows [b0, b0 + rows) of a multi-hot tile whose ids are in the scratch: every column of the
// stream. `any_long`: some row has more than kCoopRowIds ids.
template <bool BF16, class QT>
__device__ __forceinline__ void sparse_users(const DevPlan& P, const ReqParams& A, const QT& Q,
                                             const ColItem& it, UserRegs u, int64_t b0, int rows,
                                             uint32_t s_sp, uint32_t s_ids, int32_t e_base,
                                             bool any_long, uint32_t s_tmp, int lane) {
  constexpr int ELEM = BF16 ? 2 : 4;
  for (int ui = 0;;) {
    if (u.vec_bytes == 16) sparse_rows<16, BF16>(Q, u, b0, rows, s_sp, s_ids, e_base, lane);
    else sparse_rows_cold<ELEM, BF16>(Q, u, b0, rows, s_sp, s_ids, e_base, lane);
    if (any_long) {
      if (u.vec_bytes == 16) coop_rows<16, BF16>(Q, u, b0, rows, s_sp, s_ids, e_base, s_tmp, lane);
      else coop_rows<ELEM, BF16>(Q, u, b0, rows, s_sp, s_ids, e_base, s_tmp, lane);
    }
    if (++ui >= it.num_users) break;
    u = load_user(P, A, it.first_user + ui);
  }
}

// A multi-hot tile with more ids than the scratch holds: row ranges that fit, one after the
// other; single rows that do not fit go through long_row.
template <bool BF16, class QT>
__device__ __noinline__ void sparse_tile_overflow(const DevPlan& P, const ReqParams& A, const QT Q,
                                                  const ColItem it, const uint8_t* values,
                                                  const uint8_t* bytes, uint32_t s_sp,
                                                  uint32_t s_ids, int cap, uint32_t s_tmp, int lane,
                                                  int item_k) {
  constexpr int ELEM = BF16 ? 2 : 4;
  const int R = it.count;
  int r_lo = 0;
  REQUEST_PDL_WAIT(A);
  // rows longer than kDeferIds are left to long_rows_kernel when the request has a list for them
  auto deferred = [&](int r) {
    return A.long_list != nullptr && item_k >= 0 &&
           lds_i32(s_sp + 4u * (uint32_t)(r + 1)) - lds_i32(s_sp + 4u * (uint32_t)r) > A.long_ids;
  };
  while (r_lo < R) {
    const int32_t e_lo = lds_i32(s_sp + 4u * (uint32_t)r_lo);
    if (deferred(r_lo)) {
      int slot = A.long_cap;
      if (lane == 0) slot = atomicAdd(A.long_list, 1);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (slot < A.long_cap) {  // (a full list: the row is summed here after all)
        if (lane == 0) {
          A.long_list[2 + 2 * slot] = item_k;
          A.long_list[3 + 2 * slot] = it.begin + r_lo;
        }
        ++r_lo;
        continue;
      }
    }
    int r_hi = r_lo;
    while (r_hi < R && lds_i32(s_sp + 4u * (uint32_t)(r_hi + 1)) - e_lo <= cap && !(r_hi > r_lo && deferred(r_hi))) ++r_hi;
    if (r_hi == r_lo) {
      const int32_t e_hi = lds_i32(s_sp + 4u * (uint32_t)(r_lo + 1));
      for (int ui = 0; ui < it.num_users; ++ui) {
        const UserRegs u = load_user(P, A, it.first_user + ui);
        if (u.vec_bytes == 16)
          long_row<16, BF16>(P, Q, it, u, values, bytes, (int64_t)it.begin + r_lo, e_lo, e_hi, s_ids, cap, s_tmp, lane);
        else
          long_row<ELEM, BF16>(P, Q, it, u, values, bytes, (int64_t)it.begin + r_lo, e_lo, e_hi, s_ids, cap, s_tmp, lane);
      }
      ++r_lo;
      continue;
    }
    const int32_t e_hi = lds_i32(s_sp + 4u * (uint32_t)r_hi);
    transform_ids(P, it, values, bytes, e_lo, e_hi, s_ids, lane);
    __syncwarp();
    sparse_users<BF16>(P, A, Q, it, first_user(A, it), (int64_t)it.begin + r_lo,
                       r_hi - r_lo, s_sp + 4u * (uint32_t)r_lo, s_ids, e_lo, true, s_tmp, lane);
    __syncwarp();
    r_lo = r_hi;
  }
}

// Where the hashed tokens of split input k start in ws_split_ids (in ids): input k' < k takes
// payload bytes + batch + 1 slots, an upper bound of its tokens whatever the delimiter. Warp
// collective (lanes stride over k').
__device__ __forceinline__ int64_t split_ids_offset(const DevPlan& P, const ReqParams& A, int k, int lane) {
  int64_t s = 0;
  for (int j = lane; j < k; j += 32) s += A.slot_elems[P.splits[j].bytes_slot] + A.batch + 1;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  return s;
}

// The record of a column tile: 96 bytes, six 128-bit loads (lane-uniform address).
__device__ __forceinline__ void load_item(const ColItem* p, ColItem& it) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&it);
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = __ldg(q + k);
}


template <bool BF16, class QT>
__device__ __forceinline__ void column_tile(const DevPlan& P, const ReqParams& A, const QT& Q,
                                            const ColItem& it, uint32_t s_scratch, int lane, int item_k
                                            PROF_PARAM) {
  constexpr int ELEM = BF16 ? 2 : 4;
  // resolve the stream and its first column against the request: one read of the slot offsets
  // (value_slot < 0: the ids are the hashed tokens of split input -1 - value_slot, written into the
  // workspace by the split kernel in front of this launch)
  // request_pdl 2: the column tiles wait before they read anything (plans with more column tiles than
  // resident warps: their transforms, run ahead of the wait, compete with the previous request's tail -
  // model F 49.2 -> 43.9 us/step); 1: they wait behind their transform (model E 40.2 -> 37.9)
  if (A.request_pdl == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint8_t* values =
      it.value_slot >= 0 ? A.blob + __ldg(A.slot_offsets + it.value_slot)
                         : (const uint8_t*)(A.ws_split_ids + split_ids_offset(P, A, -1 - it.value_slot, lane));
  const uint8_t* bytes = it.bytes_slot >= 0 ? A.blob + __ldg(A.slot_offsets + it.bytes_slot) : nullptr;
  const int32_t* rs = nullptr;
  if (it.seg_slot >= 0) {
    rs = it.coo_index >= 0 ? A.ws_row_splits + (int64_t)it.coo_index * (A.batch + 1)
                           : (const int32_t*)(A.blob + __ldg(A.slot_offsets + it.seg_slot));
    // row_splits built by the helper kernel in front of this launch: when this grid was launched
    // with programmatic serialization it started while the helper ran (no-op otherwise)
    if (it.coo_index >= 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  UserRegs u = first_user(A, it);
  PROF_C(1, (uintptr_t)values + (uintptr_t)bytes + (uintptr_t)rs + (uintptr_t)u.table + u.row_bytes)
  const uint32_t s_sp = s_scratch, s_ids = s_sp + kColSplitBytes;
  // scratch: [row_splits | ids (cap) | 1 KB for the partial sums of long rows]
  const int cap = (P.scratch_bytes - kColSplitBytes - 1024) >> 2;
  const uint32_t s_tmp = s_sp + (uint32_t)P.scratch_bytes - 1024u;
  if (it.seg_slot < 0) {
    // one id per row; the schedule keeps count <= cap
    const int n = it.count;
    transform_ids(P, it, values, bytes, it.begin, (int64_t)it.begin + n, s_ids, lane);
    __syncwarp();
    REQUEST_PDL_WAIT(A);
    if (u.hbm && !Q.check && Q.shard_world <= 1) prefetch_rows(u, s_ids, n, lane);
    PROF_T(3)
    PROF_C(3, lds_i32(s_ids))
    for (int ui = 0;;) {
      if (u.vec_bytes == 16) gather_rows<16>(Q, u, it.begin, n, s_ids, lane);
      else gather_rows_cold<ELEM>(Q, u, it.begin, n, s_ids, lane);
      if (++ui >= it.num_users) break;
      u = load_user(P, A, it.first_user + ui);
    }
    PROF_T(4)
    PROF_C(4, 0)
  } else {
    const int R = it.count;  // <= 32
    const int32_t v = lane < R ? __ldg(rs + it.begin + lane) : 0;
    const int32_t e_end = __ldg(rs + it.begin + R);
    sts_i32(s_sp + 4u * (uint32_t)lane, lane < R ? v : e_end);
    if (lane == 0) sts_i32(s_sp + 4u * 32u, e_end);
    const int32_t e0 = __shfl_sync(0xffffffffu, v, 0);
    const int32_t v_next = __shfl_down_sync(0xffffffffu, v, 1);
    const int32_t n_row = lane < R ? (lane == R - 1 ? e_end : v_next) - v : 0;
    const bool any_long = __any_sync(0xffffffffu, n_row > kCoopRowIds);
    const bool any_deferred = A.long_list != nullptr && item_k >= 0 && __any_sync(0xffffffffu, n_row > A.long_ids);
    __syncwarp();
    PROF_C(2, e0 + e_end)
    if (e_end - e0 <= cap && !any_deferred) {
      transform_ids(P, it, values, bytes, e0, e_end, s_ids, lane);
      __syncwarp();
      REQUEST_PDL_WAIT(A);
      if (u.hbm && !Q.check && Q.shard_world <= 1) prefetch_rows(u, s_ids, e_end - e0, lane);
      PROF_T(3)
      PROF_C(3, lds_i32(s_ids))
      sparse_users<BF16>(P, A, Q, it, u, it.begin, R, s_sp, s_ids, e0, any_long, s_tmp, lane);
    } else {
      sparse_tile_overflow<BF16>(P, A, Q, it, values, bytes, s_sp, s_ids, cap, s_tmp, lane, item_k);
    }
    PROF_T(4)
    PROF_C(4, 0)
  }
  __syncwarp();  // everybody is done with the scratch
}


// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
// What warp 0 publishes about the CTA's current FAST run (shared memory, guarded by mbar_ctx).
struct RunShared {
  Tile run;
  DevChunk chunks[32];  // lane -> its chunk of the slab's span
  int32_t nruns;        // FAST runs of this CTA
  int32_t first_run;    // index of the CTA's first run in Sched::runs
  int32_t keep_bnd;     // the run's single boundary set is already resident
  int32_t staged_bnd;   // (warp 0) pool offset of the single boundary set resident in smem, or -1
};

// The CTA's shared state outside the dynamic region.
struct CtaShared {
  uint64_t mbar_tbl;    // tables of the staged slab have landed
  uint64_t mbar_bnd;    // boundaries of the staged slab have landed
  uint64_t mbar_ctx;    // the run's record and column contexts are in shared memory
  int next_fast[2];     // next FAST tile of the run (alternating between runs)
  int next_col;         // next column tile of the CTA
  int pad;
  RunShared rs;
};

// Shared addresses of the dynamic region's parts.
struct SmemMap {
  uint32_t tables, bnd, ctx, scratch;  // scratch: this warp's
};

// Warp 0: record, column contexts and TMA staging of the CTA's run `ri`.
__device__ __noinline__ void setup_run(const DevPlan& P, const Sched& S, const ReqParams& A,
                                       CtaShared* cs, uint8_t* smem, int ri, int lane) {
  const Tile run = S.runs[cs->rs.first_run + ri];
  const int staged_bnd = cs->rs.staged_bnd;
  const bool keep = ri > 0 && run.slab.num_bnds == 1 && run.slab.first_bnd_pool == staged_bnd;
  cs->rs.chunks[lane] = P.chunks[run.slab.first_chunk + lane % run.slab.num_chunks];
  __syncwarp();
  if (lane == 0) {
    cs->rs.run = run;
    cs->rs.keep_bnd = keep ? 1 : 0;
    if (run.slab.num_bnds == 1) cs->rs.staged_bnd = run.slab.first_bnd_pool;
    else if (run.slab.num_bnds > 1) cs->rs.staged_bnd = -1;
    cs->next_fast[ri & 1] = 0;
  }
  enter_run(P, A, run, reinterpret_cast<ColCtx*>(smem + P.smem_off_ctx), smem,
            reinterpret_cast<float*>(smem + P.smem_off_bnd), keep, &cs->mbar_tbl, &cs->mbar_bnd);
  __syncwarp();
  mbar_arrive(&cs->mbar_ctx);
}

// The CTA's column tiles, taken from the counter in shared memory as the warps become free
// (k >= 0: the warp's static first tile, its record `it` requested at kernel entry). Measured:
// a static deal of the tiles to the warps costs more in balance (mixed1000 156 -> 227 us) than
// the exposed load of a claimed tile's record.
template <bool BF16, class QT>
__device__ __noinline__ void column_section(const DevPlan& P, const ReqParams& A, const QT Q,
                                            const ColItem* q_items, int q_count, int k, ColItem it,
                                            int* next_col, uint32_t s_scratch, int lane PROF_PARAM) {
  auto claim = [&]() {
    int t = 0;
    if (lane == 0) t = atomicAdd(next_col, 1);
    return __shfl_sync(0xffffffffu, t, 0);
  };
  if (k < 0) {
    k = claim();
    if (k < q_count) load_item(q_items + k, it);
  }
  while (k < q_count) {
    PROF_T(0)
    PROF_C(0, it.begin + it.count)
    column_tile<BF16>(P, A, Q, it, s_scratch, lane, -1 PROF_ARG);
    k = claim();
    if (k < q_count) load_item(q_items + k, it);
  }
}

// The CTA's FAST runs: every warp takes tiles from the run's counter.
template <bool BF16, class QT>
__device__ __noinline__ void fast_section(const DevPlan& P, const Sched& S, const ReqParams& A,
                                          const QT Q, CtaShared* cs, uint8_t* smem, SmemMap sm,
                                          int warp, int lane PROF_PARAM) {
  constexpr int ELEM = BF16 ? 2 : 4;
  uint32_t tbl_par = 0, bnd_par = 0, ctx_par = 0;
  const uint32_t s_ids = sm.scratch;
  mbar_wait(&cs->mbar_ctx, ctx_par);
  ctx_par ^= 1;
  const int nruns = cs->rs.nruns;
  for (int ri = 0; ri < nruns; ++ri) {
    if (ri > 0) {
      __syncthreads();  // every warp is done with the previous run's tables and contexts
      if (warp == 0) setup_run(P, S, A, cs, smem, ri, lane);
      mbar_wait(&cs->mbar_ctx, ctx_par);
      ctx_par ^= 1;
    }
    const DevSlab slab = cs->rs.run.slab;
    const int run_rows = cs->rs.run.rows;
    const int64_t run_begin = cs->rs.run.row_begin;
    const bool staged = cs->rs.run.stage != 0 && slab.stage_bytes > 0;
    const bool wait_bnd = !cs->rs.keep_bnd && slab.bnd_bytes > 0;
    const DevChunk ch = cs->rs.chunks[lane];
    PROF_T(5)
    bool bnd_ready = !wait_bnd;
    bool tbl_ready = !staged;
    const int TR = slab.tile_rows;  // 16 or 32 rows
    const int ntiles = (run_rows + TR - 1) / TR;
    int* ctr = &cs->next_fast[ri & 1];
    auto claim = [&]() {
      int t = 0;
      if (lane == 0) t = atomicAdd(ctr, 1);
      return __shfl_sync(0xffffffffu, t, 0);
    };
    // two value buffers behind the tile's row ids: the raw values of the warp's next tile are
    // requested (asynchronous copies) before this tile is worked on
    const uint32_t val_bytes = 4u * (uint32_t)(slab.num_cols * TR);
    uint32_t s_val = s_ids + 4u * (uint32_t)(slab.num_cols * kIdPitch), s_val_next = s_val + val_bytes;
    int t = claim();
    if (t < ntiles) fast_prefetch(sm.ctx, slab, run_begin + (int64_t)t * TR, min(TR, run_rows - t * TR), s_val, lane);
    while (t < ntiles) {
      const int64_t b0 = run_begin + (int64_t)t * TR;
      const int rows = min(TR, run_rows - t * TR);
      const int tn = claim();
      if (tn < ntiles) {
        fast_prefetch(sm.ctx, slab, run_begin + (int64_t)tn * TR, min(TR, run_rows - tn * TR), s_val_next, lane);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      if (!bnd_ready) {
        mbar_wait(&cs->mbar_bnd, bnd_par);
        bnd_ready = true;
      }
      if (staged) fast_phase1<true>(Q, sm.ctx, slab, b0, rows, sm.bnd, s_ids, s_val, lane);
      else fast_phase1<false>(Q, sm.ctx, slab, b0, rows, sm.bnd, s_ids, s_val, lane);
      __syncwarp();
      PROF_T(1)
      if (!tbl_ready) {
        mbar_wait(&cs->mbar_tbl, tbl_par);
        tbl_ready = true;
      }
      REQUEST_PDL_WAIT(A);  // phase 2 stores
      if (staged) {
        if (slab.vec_bytes == 16) fast_phase2<16, true>(Q, sm.ctx, slab.num_chunks, ch, b0, rows, sm.tables, s_ids, lane);
        else fast_phase2<ELEM, true>(Q, sm.ctx, slab.num_chunks, ch, b0, rows, sm.tables, s_ids, lane);
      } else {
        if (slab.vec_bytes == 16) fast_phase2<16, false>(Q, sm.ctx, slab.num_chunks, ch, b0, rows, sm.tables, s_ids, lane);
        else fast_phase2<ELEM, false>(Q, sm.ctx, slab.num_chunks, ch, b0, rows, sm.tables, s_ids, lane);
      }
      __syncwarp();
      PROF_T(2)
      const uint32_t tmp = s_val;
      s_val = s_val_next;
      s_val_next = tmp;
      t = tn;
    }
    // a warp without tiles must still consume the phases it would have waited for
    if (wait_bnd) bnd_par ^= 1;
    if (staged) tbl_par ^= 1;
  }
}

// Sharded output (Mode A, fused exchange): every warp of the request's launches counts itself
// when its rows are stored; the last one tells every peer that this rank's slices of its rows
// are in place (and, with the barrier inside the kernel, waits for theirs).
__device__ __forceinline__ void signal_peers(const ReqParams& A, int lane) {
  if (A.peer_flags[0] == nullptr) return;
  __syncwarp();  // the lanes' row stores are ordered before lane 0's fence
  bool last = false;
  if (lane == 0) {
    __threadfence_system();
    last = atomicAdd(A.counters, 1) == A.total_warps - 1;
    if (last) {
      A.counters[0] = 0;
      __threadfence();
      // the slot may be handed to another request (host side check in recom_plan_run)
      if (A.done_seq) *(volatile uint32_t*)A.done_seq = A.seq;
    }
  }
  if (__shfl_sync(0xffffffffu, (int)last, 0)) {
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

template <bool BF16, bool SH>
#ifndef RECOM_MIN_BLOCKS
#define RECOM_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kThreads, RECOM_MIN_BLOCKS)
fused_columns_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ Sched S,
                     const __grid_constant__ ReqParams A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(16) CtaShared cs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
#ifdef RECOM_PROFILE
  uint32_t prof_acc[6] = {0, 0, 0, 0, 0, 0};
  const uint64_t prof_t0 = gtimer();
  uint32_t prof_last = (uint32_t)clock64();
  ProfRef pr{prof_acc, &prof_last};
#endif

  // Back-to-back requests on one stream (A.request_pdl, launched with programmatic stream
  // serialization): the next request's grid may become resident while this one drains - its
  // prologue (blob prefetch, work records, column contexts, TMA staging of the tables: reads of
  // ITS request and of the plan only) overlaps this grid's tail, and it executes
  // griddepcontrol.wait (below) before its first store.
  if (A.request_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // The whole request blob -> L2 while the grid starts up (one 128-byte line per thread and
  // pass): HBM is idle then, and every later input load of the launch is an L2 hit.
  for (int64_t o = ((int64_t)cta * kThreads + threadIdx.x) * 128; o < A.blob_bytes;
       o += (int64_t)G * kThreads * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.blob + o));

  if (threadIdx.x == 0) {
    mbar_init(&cs.mbar_tbl, 1);
    mbar_init(&cs.mbar_bnd, 1);
    mbar_init(&cs.mbar_ctx, 32);  // every lane of warp 0 arrives: each releases its own writes
    cs.next_fast[0] = 0;
    cs.next_fast[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- the CTA's work: FAST runs and column tiles (contiguous in the queue, most expensive
  // first). Warps [fast_warps, kWarps) start with one column tile each. ----
  const CtaWork cw = G <= kMaxParamCtas ? S.cta_param[cta] : S.cta[cta];
  const int q_count = cw.q_count;
  const ColItem* q_items = S.queue + cw.q_first;
  const bool excl = S.fast_ctas >= 0;
  const int fast_warps = excl ? (cta < S.fast_ctas ? kWarps : 0) : S.fast_warps;
  const bool fast_first = warp < fast_warps;
  // the column-first warps' first tile is static: its record is requested before anything else
  const int k0 = fast_first ? -1 : warp - fast_warps;
  ColItem it;
  if (k0 >= 0 && k0 < q_count) load_item(q_items + k0, it);
  if (threadIdx.x == 0) cs.next_col = kWarps - fast_warps;  // the static first tiles are taken
  __syncthreads();

  // ---- warp 0: the CTA's first FAST run - record, contexts, TMA staging - in flight while
  // the other warps start on their column tiles ----
  if (warp == 0) {
    if (lane == 0) {
      cs.rs.nruns = cw.count;
      cs.rs.first_run = cw.first;
      cs.rs.staged_bnd = -1;
    }
    __syncwarp();
    if (cw.count > 0) setup_run(P, S, A, &cs, smem, 0, lane);
    else mbar_arrive(&cs.mbar_ctx);
  }

  // (A.request_pdl: nothing is stored before the launch in front of this one - the previous request - is
  // complete: every warp executes griddepcontrol.wait between the reads of a tile - values, bucket searches,
  // hashes: into shared memory - and its first store: REQUEST_PDL_WAIT in fast_section / column_tile)

  ReqRegsT<SH> Q;
  Q.out = A.out;
  Q.bind(A);
  Q.out_row_bytes = A.out_row_bytes;
  Q.status = A.status;
  Q.ids_out = A.ids_out;
  Q.ids_base = A.ids_base;
  Q.check = P.check_ids != 0;
  Q.shard_world = P.shard_world;
  Q.shard_rank = P.shard_rank;
  SmemMap sm;
  sm.tables = smem_u32(smem);
  sm.bnd = sm.tables + (uint32_t)P.smem_off_bnd;
  sm.ctx = sm.tables + (uint32_t)P.smem_off_ctx;
  sm.scratch = sm.tables + (uint32_t)P.smem_off_scratch + (uint32_t)(warp * P.scratch_bytes);

  if (!fast_first)
    column_section<BF16>(P, A, Q, q_items, q_count, k0, it, &cs.next_col, sm.scratch, lane PROF_ARG);
  fast_section<BF16>(P, S, A, Q, &cs, smem, sm, warp, lane PROF_ARG);
  if (fast_first)
    column_section<BF16>(P, A, Q, q_items, q_count, k0, it, &cs.next_col, sm.scratch, lane PROF_ARG);

  // ---- sharded output: the last warp of the request (this grid's and the flat kernel's) tells
  // every peer that this rank's slices of its rows are in place ----
  if (SH) signal_peers(A, lane);
  // launched behind the flat kernel of the column tiles with programmatic serialization: this
  // grid is complete only when that one is (no-op otherwise)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef RECOM_PROFILE
  if (lane == 0 && A.prof) {
    uint64_t* o_ = A.prof + ((size_t)blockIdx.x * kWarps + warp) * 8;
#pragma unroll
    for (int i_ = 0; i_ < 6; ++i_) o_[i_] = prof_acc[i_];
    o_[6] = prof_t0;
    o_[7] = gtimer();
  }
#endif
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
                          bool bf16, cudaStream_t stream, bool pdl) {
  const size_t sm = (size_t)plan.smem_bytes;
#ifdef RECOM_TEST_NOLAUNCH
  (void)sm; (void)sched; (void)req; (void)grid; (void)stream; (void)bf16; (void)pdl;
  return;
#endif
  const bool sh = req.world > 0;  // sharded output: rows stored into their owners' tensors
  // pdl: programmatic dependent launch behind the flat kernel of the column tiles - this grid's
  // CTAs may start while that one drains (neither reads what the other writes; every thread of
  // this kernel executes griddepcontrol.wait before it exits, so this launch - and with it the
  // request - is complete only when the flat kernel is).
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = sm;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (bf16) {
    if (sh) cudaLaunchKernelEx(&cfg, fused_columns_kernel<true, true>, plan, sched, req);
    else cudaLaunchKernelEx(&cfg, fused_columns_kernel<true, false>, plan, sched, req);
  } else {
    if (sh) cudaLaunchKernelEx(&cfg, fused_columns_kernel<false, true>, plan, sched, req);
    else cudaLaunchKernelEx(&cfg, fused_columns_kernel<false, false>, plan, sched, req);
  }
}

// ---------------------------------------------------------------------------
// the flat kernel of the column tiles: ONE WARP PER TILE, no persistence
// ---------------------------------------------------------------------------
// A column tile is a chain of dependent loads (record -> slot offsets -> row_splits -> string
// offsets -> string bytes -> table rows); inside the persistent kernel a warp walks several of
// them one after the other next to its FAST tiles, at 20 warps per SM, and the launch lasts as
// long as the sum. Here every tile has its own warp, the grid holds as many warps as there are
// tiles (most expensive first: `items` is sorted by cost), the CTAs are small (4 warps, scratch
// only) and 12 - 16 of them are resident per SM: the chains of ALL tiles overlap, and the launch
// lasts about as long as one of them. The persistent kernel behind it keeps the FAST runs only.
// With programmatic dependent launch (`pdl`) the persistent kernel's CTAs become resident as
// this kernel's last wave drains; the two write disjoint slices of the concat rows.
constexpr int kFlatWarps = 4;
#ifndef RECOM_FLAT_MIN_BLOCKS
#define RECOM_FLAT_MIN_BLOCKS 8
#endif
template <bool BF16, bool SH>
__global__ void __launch_bounds__(kFlatWarps * 32, RECOM_FLAT_MIN_BLOCKS)
column_tiles_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ ReqParams A,
                    const ColItem* __restrict__ items, int num_items) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = (int)blockIdx.x * kFlatWarps + warp;
  // let the dependent launch (the persistent kernel of the FAST runs) become resident as soon
  // as every CTA of this grid has started
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // the whole request blob -> L2 (one 128-byte line per thread and pass): the tiles' second and
  // later levels of dependent loads (string bytes behind their offsets, ...) are L2 hits
  for (int64_t o = ((int64_t)blockIdx.x * (kFlatWarps * 32) + threadIdx.x) * 128; o < A.blob_bytes;
       o += (int64_t)gridDim.x * (kFlatWarps * 32) * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.blob + o));
  if (k >= num_items) {
    if (SH) signal_peers(A, lane);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }
  ColItem it;
  load_item(items + k, it);
#ifdef RECOM_PROFILE
  uint32_t prof_acc[6] = {0, 0, 0, 0, 0, 0};
  const uint64_t prof_t0 = gtimer();
  uint32_t prof_last = (uint32_t)clock64();
  ProfRef pr{prof_acc, &prof_last};
#endif
  ReqRegsT<SH> Q;
  Q.out = A.out;
  Q.bind(A);
  Q.out_row_bytes = A.out_row_bytes;
  Q.status = A.status;
  Q.ids_out = A.ids_out;
  Q.ids_base = A.ids_base;
  Q.check = P.check_ids != 0;
  Q.shard_world = P.shard_world;
  Q.shard_rank = P.shard_rank;
  PROF_T(0)
  PROF_C(0, it.begin + it.count)
  column_tile<BF16>(P, A, Q, it, smem_u32(smem) + (uint32_t)(warp * P.scratch_bytes), lane, SH ? -1 : k PROF_ARG);
  if (SH) signal_peers(A, lane);
  // behind the COO helper with programmatic serialization: complete only when the helper is
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef RECOM_PROFILE
  if (lane == 0 && A.prof && k < A.pad0) {  // (pad0: rows of the timer buffer)
    uint64_t* o_ = A.prof + (size_t)k * 8;
#pragma unroll
    for (int i_ = 0; i_ < 6; ++i_) o_[i_] = prof_acc[i_];
    o_[6] = prof_t0;
    o_[7] = gtimer();
  }
#endif
}

// ---------------------------------------------------------------------------
// long rows: the rows the flat kernel deferred, one CTA (8 warps) per row at a time
// ---------------------------------------------------------------------------
// A row of hundreds or thousands of ids (config 3: zipf lengths) is a chain of transform passes
// and gather rounds for the one warp of its tile, and the launch lasts as long as the longest
// such chain. Here the row's ids are dealt to the CTA's warps in contiguous parts; every warp
// transforms and sums its part (the same code as long_row), the partial sums meet in shared
// memory and warp 0 adds them in warp order: deterministic, no atomics on the output.
template <int VEC, bool BF16>
__device__ __forceinline__ void long_row_cta(const DevPlan& P, const ReqRegsT<false>& Q, const ColItem& it,
                                             const UserRegs& u, const uint8_t* values, const uint8_t* bytes,
                                             int64_t row, int32_t e0, int32_t e1, uint32_t s_ids, int capc,
                                             uint32_t s_tmp, float* s_part, int32_t* s_kept, int warp, int lane) {
  constexpr int NE = Acc<VEC, BF16>::kFloats;
  const int C = u.row_bytes / VEC;
  // this warp's part of the row (multiples of 4 ids)
  const int32_t n = e1 - e0;
  const int32_t per = ((n + kLongRowWarps - 1) / kLongRowWarps + 3) & ~3;
  const int32_t w0 = min(e1, e0 + warp * per), w1 = min(e1, w0 + per);
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int W = min(32, C - c0);
    const int rpp = 32 / W, sub = lane / W;
    const int chb = (c0 + lane - sub * W) * VEC;
    Acc<VEC, BF16> acc;
    acc.clear();
    int32_t kept = 0;
    for (int32_t cb = w0; cb < w1; cb += capc) {
      const int32_t cn = min(w1 - cb, capc);
      transform_ids(P, it, values, bytes, cb, cb + cn, s_ids, lane);
      __syncwarp();
      kept += count_kept(s_ids, cn, lane);
      if (sub < rpp)
        coop_accumulate<VEC, BF16>(Q, u, acc, u.table + chb, chb == 0 ? u.col : -1, rpp, sub, s_ids, cn, cb);
      __syncwarp();
    }
    // the lane groups' partial sums -> lane group 0 of the warp
#pragma unroll
    for (int i = 0; i < NE; ++i)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_tmp + 4u * (uint32_t)(lane * NE + i)), "f"(acc.at(i)) : "memory");
    __syncwarp();
    if (sub == 0) {
      for (int g = 1; g < rpp; ++g) {
#pragma unroll
        for (int i = 0; i < NE; ++i) acc.at(i) += lds_f32(s_tmp + 4u * (uint32_t)((g * W + lane) * NE + i));
      }
#pragma unroll
      for (int i = 0; i < NE; ++i) s_part[(warp * 32 + lane) * NE + i] = acc.at(i);
    }
    if (lane == 0) s_kept[warp] = kept;
    __syncthreads();
    if (warp == 0 && sub == 0) {
      int32_t total = s_kept[0];
      for (int w = 1; w < kLongRowWarps; ++w) {
        total += s_kept[w];
#pragma unroll
        for (int i = 0; i < NE; ++i) acc.at(i) += s_part[(w * 32 + lane) * NE + i];
      }
      ColCursor<false> outp(Q, row, 0, 1, u.out_byte + chb);
      flush_row<VEC, BF16>(acc, total, u.combiner, outp.next());
    }
    __syncthreads();
  }
}

#ifndef RECOM_LONG_MIN_BLOCKS
#define RECOM_LONG_MIN_BLOCKS 2
#endif
template <bool BF16>
__global__ void __launch_bounds__(kLongRowWarps * 32, RECOM_LONG_MIN_BLOCKS)
long_rows_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ ReqParams A,
                 const ColItem* __restrict__ items) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_part[kLongRowWarps * 32 * 8];
  __shared__ int32_t s_kept[kLongRowWarps];
  constexpr int ELEM = BF16 ? 2 : 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t count = min(A.long_list[0], A.long_cap);
  ReqRegsT<false> Q;
  Q.out = A.out;
  Q.bind(A);
  Q.out_row_bytes = A.out_row_bytes;
  Q.status = A.status;
  Q.ids_out = A.ids_out;
  Q.ids_base = A.ids_base;
  Q.check = P.check_ids != 0;
  Q.shard_world = P.shard_world;
  Q.shard_rank = P.shard_rank;
  const uint32_t s_scr = smem_u32(smem) + (uint32_t)(warp * P.scratch_bytes);
  const uint32_t s_ids = s_scr + kColSplitBytes;
  const int cap = (P.scratch_bytes - kColSplitBytes - 1024) >> 2;
  const uint32_t s_tmp = s_scr + (uint32_t)P.scratch_bytes - 1024u;
  for (int32_t e = (int32_t)blockIdx.x; e < count; e += (int32_t)gridDim.x) {
    const int32_t k = A.long_list[2 + 2 * e], row = A.long_list[3 + 2 * e];
    ColItem it;
    load_item(items + k, it);
    const uint8_t* values =
        it.value_slot >= 0 ? A.blob + __ldg(A.slot_offsets + it.value_slot)
                           : (const uint8_t*)(A.ws_split_ids + split_ids_offset(P, A, -1 - it.value_slot, lane));
    const uint8_t* bytes = it.bytes_slot >= 0 ? A.blob + __ldg(A.slot_offsets + it.bytes_slot) : nullptr;
    const int32_t* rs = it.coo_index >= 0 ? A.ws_row_splits + (int64_t)it.coo_index * (A.batch + 1)
                                          : (const int32_t*)(A.blob + __ldg(A.slot_offsets + it.seg_slot));
    const int32_t e0 = __ldg(rs + row), e1 = __ldg(rs + row + 1);
    for (int ui = 0; ui < it.num_users; ++ui) {
      const UserRegs u = ui == 0 ? first_user(A, it) : load_user(P, A, it.first_user + ui);
      if (u.vec_bytes == 16)
        long_row_cta<16, BF16>(P, Q, it, u, values, bytes, row, e0, e1, s_ids, cap, s_tmp, s_part, s_kept, warp, lane);
      else
        long_row_cta<ELEM, BF16>(P, Q, it, u, values, bytes, row, e0, e1, s_ids, cap, s_tmp, s_part, s_kept, warp, lane);
    }
  }
}

int long_rows_configure(bool bf16, int smem_bytes) {
  cudaError_t e = bf16 ? cudaFuncSetAttribute(long_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
                       : cudaFuncSetAttribute(long_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  return e == cudaSuccess ? 0 : -1;
}

void launch_long_rows(const DevPlan& plan, const ReqParams& req, const ColItem* items, int ctas, bool bf16,
                      cudaStream_t stream) {
#ifdef RECOM_TEST_NOLAUNCH
  (void)plan; (void)req; (void)items; (void)ctas; (void)bf16; (void)stream;
  return;
#endif
  const size_t sm = (size_t)kLongRowWarps * plan.scratch_bytes;
  if (bf16) long_rows_kernel<true><<<ctas, kLongRowWarps * 32, sm, stream>>>(plan, req, items);
  else long_rows_kernel<false><<<ctas, kLongRowWarps * 32, sm, stream>>>(plan, req, items);
}

template <bool BF16>
static int configure_flat(int smem_bytes) {
  int n = 0;
  if (cudaFuncSetAttribute(column_tiles_kernel<BF16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess ||
      cudaFuncSetAttribute(column_tiles_kernel<BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, column_tiles_kernel<BF16, false>, kFlatWarps * 32, smem_bytes) != cudaSuccess)
    return -1;
  return n;
}
int column_tiles_configure(bool bf16, int smem_bytes) {
  return bf16 ? configure_flat<true>(smem_bytes) : configure_flat<false>(smem_bytes);
}
int column_tiles_grid(int num_items) { return (num_items + kFlatWarps - 1) / kFlatWarps; }

void launch_column_tiles(const DevPlan& plan, const ReqParams& req, const ColItem* items, int num_items,
                         bool bf16, cudaStream_t stream, bool pdl) {
#ifdef RECOM_TEST_NOLAUNCH
  (void)plan; (void)req; (void)items; (void)num_items; (void)bf16; (void)stream; (void)pdl;
  return;
#endif
  if (num_items <= 0) return;
  const bool sh = req.world > 0;  // sharded output: rows stored into their owners' tensors
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((num_items + kFlatWarps - 1) / kFlatWarps));
  cfg.blockDim = dim3(kFlatWarps * 32);
  cfg.dynamicSmemBytes = (size_t)kFlatWarps * plan.scratch_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;  // behind the COO helper: the tiles over COO inputs wait for it
  if (bf16) {
    if (sh) cudaLaunchKernelEx(&cfg, column_tiles_kernel<true, true>, plan, req, items, num_items);
    else cudaLaunchKernelEx(&cfg, column_tiles_kernel<true, false>, plan, req, items, num_items);
  } else {
    if (sh) cudaLaunchKernelEx(&cfg, column_tiles_kernel<false, true>, plan, req, items, num_items);
    else cudaLaunchKernelEx(&cfg, column_tiles_kernel<false, false>, plan, req, items, num_items);
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
// tf.strings.split + StringToHashBucketFast of a string input (RECOM_SEG_SPLIT_STR; StringSplit:
// string_split_op.cc:37-99 - SplitOnChar / SplitOnCharSet, skip_empty - and :162-170). One block per
// split input: thread = row. Pass 1 counts the row's tokens, a block scan turns the counts into
// row_splits, pass 2 hashes every token (Fingerprint64 of the bytes where they lie) into its slot.
// ---------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ int32_t for_each_token(const uint8_t* __restrict__ p, int32_t len, int32_t spec, F f) {
  const int nd = (spec >> 24) & 15;
  const bool skip_empty = (spec >> 28) & 1;
  const uint32_t d0 = spec & 0xff, d1 = (spec >> 8) & 0xff, d2 = (spec >> 16) & 0xff;
  int32_t n = 0;
  if (len == 0) return 0;  // Split(): an empty string has no tokens
  if (nd == 0) {           // empty delimiter: character by character
    for (int32_t i = 0; i < len; ++i) f(n++, i, 1);
    return n;
  }
  int32_t start = 0;
  for (int32_t i = 0; i <= len; ++i) {
    bool cut = i == len;
    if (!cut) {
      const uint32_t ch = __ldg(p + i);
      cut = ch == d0 || (nd > 1 && ch == d1) || (nd > 2 && ch == d2);
    }
    if (cut) {
      if (!skip_empty || i > start) f(n++, start, i - start);
      start = i + 1;
    }
  }
  return n;
}

__global__ void __launch_bounds__(1024)
split_hash_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ ReqParams A) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DevSplit sp = P.splits[k];
  const int32_t* off = (const int32_t*)(A.blob + A.slot_offsets[sp.value_slot]);
  const uint8_t* bytes = A.blob + A.slot_offsets[sp.bytes_slot];
  int32_t* rs = A.ws_row_splits + (int64_t)(P.num_coo + k) * (A.batch + 1);
  int32_t* ids = A.ws_split_ids + split_ids_offset(P, A, k, lane);
  const int64_t B = A.batch;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  // pass 1 + scan, 1024 rows at a time: rs[r] = tokens in front of row r
  for (int64_t base = 0; base < B; base += 1024) {
    const int64_t r = base + tid;
    int32_t cnt = 0;
    if (r < B) {
      const int32_t o0 = off[r];
      cnt = for_each_token(bytes + o0, off[r + 1] - o0, sp.spec, [](int32_t, int32_t, int32_t) {});
    }
    int32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t w = s_warp[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      s_warp[lane] = wi - w;  // exclusive prefix of the warps
    }
    __syncthreads();
    const int32_t excl = s_carry + s_warp[warp] + incl - cnt;
    if (r < B) rs[r] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + cnt;
    __syncthreads();
  }
  if (tid == 0) rs[B] = s_carry;
  __syncthreads();
  // pass 2: hash every token into its slot
  const uint32_t nbk = (uint32_t)sp.num_buckets;
  for (int64_t r = tid; r < B; r += 1024) {
    const int32_t o0 = off[r];
    int32_t* dst = ids + rs[r];
    const uint8_t* p = bytes + o0;
    for_each_token(p, off[r + 1] - o0, sp.spec, [&](int32_t j, int32_t b, int32_t n) {
      dst[j] = (int32_t)fh::mod_bucket(fh::fingerprint64(p + b, (uint32_t)n), nbk, sp.bucket_magic);
    });
  }
}

void launch_split_hash(const DevPlan& plan, const ReqParams& req, cudaStream_t stream) {
  split_hash_kernel<<<(unsigned)plan.num_splits, 1024, 0, stream>>>(plan, req);
}

// ---------------------------------------------------------------------------
// COO row ids -> CSR row_splits (replaces experiment::ComputeSegmentOffsets,
// cuda_emitter.cc:768-818, which walks tiles serially in one block).
// Every boundary i in [0, nnz] fills row_splits for the rows in (row[i-1], row[i]].
// ---------------------------------------------------------------------------
// Row id of entry i of a COO input. kind 1: SparseReshape inlined (cuda_emitter.cc:1874-1922).
struct CooRows {
  const long long* coo;
  int32_t stride;
  int32_t rank;        // > 0: reshape - linear = sum_j idx[j] * mul[j]; row = linear / div
  long long mul[8];
  long long div;
  __device__ __forceinline__ int64_t row(int64_t i) const {
    if (rank == 0) return (int64_t)__ldg(coo + i * stride);
    long long lin = 0;
    for (int j = 0; j < rank; ++j) lin += __ldg(coo + i * stride + j) * mul[j];
    return (int64_t)(lin / div);
  }
};
__device__ __forceinline__ void fill_row_splits(const CooRows& c, int64_t nnz, int64_t batch,
                                                int32_t* __restrict__ rs, int64_t i) {
  const int64_t prev = i == 0 ? -1 : c.row(i - 1);
  int64_t cur = i == nnz ? batch : c.row(i);
  if (cur > batch) cur = batch;
  for (int64_t r = prev + 1; r <= cur; ++r) rs[r] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
coo_to_row_splits_kernel(const DevCoo* __restrict__ coos, const int64_t* __restrict__ slot_offsets,
                         const int64_t* __restrict__ slot_elems, const uint8_t* __restrict__ blob,
                         int64_t batch, int32_t* __restrict__ ws_row_splits) {
  // the launch behind this one may start right away (programmatic dependent launch): only its
  // tiles over COO inputs read these row_splits, and they wait (griddepcontrol.wait) before they do
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int k = blockIdx.y;
  const DevCoo c = coos[k];
  int32_t* rs = ws_row_splits + (int64_t)k * (batch + 1);
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, dt = (int64_t)gridDim.x * blockDim.x;
  if (c.kind == 2) {  // fixed row length: Sum(axis=1) over [batch, T, D]
    const int64_t n = slot_elems[c.value_slot] - (c.value_str ? 1 : 0);
    const int64_t T = c.seg_stride > 0 ? c.seg_stride : (batch > 0 ? n / batch : 0);
    for (int64_t r = t0; r <= batch; r += dt) rs[r] = (int32_t)(r * T);
    return;
  }
  CooRows R;
  R.coo = (const long long*)(blob + slot_offsets[c.seg_slot]);
  R.stride = c.seg_stride;
  R.rank = 0;
  R.div = 1;
  const int64_t nnz = slot_elems[c.seg_slot] / c.seg_stride;
  if (c.kind == 1) {
    const int rank = c.seg_stride;
    const long long* shape = (const long long*)(blob + slot_offsets[c.seg_slot + 1]);
    R.rank = rank;
    long long m = 1;
    for (int j = rank - 1; j >= 0; --j) {
      R.mul[j] = m;
      m *= __ldg(shape + j);
    }
    if (c.has_new_shape) {
      // row = linear / prod(new_shape[1:]); a -1 entry is the quotient of the total by the others
      const long long* ns = (const long long*)(blob + slot_offsets[c.seg_slot + 2]);
      const int ro = (int)slot_elems[c.seg_slot + 2];
      long long known = 1, tail = 1;
      for (int j = 0; j < ro; ++j) {
        const long long d = __ldg(ns + j);
        if (d >= 0) known *= d;
      }
      for (int j = 1; j < ro; ++j) {
        const long long d = __ldg(ns + j);
        tail *= d < 0 ? (known > 0 ? m / known : 0) : d;
      }
      R.div = tail > 0 ? tail : 1;
    } else {
      R.div = max(1ll, __ldg(shape + rank - 1));
    }
  }
  for (int64_t i = t0; i <= nnz; i += dt) fill_row_splits(R, nnz, batch, rs, i);
}

void launch_coo_to_row_splits(const DevPlan& plan, const ReqParams& req, int64_t max_nnz,
                              cudaStream_t stream) {
  int64_t bx = (max_nnz + 1 + 255) / 256;
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)plan.num_coo);
  coo_to_row_splits_kernel<<<grid, 256, 0, stream>>>(plan.coos, req.slot_offsets, req.slot_elems,
                                                    req.blob, req.batch, req.ws_row_splits);
}

__global__ void __launch_bounds__(256)
segment_ids_to_row_splits_kernel(const long long* __restrict__ coo, int64_t nnz, int32_t stride,
                                 int64_t batch, int32_t* __restrict__ rs) {
  CooRows R;
  R.coo = coo;
  R.stride = stride;
  R.rank = 0;
  R.div = 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz;
       i += (int64_t)gridDim.x * blockDim.x)
    fill_row_splits(R, nnz, batch, rs, i);
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
// Mode B, after the cross-rank reduction of the partial sums: the mean / sqrtn scaling of rows
// [row_begin, row_begin + rows) of the summed concat tensor at `out` (row 0 of `out` = row
// row_begin). One thread per (row, column); n = the row's id count from the request's
// row_splits. TF-CPU scaling as in flush_row.
// ---------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256)
scale_rows_kernel(const DevColumn* __restrict__ cols, const int32_t* __restrict__ out_offsets,
                  int32_t num_cols, const uint8_t* __restrict__ blob,
                  const int64_t* __restrict__ slot_offsets, const int32_t* __restrict__ ws_row_splits,
                  int64_t batch, int64_t row_begin, int64_t rows, int32_t out_width, uint8_t* out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * num_cols;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % num_cols);
    const int64_t r = t / num_cols;
    const DevColumn col = cols[c];
    if (col.lookup != RECOM_LOOKUP_SEGMENT || col.combiner == RECOM_COMBINE_SUM) continue;
    const int32_t* rs = col.coo_index >= 0 ? ws_row_splits + (int64_t)col.coo_index * (batch + 1)
                                           : (const int32_t*)(blob + slot_offsets[col.seg_slot]);
    const int32_t n = rs[row_begin + r + 1] - rs[row_begin + r];
    if (n <= 1) continue;
    const float m = col.combiner == RECOM_COMBINE_MEAN ? (float)n : (float)__dsqrt_rn((double)n);
    const float inv = __fdiv_rn(1.0f, m);
    uint8_t* p = out + (r * out_width + out_offsets[c]) * (BF16 ? 2 : 4);
    for (int e = 0; e < col.dim; ++e) {
      if (BF16) {
        uint16_t* q = (uint16_t*)p + e;
        float x = bf16_lo((uint32_t)*q);
        x = n < 10 ? x * inv : __fdiv_rn(x, m);
        *q = (uint16_t)(pack_bf16(x, 0.f) & 0xffffu);
      } else {
        float* q = (float*)p + e;
        *q = n < 10 ? *q * inv : __fdiv_rn(*q, m);
      }
    }
  }
}

void launch_scale_rows(const DevPlan& plan, const int32_t* d_out_offsets, const ReqParams& req,
                       int64_t row_begin, int64_t rows, int32_t out_width, void* out, bool bf16,
                       cudaStream_t stream) {
  int64_t bx = (rows * plan.num_cols + 255) / 256;
  if (bx > 148 * 16) bx = 148 * 16;
  if (bx < 1) bx = 1;
  if (bf16)
    scale_rows_kernel<true><<<(unsigned)bx, 256, 0, stream>>>(plan.cols, d_out_offsets, plan.num_cols, req.blob,
                                                             req.slot_offsets, req.ws_row_splits, req.batch,
                                                             row_begin, rows, out_width, (uint8_t*)out);
  else
    scale_rows_kernel<false><<<(unsigned)bx, 256, 0, stream>>>(plan.cols, d_out_offsets, plan.num_cols, req.blob,
                                                              req.slot_offsets, req.ws_row_splits, req.batch,
                                                              row_begin, rows, out_width, (uint8_t*)out);
}

// ---------------------------------------------------------------------------
// `output_ptrs` of the reference's op (feature_column_process_op_gpu.cu.cc:118-121 uploads them
// and synchronises): the address of every column's slice in row 0 of the concat tensor.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
output_ptrs_kernel(int32_t num_cols, const int32_t* __restrict__ out_offsets, int32_t elem_bytes,
                   uint8_t* out, long long* __restrict__ ptrs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_cols; i += gridDim.x * blockDim.x)
    ptrs[i] = (long long)(uintptr_t)(out + (int64_t)out_offsets[i] * elem_bytes);
}

void launch_output_ptrs(int32_t num_cols, const int32_t* d_out_offsets, int32_t elem_bytes, void* out,
                        int64_t* ptrs_dev, cudaStream_t stream) {
  const int bx = (num_cols + 255) / 256;
  output_ptrs_kernel<<<bx < 1 ? 1 : bx, 256, 0, stream>>>(num_cols, d_out_offsets, elem_bytes,
                                                        (uint8_t*)out, (long long*)ptrs_dev);
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
