// dense_layer.cu - SURVEY 8(f) N4: the layer on the far side of the path.
//
// The reference hands the concat tensor to TF's MatMul + BiasAdd + Relu of the top MLP
// (examples/python/dlrm.py:175-182; its post_lookup_optimizer.cc:227-309 only sketches sparse-aware
// rewrites in code that is never instantiated). Here the first dense layer is one kernel on the
// 5th-generation tensor cores that consumes the bf16 concat rows the stage has just written (they
// are still in the 126 MB L2: 33 MB for model E at batch 2048):
//
//     out[M, N] = relu(A[M, K] * W[N, K]^T + bias[N])      A, W bf16 (K-major), fp32 accumulate
//
// One CTA per 128 x 128 output tile, 6 warps with fixed roles:
//   warp 0   TMA producer: 128 x 64 bf16 boxes of A and W (128-byte swizzle) into a 4-stage ring,
//            completion on `full` mbarriers (cp.async.bulk.tensor, expect_tx)
//   warp 1   allocates 128 TMEM columns; its lane 0 issues tcgen05.mma (UMMA 128x128x16, 4 per
//            stage, operands straight from the swizzled shared-memory tiles through matrix
//            descriptors) and releases every stage with tcgen05.commit -> `empty` mbarrier; the last
//            commit signals the epilogue
//   warps 2-5 epilogue: tcgen05.ld of the warp's 32 accumulator lanes (32 columns at a time), bias,
//            ReLU, 128-byte row segments to global memory
// K needs no padding (TMA zero-fills the box past the tensor's end); M and N are guarded in the
// epilogue. Bound: tensor pipe for large M (2 * M * N * K flops), else the read of W.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/recom_b200.h"

namespace recom {
namespace {

// Two tile shapes: 128 x 128 with a 6-stage ring (small M x N: more CTAs than a 128 x 256 grid
// would give), 128 x 256 with 4 stages (large grids: a quarter less operand traffic per flop out
// of L2, which is what bounds a 128 x 128 tile at a third of the tensor peak).
constexpr int BM = 128, BK = 64, kUmmaK = 16;
constexpr int kStageBytesA = BM * BK * 2;
constexpr int kDenseThreads = 192;
template <int BN, int STAGES>
struct DenseCfg {
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kSmem = STAGES * (kStageBytesA + kStageBytesB) + 256 + 1024;  // + barriers + alignment slack
  // D fp32, A / B bf16, both K-major, M = 128, N = BN (cute::UMMA::InstrDescriptor)
  static constexpr uint32_t kInstrDesc =
      (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile with the 128-byte swizzle (rows of 64 bf16 = 128 bytes, atoms of 8 rows =
// 1024 bytes): start address, stride byte offset 1024, descriptor version 1 (sm_100), layout
// SWIZZLE_128B (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int kStages>
__global__ void __launch_bounds__(kDenseThreads, 1)
dense_relu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const float* __restrict__ bias, float* __restrict__ out, int M, int N, int K, int relu) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);  // the swizzle atoms are 1024-byte aligned
  constexpr int kStageBytesB = DenseCfg<BN, kStages>::kStageBytesB;
  constexpr int kTmemCols = BN;
  const uint32_t s_a = s32(smem), s_b = s_a + kStages * kStageBytesA;
  uint8_t* ctl = smem + kStages * (kStageBytesA + kStageBytesB);
  const uint32_t s_full = s32(ctl), s_empty = s_full + 8 * kStages, s_tfull = s_empty + 8 * kStages;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(ctl + 8 * (2 * kStages + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nk = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(s_full + 8 * s, 1);
      mbar_init(s_empty + 8 * s, 1);
    }
    mbar_init(s_tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32((const void*)tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer ----
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kStages;
      mbar_wait(s_empty + 8 * s, ((kb / kStages) & 1) ^ 1);
      mbar_expect_tx(s_full + 8 * s, kStageBytesA + kStageBytesB);
      tma_load_2d(s_a + s * kStageBytesA, &map_a, s_full + 8 * s, kb * BK, m0);
      tma_load_2d(s_b + s * kStageBytesB, &map_w, s_full + 8 * s, kb * BK, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kStages;
      mbar_wait(s_full + 8 * s, (kb / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int k = 0; k < BK / kUmmaK; ++k) {
        const uint64_t da = umma_desc(s_a + s * kStageBytesA + k * kUmmaK * 2);
        const uint64_t db = umma_desc(s_b + s * kStageBytesB + k * kUmmaK * 2);
        umma_f16(tmem, da, db, DenseCfg<BN, kStages>::kInstrDesc, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(s_empty + 8 * s);  // the stage is free once these MMAs have read it
    }
    umma_commit(s_tfull);  // the accumulator is complete
  } else if (warp >= 2) {
    // ---- epilogue: a warp reads the TMEM lanes [32 * (warp % 4), + 32) ----
    mbar_wait(s_tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    float* orow = out + (int64_t)row * N;
    const bool vec = (N & 3) == 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      const int col0 = n0 + c * 32;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]);
        if (bias != nullptr && col0 + j < N) x += __ldg(bias + col0 + j);
        v[j] = relu ? fmaxf(x, 0.f) : x;
      }
      if (row < M) {
        if (vec && col0 + 32 <= N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) orow[col0 + j] = v[j];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// [rows, K] bf16, row stride `ld` elements, boxes of box_rows rows x 64 columns, 128-byte swizzle
bool make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// -> 0 ok, 1 bad argument, 2 driver entry point / tensor map, 3 launch
int launch_dense_relu(const void* a, int64_t M, int32_t K, int64_t lda, const void* w, int32_t N,
                      const float* bias, int relu, float* out, cudaStream_t stream, const char** why) {
  static const char* msg[] = {"", "A / W must be 16-byte aligned, K and lda multiples of 8, M, N, K > 0",
                              "cuTensorMapEncodeTiled is unavailable or rejected the operand", "kernel launch failed"};
  *why = msg[0];
  if (!a || !w || !out || M <= 0 || N <= 0 || K <= 0 || (K & 7) || (lda & 7) || lda < K ||
      ((uintptr_t)a & 15) || ((uintptr_t)w & 15)) {
    *why = msg[1];
    return 1;
  }
  // 128 x 256 tiles when they still fill most of the GPU, else 128 x 128
  const int64_t tiles_m = (M + BM - 1) / BM;
  const bool wide = N >= 256 && tiles_m * ((N + 255) / 256) >= 120;
  const int bn = wide ? 256 : 128;
  CUtensorMap ma, mw;
  if (!make_map(&ma, a, M, K, lda, BM) || !make_map(&mw, w, N, K, K, bn)) {
    *why = msg[2];
    return 2;
  }
  dim3 grid((unsigned)tiles_m, (unsigned)((N + bn - 1) / bn));
  cudaError_t e;
  if (wide) {
    auto fn = dense_relu_kernel<256, 4>;
    constexpr int smem = DenseCfg<256, 4>::kSmem;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) fn<<<grid, kDenseThreads, smem, stream>>>(ma, mw, bias, out, (int)M, N, K, relu);
  } else {
    auto fn = dense_relu_kernel<128, 6>;
    constexpr int smem = DenseCfg<128, 6>::kSmem;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) fn<<<grid, kDenseThreads, smem, stream>>>(ma, mw, bias, out, (int)M, N, K, relu);
  }
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    *why = msg[3];
    return 3;
  }
  return 0;
}

}  // namespace recom
