// Calibration microbenchmarks for the B200 box (not product code):
//  1. dependent-load latency: L2-resident / DRAM-resident pointer chase, 1 thread per SM and under load
//  2. global atomicAdd round trip, one hot address, all CTAs
//  3. TMA bulk copy global->shared latency for 400 B / 3232 B / 16 x 3232 B
//  4. store throughput: 444 CTAs x 256 threads, warp writes 512 contiguous bytes, row stride 32480 B
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void chase(const uint32_t* __restrict__ next, uint32_t nwords, int steps, uint32_t start_stride, long long* cycles, uint32_t* sink) {
  if (threadIdx.x != 0) return;
  uint32_t i = (uint32_t)(((size_t)blockIdx.x * start_stride) % nwords);
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) i = __ldcg(next + i);
  long long t1 = clock64();
  cycles[blockIdx.x] = (t1 - t0) / steps;
  sink[blockIdx.x] = i;
}

__global__ void atomics(int* ctr, int reps, long long* cycles) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  int acc = 0;
  for (int s = 0; s < reps; ++s) acc += atomicAdd(ctr, 1) & 1;  // dependent on the result
  long long t1 = clock64();
  cycles[blockIdx.x] = (t1 - t0) / reps + (acc == -1);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void bulk(const uint8_t* src, int bytes_each, int copies, long long* cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x < copies) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"(smem_u32(sm + threadIdx.x * bytes_each)), "l"(src + (size_t)(blockIdx.x * copies + threadIdx.x) * bytes_each), "r"(bytes_each), "r"(smem_u32(&bar)) : "memory");
  }
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes_each * copies) : "memory");
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// every warp writes rows_per_warp rows of 512 B (lane = 16 B chunk) into its CTA's slab span
__global__ void stores(uint8_t* out, int64_t row_bytes, int rows_per_cta, int slabs, int cs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x % slabs;
  const int64_t row0 = (int64_t)(blockIdx.x / slabs) * rows_per_cta;
  uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
  for (int r = warp; r < rows_per_cta; r += 8) {
    uint8_t* p = out + (row0 + r) * row_bytes + (int64_t)slab * 512 + lane * 16;
    if (cs) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else *(uint4*)p = v;
  }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s SMs %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  const double ns_per_cyc = 1e6 / prop.clockRate;
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 4096 * 8));
  uint32_t* d_sink; CK(cudaMalloc(&d_sink, 4096 * 4));
  std::vector<long long> h(4096);
  // 1. pointer chase
  for (size_t mb : {1, 32, 1024}) {
    size_t n = mb * 1024 * 1024 / 4;
    std::vector<uint32_t> nx(n);
    // stride permutation: jump by a large odd stride modulo n (in units of 32 words = 128 B lines)
    size_t lines = n / 32;
    for (size_t l = 0; l < lines; ++l) for (int k = 0; k < 32; ++k) nx[l * 32 + k] = (uint32_t)(((l * 7919 + 12345) % lines) * 32);
    uint32_t* d_next; CK(cudaMalloc(&d_next, n * 4));
    CK(cudaMemcpy(d_next, nx.data(), n * 4, cudaMemcpyHostToDevice));
    for (int blocks : {1, 148, 444}) {
      chase<<<blocks, 32>>>(d_next, (uint32_t)n, 200, 977 * 32, d_cyc, d_sink);  // warm
      chase<<<blocks, 32>>>(d_next, (uint32_t)n, 2000, 977 * 32, d_cyc, d_sink);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_cyc, blocks * 8, cudaMemcpyDeviceToHost));
      long long mx = *std::max_element(h.begin(), h.begin() + blocks), mn = *std::min_element(h.begin(), h.begin() + blocks);
      printf("chase %4zu MB, %3d CTAs: %lld..%lld cycles/load (%.0f..%.0f ns)\n", mb, blocks, mn, mx, mn * ns_per_cyc, mx * ns_per_cyc);
    }
    cudaFree(d_next);
  }
  // 2. atomics
  int* d_ctr; CK(cudaMalloc(&d_ctr, 4)); CK(cudaMemset(d_ctr, 0, 4));
  for (int blocks : {1, 148, 444}) {
    atomics<<<blocks, 32>>>(d_ctr, 200, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_cyc, blocks * 8, cudaMemcpyDeviceToHost));
    long long mx = *std::max_element(h.begin(), h.begin() + blocks);
    printf("atomicAdd same address, %3d CTAs: up to %lld cycles (%.0f ns) per dependent atomic\n", blocks, mx, mx * ns_per_cyc);
  }
  // 3. bulk copies
  uint8_t* d_src; CK(cudaMalloc(&d_src, 64 << 20)); CK(cudaMemset(d_src, 1, 64 << 20));
  CK(cudaFuncSetAttribute(bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10));
  for (int copies : {1, 16}) for (int bytes : {400, 3232}) for (int blocks : {1, 444}) {
    bulk<<<blocks, 256, 64 << 10>>>(d_src, bytes, copies, d_cyc);
    bulk<<<blocks, 256, 64 << 10>>>(d_src, bytes, copies, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_cyc, blocks * 8, cudaMemcpyDeviceToHost));
    long long mx = *std::max_element(h.begin(), h.begin() + blocks), mn = *std::min_element(h.begin(), h.begin() + blocks);
    printf("bulk g2s %2d x %4d B, %3d CTAs: %lld..%lld cycles (%.0f..%.0f ns)\n", copies, bytes, blocks, mn, mx, mn * ns_per_cyc, mx * ns_per_cyc);
  }
  // 4. stores: [2048, 8120] fp32 rows, 62 slabs x 7 CTAs, 292 rows each (444 CTAs)
  const int64_t row_bytes = 8120 * 4; const int rows = 2048;
  uint8_t* d_out[4];
  for (int i = 0; i < 4; ++i) CK(cudaMalloc(&d_out[i], rows * row_bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int cs = 0; cs < 2; ++cs) for (int ctas_per_slab : {7, 14, 28}) {
    int rpc = rows / ctas_per_slab;
    for (int w = 0; w < 3; ++w) stores<<<62 * ctas_per_slab, 256>>>(d_out[w % 4], row_bytes, rpc, 62, cs);
    cudaEventRecord(e0);
    for (int w = 0; w < 20; ++w) stores<<<62 * ctas_per_slab, 256>>>(d_out[w % 4], row_bytes, rpc, 62, cs);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = 62.0 * ctas_per_slab * rpc * 512;
    printf("stores cs=%d, %4d CTAs: %.2f us per launch, %.0f GB/s\n", cs, 62 * ctas_per_slab, ms / 20 * 1e3, bytes / (ms / 20 * 1e-3) / 1e9);
  }
  return 0;
}
