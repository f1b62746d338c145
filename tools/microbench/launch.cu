// Calibration microbenchmark (not product code): per-step cost of the ways a request's
// argument block can reach the fused kernel, measured back to back on one stream:
//   A. cudaMemcpyAsync(20 KB pinned -> device) + kernel reading it (v3's scheme)
//   B. kernel with the block as a __grid_constant__ parameter of 4 / 12 / 31 KB
//   C. kernel with a tiny parameter only
// The kernel itself writes 64 MB (the concat tensor of model E, B=2048) so that the launch
// gaps are seen next to a realistic body.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int N>
struct Blob {
  uint8_t b[N];
};

__device__ __forceinline__ void body(uint8_t* out, int64_t bytes, uint32_t seed) {
  uint4 v = make_uint4(seed, threadIdx.x, blockIdx.x, 7);
  const int64_t n = bytes / 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(out + i * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(512) k_global(const uint8_t* args, int nargs, uint8_t* out, int64_t bytes) {
  const uint32_t s = args[(threadIdx.x * 37 + blockIdx.x) % nargs];
  body(out, bytes, s);
}
template <int N>
__global__ void __launch_bounds__(512) k_param(const __grid_constant__ Blob<N> args, uint8_t* out, int64_t bytes) {
  const uint8_t* a = args.b;
  const uint32_t s = a[(threadIdx.x * 37 + blockIdx.x) % N];
  body(out, bytes, s);
}

int main() {
  const int64_t bytes = 64ll << 20;
  const int sets = 4;
  uint8_t* out;
  CK(cudaMalloc(&out, bytes * sets));
  uint8_t *h_args, *d_args;
  CK(cudaHostAlloc(&h_args, 32 << 10, cudaHostAllocDefault));
  CK(cudaMalloc(&d_args, 32 << 10));
  for (int i = 0; i < (32 << 10); ++i) h_args[i] = (uint8_t)i;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = 148 * 2, iters = 300;
  static Blob<4000> b4;
  static Blob<12000> b12;
  static Blob<31000> b31;
  auto run = [&](const char* name, auto fn) -> int {
    for (int i = 0; i < 20; ++i) fn(i);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) fn(i);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-44s %7.2f us/step\n", name, ms * 1e3 / iters);
    return 0;
  };
  for (int64_t wb : {bytes, (int64_t)0}) {
    printf("--- kernel writes %lld MB per step\n", (long long)(wb >> 20));
    run("memcpyAsync 20 KB + kernel(global args)", [&](int i) {
      cudaMemcpyAsync(d_args, h_args, 20 << 10, cudaMemcpyHostToDevice, st);
      k_global<<<grid, 512, 0, st>>>(d_args, 20 << 10, out + (i % sets) * bytes, wb);
    });
    run("memsetAsync 16 B + kernel(global args)", [&](int i) {
      cudaMemsetAsync(d_args, 0, 16, st);
      k_global<<<grid, 512, 0, st>>>(d_args, 20 << 10, out + (i % sets) * bytes, wb);
    });
    run("kernel(global args) only", [&](int i) {
      k_global<<<grid, 512, 0, st>>>(d_args, 20 << 10, out + (i % sets) * bytes, wb);
    });
    run("kernel(4 KB param)", [&](int i) { k_param<4000><<<grid, 512, 0, st>>>(b4, out + (i % sets) * bytes, wb); });
    run("kernel(12 KB param)", [&](int i) { k_param<12000><<<grid, 512, 0, st>>>(b12, out + (i % sets) * bytes, wb); });
    run("kernel(31 KB param)", [&](int i) { k_param<31000><<<grid, 512, 0, st>>>(b31, out + (i % sets) * bytes, wb); });
  }
  return 0;
}
