// Compiled and run by tests/test_gpu_op_surface.py on a GPU box: the reference's op surface
// (Addons>ConcatInputs -> Addons>FeatureColumnProcess[WithSymbols] -> Addons>ConcatOutputs[NoHost];
// feature_column_process_op_gpu.cu.cc:65-175, concat_outputs_op_gpu.cu.cc:180-288,
// concat_inputs_ops.cc:42-88) driven through include/recom_b200_ops.hpp with exactly the
// reference's argument types: concated_inputs int8 (device), concated_offsets int32 (host),
// concated_shapes int32 (host), inputs... (device tables), symbols int32 (host) ->
// output_ptrs int64 (device), output_shapes int32 (host), buffer int8 (device).
// The expected values are computed here from the definition of the ops.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "recom_b200_ops.hpp"

#define CK(x)                                                         \
  do {                                                                \
    cudaError_t e_ = (x);                                             \
    if (e_ != cudaSuccess) {                                          \
      printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); \
      return 100;                                                     \
    }                                                                 \
  } while (0)

static void* alloc_buff(void* ctx, size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  *(void**)ctx = p;
  return p;
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    printf("OP_SURFACE_SKIP no device\n");
    return 0;
  }
  const int B = 300;
  // three columns in the concat row, plus one column (dim 2) computed elsewhere (a host input of
  // ConcatOutputs) between the second and the third:
  //   c0  bucketize(float x; boundaries 0,10,20) -> table0 [4, 8]                     DENSE
  //   c1  ids int64 (multi-hot, row_splits) -> table1 [50, 4], mean                   SEGMENT
  //   (host column, dim 2)
  //   c2  StringToHashBucketFast(str, 7) -> table2 [7, 4]                             DENSE
  const int D0 = 8, D1 = 4, DH = 2, D2 = 4, W = D0 + D1 + DH + D2;
  std::vector<float> boundaries = {0.f, 10.f, 20.f};
  recom_column_desc cols[3];
  memset(cols, 0, sizeof cols);
  cols[0].lookup = RECOM_LOOKUP_DENSE; cols[0].transform = RECOM_XFORM_BUCKETIZE; cols[0].value_dtype = RECOM_F32;
  cols[0].table = 0; cols[0].table_dtype = RECOM_F32; cols[0].table_rows = 4; cols[0].dim = D0; cols[0].out_offset = 0;
  cols[0].value_slot = 0; cols[0].bytes_slot = -1; cols[0].seg_slot = -1; cols[0].num_boundaries = 3;
  cols[1].lookup = RECOM_LOOKUP_SEGMENT; cols[1].combiner = RECOM_COMBINE_MEAN; cols[1].transform = RECOM_XFORM_NONE;
  cols[1].value_dtype = RECOM_I64; cols[1].table = 1; cols[1].table_dtype = RECOM_F32; cols[1].table_rows = 50;
  cols[1].dim = D1; cols[1].out_offset = D0; cols[1].value_slot = 1; cols[1].bytes_slot = -1; cols[1].seg_slot = 2;
  cols[1].seg_format = RECOM_SEG_ROW_SPLITS_I32;
  cols[2].lookup = RECOM_LOOKUP_DENSE; cols[2].transform = RECOM_XFORM_HASH_STR; cols[2].value_dtype = RECOM_STR;
  cols[2].table = 2; cols[2].table_dtype = RECOM_F32; cols[2].table_rows = 7; cols[2].dim = D2;
  cols[2].out_offset = D0 + D1 + DH; cols[2].value_slot = 3; cols[2].bytes_slot = 4; cols[2].seg_slot = -1;
  cols[2].num_buckets = 7;
  recom_plan_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.out_dtype = RECOM_F32; cfg.out_width = W; cfg.num_slots = 5; cfg.num_tables = 3; cfg.check_ids = 1;
  recom_plan* plan = nullptr;
  if (recom_plan_create(cols, 3, boundaries.data(), 3, &cfg, &plan) != RECOM_OK) {
    printf("plan: %s\n", recom_last_error());
    return 1;
  }
  size_t blob_bytes = 0;
  recom_plan_serialize(plan, nullptr, 0, &blob_bytes);
  std::vector<uint8_t> plan_blob(blob_bytes);
  if (recom_plan_serialize(plan, plan_blob.data(), blob_bytes, &blob_bytes) != RECOM_OK) return 2;
  recom_plan_destroy(plan);

  // ---- host inputs (what TF feeds Addons>ConcatInputs) ----
  std::vector<float> x(B);
  std::vector<int32_t> rs(B + 1, 0);
  std::vector<long long> ids;
  std::vector<int32_t> soff(B + 1, 0);
  std::vector<uint8_t> sbytes;
  for (int b = 0; b < B; ++b) {
    x[b] = (float)((b * 7) % 35) - 5.0f + 0.5f * (b % 2);
    const int n = b % 5;  // 0..4 ids, empty rows included
    for (int k = 0; k < n; ++k) ids.push_back((b * 13 + k * 7) % 50);
    rs[b + 1] = (int32_t)ids.size();
    char s[16];
    const int len = snprintf(s, sizeof s, "key%05d", b % 41);
    sbytes.insert(sbytes.end(), s, s + len);
    soff[b + 1] = (int32_t)sbytes.size();
  }
  recom::ConcatInputs ci;
  ci.Compute({{x.data(), (int64_t)x.size() * 4, {B}},
              {ids.data(), (int64_t)ids.size() * 8, {(int32_t)ids.size()}},
              {rs.data(), (int64_t)rs.size() * 4, {B + 1}},
              {soff.data(), (int64_t)soff.size() * 4, {B + 1}},
              {sbytes.data(), (int64_t)sbytes.size(), {(int32_t)sbytes.size()}}});
  if (ci.offsets.size() != 5 || ci.shapes.size() != 5) return 3;

  // ---- tables (the op's `inputs: input_types`) ----
  std::vector<float> t0(4 * D0), t1(50 * D1), t2(7 * D2);
  for (size_t i = 0; i < t0.size(); ++i) t0[i] = 0.25f * (float)i;
  for (size_t i = 0; i < t1.size(); ++i) t1[i] = 1.0f + 0.125f * (float)(i % 97);
  for (size_t i = 0; i < t2.size(); ++i) t2[i] = -3.0f + 0.5f * (float)i;
  float *d_t0, *d_t1, *d_t2;
  CK(cudaMalloc(&d_t0, t0.size() * 4)); CK(cudaMalloc(&d_t1, t1.size() * 4)); CK(cudaMalloc(&d_t2, t2.size() * 4));
  CK(cudaMemcpy(d_t0, t0.data(), t0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_t1, t1.data(), t1.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_t2, t2.data(), t2.size() * 4, cudaMemcpyHostToDevice));

  cudaStream_t stream;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  void* d_blob;
  CK(cudaMalloc(&d_blob, ci.concated.size()));
  CK(cudaMemcpyAsync(d_blob, ci.concated.data(), ci.concated.size(), cudaMemcpyHostToDevice, stream));

  try {
    recom::FeatureColumnProcess fcp(plan_blob.data(), plan_blob.size(), 0);
    if (fcp.num_outputs() != 3 || fcp.out_width() != W) return 4;
    int64_t* d_output_ptrs;
    CK(cudaMalloc(&d_output_ptrs, 3 * sizeof(int64_t)));
    std::vector<int32_t> output_shapes(fcp.output_rank_sum());
    int32_t* d_status;
    CK(cudaMalloc(&d_status, 4));
    CK(cudaMemsetAsync(d_status, 0, 4, stream));
    void* buffer = nullptr;
    const int32_t symbols[2] = {B, (int32_t)ids.size()};  // the WithSymbols variant's extra input
    auto res = fcp.Compute(d_blob, ci.concated.size(), ci.offsets.data(), ci.shapes.data(),
                           {d_t0, d_t1, d_t2}, symbols, 2, d_output_ptrs, output_shapes.data(),
                           alloc_buff, &buffer, stream, /*with_trailer=*/true, d_status);
    if (res.batch != B || res.buffer != buffer) return 5;
    const int32_t want_shapes[6] = {B, D0, B, D1, B, D2};
    if (memcmp(output_shapes.data(), want_shapes, sizeof want_shapes) != 0) return 6;

    // Addons>ConcatOutputs: concat inputs 0,1 (device) 2 (host) 3 (device)
    recom::ConcatOutputs co(/*N=*/1, {D0, D1, DH, D2}, {0, 1, 3}, {2}, /*prefix_begin=*/0, /*prefix_end=*/1);
    std::vector<float> hostcol((size_t)B * DH);
    for (size_t i = 0; i < hostcol.size(); ++i) hostcol[i] = 100.0f + (float)i;
    void* d_scratch;
    CK(cudaMalloc(&d_scratch, hostcol.size() * 4 + 64));
    float* d_out = (float*)co.Compute(d_output_ptrs, output_shapes.data(), {hostcol.data()}, res.buffer,
                                      d_scratch, stream);
    if (d_out != res.concat || co.width() != W) return 7;

    std::vector<float> got((size_t)B * W);
    std::vector<int64_t> ptrs(3);
    int32_t status = -1;
    CK(cudaMemcpyAsync(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(ptrs.data(), d_output_ptrs, 24, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (status != 0) return 8;
    if (ptrs[0] != (int64_t)(uintptr_t)d_out || ptrs[1] != (int64_t)(uintptr_t)(d_out + D0) ||
        ptrs[2] != (int64_t)(uintptr_t)(d_out + D0 + D1 + DH))
      return 9;

    // ---- expected, from the definition of the ops ----
    for (int b = 0; b < B; ++b) {
      const float* row = got.data() + (size_t)b * W;
      int bucket = 0;  // upper_bound over {0, 10, 20}
      while (bucket < 3 && !(x[b] < boundaries[bucket])) ++bucket;
      for (int e = 0; e < D0; ++e)
        if (row[e] != t0[bucket * D0 + e]) { printf("c0 row %d\n", b); return 10; }
      const int n = rs[b + 1] - rs[b];
      for (int e = 0; e < D1; ++e) {
        float acc = 0.f;
        for (int k = rs[b]; k < rs[b + 1]; ++k) acc += t1[ids[k] * D1 + e];
        if (n > 1) acc *= 1.0f / (float)n;
        if (std::fabs(row[D0 + e] - acc) > 1e-5f * (1.0f + std::fabs(acc))) { printf("c1 row %d\n", b); return 11; }
      }
      for (int e = 0; e < DH; ++e)
        if (row[D0 + D1 + e] != hostcol[(size_t)b * DH + e]) { printf("host col row %d\n", b); return 12; }
    }
    // c2: the device hash must agree with the stand-alone StringToHashBucketFast of the C ABI
    std::vector<int64_t> buckets(B);
    {
      uint8_t* d_b; int32_t* d_o; int64_t* d_h;
      CK(cudaMalloc(&d_b, sbytes.size() + 16)); CK(cudaMalloc(&d_o, soff.size() * 4)); CK(cudaMalloc(&d_h, B * 8));
      CK(cudaMemset(d_b, 0, sbytes.size() + 16));
      CK(cudaMemcpy(d_b, sbytes.data(), sbytes.size(), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_o, soff.data(), soff.size() * 4, cudaMemcpyHostToDevice));
      recom::check(recom_string_to_hash_bucket_fast(d_b, d_o, B, 7, d_h, stream));
      CK(cudaMemcpyAsync(buckets.data(), d_h, B * 8, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
    }
    for (int b = 0; b < B; ++b)
      for (int e = 0; e < D2; ++e)
        if (got[(size_t)b * W + D0 + D1 + DH + e] != t2[buckets[b] * D2 + e]) { printf("c2 row %d\n", b); return 13; }
    // "omar" % 10 == 6 (feature_column_v2_test.py:853-872) pins the hash itself in tests/test_oracle_golden.py
  } catch (const recom::Error& e) {
    printf("recom::Error %d: %s\n", (int)e.status, e.what());
    return 50;
  }
  printf("OP_SURFACE_OK\n");
  return 0;
}
