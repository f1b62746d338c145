// recom_b200_ops.hpp — C++ host side above the C ABI (include/recom_b200.h), mirroring the
// reference's operator interface for the fused embedding-column path: same op names, argument
// meaning and error behaviour, without TensorFlow types. INTEGRATION.md shows how a TF OpKernel
// forwards to these classes.
//
//   recom::ConcatInputs          Addons>ConcatInputs
//       (tensorflow_addons/custom_ops/concat_inputs/concat_inputs_ops.cc:42-88)
//   recom::ShapeConstruct        Addons>ShapeConstruct
//       (tensorflow_addons/custom_ops/shape_construct/shape_construct_ops.cc:34-183)
//   recom::FeatureColumnProcess  Addons>FeatureColumnProcess[WithSymbols] + Addons>ConcatOutputs
//       (tensorflow_addons/custom_ops/feature_column_process/feature_column_process_op_gpu.cu.cc:33-175,
//        tensorflow_addons/custom_ops/concat_outputs/concat_outputs_op_gpu.cu.cc:180-235)
//
// Errors: the reference reports attribute / argument errors as tensorflow::Status through
// OP_REQUIRES and exits the process on CUDA errors (CubDebugExit). Here every failure is a
// recom::Error carrying the C ABI status and message; nothing exits.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "recom_b200.h"

namespace recom {

class Error : public std::runtime_error {
 public:
  Error(recom_status st, const std::string& what) : std::runtime_error(what), status(st) {}
  recom_status status;
};

inline void check(recom_status st) {
  if (st != RECOM_OK) throw Error(st, recom_last_error());
}

// One host tensor handed to ConcatInputs: a contiguous buffer and its shape.
struct HostTensor {
  const void* data;
  int64_t nbytes;
  std::vector<int32_t> shape;
};

// Addons>ConcatInputs (concat_inputs_ops.cc:27-88): inputs `T` host tensors with ranks `ranks`;
// outputs `concated: int8`, `offsets: int32` (byte offset of every input inside the blob),
// `shapes: int32` (all dims of all inputs, flattened). Same three outputs, same dtypes; inside
// the blob every input starts on a 16-byte boundary and an argument trailer follows the last one
// (recom_concat_inputs_with_args), which the one H2D copy of the blob carries to the device:
// FeatureColumnProcess then uploads nothing per request.
struct ConcatInputs {
  std::vector<uint8_t> concated;
  std::vector<int32_t> offsets;
  std::vector<int32_t> shapes;

  void Compute(const std::vector<HostTensor>& inputs) {
    const int32_t n = (int32_t)inputs.size();
    std::vector<const void*> ptrs(n);
    std::vector<int64_t> nbytes(n), elems(n), off64(n > 0 ? n : 1, 0);
    shapes.clear();
    for (int32_t i = 0; i < n; ++i) {
      ptrs[i] = inputs[i].data;
      nbytes[i] = inputs[i].nbytes;
      int64_t e = 1;
      for (int32_t d : inputs[i].shape) e *= d;
      elems[i] = e;
      shapes.insert(shapes.end(), inputs[i].shape.begin(), inputs[i].shape.end());
    }
    size_t need = 0;
    int64_t args_offset = -1;
    check(recom_concat_inputs_with_args(ptrs.data(), nbytes.data(), elems.data(), n, nullptr, 0,
                                        off64.data(), &args_offset, &need));
    if (need > 0x7fffffffull) throw Error(RECOM_ERR_INVALID_ARGUMENT, "blob exceeds int32 offsets");
    concated.resize(need);
    check(recom_concat_inputs_with_args(ptrs.data(), nbytes.data(), elems.data(), n, concated.data(),
                                        concated.size(), off64.data(), &args_offset, &need));
    offsets.assign(off64.begin(), off64.begin() + n);
  }
};

// Addons>ShapeConstruct: attrs symbols / indices / exprs; Compute(inputs) -> one value per expr.
class ShapeConstruct {
 public:
  ShapeConstruct(const std::vector<std::string>& symbols, const std::vector<int>& indices,
                 const std::vector<std::string>& exprs)
      : indices_(indices), num_exprs_((int32_t)exprs.size()) {
    if (symbols.size() != indices.size())
      throw Error(RECOM_ERR_INVALID_ARGUMENT, "Attribute error! Input size inconsistent");
    std::vector<const char*> s, e;
    for (auto& x : symbols) s.push_back(x.c_str());
    for (auto& x : exprs) e.push_back(x.c_str());
    check(recom_shape_construct_create(s.data(), (int32_t)s.size(), e.data(), (int32_t)e.size(), &h_));
  }
  ~ShapeConstruct() { recom_shape_construct_destroy(h_); }
  ShapeConstruct(const ShapeConstruct&) = delete;
  ShapeConstruct& operator=(const ShapeConstruct&) = delete;

  // inputs[i] = value of input tensor i at flat index indices[i] (the caller indexes its tensors)
  std::vector<int64_t> Compute(const std::vector<int64_t>& picked_inputs) const {
    std::vector<int64_t> out(num_exprs_);
    check(recom_shape_construct_eval(h_, picked_inputs.data(), (int32_t)picked_inputs.size(), out.data(),
                                     num_exprs_));
    return out;
  }
  const std::vector<int>& indices() const { return indices_; }

 private:
  recom_shape_construct* h_ = nullptr;
  std::vector<int> indices_;
  int32_t num_exprs_;
};

// Addons>FeatureColumnProcess / Addons>FeatureColumnProcessWithSymbols
// (feature_column_process_op_gpu.cu.cc:33-175). Same inputs, outputs and attrs as the reference's
// op; the `dlpath` attr carries the serialized plan instead of the path of a JIT-compiled library.
//   inputs   concated_inputs: int8 (device), concated_offsets: int32 (host), concated_shapes: int32
//            (host), inputs: input_types (device: the embedding tables), [symbols: int32 (host)]
//   outputs  output_ptrs: int64 (device, one per column), output_shapes: int32 (host, [batch, dim]
//            per column), buffer: int8 (device: the arena that holds the results)
// `buffer` holds the concatenated [batch, out_width] tensor (every column already in its slice:
// what Addons>ConcatOutputs produces in the reference) followed by the workspace; output_ptrs[i] is
// the address of column i's slice in row 0. Like the reference's op (malloc_buff callback, :107-111)
// the arena is sized here and allocated by the caller's allocator. No host synchronisation.
class FeatureColumnProcess {
 public:
  // allocate `bytes` of device memory for the op's `buffer` output (256-byte aligned)
  using AllocFn = void* (*)(void* ctx, size_t bytes);

  FeatureColumnProcess(const void* plan_blob, size_t bytes, int device) {
    check(recom_plan_create_from_blob(plan_blob, bytes, device, &plan_));
    check(recom_plan_out_width(plan_, &out_width_));
    check(recom_plan_num_columns(plan_, &num_cols_));
    int32_t ns = 0;
    check(recom_plan_num_slots(plan_, &ns));
    slot_ranks_.resize(ns);
    slot_elem_bytes_.resize(ns);
    check(recom_plan_slot_layout(plan_, slot_ranks_.data(), slot_elem_bytes_.data()));
  }
  ~FeatureColumnProcess() { recom_plan_destroy(plan_); }
  FeatureColumnProcess(const FeatureColumnProcess&) = delete;
  FeatureColumnProcess& operator=(const FeatureColumnProcess&) = delete;

  int32_t out_width() const { return out_width_; }
  int32_t num_outputs() const { return num_cols_; }
  int32_t output_rank_sum() const { return 2 * num_cols_; }

  struct Result {
    void* buffer = nullptr;     // the `buffer` output
    size_t buffer_bytes = 0;
    void* concat = nullptr;     // = buffer: [batch, out_width] in the table dtype
    int64_t batch = 0;
  };

  // The reference's Compute, argument for argument. `concated_bytes`: size of the concated_inputs
  // tensor (with_trailer: it ends with the argument trailer of this package's ConcatInputs).
  // `output_ptrs`: device int64[num_outputs()] (may be null: not written), `output_shapes`: host
  // int32[output_rank_sum()].
  Result Compute(const void* concated_inputs, size_t concated_bytes, const int32_t* concated_offsets,
                 const int32_t* concated_shapes, const std::vector<const void*>& inputs,
                 const int32_t* symbols, int32_t num_symbols, int64_t* output_ptrs,
                 int32_t* output_shapes, AllocFn malloc_buff, void* alloc_ctx, recom_stream_t stream,
                 bool with_trailer = true, int32_t* status_out = nullptr) const {
    (void)symbols;      // the sizes the reference evaluates from symbols (ShapeConstruct) all follow
    (void)num_symbols;  // from concated_shapes here
    const size_t ns = slot_ranks_.size();
    std::vector<int64_t> offsets(ns), elems(ns);
    const int32_t* sh = concated_shapes;
    for (size_t s = 0; s < ns; ++s) {
      offsets[s] = concated_offsets[s];
      int64_t e = 1;
      for (int32_t d = 0; d < slot_ranks_[s]; ++d) e *= *sh++;
      elems[s] = e;
    }
    Result r;
    check(recom_plan_infer_batch(plan_, elems.data(), &r.batch));
    size_t ws = 0;
    check(recom_plan_workspace_bytes(plan_, r.batch, 0, &ws));
    const size_t out_bytes = ((size_t)r.batch * out_width_ * elem_bytes() + 255) / 256 * 256;
    r.buffer_bytes = out_bytes + ws;
    r.buffer = malloc_buff(alloc_ctx, r.buffer_bytes);
    if (!r.buffer) throw Error(RECOM_ERR_CUDA, "buffer allocation failed");
    r.concat = r.buffer;
    recom_request q{};
    q.batch = r.batch;
    q.blob = concated_inputs;
    q.slot_offsets = offsets.data();
    q.slot_elems = elems.data();
    q.tables = inputs.data();
    q.out = r.concat;
    q.workspace = (uint8_t*)r.buffer + out_bytes;
    q.workspace_bytes = ws;
    q.status_out = status_out;
    q.stream = stream;
    q.args_offset = with_trailer ? (int64_t)concated_bytes - (int64_t)((16 + 16 * ns + 15) / 16 * 16) : -1;
    check(recom_plan_run(plan_, &q));
    check(recom_plan_output_shapes(plan_, r.batch, output_shapes));
    if (output_ptrs) check(recom_plan_output_ptrs(plan_, r.concat, output_ptrs, stream));
    return r;
  }

  size_t elem_bytes() const { return elem_bytes_; }
  void set_elem_bytes(size_t b) { elem_bytes_ = b; }  // 4 (float tables, default) or 2 (bfloat16)

 private:
  recom_plan* plan_ = nullptr;
  int32_t out_width_ = 0, num_cols_ = 0;
  size_t elem_bytes_ = 4;
  std::vector<int32_t> slot_ranks_, slot_elem_bytes_;
};

// Addons>ConcatOutputs / Addons>ConcatOutputsNoHost (concat_outputs_op_gpu.cu.cc:137-288). Inputs:
// device_input_ptrs: int64 (device), device_input_shapes: int32 (host), [host_inputs: N * T (host)],
// tensor_buffers (kept alive; the last one is FeatureColumnProcess' `buffer`). The device columns
// are already in their slices of the concat tensor at the head of `buffer`: the output aliases
// it (in TensorFlow: forward the `buffer` input, no copy, no launch, none of the reference's two
// synchronisations); host inputs - columns computed outside the fused op - are uploaded and
// scattered into their slices.
class ConcatOutputs {
 public:
  // attrs of the reference's op: N, embedd_dims, device_concat_indices, host_concat_indices,
  // prefix_begin / prefix_end (the range of device_input_shapes that holds the batch dims)
  ConcatOutputs(int N, std::vector<int> embedd_dims, std::vector<int> device_concat_indices,
                std::vector<int> host_concat_indices, int prefix_begin, int prefix_end, size_t elem_bytes = 4)
      : N_(N), dims_(std::move(embedd_dims)), dev_idx_(std::move(device_concat_indices)),
        host_idx_(std::move(host_concat_indices)), pb_(prefix_begin), pe_(prefix_end), es_(elem_bytes) {
    if ((int)host_idx_.size() != N_) throw Error(RECOM_ERR_INVALID_ARGUMENT, "N != host_concat_indices.size()");
    if (dims_.size() != dev_idx_.size() + host_idx_.size())
      throw Error(RECOM_ERR_INVALID_ARGUMENT, "embedd_dims must cover every concat input");
    width_ = 0;
    offs_.resize(dims_.size());
    for (size_t i = 0; i < dims_.size(); ++i) {
      offs_[i] = width_;
      width_ += dims_[i];
    }
  }
  int32_t width() const { return width_; }

  // -> the [prefix..., sum(embedd_dims)] output (device). host_inputs[i]: host tensor of concat
  // input host_concat_indices[i], prefix_size x embedd_dims[...] elements. `d_scratch`: device
  // memory for the uploaded host columns (sum of their bytes + 24 * N, 16-byte aligned), null when N == 0.
  void* Compute(const int64_t* device_input_ptrs, const int32_t* device_input_shapes,
                const std::vector<const void*>& host_inputs, void* fcp_buffer, void* d_scratch,
                recom_stream_t stream) const {
    (void)device_input_ptrs;  // the slices they point to are already in place
    int64_t prefix = 1;
    for (int i = pb_; i < pe_; ++i) prefix *= device_input_shapes[i];
    if (N_ == 0) return fcp_buffer;
    if ((int)host_inputs.size() != N_ || !d_scratch)
      throw Error(RECOM_ERR_INVALID_ARGUMENT, "host_inputs / scratch missing");
    std::vector<int32_t> dims(N_), offs(N_);
    for (int i = 0; i < N_; ++i) {
      dims[i] = dims_[host_idx_[i]];
      offs[i] = offs_[host_idx_[i]];
    }
    check(recom_scatter_host_columns(host_inputs.data(), dims.data(), offs.data(), N_, prefix, width_,
                                     (int32_t)es_, fcp_buffer, d_scratch, stream));
    return fcp_buffer;
  }

 private:
  int N_;
  std::vector<int> dims_, dev_idx_, host_idx_;
  int pb_, pe_;
  size_t es_;
  int32_t width_;
  std::vector<int32_t> offs_;
};

}  // namespace recom
