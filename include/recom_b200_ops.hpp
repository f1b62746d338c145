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

// Addons>ConcatInputs: outputs `concated` (int8 blob), `offsets`, `shapes`.
struct ConcatInputs {
  std::vector<uint8_t> concated;
  std::vector<int64_t> offsets;
  std::vector<int32_t> shapes;

  void Compute(const std::vector<HostTensor>& inputs) {
    const int32_t n = (int32_t)inputs.size();
    std::vector<const void*> ptrs(n);
    std::vector<int64_t> nbytes(n);
    shapes.clear();
    for (int32_t i = 0; i < n; ++i) {
      ptrs[i] = inputs[i].data;
      nbytes[i] = inputs[i].nbytes;
      shapes.insert(shapes.end(), inputs[i].shape.begin(), inputs[i].shape.end());
    }
    offsets.assign(n > 0 ? n : 1, 0);
    size_t need = 0;
    check(recom_concat_inputs(ptrs.data(), nbytes.data(), n, nullptr, 0, offsets.data(), &need));
    concated.resize(need);
    check(recom_concat_inputs(ptrs.data(), nbytes.data(), n, concated.data(), concated.size(),
                              offsets.data(), &need));
    offsets.resize(n);
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

// Addons>FeatureColumnProcess + Addons>ConcatOutputs. Constructed from the serialized plan
// (the payload that replaces the reference's `dlpath` attr); Compute enqueues the whole stage
// on `stream` and writes the final [batch, out_width] tensor. Re-entrant, no host sync.
class FeatureColumnProcess {
 public:
  FeatureColumnProcess(const void* plan_blob, size_t bytes, int device) {
    check(recom_plan_create_from_blob(plan_blob, bytes, device, &plan_));
    check(recom_plan_out_width(plan_, &out_width_));
  }
  ~FeatureColumnProcess() { recom_plan_destroy(plan_); }
  FeatureColumnProcess(const FeatureColumnProcess&) = delete;
  FeatureColumnProcess& operator=(const FeatureColumnProcess&) = delete;

  int32_t out_width() const { return out_width_; }

  // bytes of the `buffer` output the op must allocate for this request (the reference sizes it
  // inside the generated code and allocates through a callback)
  size_t BufferBytes(int64_t batch, int64_t total_ids) const {
    size_t n = 0;
    check(recom_plan_workspace_bytes(plan_, batch, total_ids, &n));
    return n;
  }

  // concated_inputs: device blob; offsets / elems: host arrays (ConcatInputs' outputs);
  // tables: host array of device pointers (op inputs 3..); out: device [batch, out_width].
  void Compute(int64_t batch, const void* concated_inputs, const int64_t* offsets, const int64_t* elems,
               const void* const* tables, void* out, void* buffer, size_t buffer_bytes,
               recom_stream_t stream, int32_t* status_out = nullptr) const {
    recom_request r{};
    r.batch = batch;
    r.blob = concated_inputs;
    r.slot_offsets = offsets;
    r.slot_elems = elems;
    r.tables = tables;
    r.out = out;
    r.workspace = buffer;
    r.workspace_bytes = buffer_bytes;
    r.status_out = status_out;
    r.stream = stream;
    r.args_offset = -1;
    check(recom_plan_run(plan_, &r));
  }

 private:
  recom_plan* plan_ = nullptr;
  int32_t out_width_ = 0;
};

}  // namespace recom
