// Compiled and run by tests/test_cpp_surface.py: the C++ host layer on a CPU-only box.
#include <cstdio>
#include <cstring>

#include "recom_b200_ops.hpp"

int main() {
  int32_t a[5] = {1, 2, 3, 4, 5};
  float b[3] = {0.5f, 1.5f, 2.5f};
  recom::ConcatInputs ci;
  ci.Compute({{a, sizeof a, {5}}, {b, sizeof b, {3, 1}}});
  if (ci.offsets.size() != 2 || ci.offsets[0] != 0 || ci.offsets[1] != 32) return 1;
  if (memcmp(ci.concated.data() + 32, b, sizeof b) != 0) return 2;
  if (ci.shapes != std::vector<int32_t>({5, 3, 1})) return 3;

  recom::ShapeConstruct sc({"x0", "x1"}, {0, 0}, {"x0*x1", "(x0+1)/2"});
  auto out = sc.Compute({7, 3});
  if (out[0] != 21 || out[1] != 4) return 4;

  try {
    recom::ShapeConstruct bad({"x0"}, {0, 1}, {"x0"});
    return 5;
  } catch (const recom::Error& e) {
    if (e.status != RECOM_ERR_INVALID_ARGUMENT) return 6;
  }
  try {
    recom::FeatureColumnProcess fcp("not a plan", 10, 0);  // must fail loudly, never exit
    return 7;
  } catch (const recom::Error& e) {
    if (e.status != RECOM_ERR_INVALID_ARGUMENT) return 8;
  }
  printf("CPP_SURFACE_OK abi=%d\n", recom_abi_version());
  return 0;
}
