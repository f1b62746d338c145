"""The reference's op surface driven from C++ on the GPU (tests/cpp/op_surface_gpu.cc):
ConcatInputs -> FeatureColumnProcess[WithSymbols] -> ConcatOutputs with the reference's exact
argument lists (feature_column_process_op_gpu.cu.cc:140-175, concat_outputs_op_gpu.cu.cc:255-288)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_reference_op_surface_cpp(tmp_path):
    exe = str(tmp_path / "op_surface_gpu")
    lib_dir = os.path.join(ROOT, "recom_b200")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-Wno-deprecated-gpu-targets", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "op_surface_gpu.cc"), "-o", exe,
                           "-L", lib_dir, "-lrecom_b200", "-Xlinker", f"-rpath={lib_dir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "OP_SURFACE_OK" in out.stdout
