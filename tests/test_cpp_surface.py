"""The C++ host layer (include/recom_b200_ops.hpp) compiles against the C ABI and works on a
CPU-only box: ConcatInputs, ShapeConstruct, and FeatureColumnProcess failing loudly."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_layer(tmp_path):
    exe = str(tmp_path / "host_ops_smoke")
    lib_dir = os.path.join(ROOT, "recom_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_ops_smoke.cc"), "-o", exe,
                           "-L", lib_dir, "-lrecom_b200", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "CPP_SURFACE_OK" in out.stdout
