"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares.
No compute calls here (no GPU); argument validation that needs no device is exercised."""
import ctypes as C
import glob
import os
import re

import numpy as np

from recom_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"RECOM_API\s+[\w\s\*]+?\b(recom_\w+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_what_the_binding_lists():
    assert declared_symbols() == sorted(L.EXPORTS)


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(L.LIB_PATH)
    for name in declared_symbols():
        assert getattr(lib, name) is not None, name


def test_no_unexpected_exports():
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    syms = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert {s for s in syms if s.startswith("recom_")} == set(L.EXPORTS)
    # the product library must not carry (or link) the oracle
    assert not any("oracle" in s for s in syms)
    ldd = subprocess.check_output(["ldd", L.LIB_PATH], text=True)
    assert "oracle" not in ldd


def test_abi_version_and_struct_layout():
    assert L.lib.recom_abi_version() == 4
    assert L.COLUMN_DESC.itemsize == 80
    assert C.sizeof(L.PlanConfig) == 32
    assert C.sizeof(L.Request) == 96


def test_null_arguments_are_reported_not_fatal():
    assert L.lib.recom_plan_create(None, 0, None, 0, None, None) == L.ERR_INVALID_ARGUMENT
    assert b"null" in L.lib.recom_last_error()
    n = C.c_int32()
    assert L.lib.recom_plan_num_columns(None, C.byref(n)) == L.ERR_INVALID_ARGUMENT
    assert L.lib.recom_plan_destroy(None) == L.OK
    assert L.lib.recom_plan_create_from_blob(b"xx", 2, 0, C.byref(C.c_void_p())) == L.ERR_INVALID_ARGUMENT


def test_invalid_descriptors_are_rejected_before_touching_the_device():
    from recom_b200 import synth
    m = synth.microbenchmark_model(4)
    descs, bnd = m.descs()
    cfg = L.PlanConfig(out_dtype=L.F32, out_width=m.out_width, num_slots=m.num_slots,
                       num_tables=len(m.table_shapes), check_ids=0, device=0, max_ctas_per_sm=0, row_shard=0)
    h = C.c_void_p()

    def create(d, b=bnd, cfg=cfg):
        return L.lib.recom_plan_create(d.ctypes.data, len(d), b.ctypes.data, b.size, C.byref(cfg), C.byref(h))

    bad = descs.copy(); bad[1]["out_offset"] = bad[0]["out_offset"]
    assert create(bad) == L.ERR_INVALID_ARGUMENT and b"overlap" in L.lib.recom_last_error()
    bad = descs.copy(); bad[0]["table"] = 99
    assert create(bad) == L.ERR_INVALID_ARGUMENT
    bad = descs.copy(); bad[0]["transform"] = 17
    assert create(bad) == L.ERR_INVALID_ARGUMENT
    bad = descs.copy(); bad[2]["dim"] = 0
    assert create(bad) == L.ERR_INVALID_ARGUMENT
    bad = descs.copy(); bad[0]["table_dtype"] = L.BF16
    assert create(bad) == L.ERR_UNSUPPORTED
    unsorted = bnd.copy(); unsorted[3] = -1.0  # bucketize_op.cc:57-59 "Expected sorted boundaries"
    assert create(descs, unsorted) == L.ERR_INVALID_ARGUMENT and b"sorted" in L.lib.recom_last_error()
    # a well-formed plan on a box without a GPU must fail loudly (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        assert create(descs) in (L.ERR_CUDA, L.ERR_INVALID_ARGUMENT)


def test_product_package_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "recom_b200", "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
            text = open(path, errors="replace").read()
            assert "oracle" not in text.lower() or path.endswith("__init__.py") and "import oracle" not in text, path
