"""Host-side ops of the reference surface (ConcatInputs, ShapeConstruct) through the C ABI. CPU only."""
import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import host_ops, synth
from recom_b200.schema import pack_request


def test_concat_inputs_layout():
    rng = np.random.default_rng(0)
    ins = [rng.integers(0, 100, size=7).astype(np.int32), rng.random((3, 5)).astype(np.float32),
           np.zeros(0, np.int64), rng.integers(0, 9, size=(4, 2)).astype(np.int64),
           np.frombuffer(b"hello world", np.uint8)]
    blob, off, shapes = host_ops.concat_inputs(ins)
    assert blob.size % 16 == 0 and (off % 16 == 0).all()
    for a, o in zip(ins, off):
        np.testing.assert_array_equal(blob[o:o + a.nbytes].view(a.dtype).reshape(a.shape), a)
    assert shapes.tolist() == [7, 3, 5, 0, 4, 2, 11]
    # padding is zeroed (device code reads whole words of string slots)
    o, n = int(off[4]), ins[4].nbytes
    assert (blob[o + n:o + 16] == 0).all()


def test_concat_inputs_matches_python_packer():
    m = synth.model_E_small()
    feats = synth.make_features(m, 33, seed=3)
    req = pack_request(m, 33, feats)
    slot_arrays = [None] * m.num_slots
    for c in m.columns:
        f = feats[c.key]
        if c.value_dtype == L.STR:
            slot_arrays[c.value_slot], slot_arrays[c.bytes_slot] = f.values.offsets, f.values.data
        else:
            slot_arrays[c.value_slot] = f.values
        if c.seg_slot >= 0:
            slot_arrays[c.seg_slot] = f.row_splits
    blob, off, _, args_off = host_ops.concat_inputs(slot_arrays, with_args=True)
    np.testing.assert_array_equal(off, req.slot_offsets)
    assert args_off == req.args_offset and args_off % 16 == 0
    np.testing.assert_array_equal(blob, req.blob)  # slots and the argument trailer alike
    t = blob[args_off:].view(np.int64)
    assert t[1] == m.num_slots
    np.testing.assert_array_equal(t[2:2 + m.num_slots], req.slot_offsets)
    np.testing.assert_array_equal(t[2 + m.num_slots:2 + 2 * m.num_slots], req.slot_elems)
    plain = pack_request(m, 33, feats, trailer=False)
    blob2, off2, _ = host_ops.concat_inputs(slot_arrays)
    np.testing.assert_array_equal(blob2, plain.blob)
    assert plain.args_offset == -1


def test_concat_inputs_large_uses_threads_and_is_exact():
    rng = np.random.default_rng(1)
    ins = [rng.integers(0, 255, size=rng.integers(1, 200000), dtype=np.uint8) for _ in range(64)]
    blob, off, _ = host_ops.concat_inputs(ins)
    for a, o in zip(ins, off):
        np.testing.assert_array_equal(blob[o:o + a.size], a)


def test_shape_construct_expressions():
    sc = host_ops.ShapeConstruct(["x0", "x1", "x12"], [0, 1, 0],
                                 ["x0*x1", "x0 + 1", "(x12 - 1)/2", "max(x0, x12)", "-x1 + x0**2",
                                  "min(x0,x1,x12)", "x1 % 3", "floor(x0/4)", "pow(x1, 3)", "7"])
    out = sc.compute([np.array([6, 99]), np.array([0, 5], np.int64), np.array([9])])
    assert out.tolist() == [30, 7, 4, 9, 31, 5, 2, 1, 125, 7]
    assert out.dtype == np.int32
    assert sc.compute([np.array([6]), np.array([0, 5]), np.array([9])], np.int64).dtype == np.int64


def test_shape_construct_c_int_semantics():
    # the reference compiles `const int x = inputs[i]`: truncating division, int wrap of inputs
    sc = host_ops.ShapeConstruct(["a", "b"], [0, 0], ["a/b", "-a/b", "a - a/b*b"])
    assert sc.compute([np.array([7]), np.array([2])]).tolist() == [3, -3, 1]


def test_shape_construct_errors():
    with pytest.raises(ValueError):
        host_ops.ShapeConstruct(["x0"], [0, 1], ["x0"])
    with pytest.raises(L.RecomError):
        host_ops.ShapeConstruct(["x0"], [0], ["x0 + y"])
    with pytest.raises(L.RecomError):
        host_ops.ShapeConstruct(["x0"], [0], ["x0 +"])
    with pytest.raises(L.RecomError):
        host_ops.ShapeConstruct(["x0"], [0], ["sin(x0)"])
    sc = host_ops.ShapeConstruct(["x0"], [0], ["10/x0"])
    with pytest.raises(L.RecomError):
        sc.compute([np.array([0])])


def test_workspace_sizing_symbols():
    """The two symbols the fused op needs (batch, total ids) size the arena the way
    cuda_emitter.cc:2151-2179 sums alignmem(expr * sizeof): here through expressions."""
    sc = host_ops.ShapeConstruct(["B", "nnz"], [0, 0], ["B + 1", "nnz*4", "(B+1)*4 + nnz*4"])
    assert sc.compute([np.array([2048]), np.array([11000])]).tolist() == [2049, 44000, 52196]
