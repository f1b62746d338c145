"""Host-side ops of the reference's surface, over the C ABI (no GPU needed).

  ConcatInputs    Addons>ConcatInputs (tensorflow_addons/custom_ops/concat_inputs/
                  concat_inputs_ops.cc:42-88): n host tensors -> one blob + offsets + shapes.
  ShapeConstruct  Addons>ShapeConstruct (tensorflow_addons/custom_ops/shape_construct/
                  shape_construct_ops.cc:34-183): attrs `symbols`, `indices`, `exprs`; Compute
                  reads input i at flat index indices[i], binds it to symbols[i] and evaluates
                  every expression.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib as L


def concat_inputs(inputs: Sequence[np.ndarray], with_args: bool = False):
    """-> (blob uint8, offsets int64[n], shapes int32[sum of ranks]) like the reference op's three
    outputs (its offsets are int32; every slot here is 16-byte aligned). with_args: the blob also
    carries the argument trailer (recom_concat_inputs_with_args) and its byte offset is returned
    as a fourth value."""
    arrs = [np.ascontiguousarray(a) for a in inputs]
    n = len(arrs)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
    nbytes = np.array([a.nbytes for a in arrs], np.int64)
    elems = np.array([a.size for a in arrs], np.int64)
    offsets = np.zeros(max(n, 1), np.int64)
    need = C.c_size_t()
    args_off = C.c_int64(-1)

    def call(blob_ptr, blob_bytes):
        if with_args:
            L.check(L.lib.recom_concat_inputs_with_args(ptrs, nbytes.ctypes.data, elems.ctypes.data, n,
                                                        blob_ptr, blob_bytes, offsets.ctypes.data,
                                                        C.byref(args_off), C.byref(need)))
        else:
            L.check(L.lib.recom_concat_inputs(ptrs, nbytes.ctypes.data, n, blob_ptr, blob_bytes,
                                              offsets.ctypes.data, C.byref(need)))

    call(None, 0)
    blob = np.empty(need.value, np.uint8)
    call(blob.ctypes.data, blob.nbytes)
    shapes = np.array([d for a in arrs for d in a.shape], np.int32)
    if with_args:
        return blob, offsets[:n], shapes, int(args_off.value)
    return blob, offsets[:n], shapes


class ShapeConstruct:
    def __init__(self, symbols: Sequence[str], indices: Sequence[int], exprs: Sequence[str]):
        if len(symbols) != len(indices):
            # shape_construct_ops.cc:40-44 "Attribute error! Input size inconsistent"
            raise ValueError("Attribute error! Input size inconsistent")
        self.symbols, self.indices, self.exprs = list(symbols), list(indices), list(exprs)
        sym = (C.c_char_p * max(len(symbols), 1))(*[s.encode() for s in symbols])
        ex = (C.c_char_p * max(len(exprs), 1))(*[e.encode() for e in exprs])
        h = C.c_void_p()
        L.check(L.lib.recom_shape_construct_create(sym, len(symbols), ex, len(exprs), C.byref(h)))
        self._h = h

    def compute(self, inputs: Sequence[np.ndarray], dtype=np.int32) -> np.ndarray:
        vals = np.array([int(np.asarray(t).reshape(-1)[i]) for t, i in zip(inputs, self.indices)], np.int64)
        out = np.zeros(len(self.exprs), np.int64)
        L.check(L.lib.recom_shape_construct_eval(self._h, vals.ctypes.data, len(vals), out.ctypes.data, len(out)))
        return out.astype(dtype)

    def close(self):
        if getattr(self, "_h", None):
            L.lib.recom_shape_construct_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
