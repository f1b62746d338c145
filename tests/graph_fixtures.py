"""Hand-written GraphDefs shaped like what tf.compat.v1.feature_column.input_layer emits for the
reference's example models (examples/python/dlrm.py:56-137): one safe_embedding_lookup_sparse
chain per embedding column (embedding_ops.py:882-941 - SparseReshape, prune ids < 0,
SparseFillEmptyRows, Unique, GatherV2 on the table, SparseSegment{Sum,Mean,SqrtN}, zero the empty
rows with Select(Tile(Reshape(empty)), ZerosLike, result), Reshape back), all feeding one ConcatV2.
The op sequence is the one lookup_optimizer.cc:62-140 matches on real graphs. TensorFlow is not in
this image, so the graphs are assembled node by node."""
import numpy as np

from recom_b200 import graphdef as G


class Builder:
    def __init__(self, frozen=False):
        self.nodes = []
        self.frozen = frozen

    def add(self, name, op, inputs=(), **attr):
        self.nodes.append(G.Node(name, op, list(inputs), dict(attr)))
        return name

    def const(self, name, arr):
        arr = np.asarray(arr)
        dt = G.DT_STRING if arr.dtype == object else G._DT_OF_NP[arr.dtype]
        return self.add(name, "Const", value=arr, dtype=("type", dt))

    # ---- the tail shared by every embedding column ----
    def lookup(self, scope, indices, values, dense_shape, rows, dim, combiner="mean", rank=2, shared_read=None):
        a = self.add
        if shared_read is not None:  # shared_embedding_columns: another column's table
            return self._lookup_tail(scope, indices, values, dense_shape, dim, combiner, shared_read)
        if self.frozen:  # convert_variables_to_constants: the variable has become a Const holding the table
            rng = np.random.default_rng(len(self.nodes))
            table = self.const(f"{scope}/embedding_weights", (rng.standard_normal((rows, dim)) / np.sqrt(dim)).astype(np.float32))
        else:
            table = a(f"{scope}/embedding_weights", "VariableV2", shape=("shape", [rows, dim]), dtype=("type", G.DT_FLOAT))
            a(f"{scope}/embedding_weights/Assign", "Assign", [table, self.const(f"{scope}/init", np.zeros((1,), np.float32))])
            a("save/SaveV2/" + scope, "SaveV2", [table])
        read = a(f"{scope}/embedding_weights/read", "Identity", [table])
        self.last_read = read
        return self._lookup_tail(scope, indices, values, dense_shape, dim, combiner, read)

    def _lookup_tail(self, scope, indices, values, dense_shape, dim, combiner, read):
        a = self.add
        # sparse_reshape to [prod(shape[:-1]), shape[-1]]
        new_shape = a(f"{scope}/new_shape", "Pack", [dense_shape, dense_shape])
        sr = a(f"{scope}/SparseReshape", "SparseReshape", [indices, dense_shape, new_shape])
        # _prune_invalid_ids: sparse_retain(ids, ids >= 0)
        ge = a(f"{scope}/GreaterEqual", "GreaterEqual", [values, self.const(f"{scope}/zero", np.int64(0))])
        wh = a(f"{scope}/Where", "Where", [ge])
        keep = a(f"{scope}/Reshape", "Reshape", [wh, self.const(f"{scope}/m1", np.array([-1], np.int32))])
        ax = self.const(f"{scope}/axis", np.int32(0))
        pi = a(f"{scope}/GatherV2", "GatherV2", [sr, keep, ax])
        pv = a(f"{scope}/GatherV2_1", "GatherV2", [values, keep, ax])
        sfer = a(f"{scope}/SparseFillEmptyRows", "SparseFillEmptyRows",
                 [pi, pv, sr + ":1", self.const(f"{scope}/default", np.int64(0))])
        ss = a(f"{scope}/strided_slice", "StridedSlice",
               [sfer, self.const(f"{scope}/b", np.array([0, 0], np.int32)),
                self.const(f"{scope}/e", np.array([0, 1], np.int32)),
                self.const(f"{scope}/s", np.array([1, 1], np.int32))],
               begin_mask=1, end_mask=1, ellipsis_mask=0, new_axis_mask=0, shrink_axis_mask=2)
        cast = a(f"{scope}/Cast", "Cast", [ss], DstT=("type", G.DT_INT32))
        uq = a(f"{scope}/Unique", "Unique", [sfer + ":1"])
        emb = a(f"{scope}/embedding_lookup", "GatherV2", [read, uq, ax])
        emb_id = a(f"{scope}/embedding_lookup/Identity", "Identity", [emb])
        op = {"sum": "SparseSegmentSum", "mean": "SparseSegmentMean", "sqrtn": "SparseSegmentSqrtN"}[combiner]
        seed = a(f"{scope}/embedding_lookup_sparse", op, [emb_id, uq + ":1", cast])
        rs = a(f"{scope}/Reshape_1", "Reshape", [sfer + ":2", self.const(f"{scope}/m11", np.array([-1, 1], np.int32))])
        tile = a(f"{scope}/Tile", "Tile", [rs, self.const(f"{scope}/mult", np.array([1, dim], np.int32))])
        zl = a(f"{scope}/zeros_like", "ZerosLike", [seed])
        sel = a(f"{scope}/Select", "Select", [tile, zl, seed])
        out = a(f"{scope}/Reshape_2", "Reshape", [sel, self.const(f"{scope}/oshape", np.array([-1, dim], np.int32))])
        return out

    # ---- the heads ----
    def bucketized(self, key, boundaries, dim, combiner="mean"):
        a = self.add
        ph = a(key, "Placeholder", dtype=("type", G.DT_FLOAT), shape=("shape", [-1, 1]))
        bk = a(f"il/{key}/Bucketize", "Bucketize", [ph], boundaries=[float(b) for b in boundaries])
        flat = a(f"il/{key}/Reshape", "Reshape", [bk, self.const(f"il/{key}/m1", np.array([-1], np.int32))])
        mul = a(f"il/{key}/mul", "Mul", [self.const(f"il/{key}/nb", np.int32(len(boundaries) + 1)), a(f"il/{key}/i2", "Tile", [flat, flat])])
        ids = a(f"il/{key}/add", "AddV2", [flat, mul])
        ids64 = a(f"il/{key}/ids", "Cast", [ids], DstT=("type", G.DT_INT64))
        idx = a(f"il/{key}/indices", "Cast", [a(f"il/{key}/transpose", "Transpose", [flat, flat])], DstT=("type", G.DT_INT64))
        shp = a(f"il/{key}/shape", "Cast", [a(f"il/{key}/stack", "Pack", [flat, flat])], DstT=("type", G.DT_INT64))
        return self.lookup(f"il/{key}_embedding", idx, ids64, shp, len(boundaries) + 1, dim, combiner)

    def hashed_dense(self, key, buckets, dim, string=True, combiner="mean"):
        a = self.add
        dt = G.DT_STRING if string else G.DT_INT32
        ph = a(key, "Placeholder", dtype=("type", dt), shape=("shape", [-1, 1]))
        ign = self.const(f"il/{key}/ignore", np.array([b""], dtype=object) if string else np.int32(-1))
        ne = a(f"il/{key}/NotEqual", "NotEqual", [ph, ign])
        wh = a(f"il/{key}/Where", "Where", [ne])
        vals = a(f"il/{key}/GatherNd", "GatherNd", [ph, wh])
        shp = a(f"il/{key}/Shape", "Shape", [ph], out_type=("type", G.DT_INT64))
        if not string:
            vals = a(f"il/{key}/AsString", "AsString", [vals])
        ids = a(f"il/{key}/lookup", "StringToHashBucketFast", [vals], num_buckets=buckets)
        return self.lookup(f"il/{key}_embedding", wh, ids, shp, buckets, dim, combiner)

    def hashed_sparse(self, key, buckets, dim, rank=2, combiner="sum", table_rows=None, shared_read=None):
        a = self.add
        idx = a(f"{key}/indices", "Placeholder", dtype=("type", G.DT_INT64), shape=("shape", [-1, rank]))
        val = a(f"{key}/values", "Placeholder", dtype=("type", G.DT_STRING), shape=("shape", [-1]))
        shp = a(f"{key}/shape", "Placeholder", dtype=("type", G.DT_INT64), shape=("shape", [rank]))
        ids = a(f"il/{key}/lookup", "StringToHashBucketFast", [val], num_buckets=buckets)
        return self.lookup(f"il/{key}_embedding", idx, ids, shp, table_rows or buckets, dim, combiner, rank,
                           shared_read=shared_read)

    def identity_sparse(self, key, rows, dim, combiner="sqrtn"):
        a = self.add
        idx = a(f"{key}/indices", "Placeholder", dtype=("type", G.DT_INT64), shape=("shape", [-1, 2]))
        val = a(f"{key}/values", "Placeholder", dtype=("type", G.DT_INT64), shape=("shape", [-1]))
        shp = a(f"{key}/shape", "Placeholder", dtype=("type", G.DT_INT64), shape=("shape", [2]))
        return self.lookup(f"il/{key}_embedding", idx, val, shp, rows, dim, combiner)

    def finish(self, outs, with_mlp=True):
        ax = self.const("il/concat/axis", np.int32(1))
        cc = self.add("il/concat", "ConcatV2", list(outs) + [ax], N=len(outs))
        if with_mlp:  # a dense layer behind the stage: its kernel is a variable (frozen: a rank-2 Const) that is NOT a table
            w = (self.const("dense/kernel", np.zeros((64, 16), np.float32)) if self.frozen else
                 self.add("dense/kernel", "VariableV2", shape=("shape", [64, 16]), dtype=("type", G.DT_FLOAT)))
            r = self.add("dense/kernel/read", "Identity", [w])
            self.add("dense/MatMul", "MatMul", [cc, r])
        return G.encode_graphdef(self.nodes)
