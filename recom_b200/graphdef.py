"""column descriptors from a frozen TensorFlow GraphDef, without TensorFlow.

The reference finds its embedding columns ("FCs") on the live graph inside a grappler pass -
GraphInfo::ExtractFCNodes (graph_optimizers/graph_info.cc:209-365: a table is a VariableV2 whose
only consumers behind Identity are *Gather* / *SparseSegment* ops, Assign and SaveV2 aside),
LookupOptimizer::Match (lookup_optimizer.cc:62-140: the safe_embedding_lookup_sparse pattern
around a SparseSegment{Sum,Mean,SqrtN} seed - weights through GatherV2(Unique), segment ids =
[Cast <-] StridedSlice <- SparseFillEmptyRows, result through Select(Tile(Reshape(empty rows)),
ZerosLike, seed)), and CudaEmitter::Rewrite (cuda_emitter.cc:2496-2656: every column must end in
one ConcatV2, whose input order is the order of the slices). This module does the same walk on
the serialized graph and emits the plan's column descriptors (schema.Model) plus the names of the
placeholders / variables each blob slot and table is bound to, so a model exported by TF drives
recom_plan_create directly.

TensorFlow is not importable here and its .proto files are not in the image, so the protobuf
wire format is read (and, for tests and tools, written) directly; the field numbers are the
public ones of tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape}.proto.

What is recognised per column (the input_layer graph of TF 1.x/2.x compat.v1 feature columns,
feature_column_v2.py and embedding_ops.py:882-941):
  * bucketized_column(numeric_column):  Placeholder(float) -> Bucketize(boundaries) -> ... ->
    DENSE lookup, transform BUCKETIZE;
  * categorical_column_with_hash_bucket over a dense placeholder: NotEqual('' / -1) -> Where ->
    GatherNd -> [AsString ->] StringToHashBucketFast(num_buckets): DENSE lookup with transform
    HASH_STR / HASH_INT and the pre-lookup filter that drops the ignore value (the reference
    leaves that to Addons>GatherStrValueGenIndice / GatherIndiceValue, pre_lookup_optimizer.cc:310-347);
  * the same over a sparse placeholder triple (indices, values, dense_shape): SEGMENT lookup with
    COO segments (rank 2) or the SparseReshape inlined (rank > 2, RECOM_SEG_COO_RESHAPE_I64);
  * categorical_column_with_identity: transform NONE;
  * several lookup chains over ONE variable (shared_embedding_columns; BASELINE config 5): one column each, one
    table index for all of them.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import consts as L
from .schema import Column, Filter, Model

# ----------------------------------------------------------------------------------------
# protobuf wire format
# ----------------------------------------------------------------------------------------
# tensorflow DataType enum values used here
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64, DT_BOOL, DT_BFLOAT16 = 1, 2, 3, 7, 9, 10, 14
_NP_OF_DT = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64,
             DT_BOOL: np.bool_}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    r = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) of one message; length-delimited values are
    memoryview slices (zero copy: a frozen graph carries its tables as Const tensors of up to GBs)."""
    if not isinstance(buf, memoryview):
        buf = memoryview(buf)
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(b) -> List[int]:
    out, pos = [], 0
    while pos < len(b):
        v, pos = _varint(b, pos)
        out.append(_signed(v))
    return out


def _parse_shape(buf: bytes) -> Optional[List[int]]:
    dims: List[int] = []
    for fno, _, v in _fields(buf):
        if fno == 2:  # Dim
            size = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    size = _signed(v2)
            dims.append(size)
        elif fno == 3 and v:  # unknown_rank
            return None
    return dims


def _parse_tensor(buf: bytes) -> Optional[np.ndarray]:
    dtype, shape, content = 0, [], None
    vals: Dict[int, list] = {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dtype = v
        elif fno == 2:
            shape = _parse_shape(v) or []
        elif fno == 4:
            content = v
        elif fno == 5:  # float_val
            vals.setdefault(5, []).extend(struct.unpack(f"<{len(v) // 4}f", v) if wt == 2 else struct.unpack("<f", v))
        elif fno == 6:  # double_val
            vals.setdefault(6, []).extend(struct.unpack(f"<{len(v) // 8}d", v) if wt == 2 else struct.unpack("<d", v))
        elif fno in (7, 10):  # int_val, int64_val
            vals.setdefault(fno, []).extend(_packed_varints(v) if wt == 2 else [_signed(v)])
        elif fno == 8:  # string_val
            vals.setdefault(8, []).append(bytes(v))
    if dtype == DT_STRING:
        return np.array(vals.get(8, []), dtype=object)
    np_t = _NP_OF_DT.get(dtype)
    if np_t is None:
        return None
    n = int(np.prod(shape)) if shape else 1
    if content is not None:
        return np.frombuffer(content, dtype=np_t).reshape(shape)
    flat = next((vals[k] for k in (5, 6, 7, 10) if k in vals), [])
    if not flat:
        return np.zeros(shape, np_t)
    if len(flat) < n:  # a single value broadcast (proto convention)
        flat = list(flat) + [flat[-1]] * (n - len(flat))
    return np.array(flat[:n], dtype=np_t).reshape(shape)


def _parse_attr(buf: bytes):
    """AttrValue -> python value (list -> python list)."""
    for fno, wt, v in _fields(buf):
        if fno == 2:
            return bytes(v)
        if fno == 3:
            return _signed(v)
        if fno == 4:
            return struct.unpack("<f", v)[0]
        if fno == 5:
            return bool(v)
        if fno == 6:
            return ("type", v)
        if fno == 7:
            return ("shape", _parse_shape(v))
        if fno == 8:
            return _parse_tensor(v)
        if fno == 1:  # ListValue
            out: list = []
            for f2, w2, v2 in _fields(v):
                if f2 == 2:
                    out.append(bytes(v2))
                elif f2 == 3:
                    out.extend(_packed_varints(v2) if w2 == 2 else [_signed(v2)])
                elif f2 == 4:
                    out.extend(struct.unpack(f"<{len(v2) // 4}f", v2) if w2 == 2 else struct.unpack("<f", v2))
                elif f2 == 6:
                    out.extend(_packed_varints(v2) if w2 == 2 else [v2])
                elif f2 == 7:
                    out.append(_parse_shape(v2))
            return out
    return None


@dataclass
class Node:
    name: str
    op: str
    inputs: List[str] = field(default_factory=list)
    attr: Dict[str, object] = field(default_factory=dict)

    def dtype(self, key: str = "dtype") -> int:
        v = self.attr.get(key)
        return v[1] if isinstance(v, tuple) and v[0] == "type" else 0


def tensor_node(t: str) -> Tuple[str, int]:
    """'name:1' / 'name' / '^name' -> (name, output index; -1: control edge)."""
    if t.startswith("^"):
        return t[1:], -1
    if ":" in t:
        n, i = t.rsplit(":", 1)
        return n, int(i)
    return t, 0


class Graph:
    def __init__(self, nodes: List[Node]):
        self.nodes = nodes
        self.by_name = {n.name: n for n in nodes}
        self.consumers: Dict[str, List[Tuple[Node, int]]] = {n.name: [] for n in nodes}
        for n in nodes:
            for i, t in enumerate(n.inputs):
                src, port = tensor_node(t)
                if port >= 0 and src in self.consumers:
                    self.consumers[src].append((n, i))

    def input_node(self, n: Node, i: int) -> Optional[Node]:
        if i >= len(n.inputs):
            return None
        return self.by_name.get(tensor_node(n.inputs[i])[0])

    def input_port(self, n: Node, i: int) -> int:
        return tensor_node(n.inputs[i])[1]

    def const_value(self, n: Optional[Node]) -> Optional[np.ndarray]:
        while n is not None and n.op == "Identity":
            n = self.input_node(n, 0)
        if n is None or n.op != "Const":
            return None
        v = n.attr.get("value")
        return v if isinstance(v, np.ndarray) else None


def parse_graphdef(data) -> Graph:
    nodes: List[Node] = []
    for fno, _, v in _fields(memoryview(data)):
        if fno != 1:
            continue  # versions, library
        node = Node("", "")
        for f2, _, v2 in _fields(v):
            if f2 == 1:
                node.name = bytes(v2).decode()
            elif f2 == 2:
                node.op = bytes(v2).decode()
            elif f2 == 3:
                node.inputs.append(bytes(v2).decode())
            elif f2 == 5:  # map<string, AttrValue> entry
                key, val = "", None
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        key = bytes(v3).decode()
                    elif f3 == 2:
                        val = _parse_attr(v3)
                node.attr[key] = val
        nodes.append(node)
    return Graph(nodes)


# ---- writer (tests, tools): the same subset ----
def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_field(fno: int, wt: int, payload) -> bytes:
    key = _enc_varint(fno << 3 | wt)
    if wt == 0:
        return key + _enc_varint(payload)
    if wt == 2:
        return key + _enc_varint(len(payload)) + payload
    return key + payload


def _enc_shape(dims) -> bytes:
    return b"".join(_enc_field(2, 2, _enc_field(1, 0, int(d))) for d in dims)


_DT_OF_NP = {np.dtype(np.float32): DT_FLOAT, np.dtype(np.float64): DT_DOUBLE, np.dtype(np.int32): DT_INT32,
             np.dtype(np.int64): DT_INT64, np.dtype(np.bool_): DT_BOOL}


def _enc_attr(v) -> bytes:
    if isinstance(v, bytes):
        return _enc_field(2, 2, v)
    if isinstance(v, bool):
        return _enc_field(5, 0, int(v))
    if isinstance(v, int):
        return _enc_field(3, 0, v)
    if isinstance(v, float):
        return _enc_field(4, 5, struct.pack("<f", v))
    if isinstance(v, tuple) and v[0] == "type":
        return _enc_field(6, 0, v[1])
    if isinstance(v, tuple) and v[0] == "shape":
        return _enc_field(7, 2, _enc_shape(v[1]))
    if isinstance(v, np.ndarray):
        if v.dtype == object:
            t = _enc_field(1, 0, DT_STRING) + _enc_field(2, 2, _enc_shape(v.shape))
            t += b"".join(_enc_field(8, 2, bytes(s)) for s in v.reshape(-1))
        else:
            t = _enc_field(1, 0, _DT_OF_NP[v.dtype]) + _enc_field(2, 2, _enc_shape(v.shape))
            t += _enc_field(4, 2, np.ascontiguousarray(v).tobytes())
        return _enc_field(8, 2, t)
    if isinstance(v, list):
        body = b""
        if v and isinstance(v[0], float):
            body = _enc_field(4, 2, struct.pack(f"<{len(v)}f", *v))
        elif v and isinstance(v[0], int):
            body = _enc_field(3, 2, b"".join(_enc_varint(i) for i in v))
        elif v and isinstance(v[0], bytes):
            body = b"".join(_enc_field(2, 2, s) for s in v)
        return _enc_field(1, 2, body)
    raise TypeError(type(v))


def encode_graphdef(nodes: List[Node]) -> bytes:
    out = bytearray()
    for n in nodes:
        body = _enc_field(1, 2, n.name.encode()) + _enc_field(2, 2, n.op.encode())
        for i in n.inputs:
            body += _enc_field(3, 2, i.encode())
        for k, v in n.attr.items():
            body += _enc_field(5, 2, _enc_field(1, 2, k.encode()) + _enc_field(2, 2, _enc_attr(v)))
        out += _enc_field(1, 2, body)
    return bytes(out)


# ----------------------------------------------------------------------------------------
# extraction
# ----------------------------------------------------------------------------------------
@dataclass
class Binding:
    """Where a column's inputs come from in the graph (placeholder names) and which variable is
    its table."""
    column: str
    table_variable: str
    values: str
    indices: Optional[str] = None      # sparse inputs: the SparseTensor's indices placeholder
    dense_shape: Optional[str] = None  # ... and its dense_shape placeholder
    concat: str = ""
    concat_index: int = -1


SEED_OPS = {"SparseSegmentSum": L.COMBINE_SUM, "SparseSegmentMean": L.COMBINE_MEAN,
            "SparseSegmentSqrtN": L.COMBINE_SQRTN}
_TRANSPARENT = {"Identity", "Reshape", "ExpandDims", "Squeeze", "GatherV2", "GatherNd", "AddV2", "Add",
                "StopGradient", "SparseToDense", "Cast"}


def table_shape(v: Node) -> Optional[Tuple[int, int]]:
    """(rows, dim) of a table node: the VariableV2's static shape attr, or the Const tensor's shape."""
    if v.op == "Const":
        t = v.attr.get("value")
        return (int(t.shape[0]), int(t.shape[1])) if isinstance(t, np.ndarray) and t.ndim == 2 else None
    shape = v.attr.get("shape")
    dims = shape[1] if isinstance(shape, tuple) else None
    return (int(dims[0]), int(dims[1])) if dims and len(dims) == 2 and min(dims) > 0 else None


def find_tables(g: Graph) -> List[Node]:
    """graph_info.cc:209-258 (VariableV2 at grappler time); in a FROZEN graph the variables have become
    Const nodes (convert_variables_to_constants), which LookupOptimizer::Match accepts as weights just the
    same (lookup_optimizer.cc:64-65, 'VariableV2,Const,VarHandleOp'): a rank-2 float Const qualifies."""
    out = []
    for v in g.nodes:
        if v.op == "Const":
            t = v.attr.get("value")
            if not (isinstance(t, np.ndarray) and t.ndim == 2 and t.dtype in (np.float32, np.float64)):
                continue
        elif v.op != "VariableV2":
            continue
        stack = [c for c, _ in g.consumers[v.name]]
        cnt, ok = 0, True
        while stack and ok:
            n = stack.pop()
            if n.op == "Identity":
                stack.extend(c for c, _ in g.consumers[n.name])
            elif n.op in ("Assign", "SaveV2"):
                pass
            elif "Gather" not in n.op and "SparseSegment" not in n.op:
                ok = False
            else:
                cnt += 1
        if ok and cnt > 0:
            out.append(v)
    return out


def _match_seeds(g: Graph, table: Node) -> List[Tuple[Node, Node, Node]]:
    """lookup_optimizer.cc:62-140 -> every (seed, SparseFillEmptyRows, Select) that reads `table`: one per
    column; shared_embedding_columns (and BASELINE config 5) put several columns on one table."""
    found: List[Tuple[Node, Node, Node]] = []
    stack = [c for c, _ in g.consumers[table.name]]
    reads = []
    while stack:
        n = stack.pop()
        if n.op == "Identity":
            stack.extend(c for c, _ in g.consumers[n.name])
        elif n.op not in ("Assign", "SaveV2"):
            reads.append(n)
    for r in reads:
        seeds = [r] if r.op in SEED_OPS else []
        hop = [r]
        while hop and not seeds:  # GatherV2 [-> Identity ...] -> seed (input 0)
            cur = hop.pop()
            for c, i in g.consumers[cur.name]:
                if c.op in SEED_OPS and i == 0:
                    seeds.append(c)
                elif c.op == "Identity":
                    hop.append(c)
        for seed in seeds:
            seg = g.input_node(seed, 2)
            if seg is not None and seg.op == "Cast":
                seg = g.input_node(seg, 0)
            if seg is None or seg.op != "StridedSlice" or seg.attr.get("shrink_axis_mask") != 2:
                continue
            sfer = g.input_node(seg, 0)
            if sfer is None or sfer.op != "SparseFillEmptyRows":
                continue
            ids_in = g.input_node(seed, 1)
            if ids_in is not None and ids_in.op == "Unique":
                if g.input_node(ids_in, 0) is not sfer:
                    continue
            elif ids_in is not sfer:
                continue
            sel = next((c for c, i in g.consumers[seed.name] if c.op in ("Select", "SelectV2") and i == 2), None)
            if sel is None:
                continue
            zl, tile = g.input_node(sel, 1), g.input_node(sel, 0)
            if zl is None or zl.op != "ZerosLike" or g.input_node(zl, 0) is not seed:
                continue
            if tile is None or tile.op != "Tile":
                continue
            rs = g.input_node(tile, 0)
            if rs is None or rs.op != "Reshape" or g.input_node(rs, 0) is not sfer:
                continue
            if all(seed is not f[0] for f in found):
                found.append((seed, sfer, sel))
    return found


def _walk_values(g: Graph, n: Optional[Node], port: int):
    """Backwards from SparseFillEmptyRows.values to the placeholder: the transform ops on the way
    and whether the dense input was sparsified by Where(NotEqual(x, ignore))."""
    info = {"bucketize": None, "hash": None, "as_string": False, "where": None, "placeholder": None,
            "sparse_values": False}
    seen = 0
    while n is not None and seen < 64:
        seen += 1
        if n.op in ("Placeholder", "PlaceholderWithDefault"):
            info["placeholder"] = n
            return info
        if n.op == "StringToHashBucketFast":
            info["hash"] = int(n.attr.get("num_buckets", 0))
        elif n.op == "AsString":
            info["as_string"] = True
        elif n.op == "Bucketize":
            info["bucketize"] = np.array(n.attr.get("boundaries", []), np.float32)
        elif n.op == "GatherNd":
            w = g.input_node(n, 1)
            if w is not None and w.op == "Where":
                info["where"] = g.input_node(w, 0)
        elif n.op == "GatherV2":
            pass  # sparse_retain(ids >= 0) / the values of a pruned SparseTensor: params = input 0
        elif n.op not in _TRANSPARENT and n.op not in ("Mul", "Sub"):
            return info
        nxt = g.input_node(n, 0)
        port = g.input_port(n, 0) if n.inputs else 0
        n = nxt
    return info


def _walk_indices(g: Graph, n: Optional[Node]) -> Optional[Node]:
    """Backwards from SparseFillEmptyRows.indices to an indices placeholder (sparse inputs), through
    SparseReshape / the GatherV2 of sparse_retain; None when the indices are made from a dense input."""
    seen = 0
    while n is not None and seen < 64:
        seen += 1
        if n.op in ("Placeholder", "PlaceholderWithDefault"):
            return n
        if n.op in ("SparseReshape", "GatherV2", "Identity", "Cast"):
            n = g.input_node(n, 0)
            continue
        return None
    return None


def _walk_shape(g: Graph, n: Optional[Node]) -> Optional[Node]:
    """Backwards from SparseFillEmptyRows.dense_shape (SparseReshape's output 1) to the SparseTensor's
    dense_shape placeholder (SparseReshape's input 1)."""
    seen = 0
    while n is not None and seen < 64:
        seen += 1
        if n.op in ("Placeholder", "PlaceholderWithDefault"):
            return n
        if n.op == "SparseReshape":
            n = g.input_node(n, 1)
        elif n.op in ("Identity", "Cast"):
            n = g.input_node(n, 0)
        else:
            return None
    return None


def _find_concat(g: Graph, n: Node) -> Tuple[Optional[Node], int]:
    """The ConcatV2 a column's result reaches (through Reshape / Identity), and its input index."""
    frontier = [n]
    seen = set()
    while frontier:
        cur = frontier.pop()
        if cur.name in seen:
            continue
        seen.add(cur.name)
        for c, i in g.consumers[cur.name]:
            if c.op == "ConcatV2":
                return c, i
            if c.op in ("Reshape", "Identity") and i == 0:
                frontier.append(c)
    return None, -1


@dataclass
class Extracted:
    model: Model
    bindings: List[Binding]
    skipped: List[Tuple[str, str]]  # (variable, why) - tables whose column did not match
    # the tables' values where the graph carries them (frozen graphs: Const nodes; zero-copy views of
    # the serialized graph), None for VariableV2 tables (their values live in the checkpoint)
    table_values: List[Optional[np.ndarray]] = field(default_factory=list)


def extract_model(g: Graph, out_dtype: int = L.F32) -> Extracted:
    found = []
    skipped: List[Tuple[str, str]] = []
    matches = []
    for t in find_tables(g):
        ms = _match_seeds(g, t)
        if not ms:
            skipped.append((t.name, "no safe_embedding_lookup_sparse pattern around a SparseSegment op"))
        matches.extend((t, m) for m in ms)
    for t, (seed, sfer, sel) in matches:
        dims = table_shape(t)
        if dims is None:
            skipped.append((t.name, "table shape is not static rank 2"))
            continue
        rows, dim = dims
        info = _walk_values(g, g.input_node(sfer, 1), g.input_port(sfer, 1))
        ph = info["placeholder"]
        if ph is None:
            skipped.append((t.name, "ids do not come from a placeholder"))
            continue
        idx_ph = _walk_indices(g, g.input_node(sfer, 0))
        concat, ci = _find_concat(g, sel)
        if concat is None:
            skipped.append((t.name, "result does not reach a ConcatV2 (cuda_emitter.cc:2545-2580)"))
            continue
        dt = ph.dtype()
        flt = None
        if info["bucketize"] is not None:
            xform, vd = L.XFORM_BUCKETIZE, {DT_FLOAT: L.F32, DT_DOUBLE: L.F64, DT_INT32: L.I32, DT_INT64: L.I64}[dt]
        elif info["hash"]:
            if dt == DT_STRING:
                xform, vd = L.XFORM_HASH_STR, L.STR
                if info["where"] is not None:  # _to_sparse_input_and_drop_ignore_values: x != ''
                    flt = Filter(kind=L.FILTER_STR_NE, not_equal=b"")
            else:
                xform, vd = L.XFORM_HASH_INT, (L.I64 if dt == DT_INT64 else L.I32)
                if info["where"] is not None:  # ignore value -1: keep [0, max]
                    flt = Filter(kind=L.FILTER_KEEP, intervals=[(-(1 << 63), -2), (0, (1 << 63) - 1)])
        else:
            xform, vd = L.XFORM_NONE, (L.I64 if dt == DT_INT64 else L.I32)
            if info["where"] is not None:
                flt = Filter(kind=L.FILTER_KEEP, intervals=[(0, (1 << 63) - 1)])
        col = Column(name=ph.name, lookup=L.LOOKUP_DENSE, transform=xform, value_dtype=vd, table=-1,
                     table_rows=rows, dim=dim, combiner=SEED_OPS[seed.op], filter=flt)
        if info["bucketize"] is not None:
            col.boundaries = info["bucketize"]
        if info["hash"]:
            col.num_buckets = info["hash"]
        b = Binding(column=ph.name, table_variable=t.name, values=ph.name, concat=concat.name, concat_index=ci)
        if idx_ph is not None:  # a sparse input: (indices, values, dense_shape) placeholders
            shp = idx_ph.attr.get("shape")
            rank = shp[1][1] if isinstance(shp, tuple) and shp[1] and len(shp[1]) == 2 and shp[1][1] > 0 else 2
            col.lookup = L.LOOKUP_SEGMENT
            col.filter = None
            if rank == 2:
                col.seg_format, col.seg_stride = L.SEG_COO_I64, 2
            else:  # safe_embedding_lookup_sparse reshapes to [prod(shape[:-1]), shape[-1]]
                col.seg_format, col.seg_stride = L.SEG_COO_RESHAPE_I64, L.reshape_spec(rank)
            b.indices = idx_ph.name
            shape_ph = _walk_shape(g, g.input_node(sfer, 2))
            b.dense_shape = shape_ph.name if shape_ph is not None else None
        found.append((concat.name, ci, col, b))
    # one ConcatV2 is the stage's output: the one most columns reach; its input order is the slice order
    by_concat: Dict[str, int] = {}
    for cn, _, _, _ in found:
        by_concat[cn] = by_concat.get(cn, 0) + 1
    cols: List[Column] = []
    binds: List[Binding] = []
    tables: List[Tuple[int, int]] = []
    values: List[Optional[np.ndarray]] = []
    table_index: Dict[str, int] = {}
    if by_concat:
        main = max(by_concat, key=lambda k: by_concat[k])
        for cn, ci, col, b in sorted(found, key=lambda f: f[1]):
            if cn != main:
                skipped.append((b.table_variable, f"reaches another ConcatV2 ({cn})"))
                continue
            if b.table_variable in table_index:  # columns sharing one table (descriptor-level aliasing)
                col.table = table_index[b.table_variable]
            else:
                col.table = table_index[b.table_variable] = len(tables)
                tables.append((col.table_rows, col.dim))
                tn = g.by_name[b.table_variable]
                v = tn.attr.get("value") if tn.op == "Const" else None
                values.append(v if isinstance(v, np.ndarray) else None)
            cols.append(col)
            binds.append(b)
    model = Model(cols, tables, out_dtype=out_dtype, name="graphdef").finalize() if cols else Model([], [])
    return Extracted(model, binds, skipped, values)


def extract_from_bytes(data: bytes, out_dtype: int = L.F32) -> Extracted:
    return extract_model(parse_graphdef(data), out_dtype)


if __name__ == "__main__":  # python -m recom_b200.graphdef frozen_graph.pb
    import sys

    with open(sys.argv[1], "rb") as f:
        ex = extract_from_bytes(f.read())
    for c, b in zip(ex.model.columns, ex.bindings):
        print(f"{c.out_offset:6d} +{c.dim:<3d} {c.name}: lookup {c.lookup} transform {c.transform} dtype {c.value_dtype} "
              f"table {b.table_variable} [{c.table_rows}, {c.dim}] combiner {c.combiner} seg {c.seg_format}"
              + (f" filter {c.filter.kind}" if c.filter else ""))
    for name, why in ex.skipped:
        print(f"skipped {name}: {why}")
    print(f"{len(ex.model.columns)} columns, out_width {ex.model.out_width}, {ex.model.num_slots} blob slots")
