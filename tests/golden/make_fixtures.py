#!/usr/bin/env python
"""Regenerates tests/golden/stage_fixtures.npz: outputs of the CPU oracle (pinned to TF's
vectors by tests/test_oracle_golden.py) on seeded requests of small models. The TF reference
itself cannot be imported here (no TensorFlow wheel), so these are oracle outputs, frozen so
that (a) an accidental change of the oracle is caught on CPU and (b) the CUDA path is compared
with committed values, not only with a same-process oracle run.

    python tests/golden/make_fixtures.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from recom_b200 import _lib as L  # noqa: E402
from recom_b200 import synth  # noqa: E402
from recom_b200.schema import pack_request  # noqa: E402
from tests.helpers import run_oracle  # noqa: E402

CASES = {
    # name: (model factory, batch, feature kwargs)
    "e_small_b32": (lambda: synth.model_E_small(), 32, dict(seed=11)),
    "e_small_coo_b17": (lambda: synth.model_E_small(seg_format=L.SEG_COO_I64), 17, dict(seed=12, empty_frac=0.25)),
    "mixed40_f32_b24": (lambda: synth.mixed_model(40, seed=3, large_rows=1 << 12), 24, dict(seed=13, dist="u50", empty_frac=0.1)),
    "mixed40_bf16_b24": (lambda: synth.mixed_model(40, seed=3, out_dtype=L.BF16, large_rows=1 << 12), 24, dict(seed=13, dist="u50", empty_frac=0.1)),
}


def build_case(name):
    factory, batch, kw = CASES[name]
    model = factory()
    feats = synth.make_features(model, batch, **kw)
    tables = synth.make_tables_np(model, seed=7)
    req = pack_request(model, batch, feats)
    return model, req, tables


def digest(req, tables):
    h = hashlib.sha256()
    # the slots only: the argument trailer behind them repeats slot_offsets / slot_elems
    data = req.blob if req.args_offset < 0 else req.blob[:req.args_offset]
    h.update(data.tobytes())
    h.update(req.slot_offsets.tobytes())
    for t in tables:
        h.update(np.ascontiguousarray(t).tobytes())
    return h.hexdigest()


def main():
    out = {}
    for name in CASES:
        model, req, tables = build_case(name)
        ref, ids = run_oracle(model, req, tables)
        out[f"{name}/out"] = ref
        out[f"{name}/ids"] = ids.astype(np.int32)
        out[f"{name}/inputs_sha256"] = np.frombuffer(digest(req, tables).encode(), np.uint8)
        print(name, ref.shape, ids.shape)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "stage_fixtures.npz"), **out)


if __name__ == "__main__":
    main()
