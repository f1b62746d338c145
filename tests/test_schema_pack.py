"""Host logic: request packing (the Addons>ConcatInputs mirror), slot sharing, the
SURVEY 8(d) algorithmic-byte count, synthetic config shapes. CPU only."""
import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import synth
from recom_b200.schema import Strings, algorithmic_bytes, pack_request


def test_model_e_and_f_shapes():
    e, f = synth.model_E(), synth.model_F()
    # examples/python/dlrm.py:202-203 and SURVEY.md section 8
    assert len(e.columns) == 1000 and e.out_width == 8120
    assert len(f.columns) == 1200 and f.out_width == 9672
    assert e.table_bytes() == 5392513280  # 5.39 GB
    names = [c.name for c in e.columns]
    assert names == sorted(names)  # input_layer concatenates in sorted-name order
    offs = [c.out_offset for c in e.columns]
    assert offs == list(np.cumsum([0] + [c.dim for c in e.columns[:-1]]))


def test_pack_layout_and_alignment():
    m = synth.model_E_small()
    req = pack_request(m, 37, synth.make_features(m, 37, seed=1))
    assert req.blob.dtype == np.uint8 and req.blob.size % 16 == 0
    assert (req.slot_offsets % 16 == 0).all()
    for c in m.columns:
        if c.lookup == L.LOOKUP_DENSE:
            want = 38 if c.value_dtype == L.STR else 37
            assert req.slot_elems[c.value_slot] == want
        else:
            assert req.slot_elems[c.seg_slot] == 38  # row_splits
    assert req.total_ids == int(req.ids_base[-1])
    assert req.sparse_ids == sum(int(req.ids_base[i + 1] - req.ids_base[i])
                                 for i, c in enumerate(m.columns) if c.lookup != L.LOOKUP_DENSE)


def test_pack_round_trips_values():
    m = synth.model_E_small(seg_format=L.SEG_COO_I64)
    feats = synth.make_features(m, 20, seed=2, empty_frac=0.3)
    req = pack_request(m, 20, feats)
    for c in m.columns:
        f = feats[c.key]
        o, n = int(req.slot_offsets[c.value_slot]), int(req.slot_elems[c.value_slot])
        if isinstance(f.values, Strings):
            np.testing.assert_array_equal(req.blob[o:o + 4 * n].view(np.int32), f.values.offsets)
            ob, nb = int(req.slot_offsets[c.bytes_slot]), int(req.slot_elems[c.bytes_slot])
            np.testing.assert_array_equal(req.blob[ob:ob + nb], f.values.data)
        else:
            np.testing.assert_array_equal(req.blob[o:o + f.values.nbytes].view(f.values.dtype), f.values)
        if c.lookup != L.LOOKUP_DENSE:
            o, n = int(req.slot_offsets[c.seg_slot]), int(req.slot_elems[c.seg_slot])
            np.testing.assert_array_equal(req.blob[o:o + 8 * n].view(np.int64), f.coo.reshape(-1))


def test_shared_inputs_share_slots():
    m = synth.production_model(200, share=4, seed=0)
    keys = {c.key for c in m.columns}
    assert len(keys) < len(m.columns)  # duplicated (input, transform) pairs exist
    per_key = {}
    for c in m.columns:
        per_key.setdefault(c.key, set()).add((c.value_slot, c.bytes_slot, c.seg_slot))
    assert all(len(v) == 1 for v in per_key.values())
    req = pack_request(m, 16, synth.make_features(m, 16, seed=0))
    assert req.slot_offsets.size == m.num_slots


def test_algorithmic_bytes_matches_survey_worked_number():
    # model E, fp32, fused hashing, B=2048 -> ~95.9 MB (nnz ~ 5.5 B)
    m = synth.model_E()
    req = pack_request(m, 2048, synth.make_features(m, 2048, seed=0))
    b = algorithmic_bytes(m, req)
    assert 93e6 < b < 99e6, b
    out_bytes = 2048 * 8120 * 4
    assert out_bytes == 66519040 and out_bytes / b > 0.65


def test_empty_batch_packs():
    m = synth.model_E_small()
    req = pack_request(m, 0, synth.make_features(m, 0, seed=0))
    assert req.batch == 0 and req.total_ids == 0


@pytest.mark.parametrize("dist", ["u10", "u50", "zipf", "one"])
def test_row_length_distributions(dist):
    rng = np.random.default_rng(0)
    n = synth.row_lengths(rng, 1000, dist, empty_frac=0.1)
    assert n.min() == 0 and n.shape == (1000,)
    rs, coo = synth.splits_and_coo(n)
    assert rs[-1] == n.sum() and (np.diff(coo[:, 0]) >= 0).all()
