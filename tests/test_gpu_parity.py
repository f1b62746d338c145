"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import synth
from tests.helpers import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [1, 7, 256, 300])
def test_model_e_small(batch):
    m = synth.model_E_small()
    feats = synth.make_features(m, batch, seed=batch)
    assert_parity(m, batch, feats, synth.make_tables_np(m))


@pytest.mark.parametrize("fmt", [L.SEG_ROW_SPLITS_I32, L.SEG_COO_I64])
@pytest.mark.parametrize("fused", [True, False])
def test_model_e_small_formats(fmt, fused):
    m = synth.model_E_small(seg_format=fmt, fused_hash=fused)
    feats = synth.make_features(m, 129, seed=3, empty_frac=0.2)
    assert_parity(m, 129, feats, synth.make_tables_np(m))


def test_microbenchmark_100():
    m = synth.microbenchmark_model(100)
    batch = 256
    feats = synth.make_features(m, batch, seed=0)
    # microbenchmark.py:66 feeds randint(-1, 10000) cast to float
    rng = np.random.default_rng(0)
    for c in m.columns:
        feats[c.key].values = rng.integers(-1, 10000, size=batch).astype(np.float32)
    assert_parity(m, batch, feats, synth.make_tables_np(m))


@pytest.mark.parametrize("out_dtype", [L.F32, L.BF16])
def test_mixed_model(out_dtype):
    m = synth.mixed_model(120, seed=1, out_dtype=out_dtype, large_rows=1 << 15)
    feats = synth.make_features(m, 200, seed=5, dist="u50", empty_frac=0.1)
    assert_parity(m, 200, feats, synth.make_tables_np(m), check_ids=False)


def test_mixed_model_checked_ids():
    m = synth.mixed_model(60, seed=2, large_rows=1 << 14)
    feats = synth.make_features(m, 64, seed=6, dist="zipf")
    assert_parity(m, 64, feats, synth.make_tables_np(m), check_ids=True)
