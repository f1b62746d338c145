"""The two schedules of the column tiles - fused into the persistent kernel, or in the flat
one-warp-per-tile kernel in front of it (RECOM_B200_FLAT, read at plan creation) - against the
CPU oracle on the same request: both must give the oracle's ids bit for bit and its embeddings
within tolerance, whatever the cost rule (capi.cu prefer_flat) would pick."""
import os

import numpy as np
import pytest

from recom_b200 import consts as L
from recom_b200 import synth
from tests.helpers import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["0", "1"], ids=["fused", "flat"])
def flat_mode(request):
    old = os.environ.get("RECOM_B200_FLAT")
    os.environ["RECOM_B200_FLAT"] = request.param
    yield request.param
    if old is None:
        del os.environ["RECOM_B200_FLAT"]
    else:
        os.environ["RECOM_B200_FLAT"] = old


def _launches(model, batch, feats, tables):
    from recom_b200.schema import pack_request
    from tests.helpers import run_gpu

    req = pack_request(model, batch, feats)
    return run_gpu(model, req, tables, want_ids=False)[3]


def test_model_e_small_both_schedules(flat_mode):
    m = synth.model_E_small()
    for batch in (1, 37, 300):
        feats = synth.make_features(m, batch, seed=batch)
        assert_parity(m, batch, feats, synth.make_tables_np(m))
    # FAST slabs + column tiles: the flat schedule is two launches, the fused one is one
    feats = synth.make_features(m, 300, seed=1)
    assert _launches(m, 300, feats, synth.make_tables_np(m)) == (2 if flat_mode == "1" else 1)


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
def test_mixed_model_both_schedules(flat_mode, dtype):
    m = synth.mixed_model(120, seed=3, out_dtype=dtype, large_rows=1 << 14)
    for dist in ("u10", "zipf"):
        feats = synth.make_features(m, 200, seed=5, dist=dist, empty_frac=0.2)
        assert_parity(m, 200, feats, synth.make_tables_np(m))


def test_production_model_both_schedules(flat_mode):
    m = synth.production_model(300, share=4, seed=2)
    feats = synth.make_features(m, 64, seed=9)
    assert_parity(m, 64, feats, synth.make_tables_np(m), check_ids=True)


def test_cost_rule_picks_the_measured_winner():
    """No override: model E keeps one fused launch, a model without FAST slabs goes flat."""
    assert "RECOM_B200_FLAT" not in os.environ
    m = synth.model_E_small()
    feats = synth.make_features(m, 256, seed=0)
    assert _launches(m, 256, feats, synth.make_tables_np(m)) == 1
    s = synth.dlrm_model(0, 2, 2, 3, 0, name="nofast")
    feats = synth.make_features(s, 256, seed=0)
    assert _launches(s, 256, feats, synth.make_tables_np(s)) == 1  # flat only: still one launch


def test_long_rows_are_deferred_and_match_the_oracle():
    """Config 3: zipf row lengths (rows of hundreds to 2000 ids). Requests of density class >= 1 hand
    rows beyond 256 ids to long_rows_kernel (a CTA per row, partial sums added in warp order): one
    more launch, same ids, sums within the forward bound; and with the threshold out of reach
    (RECOM_B200_DEFER_IDS) the in-tile path gives the same result."""
    from recom_b200.schema import pack_request
    from tests.helpers import run_gpu

    assert "RECOM_B200_FLAT" not in os.environ
    m = synth.dlrm_model(8, 1, 1, 3, 2, large_rows=1 << 16, name="dyn_small")
    for batch, seed in ((64, 1), (700, 2)):
        feats = synth.make_features(m, batch, seed=seed, dist="zipf", empty_frac=0.2)
        req, got, ref = assert_parity(m, batch, feats, synth.make_tables_np(m))
        lengths = np.diff(next(f.row_splits for f in feats.values() if f.row_splits is not None))
        assert lengths.max() > 256, "the case must hold a long row"
    launches = run_gpu(m, req, synth.make_tables_np(m), want_ids=False)[3]
    assert launches in (2, 3)  # flat tiles [+ FAST runs] + long rows
    os.environ["RECOM_B200_DEFER_IDS"] = "1000000"
    try:
        got2 = run_gpu(m, req, synth.make_tables_np(m), want_ids=False)[0]
    finally:
        del os.environ["RECOM_B200_DEFER_IDS"]
    np.testing.assert_allclose(got2, got, rtol=1e-4, atol=1e-4)
