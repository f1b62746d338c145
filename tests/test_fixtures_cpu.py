"""The oracle against its own frozen outputs (tests/golden/stage_fixtures.npz): catches an
accidental change of the oracle or of the synthetic generators. CPU only."""
import os

import numpy as np
import pytest

from tests.golden.make_fixtures import CASES, build_case, digest
from tests.helpers import run_oracle

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage_fixtures.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_fixture(name):
    model, req, tables = build_case(name)
    assert digest(req, tables) == FX[f"{name}/inputs_sha256"].tobytes().decode(), \
        "seeded inputs changed (RNG drift?): regenerate with tests/golden/make_fixtures.py"
    ref, ids = run_oracle(model, req, tables)
    np.testing.assert_array_equal(ids.astype(np.int32), FX[f"{name}/ids"])
    np.testing.assert_array_equal(ref, FX[f"{name}/out"])


def test_oracle_threads_do_not_change_results():
    model, req, tables = build_case("mixed40_f32_b24")
    a, ia = run_oracle(model, req, tables, threads=1)
    b, ib = run_oracle(model, req, tables, threads=4)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ia, ib)
