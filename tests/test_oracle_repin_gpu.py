"""Re-pin the CPU oracle ON THE GPU HOST ("re-run this check on the GPU box's host before trusting it
there"): the oracle library is rebuilt by the box's own gcc / glibc, so the libm identity, the reference's known
answers, the three experiment_0 digests, the predicate / FK fixtures and the Philox known answers are checked again
next to the GPU parity tests.  Light subset of tests/test_oracle_golden.py (same functions, same fixtures)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_oracle_golden as tog                      # noqa: E402
from test_oracle_golden import goldens                # noqa: E402,F401  (fixture)

pytestmark = pytest.mark.gpu


def test_oracle_repinned_on_this_host(oracle, goldens, golden_dir):
    tog.test_libm_matches_numpy()
    tog.test_known_answers(oracle, goldens)
    tog.test_manual_grid_corners(oracle)
    for idx in (0, 1, 2):
        tog.test_experiment0_digest(oracle, goldens, idx)
    tog.test_predicate_cases(oracle, golden_dir)
    tog.test_fk_cases(oracle, golden_dir)
    tog.test_philox_known_answers(oracle)
