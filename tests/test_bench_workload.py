"""bench.py's workload (BASELINE config 2: z = max(a*b + c, d) - abs(a)*0.5f over four float32
RLE operands, + compaction) at 1/5000 scale: the C-ABI path, the oracle and the UNMODIFIED
reference extension (what `bench.py --impl reference` times) agree run for run and bit for bit,
so the two bench arms measure the same computation."""
import os
import sys

import numpy as np
import pytest

import helpers
from helpers import A, O, assert_runs_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_bench_dag_matches_oracle_and_reference(lib):
    n, r = 200_000, 2_000
    ops = [bench.gen_operand(2000 + j, n, r) for j in range(bench.N_OPERANDS)]
    for e, v in ops:                              # generator contract
        assert e.dtype == np.int32 and v.dtype == np.float32 and len(e) == r and e[-1] == n
        assert np.all(np.diff(e) > 0) and np.all(v[1:] != v[:-1])
    stores = [(lib.store_from_runs(e, v), []) for e, v in ops]
    got = lib.eval(stores, bench.dag_program(A), A.F32, A.EQ_BITWISE).runs()

    a, b, c, d = (O.to_dense(e, v) for e, v in ops)
    f32 = np.float32
    want_dense = (np.maximum((a * b).astype(f32) + c, d) - np.abs(a) * f32(0.5)).astype(f32)
    assert_runs_equal(got, O.to_store(want_dense), "numpy restatement")

    if not helpers.have_ref():
        pytest.skip("oracle/_ref not built")
    R = helpers.ref_ext()
    ra, rb, rc, rd = (R.from_numpy(x) for x in (a, b, c, d))
    z = ((ra * rb + rc).max(rd) - ra.abs() * 0.5).eval()      # the expression bench.py --impl reference runs
    re_, rv = z.runs()
    assert_runs_equal(got, (np.asarray(re_), np.asarray(rv)), "reference extension")
