"""Known-answer tests of the reference restated against the C ABI (the ones round 1 left out):

  tests/eval_test.cpp:66-88    two stores stacked by fixed / scaled step functions
  tests/eval_test.cpp:90-156   4-way pool partition: ends and values of every part
  tests/step_test.cpp:22-64    strided / offset-strided / scaled / fixed / sliced step tables
  tests/step_test.cpp:66-113   3-D 4^3 sub-box gather: cnt == 8 and sum == 28

plus regression tests for gather views that collapse long groups of runs onto one output
coordinate (the tile partition must cut such groups), for queued sources (psk_eval_async
results consumed without a host round trip) and for NaN payloads under BITWISE equality."""
import numpy as np
import pytest

from helpers import A, O, assert_runs_equal, oracle_eval, rand_runs, run_eval


def test_eval_stack_kat(lib):
    # eval_test.cpp:66-88.  s_step_fn = stack(scaled<1>(2), fixed<5>(1)): elements 0, 1 of s keep
    # their place, element 2 is stretched to the end -> a SCATTER of s onto positions {0, 1, 2}
    # of 7; t_step_fn = stack(fixed<4>(1), scaled<1>(3)) -> SCATTER of t onto {3, 4, 5, 6}.
    m = O.to_store(np.array([0, 0, 0, 1, 1, 1, 1], np.int32))
    s = O.to_store(np.frombuffer(b"abc", dtype=np.int8))
    t = O.to_store(np.frombuffer(b"defg", dtype=np.int8))
    sv = [("scatter", (7,), [(0, 3, 1)])]
    tv = [("scatter", (7,), [(3, 7, 1)])]
    e, v = run_eval(lib, [m + ([],), s + (sv,), t + (tv,)],
                    ("select", ("cast", np.bool_, ("src", 0)), ("src", 2), ("src", 1)))
    assert O.to_string(e, v) == "1=>a, 2=>b, 3=>c, 4=>d, 5=>e, 6=>f, 7=>g"


def test_pool_partition_kat(lib):
    # eval_test.cpp:90-156: partition_pool(pool, 4) of two 8-run stores: part i covers output
    # coordinates [2i, 2i + 2) of every source (psk_store_slice = SimpleSource::split).
    s = lib.store_from_runs(np.arange(1, 9, dtype=np.int32), np.frombuffer(b"abcdefgh", dtype=np.int8))
    t = lib.store_from_runs(np.arange(1, 9, dtype=np.int32), np.arange(1, 9, dtype=np.int32))
    for part in range(4):
        lo, hi = (part * 8) // 4, ((part + 1) * 8) // 4          # eval.hpp:452-456
        assert hi - lo == 2
        for store, want in ((s, b"abcdefgh"), (t, list(range(1, 9)))):
            e, v = lib.store_slice(store, lo, hi).runs()
            assert (e + lo).tolist() == [lo + 1, lo + 2]           # the reference reports global ends
            assert list(v) == list(want[lo:hi])


def test_store_slice_matches_a_gather_view(lib):
    """psk_store_slice (two searches + a copy on the device) gives the store an evaluation through
    a one-dimensional gather view [start, stop) gives: random stores, cuts on and off run ends."""
    rng = np.random.default_rng(91)
    for n, r in ((1, 1), (40, 7), (5000, 900), (100_000, 3000)):
        ends = np.sort(rng.choice(np.arange(1, n), size=min(r, n) - 1, replace=False)).astype(np.int32) if n > 1 else np.zeros(0, np.int32)
        ends = np.append(ends, np.int32(n))
        vals = np.arange(ends.size, dtype=np.int32) % 97
        vals[1:][vals[1:] == vals[:-1]] += 1
        s = lib.store_from_runs(ends, vals)
        cuts = [(0, n), (0, 1), (n - 1, n)] + [tuple(sorted(rng.choice(n + 1, size=2, replace=False))) for _ in range(12)]
        cuts += [(int(ends[len(ends) // 2]) - 1, n)] if ends.size > 1 and ends[len(ends) // 2] > 1 else []
        for a, b in cuts:
            a, b = int(a), int(b)
            if a >= b:
                continue
            ge, gv = lib.store_slice(s, a, b).runs()
            view = [("gather", (n,), [(a, b, 1)])]
            we, wv = oracle_eval([(ends, vals, view)], ("src", 0))
            assert_runs_equal((ge, gv), (we, wv), f"slice [{a}, {b}) of {n}")


def _step_table(lib, views, span_in, args):
    """f(E) for every E in args: the end of run E of an all-distinct store of span_in elements,
    pushed through the view chain, is where that run ends in the output (0 when it vanishes)."""
    ends = np.arange(1, span_in + 1, dtype=np.int32)
    vals = np.arange(span_in, dtype=np.int32)
    got, _ = run_eval(lib, [(ends, vals, views)], ("src", 0))
    out_v = run_eval(lib, [(ends, vals, views)], ("src", 0))[1]
    # output run j ends at got[j] and carries the value (= index) of the source element there
    f = {0: 0}
    last = 0
    for E in range(1, span_in + 1):
        hit = np.nonzero(out_v == E - 1)[0]
        if len(hit):
            last = int(got[hit[0]])
        f[E] = last
    return [f[a] for a in args]


def test_step_tables_kat(lib):
    # step_test.cpp:22-64
    # strided<4>(10): every 4th of 10 elements
    assert _step_table(lib, [("gather", (10,), [(0, 10, 4)])], 10, range(10)) == [0, 1, 1, 1, 1, 2, 2, 2, 2, 3]
    # stack(fixed<0>(2), strided<2>(8)): skip 2, then every 2nd of 8 (inputs beyond 10 stay at 4)
    assert _step_table(lib, [("gather", (14,), [(2, 10, 2)])], 14, range(15)) == \
        [0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4, 4, 4]
    # scaled<4>(10): every element stretched over 4
    assert _step_table(lib, [("affine", 4, 0, 10)], 10, range(10)) == [0, 4, 8, 12, 16, 20, 24, 28, 32, 36]
    # fixed<4>(10): ten elements, only the first one is visible, stretched over 4
    # (closed form: a gather of element 0 followed by a scale by 4)
    assert _step_table(lib, [("gather", (10,), [(0, 1, 1)]), ("affine", 4, 0, 1)], 10, range(10)) == \
        [0, 4, 4, 4, 4, 4, 4, 4, 4, 4]
    # build(3, 7, scaled<1>(10)): the slice [3, 7)
    assert _step_table(lib, [("gather", (10,), [(3, 7, 1)])], 10, range(10)) == [0, 0, 0, 0, 1, 2, 3, 4, 4, 4]


def test_subbox_gather_sum_kat(lib):
    # step_test.cpp:66-113: the step function of the 2x2x2 box [1,3)^3 of a 4^3 grid is walked
    # over the source positions; it takes cnt == 8 distinct values `end`, and with v[i] = i % 10
    # the sum of v[end - 1] is 28 (the ends are 1..8).  Here: the mapped ends of an all-distinct
    # source are the run ends of the gathered result.
    d = 4
    n = d ** 3
    view = [("gather", (d, d, d), [(1, 3, 1)] * 3)]
    src = (np.arange(1, n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), view)
    e, v = run_eval(lib, [src], ("src", 0))
    vtab = np.arange(n) % 10
    assert len(e) == 8 and int(vtab[e - 1].sum()) == 28
    want = np.arange(n).reshape(d, d, d)[1:3, 1:3, 1:3].reshape(-1)
    assert np.array_equal(O.to_dense(e, v), want)


@pytest.mark.parametrize("n,step", [(60000, 20000), (60000, 59999), (30000, 7)])
def test_gather_collapsing_long_groups(lib, n, step):
    """x[0:n:step] of an array with one run per element: all raw runs between two selected
    positions map onto the same output coordinate.  The partition cuts tiles in the total
    order (mapped end, source, index), so such a group is split like any other run sequence."""
    if lib.is_simulator():
        n, step = n // 6, max(1, step // 6)
    ends = np.arange(1, n + 1, dtype=np.int32)
    vals = (np.arange(n) % 1000).astype(np.int32)
    sel = np.arange(0, n, step)
    view = [("gather", (n,), [(0, n, step)])]
    srcs = [(ends, vals, view)]
    got = run_eval(lib, srcs, ("src", 0))
    assert_runs_equal(got, oracle_eval(srcs, ("src", 0)), "strided pick")
    assert np.array_equal(O.to_dense(*got), vals[sel])
    # two such sources and a plain one in the same evaluation
    e2, v2 = rand_runs(np.random.default_rng(7), len(sel), min(len(sel), 3))
    srcs = [(ends, vals, view), (ends, vals[::-1].copy(), view), (e2, v2, [])]
    expr = ("add", ("mul", ("src", 0), ("src", 2)), ("src", 1))
    assert_runs_equal(run_eval(lib, srcs, expr), oracle_eval(srcs, expr), "collapsed groups in two sources")


def test_row_pick_of_run_dense_tensor(lib):
    # a 2-D box that keeps a narrow column band of every row: thousands of runs between two
    # selected positions
    w, h = (10000, 6) if not lib.is_simulator() else (2000, 6)
    n = w * h
    ends = np.arange(1, n + 1, dtype=np.int32)
    vals = np.arange(n, dtype=np.int32)
    for box in ([(7, 8, 1), (0, h, 1)], [(7, 9, 1), (1, 5, 1)]):
        view = [("gather", (w, h), box)]
        got = run_eval(lib, [(ends, vals, view)], ("src", 0))
        want = vals.reshape(h, w)[box[1][0]:box[1][1], box[0][0]:box[0][1]].reshape(-1)
        assert np.array_equal(O.to_dense(*got), want)


def test_queued_sources_chain_without_host_sync(lib):
    """An evaluation whose source is itself a queued result (psk_eval_async) reads that
    source's run count on the device; a chain of four dependent steps equals the oracle."""
    rng = np.random.default_rng(77)
    n = 50_000 if not lib.is_simulator() else 4000
    ops = [rand_runs(rng, n, n // 20, lo=-5, hi=5) for _ in range(5)]
    stores = [lib.store_from_runs(e, v) for e, v in ops]
    I = A.OP_I
    add = [(A.OP_SRC, 0), (A.OP_SRC, 1), (I["add"], 0)]
    acc = stores[0]
    for s in stores[1:]:
        acc = lib.prepare_eval([(acc, []), (s, [])], add, A.I32, deferred=True)()
    # a queued result also feeds a reduction and a slice without being completed first
    total = lib.reduce(acc, A.RED_SUM)
    e, v = acc.runs()
    want = O.eval_sources([(ee, vv, []) for ee, vv in ops], lambda a, b, c, d, f: O.add(O.add(O.add(O.add(a, b), c), d), f))
    assert np.array_equal(e, want[0]) and np.array_equal(v, want[1])
    assert int(np.int32(total)) == int(O.reduce_runs(want[0], want[1], "sum"))


def test_nan_payloads_under_bitwise_equality(lib):
    """Documented deviation (float semantics): x86 propagates the payload of a NaN
    operand through arithmetic, the GPU returns the canonical NaN.  Under BITWISE equality
    (Array.eval(), box.hpp:84-90) two adjacent NaN spans with DIFFERENT payloads therefore stay
    two runs on the CPU when they pass through arithmetic, and merge on the GPU.  A NaN that is
    only copied (no arithmetic) keeps its bits on both, and so does the run structure."""
    bits = np.array([0x7FC00001, 0x7FC00002, 0x3F800000], dtype=np.uint32)
    vals = bits.view(np.float32)
    ends = np.array([2, 4, 6], dtype=np.int32)
    # copy only: payloads and run structure are preserved exactly
    e, v = run_eval(lib, [(ends, vals, [])], ("src", 0), "bitwise")
    assert e.tolist() == [2, 4, 6] and np.array_equal(v.view(np.uint32), bits)
    # through arithmetic: the GPU canonicalises, so the two NaN runs compare equal bit for bit
    e, v = run_eval(lib, [(ends, vals, [])], ("mul", ("src", 0), ("const", 1.0, np.float32)), "bitwise")
    if lib.is_simulator():           # host arithmetic keeps the payloads: reference behaviour
        assert e.tolist() == [2, 4, 6]
    else:                            # device arithmetic: one canonical NaN run
        assert e.tolist() == [4, 6] and np.isnan(v[0]) and v[1] == 1.0
    # TYPED equality never merges NaNs, on either side (operator== on float)
    e, v = run_eval(lib, [(ends, vals, [])], ("mul", ("src", 0), ("const", 1.0, np.float32)), "typed")
    assert e.tolist() == [2, 4, 6]
