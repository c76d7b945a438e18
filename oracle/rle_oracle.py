"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the run-length-encoded (run-end encoded) array evaluation path of
turtlesoupy/pyskip.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg may import this module.  Nothing under
``pyskip_b200/`` imports it, and the product path never falls back to it.

Parity status: PINNED.  ``tests/test_oracle_*.py`` check this restatement against
  * every known-answer test the reference's own tests hold for the path, and
  * outputs of the UNMODIFIED reference itself (``oracle/_ref``, built by ``oracle/build_ref.sh``
    from the sources under /root/reference) plus the committed fixtures in ``tests/golden/``
    that ``tests/golden/make_golden.py`` generated from it.

Every function cites the reference file:line (relative to /root/reference) whose observable
behaviour it restates.  Nothing here is copied from the reference: the reference walks run
ends with a tournament tree one at a time; this file states the same input->output relation
with whole-array numpy operations (sorted union of mapped run ends, one evaluation per
distinct output span, drop-left compaction of equal neighbours).
"""
from __future__ import annotations

import numpy as np

INT32_MIN = np.int32(-2 ** 31)

# --------------------------------------------------------------------------------------
# Store <-> dense          include/pyskip/detail/core.hpp:12-62, detail/conv.hpp:15-78
# --------------------------------------------------------------------------------------


def typed_ne(a, b):
    """C++ ``a != b`` on the element type (float: NaN != NaN, -0.0 == +0.0)."""
    return a != b


def bits32(v):
    """The 32-bit payload a value occupies inside the reference's Box
    (detail/box.hpp:71-90): int32 as is, float32 bit pattern, bool 0/1, char low byte."""
    v = np.asarray(v)
    if v.dtype == np.float32:
        return v.view(np.uint32)
    if v.dtype == np.int32:
        return v.view(np.uint32)
    if v.dtype == np.bool_:
        return v.astype(np.uint32)
    if v.dtype == np.int8:
        return v.view(np.uint8).astype(np.uint32)
    raise TypeError(v.dtype)


def to_store(dense):
    """dense buffer -> (ends, vals).  detail/conv.hpp:15-44: a boundary wherever
    ``buffer[i] != buffer[i-1]`` (typed compare) and the stored value of a run is
    ``buffer[i-1]``, i.e. the run's LAST element (matters for -0.0/+0.0)."""
    dense = np.asarray(dense)
    n = dense.shape[0]
    assert n > 0
    cut = np.flatnonzero(typed_ne(dense[1:], dense[:-1])) + 1
    ends = np.concatenate([cut, [n]]).astype(np.int32)
    vals = dense[ends - 1].copy()
    return ends, vals


def to_dense(ends, vals):
    """(ends, vals) -> dense buffer.  detail/conv.hpp:63-78."""
    ends = np.asarray(ends, dtype=np.int64)
    lens = np.diff(np.concatenate([[0], ends]))
    return np.repeat(np.asarray(vals), lens)


def to_string(ends, vals):
    """The golden text format every reference C++ test asserts on,
    ``"end=>val, end=>val"`` (detail/conv.hpp:102-115)."""
    def fmt(v):
        if isinstance(v, (np.bool_, bool)):
            return "true" if v else "false"
        if isinstance(v, np.int8):
            return chr(int(v))
        if isinstance(v, (np.floating, float)):
            f = float(v)
            return repr(f) if f != int(f) else f"{f:.1f}"
        return str(int(v))
    return ", ".join(f"{int(e)}=>{fmt(v)}" for e, v in zip(ends, vals))


def canonical(ends, vals, eq="bitwise"):
    """Canonical maximal-run form of a run list that may hold zero-length runs and equal
    neighbours.  Rule (detail/eval.hpp:189-202 rewind+emit, :581-588 fuse_stores):
    neighbours equal under ``eq`` merge and the LATER run's value is kept."""
    ends = np.asarray(ends, dtype=np.int64)
    vals = np.asarray(vals)
    prev = np.concatenate([[0], ends[:-1]])
    live = ends > prev
    ends, vals = ends[live], vals[live]
    keep = np.ones(len(ends), dtype=bool)
    if len(ends) > 1:
        keep[:-1] = ~_equal(vals[:-1], vals[1:], eq)
    return ends[keep].astype(np.int32), vals[keep].copy()


def _equal(a, b, eq):
    if eq == "bitwise":   # detail/box.hpp:84-90
        return bits32(a) == bits32(b)
    if eq == "typed":     # detail/eval.hpp:191 with Val != Box (lang.hpp:899)
        return a == b
    raise ValueError(eq)


# --------------------------------------------------------------------------------------
# Views (step functions)     include/pyskip/detail/step.hpp:85-172,271-598
#                            include/pyskip/tensor.hpp:99-198, array.hpp:54-64
# Closed forms of SURVEY.md Appendix A.3.  A strided box over a flat array of shape S
# (dim 0 fastest) selects positions sel(0) < ... < sel(m-1).
# --------------------------------------------------------------------------------------


def _box(shape, comps):
    shape = [int(s) for s in shape]
    comps = [(int(a), int(b), int(c)) for a, b, c in comps]
    assert len(shape) == len(comps)
    scale, acc = [], 1
    for s in shape:
        scale.append(acc)
        acc *= s
    cnt = [max(0, -(-(c1 - c0) // cs)) for c0, c1, cs in comps]
    return shape, comps, scale, acc, cnt


def gather_map(E, shape, comps):
    """f(E) = |{ j : sel(j) < E }| -- where a source run end E lands in the output of a
    gather view.  Restates TensorSlice::get_fn (tensor.hpp:99-125) / Slice::get_fn
    (array.hpp:54-56) evaluated through cyclic::StepFn::operator() (step.hpp:105-117)."""
    shape, comps, scale, n, cnt = _box(shape, comps)
    E = np.clip(np.asarray(E, dtype=np.int64), 0, n)
    D = len(shape)
    out = np.zeros_like(E)
    ok = np.ones(E.shape, dtype=bool)
    for d in range(D - 1, -1, -1):
        x = E // scale[d]
        if d < D - 1:
            x = x % shape[d]
        c0, c1, cs = comps[d]
        below = np.clip(-(-(x - c0) // cs), 0, cnt[d])        # selected coords < x
        inner = 1
        for dd in range(d):
            inner *= cnt[dd]
        out += np.where(ok, below * inner, 0)
        ok &= (x >= c0) & (x < c1) & ((x - c0) % cs == 0)
    return out


def scatter_map(E, shape, comps):
    """g(0)=0, g(E)=sel(E) for 0<E<m, g(m)=n -- where run end E of the small array lands
    when it is scattered into the big one.  Restates TensorSlice::set_fn
    (tensor.hpp:127-169) / cyclic::insert_fn (step.hpp:583-598)."""
    shape, comps, scale, n, cnt = _box(shape, comps)
    m = int(np.prod(cnt, dtype=np.int64))
    E = np.clip(np.asarray(E, dtype=np.int64), 0, m)
    flat = np.zeros_like(E)
    rem = E.copy()
    for d in range(len(shape)):
        j = rem % cnt[d] if cnt[d] else rem * 0
        rem = rem // cnt[d] if cnt[d] else rem
        c0, _, cs = comps[d]
        flat += (c0 + j * cs) * scale[d]
    flat = np.where(E <= 0, 0, flat)
    flat = np.where(E >= m, n, flat)
    return flat


def affine_map(E, scale, offset, span_in):
    """f(E) = offset + scale*E clipped to E in [0, span_in] (cyclic::scaled / shift,
    step.hpp:319-330,346-355; used by tests/eval_test.cpp:45-64 to weave two stores)."""
    E = np.clip(np.asarray(E, dtype=np.int64), 0, span_in)
    return np.where(E <= 0, 0, offset + scale * E)


def box_len(shape, comps):
    return int(np.prod(_box(shape, comps)[4], dtype=np.int64))


def map_ends(ends, views):
    """Apply a chain of views, innermost first (lang::slice_eval, lang.hpp:37-42)."""
    m = np.asarray(ends, dtype=np.int64)
    for v in views:
        kind = v[0]
        if kind == "gather":
            m = gather_map(m, v[1], v[2])
        elif kind == "scatter":
            m = scatter_map(m, v[1], v[2])
        elif kind == "affine":
            m = affine_map(m, v[1], v[2], v[3])
        else:
            raise ValueError(kind)
    return m


def view_span(span_in, views):
    return int(map_ends(np.array([span_in]), views)[0])


def mask_store(shape, comps):
    """RLE indicator (0/1, dtype bool) of the positions a strided box selects inside a flat
    array -- TensorSlice::mask (tensor.hpp:171-198), mask::stride_mask / mask::build
    (detail/mask.hpp:132-200).  Built from row segments, then canonicalised."""
    shape, comps, scale, n, cnt = _box(shape, comps)
    m = int(np.prod(cnt, dtype=np.int64))
    if m == 0:
        return np.array([n], np.int32), np.array([False])
    if comps[0][2] == 1:                       # dim-0 stride 1: one segment per row
        seg_len = cnt[0]
        starts = scatter_map(np.arange(0, m, seg_len, dtype=np.int64), shape, comps)
        starts[0] = _sel0(shape, comps)
    else:
        seg_len = 1
        starts = scatter_map(np.arange(0, m, dtype=np.int64), shape, comps)
        starts[0] = _sel0(shape, comps)
    ends = np.empty(2 * len(starts) + 1, dtype=np.int64)
    vals = np.empty(2 * len(starts) + 1, dtype=bool)
    ends[0:-1:2] = starts
    vals[0:-1:2] = False
    ends[1:-1:2] = starts + seg_len
    vals[1:-1:2] = True
    ends[-1] = n
    vals[-1] = False
    return canonical(ends, vals, "typed")


def _sel0(shape, comps):
    _, comps, scale, _, _ = _box(shape, comps)
    return sum(c[0] * s for c, s in zip(comps, scale))


# --------------------------------------------------------------------------------------
# Scalar semantics of the operator catalogue    include/pyskip/array.hpp:276-338
# (observed through the reference: SURVEY.md A.5).  int32 wraps, / and % are C
# truncating, float ops are IEEE binary32 round-to-nearest with no contraction.
# --------------------------------------------------------------------------------------

_err = dict(over="ignore", invalid="ignore", divide="ignore", under="ignore")


def _i32(x):
    return np.asarray(x).astype(np.int64).astype(np.uint32).view(np.int32) \
        if np.asarray(x).dtype != np.int32 else np.asarray(x)


def _wrap(x64):
    return (np.asarray(x64, dtype=np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


def _isf(a):
    return np.asarray(a).dtype == np.float32


def add(a, b):
    with np.errstate(**_err):
        return (a + b) if _isf(a) else _wrap(a.astype(np.int64) + b.astype(np.int64))


def sub(a, b):
    with np.errstate(**_err):
        return (a - b) if _isf(a) else _wrap(a.astype(np.int64) - b.astype(np.int64))


def mul(a, b):
    with np.errstate(**_err):
        return (a * b) if _isf(a) else _wrap(a.astype(np.int64) * b.astype(np.int64))


def div(a, b):
    """array.hpp:292 -- float IEEE divide; int C truncating divide (b==0 undefined)."""
    with np.errstate(**_err):
        if _isf(a):
            return a / b
        a64, b64 = a.astype(np.int64), b.astype(np.int64)
        safe = np.where(b64 == 0, 1, b64)
        q = np.abs(a64) // np.abs(safe)
        return _wrap(np.where((a64 < 0) != (safe < 0), -q, q))


def mod(a, b):
    """array.hpp:293-298 -- int C remainder (sign of dividend); float fmodf."""
    with np.errstate(**_err):
        if _isf(a):
            return np.fmod(a, b).astype(np.float32)
        a64, b64 = a.astype(np.int64), b.astype(np.int64)
        safe = np.where(b64 == 0, 1, b64)
        r = np.abs(a64) % np.abs(safe)
        return _wrap(np.where(a64 < 0, -r, r))


def neg(a):
    return (-a) if _isf(a) else _wrap(-a.astype(np.int64))


def pos(a):
    return a


def bnot(a):
    return ~a


def band(a, b):
    return a & b


def bor(a, b):
    return a | b


def bxor(a, b):
    return a ^ b


def shl(a, b):   # x86 masks the count to 5 bits
    return _wrap(a.astype(np.int64) << (b.astype(np.int64) & 31))


def shr(a, b):
    return (a >> (b & 31)).astype(np.int32)


def absv(a):
    return np.abs(a) if _isf(a) else _wrap(np.abs(a.astype(np.int64)))


def sqrtv(a):
    """array.hpp:310-312 -- static_cast<Val>(std::sqrt(a)): sqrtf for float, double sqrt
    then truncation for int."""
    with np.errstate(**_err):
        if _isf(a):
            return np.sqrt(a)
        return _f64_to_i32(np.sqrt(a.astype(np.float64)))


def expv(a):
    with np.errstate(**_err):
        if _isf(a):
            return np.exp(a.astype(np.float32)).astype(np.float32)
        return _f64_to_i32(np.exp(a.astype(np.float64)))


def powv(a, b):
    with np.errstate(**_err):
        if _isf(a):
            return np.power(a, b).astype(np.float32)
        return _f64_to_i32(np.power(a.astype(np.float64), b.astype(np.float64)))


def _f64_to_i32(x):
    """x86 cvttsd2si: truncation, 0x80000000 when out of range or NaN."""
    ok = np.isfinite(x) & (x > -2147483649.0) & (x < 2147483648.0)
    return np.where(ok, np.trunc(np.where(ok, x, 0)), -2147483648.0).astype(np.int64).astype(np.int32)


def _f32_to_i32(x):
    ok = np.isfinite(x) & (x >= -2147483648.0) & (x < 2147483648.0)
    return np.where(ok, np.trunc(np.where(ok, x, 0)), -2147483648.0).astype(np.int64).astype(np.int32)


def minv(a, b):   # std::min(a,b) = (b < a) ? b : a      array.hpp:315
    return np.where(b < a, b, a)


def maxv(a, b):   # std::max(a,b) = (a < b) ? b : a      array.hpp:316
    return np.where(a < b, b, a)


def coalesce(a, b):   # a ? a : b     array.hpp:320
    return np.where(a != 0, a, b)


def select(m, a, b):  # splat / Tensor::set lambda, array.hpp:323-325, tensor.hpp:295
    return np.where(m, a, b)


def lnot(a):
    return ~(a != 0)


def land(a, b):
    return (a != 0) & (b != 0)


def lor(a, b):
    return (a != 0) | (b != 0)


def eq(a, b):
    return a == b


def ne(a, b):
    return a != b


def lt(a, b):
    return a < b


def gt(a, b):
    return a > b


def le(a, b):
    return a <= b


def ge(a, b):
    return a >= b


def cast(a, dtype):
    """static_cast<Out>(a) (array.hpp:277-281) with x86 float->int conversion."""
    a = np.asarray(a)
    dtype = np.dtype(dtype)
    if dtype == a.dtype:
        return a
    if dtype == np.bool_:
        return a != 0
    if dtype == np.float32:
        return a.astype(np.float32)
    if dtype == np.int32:
        return _f32_to_i32(a) if a.dtype == np.float32 else a.astype(np.int32)
    if dtype == np.int8:
        if a.dtype == np.float32:
            return _f32_to_i32(a).astype(np.int8)
        return a.astype(np.int32).astype(np.int8)
    raise TypeError(dtype)


OPS = dict(add=add, sub=sub, mul=mul, div=div, mod=mod, neg=neg, pos=pos, bnot=bnot,
           band=band, bor=bor, bxor=bxor, shl=shl, shr=shr, abs=absv, sqrt=sqrtv, exp=expv,
           pow=powv, min=minv, max=maxv, coalesce=coalesce, select=select, lnot=lnot,
           land=land, lor=lor, eq=eq, ne=ne, lt=lt, gt=gt, le=le, ge=ge)
ARITY = dict(neg=1, pos=1, bnot=1, abs=1, sqrt=1, exp=1, lnot=1, select=3)


def run_program(prog, src_vals):
    """Stack machine over a postfix program (lang::execute_plan_fixed, lang.hpp:754-777).
    prog items: ("src", i) | ("const", numpy scalar) | ("cast", dtype) | (opname,)."""
    st = []
    n = len(src_vals[0]) if src_vals else 1
    for node in prog:
        k = node[0]
        if k == "src":
            st.append(src_vals[node[1]])
        elif k == "const":
            st.append(np.full(n, node[1]))
        elif k == "cast":
            st.append(cast(st.pop(), node[1]))
        else:
            ar = ARITY.get(k, 2)
            args = st[-ar:]
            del st[-ar:]
            st.append(np.asarray(OPS[k](*args)))
    assert len(st) == 1
    return st[0]


# --------------------------------------------------------------------------------------
# The hot path: k-way merge of mapped run ends + one evaluation per output span +
# equal-neighbour compaction.   include/pyskip/detail/eval.hpp:177-231 (eval),
# :469-516 (SimpleEvaluator), :568-607 (fuse_stores); driven by lang.hpp:749-784.
# --------------------------------------------------------------------------------------


def eval_sources(sources, fn, eq="bitwise", span=None):
    """sources: list of (ends, vals, views).  fn: callable over k value arrays (or a postfix
    program for run_program).  Returns canonical (ends, vals).

    Semantics: dense_out[p] = fn(dense(s_1)[p], ..., dense(s_k)[p]) and the
    result is the canonical RLE of dense_out under ``eq`` with the later value kept."""
    mapped = []
    for ends, vals, views in sources:
        mapped.append(map_ends(ends, views))
    spans = {int(m[-1]) for m in mapped}
    assert len(spans) == 1, f"source spans differ: {spans}"   # eval.hpp:400-403
    if span is None:
        span = spans.pop()
    assert span > 0                                             # eval.hpp:611
    allm = np.concatenate(mapped)
    U = np.unique(allm[(allm > 0) & (allm <= span)])            # distinct output span ends
    cur = []
    for (ends, vals, views), m in zip(sources, mapped):
        # value of a source on the span ending at e: the first run whose mapped end is
        # >= e (zero-length mapped runs vanish, eval.hpp:213-217)
        idx = np.searchsorted(m, U, side="left")
        cur.append(np.asarray(vals)[idx])
    out = run_program(fn, cur) if isinstance(fn, (list, tuple)) else np.asarray(fn(*cur))
    keep = np.ones(len(U), dtype=bool)
    if len(U) > 1:
        keep[:-1] = ~_equal(out[:-1], out[1:], eq)              # eval.hpp:191-193
    return U[keep].astype(np.int32), out[keep].copy()


def eval_dense(sources, fn, eq="bitwise"):
    """Independent dense statement of the same relation (small inputs only): materialise
    every source view densely, apply fn elementwise, re-encode."""
    dense = []
    for ends, vals, views in sources:
        m = map_ends(ends, views)
        dense.append(to_dense(*canonical_keep_first(m, vals)))
    out = run_program(fn, dense) if isinstance(fn, (list, tuple)) else np.asarray(fn(*dense))
    n = len(out)
    cut = np.flatnonzero(~_equal(out[:-1], out[1:], eq)) + 1
    ends = np.concatenate([cut, [n]]).astype(np.int32)
    return ends, out[ends - 1].copy()


def canonical_keep_first(mapped_ends, vals):
    """Drop zero-length mapped runs (keep the first run of each plateau)."""
    m = np.asarray(mapped_ends, dtype=np.int64)
    prev = np.concatenate([[0], m[:-1]])
    live = m > prev
    return m[live], np.asarray(vals)[live]


def fuse_stores(parts, eq="bitwise"):
    """Concatenate partial stores of consecutive coordinate ranges (ends already global),
    dropping a part's last run when the next part's first value is equal
    (eval::fuse_stores, eval.hpp:568-607).  This is the multi-GPU stitch rule."""
    ends = np.concatenate([p[0] for p in parts]).astype(np.int64)
    vals = np.concatenate([p[1] for p in parts])
    # only the last run of each part can be dropped (eval.hpp:581-588)
    last = np.cumsum([len(p[0]) for p in parts])[:-1] - 1
    mask = np.ones(len(ends), dtype=bool)
    mask[last] = ~_equal(vals[last], vals[last + 1], eq)
    return ends[mask], vals[mask]


# --------------------------------------------------------------------------------------
# Callers either side of the path (restated with the primitives above)
# --------------------------------------------------------------------------------------


def slice_get(store, span, start, stop, stride=1, eq="typed"):
    """Array::get(Slice) materialised (array.hpp:116-123)."""
    ends, vals = store
    view = ("gather", [span], [(start, stop, stride)])
    return eval_sources([(ends, vals, [view])], lambda a: a, eq)


def box_get(store, shape, comps, eq="typed"):
    """Tensor::get(TensorSlice) materialised (tensor.hpp:271-279)."""
    ends, vals = store
    return eval_sources([(ends, vals, [("gather", list(shape), list(comps))])], lambda a: a, eq)


def box_set(big, shape, comps, small, eq="typed"):
    """Tensor::set(TensorSlice, other) materialised (tensor.hpp:291-304):
    select(mask, scatter(small), big)."""
    m_ends, m_vals = mask_store(shape, comps)
    return eval_sources(
        [(m_ends, m_vals, []), (small[0], small[1], [("scatter", list(shape), list(comps))]),
         (big[0], big[1], [])],
        lambda m, a, b: np.where(m, a, b), eq)


def conv3d_dense(t_np, k_np, padding=0, fill=0):
    """pyskip.functional.conv_3d (src/pyskip/convolve.py:39-59) on numpy arrays whose axes
    are reversed w.r.t. pyskip order (tensor.py:78-81): t_np[z,y,x], k_np[z,y,x].
    Accumulates in the tensor dtype in tap order z -> y -> x (convolve.py:52-58)."""
    p = (padding,) * 3 if np.isscalar(padding) else tuple(padding)   # (pw, ph, pd) = x,y,z
    pz, py, px = p[2], p[1], p[0]
    P = np.full(tuple(s + 2 * q for s, q in zip(t_np.shape, (pz, py, px))), fill, dtype=t_np.dtype)
    P[pz:P.shape[0] - pz, py:P.shape[1] - py, px:P.shape[2] - px] = t_np
    kd, kh, kw = k_np.shape
    oz, oy, ox = P.shape[0] - kd + 1, P.shape[1] - kh + 1, P.shape[2] - kw + 1
    out = np.zeros((oz, oy, ox), dtype=t_np.dtype)
    for z in range(kd):
        for y in range(kh):
            for x in range(kw):
                w = np.full(1, k_np[z, y, x], dtype=t_np.dtype)
                out = add(out, mul(w, P[z:z + oz, y:y + oy, x:x + ox]))
    return out


def conv2d_dense(t_np, k_np, padding=0, fill=0):
    """pyskip.functional.conv_2d (src/pyskip/convolve.py:5-24); numpy axes [y,x]."""
    p = (padding,) * 2 if np.isscalar(padding) else tuple(padding)
    py, px = p[1], p[0]
    P = np.full((t_np.shape[0] + 2 * py, t_np.shape[1] + 2 * px), fill, dtype=t_np.dtype)
    P[py:P.shape[0] - py, px:P.shape[1] - px] = t_np
    kh, kw = k_np.shape
    oy, ox = P.shape[0] - kh + 1, P.shape[1] - kw + 1
    out = np.zeros((oy, ox), dtype=t_np.dtype)
    for y in range(kh):
        for x in range(kw):
            w = np.full(1, k_np[y, x], dtype=t_np.dtype)
            out = add(out, mul(w, P[y:y + oy, x:x + ox]))
    return out


def reduce_pairwise(op, dense):
    """pyskip.reduce.reduce over a flat array (src/pyskip/reduce.py:8-33): odd n ->
    op(reduce(t[:-1]), t[-1]); even n -> reduce(op(t[0::2], t[1::2]))."""
    t = np.asarray(dense)
    pending = []
    while len(t) > 1:
        if len(t) % 2 == 1:
            pending.append(t[-1:])
            t = t[:-1]
        else:
            t = np.asarray(op(t[0::2], t[1::2]))
    acc = t
    for last in reversed(pending):
        acc = np.asarray(op(acc, last))
    return acc[0]


def reduce_runs(ends, vals, kind):
    """Closed form over runs used for int32 (order independent mod 2^32) and max/min:
    sum = sum(val * len) mod 2^32."""
    ends = np.asarray(ends, dtype=np.int64)
    lens = np.diff(np.concatenate([[0], ends]))
    vals = np.asarray(vals)
    if kind == "sum":
        if vals.dtype == np.float32:
            return np.float64(np.sum(vals.astype(np.float64) * lens))
        return _wrap(np.sum((vals.astype(np.int64) * lens) & 0xFFFFFFFF))[()]
    if kind == "max":
        return vals.max()
    if kind == "min":
        return vals.min()
    raise ValueError(kind)
