"""The golden cases: each is a function of (ext module, Tensor class, convolve module, reduce
module, lazy scope) so that the SAME code drives the reference (tests/golden/make_golden.py)
and this repository (tests/test_golden.py).  Inputs are generated from fixed seeds; outputs
are run arrays / dense arrays / scalars."""
import numpy as np


def dense_operand(seed, n, nruns, dtype=np.int32, lo=0, hi=1000):
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.choice(np.arange(1, n), size=nruns - 1, replace=False))
    lens = np.diff(np.concatenate([[0], cuts, [n]]))
    if dtype == np.float32:
        vals = ((1 + rng.integers(0, 64, size=nruns)) * 0.25).astype(np.float32)
    else:
        vals = rng.integers(lo, hi, size=nruns).astype(np.int32)
    return np.repeat(vals, lens)


def _runs(arr):
    e, v = arr.runs()
    return np.asarray(e), np.asarray(v)


def case_c1_int(R, Tensor, conv, red, lazy):
    # BASELINE config 1 shape at 1/10 scale: int32 a + b*c, ~1e3 runs per operand
    a, b, c = (R.from_numpy(dense_operand(1000 + j, 1_000_000, 1000)) for j in range(3))
    e, v = _runs(a + b * c)
    ee, ev = _runs((a + b * c).eval())
    return dict(ends=e, vals=v, eval_ends=ee, eval_vals=ev)


def case_c2_float(R, Tensor, conv, red, lazy):
    # BASELINE config 2 DAG at reduced size
    a, b, c, d = (R.from_numpy(dense_operand(2000 + j, 400_000, 4000, np.float32)) for j in range(4))
    z = (a * b + c).max(d) - a.abs() * 0.5
    e, v = _runs(z)
    return dict(ends=e, vals=v)


def case_int_ops(R, Tensor, conv, red, lazy):
    a = R.from_numpy(dense_operand(11, 5000, 300, lo=-50, hi=50))
    b = R.from_numpy(dense_operand(12, 5000, 200, lo=1, hi=9))
    out = {}
    exprs = {
        "add": a + b, "sub": a - b, "mul": a * b, "floordiv": a // b, "mod": a % b, "and": a & b, "or": a | b,
        "xor": a ^ b, "lshift": a << b, "rshift": a >> b, "min": a.min(b), "max": a.max(b), "neg": -a,
        "abs": abs(a), "invert": ~a, "coalesce": a.coalesce(b), "eq": a == b, "lt": a < b, "ge": a >= b,
        "truediv": a / b, "scalar": 3 * a - 7, "rsub": 5 - a, "tofloat": a.float() * 0.5, "tobool": a.bool(),
        "pow": b ** 3,
    }
    for k, x in exprs.items():
        e, v = _runs(x)
        out[k + "_ends"], out[k + "_vals"] = e, v
    return out


def case_slices(R, Tensor, conv, red, lazy):
    # tests/array_test.cpp:9-95 style sequences on a larger array
    x = R.from_numpy(dense_operand(21, 2000, 150, lo=0, hi=9))
    y = R.from_numpy(dense_operand(22, 2000, 90, lo=0, hi=9))
    out = {}
    out["get_ends"], out["get_vals"] = _runs(x[100:1900:3])
    x[10:1500:7] = y[0:213]
    out["set1_ends"], out["set1_vals"] = _runs(x)
    x[500:700] = 4
    x[3] = 5
    out["set2_ends"], out["set2_vals"] = _runs(x)
    z = x[0:1000] * y[1000:2000] + x[1000:2000]
    out["mix_ends"], out["mix_vals"] = _runs(z)
    out["dense"] = np.asarray(z.to_numpy())
    return out


def case_tensor_views(R, Tensor, conv, red, lazy):
    # tests/tensor_test.cpp / lang_test.cpp:148-172 style sub-box gathers and assigns
    t = Tensor.from_numpy(dense_operand(31, 24 * 20 * 16, 400, lo=0, hi=5).reshape(16, 20, 24))
    out = {}
    sub = t[3:21:2, 1:19, 2:14:3]
    out["sub"] = sub.to_numpy()
    out["sub_ends"], out["sub_vals"] = (np.asarray(q) for q in sub.to_runs())
    t[1:23, 2:18:2, 0:16:5] = 7
    t[0:12, 0:10, 0:8] = t[12:24, 10:20, 8:16] * 2
    out["assigned"] = t.to_numpy()
    out["assigned_ends"], out["assigned_vals"] = (np.asarray(q) for q in t.to_runs())
    return out


def case_conv3d(R, Tensor, conv, red, lazy):
    rng = np.random.default_rng(3001)
    vol = dense_operand(3000, 20 * 18 * 16, 120, lo=0, hi=6).reshape(16, 18, 20)
    ker = rng.integers(0, 8, size=(3, 3, 3)).astype(np.int32)
    t, k = Tensor.from_numpy(vol), Tensor.from_numpy(ker)
    o = conv.conv_3d(t, k, padding=1, fill=0)
    e, v = o.to_runs()
    o2 = conv.conv_3d(t, k, padding=0, fill=0)
    return dict(vol=vol, ker=ker, out=o.to_numpy(), ends=np.asarray(e), vals=np.asarray(v), out_nopad=o2.to_numpy())


def case_conv2d_sobel(R, Tensor, conv, red, lazy):
    # src/pyskip/tests/convolve_test.py:22-36 at D = 256: sum(abs(sobel (*) disc)) = 8 * D
    D = 256
    Rr = D // 4
    disc = Tensor((D, D), 0)
    for y in range(D):
        disc_ = Rr ** 2 - (y - D // 2 + 0.5) ** 2
        if disc_ < 0:
            continue
        x0 = int(D // 2 - 0.5 - disc_ ** 0.5)
        x1 = int(D // 2 - 0.5 + disc_ ** 0.5)
        disc[x0:x1, y] = 1
    sobel = Tensor.from_list([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]])
    edge = conv.conv_2d(disc, sobel, padding=1)
    e, v = edge.to_runs()
    return dict(ends=np.asarray(e), vals=np.asarray(v), total=np.array([red.sum(edge.abs())]),
                disc_sum=np.array([red.sum(disc)]))


def case_reduce(R, Tensor, conv, red, lazy):
    t = Tensor.from_numpy(dense_operand(41, 12 * 10 * 7, 90, lo=-20, hi=20).reshape(7, 10, 12))
    out = dict(total=np.array([red.sum(t)]))
    for axis in range(3):
        out[f"axis{axis}"] = red.sum(t, axis=axis, keepdims=True).to_numpy()
    # max through the pairwise tree only works for power-of-two extents in the reference
    p2 = Tensor.from_numpy(dense_operand(42, 1024, 60, lo=-500, hi=500))
    out["max"] = np.array([red.reduce(lambda a, b: a.max(b), p2)])
    out["min"] = np.array([red.reduce(lambda a, b: a.min(b), p2)])
    return out


CASES = {
    "c1_int": case_c1_int,
    "c2_float": case_c2_float,
    "int_ops": case_int_ops,
    "slices": case_slices,
    "tensor_views": case_tensor_views,
    "conv3d": case_conv3d,
    "conv2d_sobel": case_conv2d_sobel,
    "reduce": case_reduce,
}
