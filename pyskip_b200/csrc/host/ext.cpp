// ext.cpp -- pybind11 module `_pyskip_b200_ext`: the Python-visible surface of the
// reference's `_pyskip_cpp_ext` (src/cpp_ext/pyskip_ext.cpp:20-73, binds/type_binds.hpp:
// 96-430) over the device-backed host layer: {Int,Float,Char,Bool}{Array,Builder},
// Tensor{1..4}{i,f,c,b}, from_numpy, and the `config` submodule -- same class names, method
// names, argument meaning and error mapping (std::invalid_argument -> ValueError,
// std::logic_error -> RuntimeError).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <optional>

#include "array.hpp"
#include "builder.hpp"

namespace py = pybind11;
using namespace psk_host;

namespace {

template <int DT> struct CType;
template <> struct CType<PSK_I32> { using type = int32_t; using py_scalar = long long; };
template <> struct CType<PSK_F32> { using type = float; using py_scalar = double; };
template <> struct CType<PSK_BOOL> { using type = bool; using py_scalar = bool; };
template <> struct CType<PSK_CHAR> { using type = int8_t; using py_scalar = char; };

// distinct C++ types per element type so that pybind11 sees distinct Python classes
template <int DT> struct TArray { Array a; };
template <int DT> struct TBuilder { ArrayBuilder b; };
template <int DIM, int DT> struct TTensor { Tensor t; };

template <int DT>
TArray<DT> wrap(Array a) {
  return TArray<DT>{std::move(a)};
}

template <int DT>
py::object to_py_scalar(double v) {
  if (DT == PSK_F32) return py::float_((float)v);
  if (DT == PSK_BOOL) return py::bool_(v != 0);
  if (DT == PSK_CHAR) return py::str(std::string(1, (char)(int)v));
  return py::int_((long long)v);
}

template <int DT>
double from_py_scalar(typename CType<DT>::py_scalar v) {
  return (double)v;
}

// convert_slice (type_binds.hpp:42-71)
Slice convert_slice(Pos length, const py::slice& sl) {
  long long start = 0, stop = length, stride = 1;
  if (!sl.attr("start").is_none()) start = sl.attr("start").cast<long long>();
  if (!sl.attr("stop").is_none()) stop = sl.attr("stop").cast<long long>();
  if (!sl.attr("step").is_none()) stride = sl.attr("step").cast<long long>();
  if (start < 0) start += length;
  if (stop < 0) stop += length;
  PSK_CHECK_ARGUMENT(stride > 0);
  stop = std::min<long long>(stop, length);
  start = std::min(start, stop);
  PSK_CHECK_ARGUMENT(start >= -0x7fffffffLL && stride <= 0x7fffffffLL);
  return Slice((Pos)start, (Pos)stop, (Pos)stride);
}

// convert_band (type_binds.hpp:15-40)
Band convert_band(Pos length, const py::slice& sl) {
  PSK_CHECK_ARGUMENT(sl.attr("step").is_none());
  long long start = 0, stop = length;
  if (!sl.attr("start").is_none()) start = sl.attr("start").cast<long long>();
  if (!sl.attr("stop").is_none()) stop = sl.attr("stop").cast<long long>();
  if (start < 0) start += length;
  if (stop < 0) stop += length;
  stop = std::min<long long>(stop, length);
  start = std::min(start, stop);
  return Band((Pos)start, (Pos)stop);
}

template <int DT>
void bind_array(py::module& m, const char* name) {
  using A = TArray<DT>;
  using S = typename CType<DT>::py_scalar;
  using T = typename CType<DT>::type;
  auto cls = py::class_<A>(m, name);
  cls.def(py::init([](Pos span, S fill) { return wrap<DT>(Array::filled(DT, span, from_py_scalar<DT>(fill))); }))
      .def("__len__", [](const A& s) { return s.a.len(); })
      .def("__repr__", [](const A& s) { return s.a.repr(); })
      .def("clone", [](const A& s) { return wrap<DT>(s.a.clone()); })
      .def("eval", [](const A& s) { return wrap<DT>(s.a.eval()); })
      .def("dumps", [](const A& s) { return s.a.str(); })
      .def("tensor", [](const A&) { /* type_binds.hpp:134-136 builds a tensor and drops it */ })
      .def("to_numpy",
           [](const A& s) {
             py::array_t<T> out((py::ssize_t)s.a.len());
             if (!s.a.empty()) s.a.to_dense(out.mutable_data());
             return out;
           })
      .def("runs",
           [](const A& s) {
             std::vector<int32_t> ends;
             std::vector<uint8_t> vals;
             s.a.runs(&ends, &vals);
             py::array_t<int32_t> e((py::ssize_t)ends.size());
             py::array_t<T> v((py::ssize_t)ends.size());
             std::memcpy(e.mutable_data(), ends.data(), ends.size() * 4);
             std::memcpy(v.mutable_data(), vals.data(), vals.size());
             return py::make_tuple(e, v);
           })
      .def("rle_length", [](const A& s) { return s.a.empty() ? (int64_t)0 : s.a.store().nruns(); })
      .def("__getitem__", [](const A& s, Pos pos) { return to_py_scalar<DT>(s.a.get(pos)); })
      .def("__getitem__", [](const A& s, const py::slice& sl) { return wrap<DT>(s.a.get(convert_slice(s.a.len(), sl))); })
      .def("__setitem__", [](A& s, Pos pos, S val) { s.a.set(pos, from_py_scalar<DT>(val)); })
      .def("__setitem__",
           [](A& s, const py::slice& sl, S val) { s.a.set(convert_slice(s.a.len(), sl), from_py_scalar<DT>(val)); })
      .def("__setitem__",
           [](A& s, const py::slice& sl, const A& other) { s.a.set(convert_slice(s.a.len(), sl), other.a); });

  auto bin = [&cls](const char* fwd, const char* rev, OpName op) {
    cls.def(fwd, [op](const A& s, S v) { return wrap<DT>(s.a.binary(op, from_py_scalar<DT>(v))); });
    if (rev) cls.def(rev, [op](const A& s, S v) { return wrap<DT>(s.a.rbinary(op, from_py_scalar<DT>(v))); });
    cls.def(fwd, [op](const A& s, const A& o) { return wrap<DT>(s.a.binary(op, o.a)); });
  };

  // numerical operations (type_binds.hpp:174-236)
  if constexpr (DT == PSK_I32 || DT == PSK_F32) {
    cls.def("__neg__", [](const A& s) { return wrap<DT>(s.a.unary(OpName::neg)); })
        .def("__pos__", [](const A& s) { return wrap<DT>(s.a.pos()); })
        .def("__abs__", [](const A& s) { return wrap<DT>(s.a.unary(OpName::abs)); })
        .def("abs", [](const A& s) { return wrap<DT>(s.a.unary(OpName::abs)); })
        .def("sqrt", [](const A& s) { return wrap<DT>(s.a.unary(OpName::sqrt)); })
        .def("exp", [](const A& s) { return wrap<DT>(s.a.unary(OpName::exp)); });
    bin("__add__", "__radd__", OpName::add);
    bin("__sub__", "__rsub__", OpName::sub);
    bin("__mul__", "__rmul__", OpName::mul);
    bin("__floordiv__", "__rfloordiv__", OpName::div);  // C division (type_binds.hpp:193-201)
    bin("__mod__", "__rmod__", OpName::mod);
    bin("__pow__", "__rpow__", OpName::pow);
    // true division always goes through float (type_binds.hpp:202-216)
    cls.def("__truediv__",
            [](const A& s, S v) { return wrap<PSK_F32>(s.a.cast(PSK_F32).binary(OpName::div, (double)(float)v)); })
        .def("__rtruediv__",
             [](const A& s, S v) { return wrap<PSK_F32>(s.a.cast(PSK_F32).rbinary(OpName::div, (double)(float)v)); })
        .def("__truediv__", [](const A& s, const A& o) {
          return wrap<PSK_F32>(s.a.cast(PSK_F32).binary(OpName::div, o.a.cast(PSK_F32)));
        });
  }
  // bitwise (int) / logical (bool) operations (type_binds.hpp:239-299)
  if constexpr (DT == PSK_I32) {
    cls.def("__invert__", [](const A& s) { return wrap<DT>(s.a.unary(OpName::bnot)); });
    bin("__and__", "__rand__", OpName::band);
    bin("__or__", "__ror__", OpName::bor);
    bin("__xor__", "__rxor__", OpName::bxor);
    bin("__lshift__", "__rlshift__", OpName::shl);
    bin("__rshift__", "__rrshift__", OpName::shr);
  } else if constexpr (DT == PSK_BOOL) {
    cls.def("__invert__", [](const A& s) { return wrap<PSK_BOOL>(s.a.unary(OpName::lnot)); });
    auto logical = [&cls](const char* fwd, const char* rev, OpName op) {
      cls.def(fwd, [op](const A& s, S v) { return wrap<PSK_BOOL>(s.a.binary(op, from_py_scalar<DT>(v))); });
      cls.def(rev, [op](const A& s, S v) { return wrap<PSK_BOOL>(s.a.rbinary(op, from_py_scalar<DT>(v))); });
      cls.def(fwd, [op](const A& s, const A& o) { return wrap<PSK_BOOL>(s.a.binary(op, o.a)); });
    };
    logical("__and__", "__rand__", OpName::land);
    logical("__or__", "__ror__", OpName::lor);
  }
  // comparisons + coalesce + min/max (type_binds.hpp:302-357)
  auto cmp = [&cls](const char* fwd, OpName op) {
    cls.def(fwd, [op](const A& s, S v) { return wrap<PSK_BOOL>(s.a.binary(op, from_py_scalar<DT>(v))); });
    cls.def(fwd, [op](const A& s, const A& o) { return wrap<PSK_BOOL>(s.a.binary(op, o.a)); });
  };
  cmp("__eq__", OpName::eq);
  cmp("__ne__", OpName::ne);
  cmp("__le__", OpName::le);
  cmp("__lt__", OpName::lt);
  cmp("__ge__", OpName::ge);
  cmp("__gt__", OpName::gt);
  cls.attr("__hash__") = py::none();
  bin("coalesce", nullptr, OpName::coalesce);
  bin("min", nullptr, OpName::min);
  bin("max", nullptr, OpName::max);
  // conversions (type_binds.hpp:360-363)
  cls.def("int", [](const A& s) { return wrap<PSK_I32>(s.a.cast(PSK_I32)); })
      .def("char", [](const A& s) { return wrap<PSK_CHAR>(s.a.cast(PSK_CHAR)); })
      .def("float", [](const A& s) { return wrap<PSK_F32>(s.a.cast(PSK_F32)); })
      .def("bool", [](const A& s) { return wrap<PSK_BOOL>(s.a.cast(PSK_BOOL)); });
  // B200 additions: one-pass reductions over runs (src/pyskip/reduce.py collapsed)
  cls.def("_reduce", [](const A& s, int op) -> py::object {
    PSK_CHECK_ARGUMENT(!s.a.empty());
    StoreRef st = s.a.store();
    uint64_t bits = 0;
    check(backend().psk_reduce(st.get(), (psk_reduce_op)op, &bits));
    if (DT == PSK_F32 && (op == PSK_RED_SUM || op == PSK_RED_PROD)) {
      double d;
      std::memcpy(&d, &bits, 8);
      return py::float_((float)d);
    }
    return to_py_scalar<DT>(Array::value_of(DT, (uint32_t)bits));
  });
  // out[j] = reduction of the flat coordinates [j * seg_len, (j + 1) * seg_len)  (psk_reduce_segments)
  if (DT == PSK_I32 || DT == PSK_F32) {
    cls.def("_reduce_segments", [](const A& s, int op, int32_t seg_len) {
      PSK_CHECK_ARGUMENT(!s.a.empty());
      StoreRef st = s.a.store();
      psk_store_t h = nullptr;
      check(backend().psk_reduce_segments(st.get(), seg_len, (psk_reduce_op)op, &h));
      return wrap<DT>(Array::from_store(StoreRef(h)));
    });
  }
  // interop with the ctypes binding of the same library (pyskip_b200._abi): the psk_store_t of
  // the materialised array (retained for the caller, who releases it), and an array that
  // adopts a handle (takes over the caller's reference)
  cls.def("_store_handle", [](const A& s) {
    PSK_CHECK_ARGUMENT(!s.a.empty());
    StoreRef st = s.a.eval().store();
    backend().psk_store_retain(st.get());
    return (uintptr_t)st.get();
  });
  cls.def_static("_adopt_store", [](uintptr_t handle) {
    return wrap<DT>(Array::from_store(StoreRef(reinterpret_cast<psk_store_t>(handle))));
  });
  cls.def_static("_from_runs", [](py::array_t<int32_t, py::array::c_style | py::array::forcecast> ends,
                                  py::array_t<T, py::array::c_style | py::array::forcecast> vals) {
    PSK_CHECK_ARGUMENT(ends.size() == vals.size());
    return wrap<DT>(Array::from_runs(DT, ends.size(), ends.data(), vals.data()));
  });
}

template <int DT>
void bind_builder(py::module& m, const char* name) {
  using B = TBuilder<DT>;
  using A = TArray<DT>;
  using S = typename CType<DT>::py_scalar;
  py::class_<B>(m, name)
      .def(py::init([](Pos len, S val) { return B{ArrayBuilder(DT, len, from_py_scalar<DT>(val))}; }))
      .def(py::init([](const A& a) { return B{ArrayBuilder(a.a)}; }))
      .def("__len__", [](const B& s) { return s.b.len(); })
      .def("__repr__", [](const B& s) { return s.b.repr(); })
      .def("__setitem__", [](B& s, Pos pos, S val) { s.b.set(pos, from_py_scalar<DT>(val)); })
      .def("__setitem__",
           [](B& s, const py::slice& sl, S val) { s.b.set(convert_band(s.b.len(), sl), from_py_scalar<DT>(val)); })
      .def("__setitem__", [](B& s, const py::slice& sl, const A& o) { s.b.set(convert_band(s.b.len(), sl), o.a); })
      .def("dumps", [](const B& s) { return s.b.str(); })
      .def("build", [](const B& s) { return wrap<DT>(s.b.build()); });
}

template <int DIM>
TensorShape to_shape(const std::array<Pos, DIM>& s) {
  TensorShape r;
  r.dim = DIM;
  for (int i = 0; i < DIM; ++i) {
    PSK_CHECK_ARGUMENT(s[i] >= 0);
    r.s[i] = s[i];
  }
  return r;
}

template <int DIM>
TensorSlice to_tensor_slice(const TensorShape& shape, const std::array<py::slice, DIM>& slices) {
  TensorSlice r;  // convert_tensor_slice (type_binds.hpp:74-85)
  r.dim = DIM;
  for (int i = 0; i < DIM; ++i) {
    Slice s = convert_slice(shape.s[i], slices[i]);
    r.c[i] = {s.start, s.stop, s.stride};
  }
  return r;
}

template <int DIM, int DT>
void bind_tensor(py::module& m, const std::string& name) {
  using TT = TTensor<DIM, DT>;
  using A = TArray<DT>;
  using S = typename CType<DT>::py_scalar;
  auto to_pos = [](const std::array<Pos, DIM>& p) {
    std::array<Pos, kMaxTensorDim> r{{0, 0, 0, 0}};
    for (int i = 0; i < DIM; ++i) r[i] = p[i];
    return r;
  };
  py::class_<TT>(m, name.c_str())
      .def(py::init([](const std::array<Pos, DIM>& s, S v) {
        return TT{Tensor::filled(DT, to_shape<DIM>(s), from_py_scalar<DT>(v))};
      }))
      .def(py::init([](const std::array<Pos, DIM>& s, const A& a) { return TT{Tensor(to_shape<DIM>(s), a.a)}; }))
      .def("__len__", [](const TT& s) { return s.t.len(); })
      .def("__repr__", [](const TT& s) { return s.t.repr(); })
      .def("clone", [](const TT& s) { return TT{s.t.clone()}; })
      .def("eval", [](const TT& s) { return TT{s.t.eval()}; })
      .def("dumps", [](const TT& s) { return s.t.str(); })
      .def("array", [](const TT& s) { return wrap<DT>(s.t.array()); })
      .def("shape",
           [](const TT& s) {
             py::tuple t(DIM);
             for (int i = 0; i < DIM; ++i) t[i] = py::int_(s.t.shape().s[i]);
             return t;
           })
      .def("__getitem__",
           [to_pos](const TT& s, const std::array<Pos, DIM>& pos) { return to_py_scalar<DT>(s.t.get(to_pos(pos))); })
      .def("__getitem__",
           [](const TT& s, const std::array<py::slice, DIM>& sl) {
             return TT{s.t.get(to_tensor_slice<DIM>(s.t.shape(), sl))};
           })
      .def("__setitem__",
           [to_pos](TT& s, const std::array<Pos, DIM>& pos, S v) { s.t.set(to_pos(pos), from_py_scalar<DT>(v)); })
      .def("__setitem__",
           [](TT& s, const std::array<py::slice, DIM>& sl, S v) {
             s.t.set(to_tensor_slice<DIM>(s.t.shape(), sl), from_py_scalar<DT>(v));
           })
      .def("__setitem__", [](TT& s, const std::array<py::slice, DIM>& sl, const TT& o) {
        s.t.set(to_tensor_slice<DIM>(s.t.shape(), sl), o.t);
      });
}

// type_binds (type_binds.hpp:414-430)
template <int DT>
void type_binds(py::module& m, const std::string& friendly, const std::string& postfix) {
  using T = typename CType<DT>::type;
  bind_array<DT>(m, (friendly + "Array").c_str());
  bind_builder<DT>(m, (friendly + "Builder").c_str());
  bind_tensor<1, DT>(m, "Tensor1" + postfix);
  bind_tensor<2, DT>(m, "Tensor2" + postfix);
  bind_tensor<3, DT>(m, "Tensor3" + postfix);
  bind_tensor<4, DT>(m, "Tensor4" + postfix);
  m.def("from_numpy", [](py::array_t<T, py::array::c_style | py::array::forcecast>& array) {
    return wrap<DT>(Array::from_dense(DT, array.size(), array.data()));
  });
  // stencil-over-runs convolution (psk_conv_nd): `weights` is the dense kernel flattened in
  // pyskip order (x fastest); returns the flat result array, or None when the kernel shape is
  // not one the stencil kernel handles (the caller then runs the reference's tap loop)
  if (DT == PSK_I32 || DT == PSK_F32) {
    m.def("_conv_nd", [](const TArray<DT>& a, const std::vector<int32_t>& shape, const std::vector<int32_t>& kshape,
                         py::array_t<T, py::array::c_style | py::array::forcecast>& weights,
                         const std::vector<int32_t>& pads, typename CType<DT>::py_scalar fill) -> py::object {
      const int nd = (int)shape.size();
      PSK_CHECK_ARGUMENT((int)kshape.size() == nd && (int)pads.size() == nd);
      if (!backend().psk_conv_supported(nd, kshape.data())) return py::none();
      int64_t taps = 1;
      for (int v : kshape) taps *= v;
      PSK_CHECK_ARGUMENT(weights.size() == taps);
      std::vector<uint32_t> w((size_t)taps);
      memcpy(w.data(), weights.data(), (size_t)taps * 4);
      StoreRef st = a.a.eval().store();
      psk_store_t h = nullptr;
      check(backend().psk_conv_nd(st.get(), nd, shape.data(), kshape.data(), w.data(), pads.data(),
                                  bits_of(DT, from_py_scalar<DT>(fill)), PSK_EQ_BITWISE, &h));
      return py::cast(wrap<DT>(Array::from_store(StoreRef(h))));
    });
  }
}

}  // namespace

PYBIND11_MODULE(_pyskip_b200_ext, m) {
  m.doc() = "Space-optimized arrays on B200 (drop-in surface of pyskip's _pyskip_cpp_ext)";
  m.attr("__version__") = "0.1.5+b200";

  // Binding of the C-ABI library (pyskip_b200/libpskb200.so).  `force` exists for the
  // repository's kernel-logic simulator tests only; the package never passes it.
  m.def("_bind_backend", [](const std::string& path, bool force) { bind_backend(path, force); },
        py::arg("path"), py::arg("force") = false);
  m.def("_backend_path", [] { return backend().path; });
  m.def("_backend_is_simulator", [] { return backend().psk_is_simulator ? backend().psk_is_simulator() : -1; });
  m.def("_init_device", [](int device) { check(backend().psk_init(device)); });
  m.def("synchronize", [] { check(backend().psk_synchronize()); });

  // same registration order as pyskip_ext.cpp:24-27 (it decides which from_numpy overload
  // force-casts an unmatched dtype)
  type_binds<PSK_I32>(m, "Int", "i");
  type_binds<PSK_F32>(m, "Float", "f");
  type_binds<PSK_CHAR>(m, "Char", "c");
  type_binds<PSK_BOOL>(m, "Bool", "b");

  auto cfg = m.def_submodule("config", "Methods related to pyskip configuration");
  cfg.def("set_int_value", [](const std::string& key, std::optional<long> val) {
    if (val) config_set<int64_t>(key, *val); else config_clear(key);
  });
  cfg.def("set_bool_value", [](const std::string& key, std::optional<bool> val) {
    if (val) config_set<bool>(key, *val); else config_clear(key);
  });
  cfg.def("set_float_value", [](const std::string& key, std::optional<double> val) {
    if (val) config_set<double>(key, *val); else config_clear(key);
  });
  cfg.def("set_string_value", [](const std::string& key, std::optional<std::string> val) {
    if (val) config_set<std::string>(key, *val); else config_clear(key);
  });
  cfg.def("get_all_values", [] { return config_get_all(); });
  cfg.def("set_all_values", [](const ConfigMap& map) { config_set_all(map); });
}
