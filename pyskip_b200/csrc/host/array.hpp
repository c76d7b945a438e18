// array.hpp -- Array / Tensor handles of the host layer.
//
// Same public surface and semantics as include/pyskip/array.hpp (Slice :28-65, Array
// :67-223, operators :276-338 via macros.hpp:3-64) and include/pyskip/tensor.hpp
// (TensorSlice :57-199, Tensor :210-338), re-stated over the device-backed expression DAG of
// lang.hpp.  One dtype-tagged class replaces the reference's Array<Val> template; the typed
// facades (IntArray, FloatArray, ...) live in the pybind module.
#pragma once

#include <charconv>
#include <cmath>
#include <cstdio>

#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "config.hpp"
#include "lang.hpp"

namespace psk_host {

using Pos = int32_t;  // core::Pos (detail/core.hpp:10)

inline uint32_t bits_of(int dtype, double v) {
  switch (dtype) {
    case PSK_F32: {
      float f = (float)v;
      uint32_t u;
      std::memcpy(&u, &f, 4);
      return u;
    }
    case PSK_BOOL: return v != 0 ? 1u : 0u;
    case PSK_CHAR: return (uint32_t)(int32_t)(int8_t)(long long)v;
    default: return (uint32_t)(int32_t)(long long)v;
  }
}

inline int elem_size(int dtype) { return (dtype == PSK_I32 || dtype == PSK_F32) ? 4 : 1; }
inline const char* type_letter(int dtype) {  // typeid(Val).name() on the Itanium ABI
  switch (dtype) {
    case PSK_F32: return "f";
    case PSK_BOOL: return "b";
    case PSK_CHAR: return "c";
    default: return "i";
  }
}

// array.hpp:28-65
struct Slice {
  Pos start, stop, stride;
  Slice(Pos start, Pos stop, Pos stride = 1) : start(start), stop(stop), stride(stride) {
    PSK_CHECK_ARGUMENT(0 <= start);
    PSK_CHECK_ARGUMENT(start <= stop);
    PSK_CHECK_ARGUMENT(stride > 0);
  }
  Pos len() const { return (stop - start + stride - 1) / stride; }
};

inline psk_view box_view(int kind, int ndim, const Pos* shape, const Pos* c0, const Pos* c1, const Pos* cs) {
  psk_view v;
  std::memset(&v, 0, sizeof(v));
  v.kind = kind;
  v.ndim = ndim;
  for (int i = 0; i < ndim; ++i) {
    v.shape[i] = shape[i];
    v.start[i] = c0[i];
    v.stop[i] = c1[i];
    v.stride[i] = cs[i];
  }
  return v;
}

enum class OpName {
  add, sub, mul, div, mod, neg, abs, bnot, band, bor, bxor, shl, shr, min, max, pow, sqrt, exp, coalesce,
  eq, ne, lt, gt, le, ge, lnot, land, lor
};

inline int opcode_for(OpName n, int dtype) {
  const bool f = dtype == PSK_F32;
  switch (n) {
    case OpName::add: return f ? PSK_OP_ADD_F : PSK_OP_ADD_I;
    case OpName::sub: return f ? PSK_OP_SUB_F : PSK_OP_SUB_I;
    case OpName::mul: return f ? PSK_OP_MUL_F : PSK_OP_MUL_I;
    case OpName::div: return f ? PSK_OP_DIV_F : PSK_OP_DIV_I;
    case OpName::mod: return f ? PSK_OP_MOD_F : PSK_OP_MOD_I;
    case OpName::neg: return f ? PSK_OP_NEG_F : PSK_OP_NEG_I;
    case OpName::abs: return f ? PSK_OP_ABS_F : PSK_OP_ABS_I;
    case OpName::min: return f ? PSK_OP_MIN_F : PSK_OP_MIN_I;
    case OpName::max: return f ? PSK_OP_MAX_F : PSK_OP_MAX_I;
    case OpName::pow: return f ? PSK_OP_POW_F : PSK_OP_POW_I;
    case OpName::sqrt: return f ? PSK_OP_SQRT_F : PSK_OP_SQRT_I;
    case OpName::exp: return f ? PSK_OP_EXP_F : PSK_OP_EXP_I;
    case OpName::coalesce: return f ? PSK_OP_COALESCE_F : PSK_OP_COALESCE_I;
    case OpName::eq: return f ? PSK_OP_EQ_F : PSK_OP_EQ_I;
    case OpName::ne: return f ? PSK_OP_NE_F : PSK_OP_NE_I;
    case OpName::lt: return f ? PSK_OP_LT_F : PSK_OP_LT_I;
    case OpName::gt: return f ? PSK_OP_GT_F : PSK_OP_GT_I;
    case OpName::le: return f ? PSK_OP_LE_F : PSK_OP_LE_I;
    case OpName::ge: return f ? PSK_OP_GE_F : PSK_OP_GE_I;
    case OpName::lnot: return f ? PSK_OP_LNOT_F : PSK_OP_LNOT_I;
    case OpName::land: return f ? PSK_OP_LAND_F : PSK_OP_LAND_I;
    case OpName::lor: return f ? PSK_OP_LOR_F : PSK_OP_LOR_I;
    case OpName::bnot: PSK_CHECK_ARGUMENT(!f); return PSK_OP_BNOT_I;
    case OpName::band: PSK_CHECK_ARGUMENT(!f); return PSK_OP_AND_I;
    case OpName::bor: PSK_CHECK_ARGUMENT(!f); return PSK_OP_OR_I;
    case OpName::bxor: PSK_CHECK_ARGUMENT(!f); return PSK_OP_XOR_I;
    case OpName::shl: PSK_CHECK_ARGUMENT(!f); return PSK_OP_SHL_I;
    case OpName::shr: PSK_CHECK_ARGUMENT(!f); return PSK_OP_SHR_I;
  }
  throw std::invalid_argument("unknown operation");
}

inline bool yields_bool(OpName n) {
  switch (n) {
    case OpName::eq: case OpName::ne: case OpName::lt: case OpName::gt: case OpName::le:
    case OpName::ge: case OpName::lnot: case OpName::land: case OpName::lor:
      return true;
    default:
      return false;
  }
}

class Array {
 public:
  Array() = default;  // empty array (array.hpp:70)
  Array(int dtype, ExprPtr op) : dtype_(dtype), op_(std::move(op)) { maybe_flush(); }

  // make_array(len, val) (array.hpp:196-201, 227-229)
  static Array filled(int dtype, Pos len, double val) {
    PSK_CHECK_ARGUMENT(len >= 0);
    if (len == 0) return Array(dtype);
    return Array(dtype, make_fill_expr(dtype, len, bits_of(dtype, val)));
  }
  static Array from_store(StoreRef st) {
    int dt = st.dtype();
    return Array(dt, make_store_expr(std::move(st), dt));
  }
  // from_buffer (array.hpp:243-249): dense host buffer -> RLE on device
  static Array from_dense(int dtype, int64_t n, const void* data) {
    if (n == 0) return Array(dtype);
    psk_store_t h = nullptr;
    check(backend().psk_store_from_dense((psk_dtype)dtype, n, data, &h));
    return from_store(StoreRef(h));
  }
  static Array from_runs(int dtype, int64_t nruns, const int32_t* ends, const void* vals) {
    psk_store_t h = nullptr;
    check(backend().psk_store_from_runs((psk_dtype)dtype, nruns, ends, vals, &h));
    return from_store(StoreRef(h));
  }

  int dtype() const { return dtype_; }
  Pos len() const { return op_ ? op_->span : 0; }
  bool empty() const { return len() == 0; }
  const ExprPtr& expr() const { return op_; }

  // array.hpp:102-110
  StoreRef store() const {
    PSK_CHECK_STATE(!empty());
    return materialize_typed(op_);
  }
  Array clone() const { return *this; }
  Array eval() const {
    if (empty()) return *this;
    StoreRef st = materialize_bitwise(op_);
    Array out(dtype_);
    out.op_ = wrap_store(st, dtype_);
    return out;
  }

  // array.hpp:113-123
  double get(Pos pos) const {
    Array one = get(Slice(pos, pos + 1));
    StoreRef st = one.store();
    uint32_t bits = 0;
    check(backend().psk_store_first_value(st.get(), &bits));
    return value_of(dtype_, bits);
  }
  Array get(const Slice& s) const {
    PSK_CHECK_ARGUMENT(s.stop <= len());
    if (s.len() == 0) return Array(dtype_);
    Pos n = len();
    psk_view v = box_view(PSK_VIEW_GATHER, 1, &n, &s.start, &s.stop, &s.stride);
    return Array(dtype_, make_view_expr(op_, v));
  }

  // array.hpp:126-146
  void set(Pos pos, double val) { set(Slice(pos, pos + 1), val); }
  void set(const Slice& s, double val) {
    PSK_CHECK_ARGUMENT(s.stop <= len());
    if (s.len() > 0) set(s, filled(dtype_, s.len(), val));
  }
  void set(const Slice& s, const Array& other) {
    PSK_CHECK_ARGUMENT(s.stop <= len());
    PSK_CHECK_ARGUMENT(s.len() == other.len());
    PSK_CHECK_ARGUMENT(other.empty() || other.dtype_ == dtype_);
    if (s.len() > 0) {
      Pos n = len();
      // mask store + scattered source + select   (array.hpp:139-144)
      psk_store_t mh = nullptr;
      check(backend().psk_mask_nd(1, &n, &s.start, &s.stop, &s.stride, &mh));
      ExprPtr mask = make_store_expr(StoreRef(mh), PSK_BOOL);
      psk_view sv = box_view(PSK_VIEW_SCATTER, 1, &n, &s.start, &s.stop, &s.stride);
      ExprPtr src = make_view_expr(other.op_, sv);
      *this = Array(dtype_, make_op_expr(PSK_OP_SELECT, dtype_, mask, src, op_));
    }
  }

  // ---- merges (array.hpp:149-190, operators :276-338) ----------------------------------
  Array unary(OpName n) const {
    if (empty()) return Array(yields_bool(n) ? PSK_BOOL : dtype_);
    const int out = yields_bool(n) ? PSK_BOOL : dtype_;
    return Array(out, make_op_expr(opcode_for(n, dtype_), out, op_));
  }
  Array pos() const {  // unary + : same values, one more node (array.hpp:284)
    if (empty()) return *this;
    return Array(dtype_, make_op_expr(PSK_OP_SRC, dtype_, op_));
  }
  Array binary(OpName n, const Array& rhs) const {
    PSK_CHECK_ARGUMENT(len() == rhs.len());  // array.hpp:158
    PSK_CHECK_ARGUMENT(empty() || dtype_ == rhs.dtype_);
    const int out = yields_bool(n) ? PSK_BOOL : dtype_;
    if (empty()) return Array(out);
    return Array(out, make_op_expr(opcode_for(n, dtype_), out, op_, rhs.op_));
  }
  // scalar broadcast: make_array(len, scalar) (macros.hpp:16-23)
  Array binary(OpName n, double rhs) const { return binary(n, filled(dtype_, len(), rhs)); }
  Array rbinary(OpName n, double lhs) const { return filled(dtype_, len(), lhs).binary(n, *this); }

  // splat(m, a, b) = m ? a : b  (array.hpp:323-325)
  static Array select(const Array& m, const Array& a, const Array& b) {
    PSK_CHECK_ARGUMENT(m.dtype_ == PSK_BOOL);
    PSK_CHECK_ARGUMENT(m.len() == a.len() && m.len() == b.len());
    PSK_CHECK_ARGUMENT(a.dtype_ == b.dtype_);
    if (m.empty()) return Array(a.dtype_);
    return Array(a.dtype_, make_op_expr(PSK_OP_SELECT, a.dtype_, m.op_, a.op_, b.op_));
  }

  // cast<Out> (array.hpp:277-281)
  Array cast(int out) const {
    if (empty()) return Array(out);
    const bool src_f = dtype_ == PSK_F32;
    int op = PSK_OP_SRC;  // no-op
    if (out != dtype_) {
      if (out == PSK_F32) op = src_f ? PSK_OP_SRC : PSK_OP_I2F;
      else if (out == PSK_I32) op = src_f ? PSK_OP_F2I : PSK_OP_SRC;
      else if (out == PSK_BOOL) op = src_f ? PSK_OP_F2B : PSK_OP_I2B;
      else op = src_f ? PSK_OP_F2C : PSK_OP_I2C;
    }
    return Array(out, make_op_expr(op, out, op_));
  }

  // ---- egress ---------------------------------------------------------------------------
  // conv::to_string (detail/conv.hpp:102-119): "end=>val, end=>val"
  std::string str() const {
    if (empty()) return "";
    std::vector<int32_t> ends;
    std::vector<uint8_t> vals;
    runs(&ends, &vals);
    std::ostringstream os;
    for (size_t i = 0; i < ends.size(); ++i) {
      if (i) os << ", ";
      os << ends[i] << "=>" << format_value(vals.data() + i * elem_size(dtype_));
    }
    return os.str();
  }
  // array.hpp:84-99
  std::string repr() const {
    std::ostringstream os;
    os << "Array<" << type_letter(dtype_) << ">([";
    if (!empty()) {
      if (len() <= 10) {
        os << join_dense(*this);
      } else {
        os << join_dense(get(Slice(0, 4))) << ", ..., " << format_scalar(get(len() - 1));
      }
    }
    os << "])";
    return os.str();
  }
  void runs(std::vector<int32_t>* ends, std::vector<uint8_t>* vals) const {
    StoreRef st = store();
    const size_t n = (size_t)st.nruns();
    ends->resize(n);
    vals->resize(n * elem_size(dtype_));
    check(backend().psk_store_runs(st.get(), ends->data(), vals->data()));
  }
  void to_dense(void* out) const {
    StoreRef st = store();
    check(backend().psk_store_to_dense(st.get(), out));
  }

  static double value_of(int dtype, uint32_t bits) {
    if (dtype == PSK_F32) {
      float f;
      std::memcpy(&f, &bits, 4);
      return f;
    }
    return (double)(int32_t)bits;
  }

  // float32 the way the reference prints it (fmt 6 "{}": shortest digits that round-trip, fixed
  // notation with at least one decimal for exponents in [-4, 16), exponent notation otherwise)
  static std::string format_float(float f) {
    if (std::isnan(f)) return std::signbit(f) ? "-nan" : "nan";
    if (std::isinf(f)) return f < 0 ? "-inf" : "inf";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);
    std::string sci(buf, r.ptr);  // [-]d[.ddd]e[+-]XX
    std::string out;
    size_t pos = 0;
    if (sci[0] == '-') {
      out = "-";
      pos = 1;
    }
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = pos; i < epos; ++i)
      if (sci[i] != '.') digits += sci[i];
    const int exp10 = std::atoi(sci.c_str() + epos + 1);
    if (exp10 >= -4 && exp10 < 16) {
      if (exp10 >= 0) {
        if ((int)digits.size() <= exp10 + 1) {
          out += digits + std::string((size_t)(exp10 + 1) - digits.size(), '0') + ".0";
        } else {
          out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
        }
      } else {
        out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
      }
    } else {
      out += digits.substr(0, 1);
      if (digits.size() > 1) out += "." + digits.substr(1);
      char eb[16];
      std::snprintf(eb, sizeof eb, "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
      out += eb;
    }
    return out;
  }

  std::string format_scalar(double v) const {
    std::ostringstream os;
    if (dtype_ == PSK_F32) os << format_float((float)v);
    else if (dtype_ == PSK_BOOL) os << (v != 0 ? "true" : "false");
    else if (dtype_ == PSK_CHAR) os << (char)(int)v;
    else os << (long long)v;
    return os.str();
  }

 private:
  explicit Array(int dtype) : dtype_(dtype) {}

  static ExprPtr wrap_store(const StoreRef& st, int dtype) {
    ExprPtr e = make_store_expr(st, dtype);
    if (st.nruns() == 1) {
      uint32_t bits = 0;
      check(backend().psk_store_first_value(st.get(), &bits));
      e->is_const = true;
      e->const_bits = bits;
    }
    return e;
  }

  // array.hpp:204-214: expressions above flush_tree_size_threshold are evaluated eagerly
  void maybe_flush() {
    if (!op_) return;
    int64_t threshold = config_get_int("flush_tree_size_threshold", 32);
    if (threshold > 4096) threshold = 4096;  // bound planner recursion in "lazy" scopes
    if (op_->size > threshold) op_ = wrap_store(materialize_bitwise(op_), dtype_);
  }

  std::string format_value(const uint8_t* p) const {
    if (dtype_ == PSK_F32) {
      float f;
      std::memcpy(&f, p, 4);
      return format_scalar(f);
    }
    if (dtype_ == PSK_I32) {
      int32_t i;
      std::memcpy(&i, p, 4);
      return format_scalar(i);
    }
    return format_scalar((double)(int8_t)*p);
  }
  std::string join_dense(const Array& a) const {
    std::vector<uint8_t> buf((size_t)a.len() * elem_size(dtype_));
    a.to_dense(buf.data());
    std::ostringstream os;
    for (Pos i = 0; i < a.len(); ++i) {
      if (i) os << ", ";
      os << format_value(buf.data() + (size_t)i * elem_size(dtype_));
    }
    return os.str();
  }

  int dtype_ = PSK_I32;
  ExprPtr op_;
};

// ---- tensors (tensor.hpp) -------------------------------------------------------------
constexpr int kMaxTensorDim = 4;

struct TensorShape {
  int dim = 1;
  std::array<Pos, kMaxTensorDim> s{{1, 1, 1, 1}};
  // tensor.hpp:38-44
  Pos len() const {
    long long n = 1;
    for (int i = 0; i < dim; ++i) n *= s[i];
    PSK_CHECK_ARGUMENT(n <= 0x7fffffffLL);
    return (Pos)n;
  }
  bool operator==(const TensorShape& o) const {
    if (dim != o.dim) return false;
    for (int i = 0; i < dim; ++i)
      if (s[i] != o.s[i]) return false;
    return true;
  }
};

// tensor.hpp:57-199
struct TensorSlice {
  int dim = 1;
  std::array<std::array<Pos, 3>, kMaxTensorDim> c{};
  void validate() const {
    for (int i = 0; i < dim; ++i) {
      PSK_CHECK_ARGUMENT(0 <= c[i][0]);
      PSK_CHECK_ARGUMENT(c[i][0] <= c[i][1]);
      PSK_CHECK_ARGUMENT(c[i][2] > 0);
    }
  }
  bool valid(const TensorShape& shape) const {
    for (int i = 0; i < dim; ++i)
      if (c[i][1] > shape.s[i]) return false;
    return true;
  }
  TensorShape shape() const {
    TensorShape r;
    r.dim = dim;
    for (int i = 0; i < dim; ++i) r.s[i] = (c[i][1] - c[i][0] + c[i][2] - 1) / c[i][2];
    return r;
  }
  Pos len() const { return shape().len(); }
  psk_view view(int kind, const TensorShape& shape) const {
    Pos c0[kMaxTensorDim], c1[kMaxTensorDim], cs[kMaxTensorDim];
    for (int i = 0; i < dim; ++i) {
      c0[i] = c[i][0];
      c1[i] = c[i][1];
      cs[i] = c[i][2];
    }
    return box_view(kind, dim, shape.s.data(), c0, c1, cs);
  }
};

class Tensor {
 public:
  Tensor() = default;
  Tensor(const TensorShape& shape, Array array) : shape_(shape), array_(std::move(array)) {
    PSK_CHECK_ARGUMENT(array_.empty() || shape_.len() == array_.len());
  }
  static Tensor filled(int dtype, const TensorShape& shape, double val) {
    return Tensor(shape, Array::filled(dtype, shape.len(), val));
  }

  const TensorShape& shape() const { return shape_; }
  Pos len() const { return array_.len(); }
  bool empty() const { return array_.empty(); }
  int dtype() const { return array_.dtype(); }
  const Array& array() const { return array_; }
  Tensor clone() const { return *this; }
  Tensor eval() const { return Tensor(shape_, array_.eval()); }
  std::string str() const { return array_.str(); }
  std::string repr() const {
    std::string r = array_.repr();  // "Array<i>([..])" -> "Tensor<dim, i>([..])"
    std::ostringstream os;
    os << "Tensor<" << shape_.dim << ", " << type_letter(dtype()) << ">" << r.substr(r.find('('));
    return os.str();
  }

  // tensor.hpp:268-279
  double get(const std::array<Pos, kMaxTensorDim>& pos) const {
    TensorSlice sl;
    sl.dim = shape_.dim;
    for (int i = 0; i < shape_.dim; ++i) sl.c[i] = {pos[i], pos[i] + 1, 1};
    Tensor one = get(sl);
    StoreRef st = one.array().store();
    uint32_t bits = 0;
    check(backend().psk_store_first_value(st.get(), &bits));
    return Array::value_of(dtype(), bits);
  }
  Tensor get(const TensorSlice& sl) const {
    sl.validate();
    PSK_CHECK_ARGUMENT(sl.valid(shape_));
    if (sl.len() == 0) return Tensor();
    psk_view v = sl.view(PSK_VIEW_GATHER, shape_);
    return Tensor(sl.shape(), Array(dtype(), make_view_expr(array_.expr(), v)));
  }

  // tensor.hpp:282-304
  void set(const std::array<Pos, kMaxTensorDim>& pos, double val) {
    TensorSlice sl;
    sl.dim = shape_.dim;
    for (int i = 0; i < shape_.dim; ++i) sl.c[i] = {pos[i], pos[i] + 1, 1};
    set(sl, val);
  }
  void set(const TensorSlice& sl, double val) {
    sl.validate();
    PSK_CHECK_ARGUMENT(sl.valid(shape_));
    if (sl.len() > 0) set(sl, Tensor::filled(dtype(), sl.shape(), val));
  }
  void set(const TensorSlice& sl, const Tensor& other) {
    sl.validate();
    PSK_CHECK_ARGUMENT(sl.valid(shape_));
    PSK_CHECK_ARGUMENT(sl.shape() == other.shape());
    PSK_CHECK_ARGUMENT(other.empty() || other.dtype() == dtype());
    if (sl.len() > 0) {
      Pos c0[kMaxTensorDim], c1[kMaxTensorDim], cs[kMaxTensorDim];
      for (int i = 0; i < sl.dim; ++i) {
        c0[i] = sl.c[i][0];
        c1[i] = sl.c[i][1];
        cs[i] = sl.c[i][2];
      }
      psk_store_t mh = nullptr;
      check(backend().psk_mask_nd(sl.dim, shape_.s.data(), c0, c1, cs, &mh));
      ExprPtr mask = make_store_expr(StoreRef(mh), PSK_BOOL);
      ExprPtr src = make_view_expr(other.array().expr(), sl.view(PSK_VIEW_SCATTER, shape_));
      array_ = Array(dtype(), make_op_expr(PSK_OP_SELECT, dtype(), mask, src, array_.expr()));
    }
  }

 private:
  TensorShape shape_;
  Array array_;
};

}  // namespace psk_host
