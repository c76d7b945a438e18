// psk_ops.h -- scalar semantics of the fused element-wise program, on 32-bit payloads.
//
// Restates the operator catalogue of include/pyskip/array.hpp:276-338 (lambdas wrapped by
// lang::merge_fn, detail/lang.hpp:65-95) as the reference's x86-64 build behaves
// (SURVEY.md A.5): int32 + - * wrap, / and % are C truncating, shifts mask the count to
// 5 bits, float ops are IEEE binary32 round-to-nearest WITHOUT contraction (this file is
// compiled with -fmad=false), std::min/max are the (b<a)?b:a / (a<b)?b:a selects,
// float->int casts follow cvttss2si (0x80000000 when out of range).
#pragma once

#include "../../include/pskb200.h"
#include "psk_platform.h"

#if !defined(PSK_CPU_SIM) && !defined(__CUDACC_RTC__)
#include <math.h>
#endif

namespace psk {

PSK_HD float u2f(uint32_t u) {
#ifdef PSK_CPU_SIM
  float f;
  memcpy(&f, &u, 4);
  return f;
#else
  union { uint32_t u; float f; } c;
  c.u = u;
  return c.f;
#endif
}
PSK_HD uint32_t f2u(float f) {
#ifdef PSK_CPU_SIM
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#else
  union { uint32_t u; float f; } c;
  c.f = f;
  return c.u;
#endif
}

PSK_HD uint32_t f64_to_i32(double x) {  // cvttsd2si
  if (!(x > -2147483649.0 && x < 2147483648.0)) return 0x80000000u;
  return (uint32_t)(int32_t)x;
}
PSK_HD uint32_t f32_to_i32(float x) {  // cvttss2si
  if (!(x >= -2147483648.0f && x < 2147483648.0f)) return 0x80000000u;
  return (uint32_t)(int32_t)x;
}

PSK_HD uint32_t idiv(uint32_t a, uint32_t b) {
  int32_t x = (int32_t)a, y = (int32_t)b;
  if (y == 0) return 0;                      // traps on the CPU; unspecified here
  if (y == -1) return (uint32_t)0 - a;       // INT_MIN / -1 traps on the CPU
  return (uint32_t)(x / y);
}
PSK_HD uint32_t imod(uint32_t a, uint32_t b) {
  int32_t x = (int32_t)a, y = (int32_t)b;
  if (y == 0 || y == -1) return 0;
  return (uint32_t)(x % y);
}

// static_cast<int>(std::pow(int,int)) goes through double pow (array.hpp:317-319); glibc
// returns the exact value whenever it is representable, so integer powers are formed by
// exact repeated multiplication while they fit in 2^53.
PSK_HD uint32_t ipow(uint32_t a, uint32_t b) {
  int32_t base = (int32_t)a, ex = (int32_t)b;
  if (ex >= 0) {
    double r = 1.0, p = (double)base;
    int32_t e = ex;
    bool exact = true;
    while (e) {
      if (e & 1) { r *= p; if (r > 9007199254740992.0 || r < -9007199254740992.0) exact = false; }
      e >>= 1;
      if (e) { p *= p; if (p > 9007199254740992.0) exact = false; }
    }
    if (!exact) r = pow((double)base, (double)ex);
    return f64_to_i32(r);
  }
  return f64_to_i32(pow((double)base, (double)ex));
}

// Apply one node.  st points at the stack top (one past the last element).
PSK_HD uint32_t apply_unary(int op, uint32_t a) {
  switch (op) {
    case PSK_OP_NEG_I: return (uint32_t)0 - a;
    case PSK_OP_ABS_I: return ((int32_t)a < 0) ? (uint32_t)0 - a : a;
    case PSK_OP_BNOT_I: return ~a;
    case PSK_OP_SQRT_I: return f64_to_i32(sqrt((double)(int32_t)a));
    case PSK_OP_EXP_I: return f64_to_i32(exp((double)(int32_t)a));
    case PSK_OP_LNOT_I: return a == 0 ? 1u : 0u;
    case PSK_OP_NEG_F: return a ^ 0x80000000u;
    case PSK_OP_ABS_F: return a & 0x7fffffffu;
    case PSK_OP_SQRT_F: return f2u(sqrtf(u2f(a)));
    case PSK_OP_EXP_F: return f2u(expf(u2f(a)));
    case PSK_OP_LNOT_F: return u2f(a) == 0.0f ? 1u : 0u;
    case PSK_OP_I2F: return f2u((float)(int32_t)a);
    case PSK_OP_F2I: return f32_to_i32(u2f(a));
    case PSK_OP_I2B: return a != 0 ? 1u : 0u;
    case PSK_OP_F2B: return u2f(a) != 0.0f ? 1u : 0u;
    case PSK_OP_I2C: return (uint32_t)(int32_t)(int8_t)(a & 0xff);
    case PSK_OP_F2C: return (uint32_t)(int32_t)(int8_t)(f32_to_i32(u2f(a)) & 0xff);
    default: return a;
  }
}

PSK_HD uint32_t apply_binary(int op, uint32_t a, uint32_t b) {
  switch (op) {
    case PSK_OP_ADD_I: return a + b;
    case PSK_OP_SUB_I: return a - b;
    case PSK_OP_MUL_I: return a * b;
    case PSK_OP_DIV_I: return idiv(a, b);
    case PSK_OP_MOD_I: return imod(a, b);
    case PSK_OP_AND_I: return a & b;
    case PSK_OP_OR_I: return a | b;
    case PSK_OP_XOR_I: return a ^ b;
    case PSK_OP_SHL_I: return a << (b & 31);
    case PSK_OP_SHR_I: return (uint32_t)((int32_t)a >> (b & 31));
    case PSK_OP_MIN_I: return ((int32_t)b < (int32_t)a) ? b : a;
    case PSK_OP_MAX_I: return ((int32_t)a < (int32_t)b) ? b : a;
    case PSK_OP_POW_I: return ipow(a, b);
    case PSK_OP_COALESCE_I: return a ? a : b;
    case PSK_OP_EQ_I: return a == b;
    case PSK_OP_NE_I: return a != b;
    case PSK_OP_LT_I: return (int32_t)a < (int32_t)b;
    case PSK_OP_GT_I: return (int32_t)a > (int32_t)b;
    case PSK_OP_LE_I: return (int32_t)a <= (int32_t)b;
    case PSK_OP_GE_I: return (int32_t)a >= (int32_t)b;
    case PSK_OP_LAND_I: return (a != 0) && (b != 0);
    case PSK_OP_LOR_I: return (a != 0) || (b != 0);
    case PSK_OP_ADD_F: return f2u(u2f(a) + u2f(b));
    case PSK_OP_SUB_F: return f2u(u2f(a) - u2f(b));
    case PSK_OP_MUL_F: return f2u(u2f(a) * u2f(b));
    case PSK_OP_DIV_F: return f2u(u2f(a) / u2f(b));
    case PSK_OP_MOD_F: return f2u(fmodf(u2f(a), u2f(b)));
    case PSK_OP_MIN_F: return (u2f(b) < u2f(a)) ? b : a;
    case PSK_OP_MAX_F: return (u2f(a) < u2f(b)) ? b : a;
    case PSK_OP_POW_F: return f2u(powf(u2f(a), u2f(b)));
    case PSK_OP_COALESCE_F: return (u2f(a) != 0.0f) ? a : b;
    case PSK_OP_EQ_F: return u2f(a) == u2f(b);
    case PSK_OP_NE_F: return u2f(a) != u2f(b);
    case PSK_OP_LT_F: return u2f(a) < u2f(b);
    case PSK_OP_GT_F: return u2f(a) > u2f(b);
    case PSK_OP_LE_F: return u2f(a) <= u2f(b);
    case PSK_OP_GE_F: return u2f(a) >= u2f(b);
    case PSK_OP_LAND_F: return (u2f(a) != 0.0f) && (u2f(b) != 0.0f);
    case PSK_OP_LOR_F: return (u2f(a) != 0.0f) || (u2f(b) != 0.0f);
    default: return a;
  }
}

PSK_HD int op_arity(int op) {
  if (op == PSK_OP_SRC || op == PSK_OP_CONST) return 0;
  if (op == PSK_OP_SELECT) return 3;
  switch (op) {
    case PSK_OP_NEG_I: case PSK_OP_ABS_I: case PSK_OP_BNOT_I: case PSK_OP_SQRT_I:
    case PSK_OP_EXP_I: case PSK_OP_LNOT_I: case PSK_OP_NEG_F: case PSK_OP_ABS_F:
    case PSK_OP_SQRT_F: case PSK_OP_EXP_F: case PSK_OP_LNOT_F: case PSK_OP_I2F:
    case PSK_OP_F2I: case PSK_OP_I2B: case PSK_OP_F2B: case PSK_OP_I2C: case PSK_OP_F2C:
      return 1;
    default:
      return 2;
  }
}

PSK_HD bool op_known(int op) {
  return op == PSK_OP_SRC || op == PSK_OP_CONST || (op >= PSK_OP_ADD_I && op <= PSK_OP_LOR_I) ||
         (op >= PSK_OP_ADD_F && op <= PSK_OP_LOR_F) || op == PSK_OP_SELECT ||
         (op >= PSK_OP_I2F && op <= PSK_OP_F2C);
}

constexpr int kMaxStack = 128;  // lang::execute_plan_fixed's stack (lang.hpp:753)

// The stack machine of lang::execute_plan_fixed (detail/lang.hpp:754-777) with op-codes in
// place of function pointers.  `val` holds the current value of every source.
template <typename ValArray>
PSK_HD uint32_t run_program(const psk_node* prog, int nprog, const ValArray& val) {
  uint32_t st[kMaxStack];
  int sp = 0;
  for (int i = 0; i < nprog; ++i) {
    const int op = prog[i].op;
    if (op == PSK_OP_SRC) {
      st[sp++] = val[prog[i].arg];
    } else if (op == PSK_OP_CONST) {
      st[sp++] = prog[i].arg;
    } else if (op == PSK_OP_SELECT) {
      sp -= 2;
      st[sp - 1] = st[sp - 1] ? st[sp] : st[sp + 1];
    } else if (op_arity(op) == 1) {
      st[sp - 1] = apply_unary(op, st[sp - 1]);
    } else {
      sp -= 1;
      st[sp - 1] = apply_binary(op, st[sp - 1], st[sp]);
    }
  }
  return st[sp - 1];
}

PSK_HD bool payload_equal(uint32_t a, uint32_t b, bool typed_float) {
  // bitwise: Box::operator== (detail/box.hpp:84-90); typed float: operator== on float
  return typed_float ? (u2f(a) == u2f(b)) : (a == b);
}

}  // namespace psk
