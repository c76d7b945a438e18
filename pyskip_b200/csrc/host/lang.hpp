// lang.hpp -- lazy expression DAG and planner of the host layer.
//
// Mirrors the observable behaviour of include/pyskip/detail/lang.hpp (Expr DAG of
// STORE | SLICE | MERGE_1..3, lang.hpp:98-105; materialize/evaluate, :860-905) with a fresh
// design for the GPU backend:
//   * a MERGE node carries an op-code (the reference's MergeFn is an opaque host function
//     pointer, lang.hpp:56) and a SLICE node carries a closed-form view descriptor (the
//     reference's cyclic::StepFn is a LUT tree);
//   * lowering walks the expression once, pushing views down to the leaves on the way
//     (what normalize does by graph surgery, lang.hpp:530-600) and emitting the postfix
//     program + source list that build_plan produces (lang.hpp:616-660);
//   * 1-run stores (broadcast scalars, macros.hpp:16-23) become immediates, equal
//     (store, view-chain) leaves share one source, and sub-expressions with several
//     consumers are materialised once (the reference re-expands them into a tree,
//     lang.hpp:622-625) -- none of which is observable in the results;
//   * when a step would need more than kMaxWidth sources, deeper sub-expressions are
//     materialised first with BITWISE equality, like the reference's schedule()
//     (lang.hpp:466-527; its limit is 16, execute_plan accepts up to 32).
#pragma once

#include <cstring>
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>

#include "backend.hpp"

namespace psk_host {

constexpr int kMaxWidth = 30;    // sources per launch (PSK_MAX_SOURCES = 32)
constexpr int kDefaultWidth = 8;   // measured optimum on the 27-tap voxel convolution (profiles/r1_summary.md)
constexpr int kMaxProgram = 160; // nodes per launch  (PSK_MAX_PROGRAM = 192)
constexpr int kMaxDepth = 124;   // stack depth       (kMaxStack = 128 in the kernels, as lang.hpp:753)

struct Expr;
using ExprPtr = std::shared_ptr<Expr>;

struct Expr {
  enum Kind { STORE, VIEW, OP } kind = STORE;
  int dtype = PSK_I32;   // element type of this node's value
  int32_t span = 0;      // output length (ExprData::span, lang.hpp:100)
  int size = 1;          // node count of the subtree as a tree (ExprData::size, lang.hpp:99)
  // STORE
  mutable StoreRef store;      // a constant's device store is only created when something needs it
  bool is_const = false;       // 1-run store: value known on the host
  uint32_t const_bits = 0;
  // VIEW
  psk_view view{};
  // OP
  int op = 0;                  // psk_opcode; PSK_OP_SELECT for arity 3
  int arity = 0;
  ExprPtr deps[3];
};

inline ExprPtr make_store_expr(StoreRef store, int dtype) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::STORE;
  e->dtype = dtype;
  e->span = store.span();
  e->store = std::move(store);
  return e;
}

// A filled array (broadcast scalar, macros.hpp:16-23; Array(span, fill), array.hpp:82-84) is a
// constant of the expression language: it lowers to an immediate of the fused program, so its
// 1-run device store is created lazily (expr_store) -- most scalars never need one.
inline ExprPtr make_fill_expr(int dtype, int32_t span, uint32_t bits) {
  PSK_CHECK_ARGUMENT(span > 0);
  auto e = std::make_shared<Expr>();
  e->kind = Expr::STORE;
  e->dtype = dtype;
  e->span = span;
  e->is_const = true;
  e->const_bits = bits;
  return e;
}

// the device store of a STORE node
inline const StoreRef& expr_store(const Expr* e) {
  if (!e->store && e->is_const) {
    psk_store_t h = nullptr;
    check(backend().psk_store_fill((psk_dtype)e->dtype, e->span, e->const_bits, &h));
    e->store = StoreRef(h);
  }
  return e->store;
}

inline int32_t view_out_span(int32_t span_in, const psk_view& v) {
  int32_t out = 0;
  check(backend().psk_view_span(span_in, 1, &v, &out));
  return out;
}

// lang::slice (lang.hpp:158-167)
inline ExprPtr make_view_expr(ExprPtr in, const psk_view& v) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::VIEW;
  e->dtype = in->dtype;
  e->span = view_out_span(in->span, v);
  e->size = 1 + in->size;
  e->view = v;
  e->deps[0] = std::move(in);
  return e;
}

// lang::merge (lang.hpp:179-242); span equality is CHECK_ARGUMENT there (:198, :223-224)
inline ExprPtr make_op_expr(int op, int out_dtype, ExprPtr a, ExprPtr b = nullptr, ExprPtr c = nullptr) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::OP;
  e->op = op;
  e->dtype = out_dtype;
  e->span = a->span;
  e->size = 1 + a->size;
  e->arity = 1;
  if (b) {
    PSK_CHECK_ARGUMENT(a->span == b->span);
    e->size += b->size;
    e->arity = 2;
  }
  if (c) {
    PSK_CHECK_ARGUMENT(a->span == c->span);
    e->size += c->size;
    e->arity = 3;
  }
  e->deps[0] = std::move(a);
  e->deps[1] = std::move(b);
  e->deps[2] = std::move(c);
  return e;
}

// ---------------------------------------------------------------------------------------
class Planner {
 public:
  // lang::materialize (lang.hpp:860-900): final step under `eq`, intermediates BITWISE.
  StoreRef materialize(const ExprPtr& root, psk_eq eq) {
    count_parents(root);
    return run(root, eq, /*is_root=*/true);
  }

 private:
  struct ViewChain {
    std::vector<psk_view> v;  // innermost first (order of psk_source::views)
  };

  struct Step {
    std::vector<psk_source> sources;
    std::vector<const Expr*> source_nodes;
    std::vector<ViewChain> source_chains;
    std::vector<StoreRef> keep_alive;
    std::vector<psk_node> prog;
    int depth = 0, max_depth = 0;
  };

  std::unordered_map<const Expr*, int> parents_;
  std::unordered_map<const Expr*, ExprPtr> replaced_;  // sub-expressions materialised on the way

  void count_parents(const ExprPtr& root) {
    std::vector<const Expr*> stack{root.get()};
    std::unordered_map<const Expr*, bool> seen;
    while (!stack.empty()) {
      const Expr* e = stack.back();
      stack.pop_back();
      for (int i = 0; i < 3; ++i) {
        const Expr* d = e->deps[i].get();
        if (!d) continue;
        parents_[d] += 1;
        if (!seen[d]) {
          seen[d] = true;
          stack.push_back(d);
        }
      }
    }
  }

  const Expr* resolve(const Expr* e) const {
    auto it = replaced_.find(e);
    return it == replaced_.end() ? e : it->second.get();
  }

  // number of leaf sources the subtree would contribute (after replacements)
  int width(const Expr* e0) const {
    const Expr* e = resolve(e0);
    if (e->kind == Expr::STORE) return e->is_const ? 0 : 1;
    int w = 0;
    for (int i = 0; i < 3; ++i)
      if (e->deps[i]) w += width(e->deps[i].get());
    return w;
  }
  int prog_len(const Expr* e0) const {
    const Expr* e = resolve(e0);
    if (e->kind == Expr::STORE) return 1;
    int n = e->kind == Expr::OP ? 1 : 0;
    for (int i = 0; i < 3; ++i)
      if (e->deps[i]) n += prog_len(e->deps[i].get());
    return n;
  }
  int depth(const Expr* e0) const {  // operand-stack depth needed to evaluate the subtree
    const Expr* e = resolve(e0);
    if (e->kind == Expr::STORE) return 1;
    if (e->kind == Expr::VIEW) return depth(e->deps[0].get());
    int d = 0;
    for (int i = 0; i < e->arity; ++i) d = std::max(d, i + depth(e->deps[i].get()));
    return d;
  }
  int chain_len(const Expr* e0) const {  // longest run of nested views
    const Expr* e = resolve(e0);
    if (e->kind == Expr::STORE) return 0;
    if (e->kind == Expr::VIEW) return 1 + chain_len(e->deps[0].get());
    int d = 0;
    for (int i = 0; i < e->arity; ++i) d = std::max(d, chain_len(e->deps[i].get()));
    return d;
  }

  void replace_with_store(const Expr* e, StoreRef st) {
    ExprPtr s = make_store_expr(st, resolve(e)->dtype);
    // (a queued result is not waited for just to learn whether it is a constant)
    if (!backend().psk_store_pending(st.get()) && st.nruns() == 1) {
      uint32_t bits = 0;
      check(backend().psk_store_first_value(st.get(), &bits));
      s->is_const = true;
      s->const_bits = bits;
    }
    replaced_[e] = s;
  }

  std::unordered_map<const Expr*, bool> feasible_;
  // sources per launch: the merge step costs O(k) per run, so a wide expression is cheaper as
  // a short chain of narrower launches (config key "gpu_eval_max_sources", 2..30)
  int max_width_ = (int)std::min<int64_t>(kMaxWidth, std::max<int64_t>(2, config_get_int("gpu_eval_max_sources", kDefaultWidth)));

  // Materialise sub-expressions until `e` fits one launch.
  void make_feasible(const Expr* e0) {
    const Expr* e = resolve(e0);
    if (e->kind == Expr::STORE) return;
    if (feasible_[e]) return;
    feasible_[e] = true;
    for (int i = 0; i < 3; ++i)
      if (e->deps[i]) make_feasible(e->deps[i].get());
    // shared sub-expression: evaluate once (only when it is real work)
    for (int i = 0; i < 3; ++i) {
      const Expr* d = e->deps[i].get();
      if (!d) continue;
      const Expr* r = resolve(d);
      if (r->kind != Expr::STORE && parents_[d] > 1 && width(d) >= 2) replace_with_store(d, run_sub(d));
    }
    while (width(e) > max_width_ || prog_len(e) > kMaxProgram || depth(e) > kMaxDepth ||
           chain_len(e) > PSK_MAX_VIEWS) {
      // materialise the heaviest child that is not already a store
      int best = -1, best_w = -1;
      for (int i = 0; i < 3; ++i) {
        const Expr* d = e->deps[i].get();
        if (!d || resolve(d)->kind == Expr::STORE) continue;
        int w = width(d) * 1000 + prog_len(d) + chain_len(d);
        if (w > best_w) {
          best_w = w;
          best = i;
        }
      }
      PSK_CHECK_STATE(best >= 0);
      replace_with_store(e->deps[best].get(), run_sub(e->deps[best].get()));
    }
  }

  // intermediates are queued (psk_eval_async): their consumers read the run count on the device,
  // so a cut expression costs one host round trip in total, not one per step
  StoreRef run_sub(const Expr* e) { return run_node(e, PSK_EQ_BITWISE, /*queued=*/true); }

  StoreRef run(const ExprPtr& root, psk_eq eq, bool) {
    make_feasible(root.get());
    return run_node(root.get(), eq, /*queued=*/false);
  }

  void lower(const Expr* e0, ViewChain chain, Step& st) {
    const Expr* e = resolve(e0);
    switch (e->kind) {
      case Expr::STORE: {
        if (e->is_const) {
          // any view of a constant array is the same constant
          st.prog.push_back(psk_node{PSK_OP_CONST, e->const_bits});
        } else {
          int idx = -1;
          for (size_t i = 0; i < st.source_nodes.size(); ++i) {
            if (st.source_nodes[i]->store.get() == e->store.get() && same_chain(st.source_chains[i], chain)) {
              idx = (int)i;
              break;
            }
          }
          if (idx < 0) {
            psk_source s;
            std::memset(&s, 0, sizeof(s));
            s.store = e->store.get();
            s.nviews = (int32_t)chain.v.size();
            PSK_CHECK_STATE(s.nviews <= PSK_MAX_VIEWS);
            for (size_t i = 0; i < chain.v.size(); ++i) s.views[i] = chain.v[i];
            idx = (int)st.sources.size();
            st.sources.push_back(s);
            st.source_nodes.push_back(e);
            st.source_chains.push_back(chain);
            st.keep_alive.push_back(e->store);
          }
          st.prog.push_back(psk_node{PSK_OP_SRC, (uint32_t)idx});
        }
        st.depth += 1;
        st.max_depth = std::max(st.max_depth, st.depth);
        break;
      }
      case Expr::VIEW: {
        // a view applied on top of the views already pending above it: the pending ones are
        // applied after this one (lang::slice_compose, lang.hpp:50-54)
        ViewChain inner;
        inner.v.push_back(e->view);
        inner.v.insert(inner.v.end(), chain.v.begin(), chain.v.end());
        lower(e->deps[0].get(), inner, st);
        break;
      }
      case Expr::OP: {
        for (int i = 0; i < e->arity; ++i) lower(e->deps[i].get(), chain, st);
        if (e->op != PSK_OP_SRC) st.prog.push_back(psk_node{e->op, 0});  // PSK_OP_SRC marks a no-op (unary +)
        st.depth -= e->arity - 1;
        break;
      }
    }
  }

  static bool same_chain(const ViewChain& a, const ViewChain& b) {
    if (a.v.size() != b.v.size()) return false;
    for (size_t i = 0; i < a.v.size(); ++i)
      if (std::memcmp(&a.v[i], &b.v[i], sizeof(psk_view)) != 0) return false;
    return true;
  }

  StoreRef run_node(const Expr* e0, psk_eq eq, bool queued) {
    const Expr* e = resolve(e0);
    Step st;
    lower(e, ViewChain{}, st);
    if (st.sources.empty()) {
      // every leaf is a constant: keep one real 1-run source so that the kernel has a span
      psk_store_t h = nullptr;
      check(backend().psk_store_fill((psk_dtype)PSK_I32, e->span, 0, &h));
      StoreRef tmp(h);
      psk_source s;
      std::memset(&s, 0, sizeof(s));
      s.store = tmp.get();
      st.sources.push_back(s);
      st.keep_alive.push_back(tmp);
    }
    psk_store_t out = nullptr;
    check((queued ? backend().psk_eval_async : backend().psk_eval)(
        (int32_t)st.sources.size(), st.sources.data(), (int32_t)st.prog.size(), st.prog.data(), (psk_dtype)e->dtype,
        eq, &out));
    return StoreRef(out);
  }
};

// lang::materialize with the element type's own equality for the final step (lang.hpp:899)
inline StoreRef materialize_typed(const ExprPtr& e) {
  // A bare store still takes the final pass in the reference (normalize wraps it in a slice,
  // lang.hpp:532-540).  For int/bool/char both equalities coincide and a store is already
  // canonical, so only multi-run float stores need it (-0/+0, NaN: SURVEY.md A.2).
  if (e->kind == Expr::STORE && (e->is_const || e->dtype != PSK_F32)) return expr_store(e.get());
  Planner p;
  return p.materialize(e, PSK_EQ_TYPED);
}

// lang::evaluate (lang.hpp:902-905): Box-typed (BITWISE) materialisation
inline StoreRef materialize_bitwise(const ExprPtr& e) {
  if (e->kind == Expr::STORE && (e->is_const || e->dtype != PSK_F32)) return expr_store(e.get());
  Planner p;
  return p.materialize(e, PSK_EQ_BITWISE);
}

}  // namespace psk_host
