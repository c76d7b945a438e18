// backend.hpp -- run-time binding of the C ABI (include/pskb200.h).
//
// The host layer has no link-time CUDA dependency: it dlopens the library the Python
// package points it at (pyskip_b200/libpskb200.so) and calls it through this table.
#pragma once

#include <dlfcn.h>

#include <memory>
#include <stdexcept>
#include <string>

#include "pskb200.h"

namespace psk_host {

struct Backend {
  void* handle = nullptr;
  std::string path;
#define PSK_FN(name) decltype(&::name) name = nullptr;
  PSK_FN(psk_last_error)
  PSK_FN(psk_is_simulator)
  PSK_FN(psk_init)
  PSK_FN(psk_synchronize)
  PSK_FN(psk_store_from_runs)
  PSK_FN(psk_store_fill)
  PSK_FN(psk_store_from_dense)
  PSK_FN(psk_store_to_dense)
  PSK_FN(psk_store_runs)
  PSK_FN(psk_store_nruns)
  PSK_FN(psk_store_span)
  PSK_FN(psk_store_dtype)
  PSK_FN(psk_store_first_value)
  PSK_FN(psk_store_retain)
  PSK_FN(psk_store_release)
  PSK_FN(psk_eval)
  PSK_FN(psk_view_span)
  PSK_FN(psk_mask_nd)
  PSK_FN(psk_reduce)
  PSK_FN(psk_conv_supported)
  PSK_FN(psk_conv_nd)
  PSK_FN(psk_eval_async)
  PSK_FN(psk_store_pending)
  PSK_FN(psk_reduce_segments)
#undef PSK_FN
};

inline Backend& backend() {
  static Backend b;
  return b;
}

inline void bind_backend(const std::string& path, bool rebind) {
  Backend& b = backend();
  if (b.handle && !rebind) {
    if (b.path == path) return;
    throw std::logic_error("pyskip_b200 backend already bound to " + b.path);
  }
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) throw std::runtime_error(std::string("cannot load pskb200 library: ") + dlerror());
  Backend nb;
  nb.handle = h;
  nb.path = path;
#define PSK_FN(name)                                                                  \
  nb.name = reinterpret_cast<decltype(nb.name)>(dlsym(h, #name));                     \
  if (!nb.name) throw std::runtime_error(std::string("pskb200 library lacks ") + #name);
  PSK_FN(psk_last_error)
  PSK_FN(psk_is_simulator)
  PSK_FN(psk_init)
  PSK_FN(psk_synchronize)
  PSK_FN(psk_store_from_runs)
  PSK_FN(psk_store_fill)
  PSK_FN(psk_store_from_dense)
  PSK_FN(psk_store_to_dense)
  PSK_FN(psk_store_runs)
  PSK_FN(psk_store_nruns)
  PSK_FN(psk_store_span)
  PSK_FN(psk_store_dtype)
  PSK_FN(psk_store_first_value)
  PSK_FN(psk_store_retain)
  PSK_FN(psk_store_release)
  PSK_FN(psk_eval)
  PSK_FN(psk_view_span)
  PSK_FN(psk_mask_nd)
  PSK_FN(psk_reduce)
  PSK_FN(psk_conv_supported)
  PSK_FN(psk_conv_nd)
  PSK_FN(psk_eval_async)
  PSK_FN(psk_store_pending)
  PSK_FN(psk_reduce_segments)
#undef PSK_FN
  b = nb;
}

// status -> the exception the reference would throw (detail/errors.hpp:7-21)
inline void check(psk_status st) {
  if (st == PSK_OK) return;
  std::string msg = backend().psk_last_error ? backend().psk_last_error() : "pskb200 backend not bound";
  if (st == PSK_EINVAL) throw std::invalid_argument(msg);
  if (st == PSK_ESTATE) throw std::logic_error(msg);
  throw std::runtime_error(msg);
}

#define PSK_CHECK_ARGUMENT(cond)                                                       \
  do {                                                                                 \
    if (!(cond)) throw std::invalid_argument(std::string("CHECK_ARGUMENT failed: ") + #cond); \
  } while (0)
#define PSK_CHECK_STATE(cond)                                                          \
  do {                                                                                 \
    if (!(cond)) throw std::logic_error(std::string("CHECK_STATE failed: ") + #cond);  \
  } while (0)

// Owning handle of a device store.
class StoreRef {
 public:
  StoreRef() = default;
  explicit StoreRef(psk_store_t h) : h_(h, [](psk_store_s* p) {
                                        if (p && backend().psk_store_release) backend().psk_store_release(p);
                                      }) {}
  psk_store_t get() const { return h_.get(); }
  explicit operator bool() const { return !!h_; }
  int64_t nruns() const { return backend().psk_store_nruns(get()); }
  int32_t span() const { return backend().psk_store_span(get()); }
  int dtype() const { return backend().psk_store_dtype(get()); }

 private:
  std::shared_ptr<psk_store_s> h_;
};

}  // namespace psk_host
