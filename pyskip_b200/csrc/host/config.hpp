// config.hpp -- process-wide configuration map (include/pyskip/detail/config.hpp:13-72):
// string key -> string | int64 | double | bool behind a mutex.  The reference reads four
// keys; on the GPU backend
//   flush_tree_size_threshold  keeps its meaning (array.hpp:210, tensor.hpp:330);
//   accelerated_eval, parallelize_threshold, parallelize_parts are accepted and stored but
//   have no effect: tournament-tree vs linear min and the CPU thread partition
//   (eval.hpp:621, :634-637) have no counterpart -- results never depended on them.
#pragma once

#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <variant>

namespace psk_host {

using ConfigValue = std::variant<std::string, int64_t, double, bool>;
using ConfigMap = std::unordered_map<std::string, ConfigValue>;

struct GlobalConfig {
  std::mutex mu;
  ConfigMap map;
  static GlobalConfig& get() {
    static GlobalConfig g;
    return g;
  }
};

template <typename T>
inline void config_set(const std::string& key, T val) {
  auto& g = GlobalConfig::get();
  std::lock_guard<std::mutex> l(g.mu);
  g.map[key] = ConfigValue(std::move(val));
}

inline void config_clear(const std::string& key) {
  auto& g = GlobalConfig::get();
  std::lock_guard<std::mutex> l(g.mu);
  g.map.erase(key);
}

inline ConfigMap config_get_all() {
  auto& g = GlobalConfig::get();
  std::lock_guard<std::mutex> l(g.mu);
  return g.map;
}

inline void config_set_all(const ConfigMap& m) {
  auto& g = GlobalConfig::get();
  std::lock_guard<std::mutex> l(g.mu);
  g.map = m;
}

inline int64_t config_get_int(const std::string& key, int64_t dflt) {
  auto& g = GlobalConfig::get();
  std::lock_guard<std::mutex> l(g.mu);
  auto it = g.map.find(key);
  if (it == g.map.end()) return dflt;
  if (auto p = std::get_if<int64_t>(&it->second)) return *p;
  return dflt;
}

}  // namespace psk_host
