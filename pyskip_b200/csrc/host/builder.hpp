// builder.hpp -- ArrayBuilder: cheap random point / band writes on the host, uploaded once.
//
// Same surface as include/pyskip/builder.hpp:39-152 (set(pos,val), set(band,val),
// set(band,array), build, len, str, repr).  The reference keeps 4096-element blocks of small
// stores edited in place with core::set / core::insert (detail/core.hpp:133-270); here the
// pending runs live in one ordered map (run start -> payload), which gives the same
// O(log n) point update without the block bookkeeping, and build() uploads the canonical
// run list to the device in one copy.
#pragma once

#include <iterator>
#include <map>

#include "array.hpp"

namespace psk_host {

struct Band {  // builder.hpp:20-34
  Pos start, stop;
  Band(Pos start, Pos stop) : start(start), stop(stop) {
    PSK_CHECK_ARGUMENT(0 <= start);
    PSK_CHECK_ARGUMENT(start <= stop);
  }
  Pos len() const { return stop - start; }
};

class ArrayBuilder {
 public:
  ArrayBuilder(int dtype, Pos len, double val) : dtype_(dtype), len_(len) {
    PSK_CHECK_ARGUMENT(len_ > 0);  // builder.hpp:45
    starts_[0] = bits_of(dtype, val);
  }
  explicit ArrayBuilder(const Array& array) : ArrayBuilder(array.dtype(), array.len(), 0) {
    set(Band(0, len_), array);
  }

  Pos len() const { return len_; }
  int dtype() const { return dtype_; }

  void set(Pos pos, double val) {
    PSK_CHECK_ARGUMENT(0 <= pos && pos < len_);  // builder.hpp:87
    assign(pos, pos + 1, bits_of(dtype_, val));
  }
  void set(const Band& band, double val) {
    PSK_CHECK_ARGUMENT(band.stop <= len_);
    if (band.len() > 0) assign(band.start, band.stop, bits_of(dtype_, val));
  }
  void set(const Band& band, const Array& other) {
    PSK_CHECK_ARGUMENT(band.stop <= len_);          // builder.hpp:106-107
    PSK_CHECK_ARGUMENT(band.len() == other.len());
    if (band.len() == 0) return;
    PSK_CHECK_ARGUMENT(other.dtype() == dtype_);
    std::vector<int32_t> ends;
    std::vector<uint8_t> vals;
    other.runs(&ends, &vals);
    Pos prev = 0;
    const int es = elem_size(dtype_);
    for (size_t i = 0; i < ends.size(); ++i) {
      uint32_t bits;
      if (es == 4) std::memcpy(&bits, vals.data() + 4 * i, 4);
      else bits = dtype_ == PSK_BOOL ? (vals[i] ? 1u : 0u) : (uint32_t)(int32_t)(int8_t)vals[i];
      assign(band.start + prev, band.start + ends[i], bits);
      prev = ends[i];
    }
  }

  // builder.hpp:126-133: concatenate with equal-neighbour compaction
  Array build() const {
    std::vector<int32_t> ends;
    std::vector<uint32_t> bits;
    for (auto it = starts_.begin(); it != starts_.end(); ++it) {
      auto nx = std::next(it);
      Pos end = nx == starts_.end() ? len_ : nx->first;
      if (!bits.empty() && bits.back() == it->second) ends.back() = end;
      else {
        ends.push_back(end);
        bits.push_back(it->second);
      }
    }
    if (elem_size(dtype_) == 4) return Array::from_runs(dtype_, (int64_t)ends.size(), ends.data(), bits.data());
    std::vector<uint8_t> narrow(bits.size());
    for (size_t i = 0; i < bits.size(); ++i) narrow[i] = (uint8_t)bits[i];
    return Array::from_runs(dtype_, (int64_t)ends.size(), ends.data(), narrow.data());
  }
  std::string str() const { return build().str(); }
  std::string repr() const {
    std::string r = build().repr();  // "Array<i>([...])" -> "Builder<i>([...])"
    return "Builder" + r.substr(5);
  }

 private:
  // value in force at position p
  uint32_t at(Pos p) const { return std::prev(starts_.upper_bound(p))->second; }

  void assign(Pos start, Pos stop, uint32_t bits) {
    const bool restore = stop < len_;
    const uint32_t after = restore ? at(stop) : 0;
    starts_.erase(starts_.lower_bound(start), starts_.lower_bound(stop));
    starts_[start] = bits;
    if (restore && starts_.find(stop) == starts_.end()) starts_[stop] = after;
  }

  int dtype_;
  Pos len_;
  std::map<Pos, uint32_t> starts_;  // run start -> payload
};

}  // namespace psk_host
