// TEST INFRASTRUCTURE ONLY. Minimal stand-in for <boost/circular_buffer.hpp> (Boost is not in
// this image) so the reference's dsd_sample_reader.{h,cpp} and dsd_decimator.cpp compile
// unmodified into oracle/_ref.  Only the members those files use are provided:
// ctor(capacity), default ctor, copy, set_capacity, push_front, operator[] (0 = newest).
#pragma once
#include <cstddef>
#include <vector>

namespace boost {
template <typename T>
class circular_buffer {
public:
    circular_buffer() : head_(0), count_(0) {}
    explicit circular_buffer(std::size_t cap) : slots_(cap), head_(0), count_(0) {}
    void set_capacity(std::size_t cap) { slots_.assign(cap, T()); head_ = 0; count_ = 0; }
    std::size_t capacity() const { return slots_.size(); }
    std::size_t size() const { return count_; }
    void push_front(const T& v) {
        const std::size_t cap = slots_.size();
        if (cap == 0) return;
        head_ = (head_ == 0 ? cap : head_) - 1;
        slots_[head_] = v;
        if (count_ < cap) ++count_;
    }
    // Boost indexes with a compare-and-subtract wrap, not a division; do the same so that the
    // reference timed through this stand-in is not slowed down by it.
    T& operator[](std::size_t i) { return slots_[wrap(head_ + i)]; }
    const T& operator[](std::size_t i) const { return slots_[wrap(head_ + i)]; }
private:
    std::size_t wrap(std::size_t k) const { return k < slots_.size() ? k : k - slots_.size(); }
    std::vector<T> slots_;
    std::size_t head_, count_;
};
}  // namespace boost
