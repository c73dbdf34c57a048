// TEST INFRASTRUCTURE ONLY. Stand-in for boost::ptr_vector with the members the reference's DSDIFF
// reader uses for its marker list (dsdiff_file_reader.h:43, .cpp: push_back(new T), size(), [],
// sort()): an owning vector of pointers that hands out references.
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>
namespace boost {
template <class T>
class ptr_vector {
public:
    ptr_vector() {}
    ptr_vector(const ptr_vector&) = delete;
    ptr_vector& operator=(const ptr_vector&) = delete;
    ~ptr_vector() { clear(); }
    void push_back(T* p) { v.push_back(p); }
    size_t size() const { return v.size(); }
    bool empty() const { return v.empty(); }
    T& operator[](size_t i) { return *v[i]; }
    const T& operator[](size_t i) const { return *v[i]; }
    void clear() { for (T* p : v) delete p; v.clear(); }
    void sort() { std::sort(v.begin(), v.end(), [](const T* a, const T* b) { return *a < *b; }); }
    template <class C> void sort(C c) { std::sort(v.begin(), v.end(), [&](const T* a, const T* b) { return c(*a, *b); }); }
private:
    std::vector<T*> v;
};
}  // namespace boost
