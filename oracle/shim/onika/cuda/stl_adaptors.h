#pragma once
#include <cstring>
#include <span>
#include <vector>
namespace onika { namespace cuda {
template <class T, class A> inline T* vector_data(std::vector<T, A>& v) { return v.data(); }
template <class T, class A> inline const T* vector_data(const std::vector<T, A>& v) { return v.data(); }
template <class T, class A> inline size_t vector_size(const std::vector<T, A>& v) { return v.size(); }
template <class T> using span = std::span<T>;   // rshape_grid.hpp:41-48, shapes.hpp:38 (non-owning view)
template <class T, class A> inline span<T> make_span(std::vector<T, A>& v) { return span<T>(v.data(), v.size()); }
}}
