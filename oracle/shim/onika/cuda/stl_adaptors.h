#pragma once
#include <vector>
namespace onika { namespace cuda {
template <class T, class A> inline T* vector_data(std::vector<T, A>& v) { return v.data(); }
template <class T, class A> inline const T* vector_data(const std::vector<T, A>& v) { return v.data(); }
template <class T, class A> inline size_t vector_size(const std::vector<T, A>& v) { return v.size(); }
}}
