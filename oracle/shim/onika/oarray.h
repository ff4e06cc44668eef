#pragma once
#include <array>
namespace onika { template <class T, size_t N> using oarray_t = std::array<T, N>; }
