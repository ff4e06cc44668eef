#pragma once
#include <vector>
namespace onika { namespace memory { template <class T> using CudaMMVector = std::vector<T>;
template <class T> struct MemoryUsage { static inline size_t memory_bytes(const T&) { return sizeof(T); } };
template <class... T> inline size_t memory_bytes(const T&...) { return 0; } } }
