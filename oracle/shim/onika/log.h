#pragma once
#include <iostream>
namespace onika { struct NullLog { template <class T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } };
inline NullLog lout, ldbg, lerr; }
namespace exanb { using onika::lout; using onika::ldbg; using onika::lerr; }
