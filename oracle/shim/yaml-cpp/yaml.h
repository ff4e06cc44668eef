#pragma once
#include <string>
namespace YAML {
struct Node { bool IsMap() const { return false; } Node operator[](const char*) const { return {}; } Node operator[](const std::string&) const { return {}; } explicit operator bool() const { return false; }
  bool operator!() const { return true; } template <class T> T as() const { return T{}; } };
template <class T> struct convert;
}
