#pragma once
#include <string>
namespace onika { template <class... A> inline std::string format_string(const char*, A...) { return {}; } }
