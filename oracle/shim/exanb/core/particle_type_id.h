#pragma once
// stand-in: only the include is needed by shape_reader.hpp / shape_writer.hpp for the parts compiled here
#include <map>
#include <string>
namespace exanb { using ParticleTypeMap = std::map<std::string, long>; }
