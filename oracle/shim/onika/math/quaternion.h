#pragma once
#include <onika/math/basic_types.h>
