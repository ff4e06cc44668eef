#pragma once
#include <onika/math/basic_types.h>
#define ONIKA_CU_ATOMIC_ADD(a, b) ((a) += (b))
