// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
#pragma once
#include <cmath>
#define CUDART_INF_F (__builtin_inff())
#define CUDART_NAN_F (__builtin_nanf(""))
