/* tests/emu stub of CUDA's math_constants.h (host emulation of the solve kernel; test infrastructure only) */
#pragma once
#include <limits>
#define CUDART_INF (std::numeric_limits<double>::infinity())
