// Stand-in for MATLAB's "mexAdapter.hpp" (see mex.hpp in this directory).
#pragma once
#include "mex.hpp"
