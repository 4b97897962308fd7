// Shared declarations of the dense (MinCut / DiffPool) path.
#pragma once
#include <limits.h>

#include "common.cuh"
