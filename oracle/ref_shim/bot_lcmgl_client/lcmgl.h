// stand-in: lcmgl is not used by the files compiled here; libbot's bot_tictoc (a profiling timer pulled in through
// this include chain in the real build) is a no-op
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
static inline int64_t bot_tictoc(const char*) { return 0; }
