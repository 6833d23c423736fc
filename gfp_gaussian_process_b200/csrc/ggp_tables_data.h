// ggp_tables_data.h — one initialiser for GgpMathTables (exp / pow-log / log tables of glibc 2.39
// and the Dawson Chebyshev rows), usable for a host static or a __device__ global.
#pragma once
#include "ggp_libm.cuh"
#include "ggp_dawson_tables.h"
#define GGP_MATH_TABLES_INIT { GGP_EXP_TAB_INIT, GGP_POWLOG_TAB_INIT, GGP_LOG_TAB_INIT, GGP_DAWSON_TAB_INIT }
