// float32 instantiation of abi_solve_sqrt.cu (same source, RODEO_REAL = float): float32 factors, double means
#define RODEO_REAL float
#define RODEO_SUFFIX _f32
#include "abi_solve_sqrt.cu"
