// float32 instantiation of abi_dalton.cu (same source, RODEO_REAL = float)
#define RODEO_REAL float
#define RODEO_SUFFIX _f32
#include "abi_dalton.cu"
