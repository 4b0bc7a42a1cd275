// float32 instantiation of abi_fenrir.cu (same source, RODEO_REAL = float)
#define RODEO_REAL float
#define RODEO_SUFFIX _f32
#include "abi_fenrir.cu"
