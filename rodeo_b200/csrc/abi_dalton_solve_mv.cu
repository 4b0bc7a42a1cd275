// data-adaptive solve_mv translation unit (see abi_dalton_solve.cu)
#define RODEO_ONLY_MV
#include "abi_dalton_solve.cu"
