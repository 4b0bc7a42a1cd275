// data-adaptive solve_sim translation unit (see abi_dalton_solve.cu)
#define RODEO_ONLY_SIM
#include "abi_dalton_solve.cu"
