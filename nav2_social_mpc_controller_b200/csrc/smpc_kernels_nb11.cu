// one translation unit per number of parameter blocks (parallel build)
#define SMPC_NB 11
#include "smpc_kernels_nb.inc"
