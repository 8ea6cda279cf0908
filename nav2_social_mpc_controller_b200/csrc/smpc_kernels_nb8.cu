// one translation unit per number of parameter blocks (parallel build)
#define SMPC_NB 8
#include "smpc_kernels_nb.inc"
