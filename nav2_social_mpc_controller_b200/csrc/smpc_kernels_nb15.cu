// one translation unit per number of parameter blocks (parallel build)
#define SMPC_NB 15
#include "smpc_kernels_nb.inc"
