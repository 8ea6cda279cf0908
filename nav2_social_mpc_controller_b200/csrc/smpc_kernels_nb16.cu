// one translation unit per number of parameter blocks (parallel build)
#define SMPC_NB 16
#include "smpc_kernels_nb.inc"
