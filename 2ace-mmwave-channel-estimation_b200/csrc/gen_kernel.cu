// Translation unit of the general stage kernel (see the note at the end of admm_stage.cuh).
#define TWOACE_GEN_KERNEL_TU
#include "admm_stage.cuh"
