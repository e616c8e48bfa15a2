// Translation unit of the PhaseLift kernel (see the note at the end of phaselift.cuh).
#define TWOACE_PL_KERNEL_TU
#include "phaselift.cuh"
