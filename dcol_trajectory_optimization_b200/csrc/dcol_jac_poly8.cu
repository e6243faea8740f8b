/* solution-Jacobian pair kernels whose first primitive is of class CLS_POLY8 */
#include "dcol_kernels.cuh"

DCOL_DEFINE_FIRST_CLASS_JAC(CLS_POLY8)
