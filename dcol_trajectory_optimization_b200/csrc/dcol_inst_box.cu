/* pair kernels whose first primitive is of class CLS_BOX (see dcol_classes.cuh) */
#include "dcol_kernels.cuh"

DCOL_DEFINE_FIRST_CLASS(CLS_BOX)
