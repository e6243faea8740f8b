/* pair kernels whose first primitive is of class CLS_ELLIPSOID (see dcol_classes.cuh) */
#include "dcol_kernels.cuh"

DCOL_DEFINE_FIRST_CLASS(CLS_ELLIPSOID)
