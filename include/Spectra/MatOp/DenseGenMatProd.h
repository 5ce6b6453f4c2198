// B200 shim of Spectra/MatOp/DenseGenMatProd.h: see DenseSymMatProd.h (both dense wrappers live there).
#ifndef SPECTRA_B200_DENSE_GEN_MAT_PROD_H
#define SPECTRA_B200_DENSE_GEN_MAT_PROD_H
#include "DenseSymMatProd.h"
#endif
