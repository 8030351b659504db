// ob_kern_large.cu — the grid-wide path for one large world (ob_large_kernels.cuh)
#include "ob_backend_cuda.h"
#include "ob_large_kernels.cuh"
