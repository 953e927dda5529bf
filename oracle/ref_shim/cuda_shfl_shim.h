// TEST INFRASTRUCTURE.  Force-included when the reference's own CUDA Poisson backend (poisson_solver/BackendCUDA.cu) is
// compiled for sm_100a as a timing baseline: the pre-Volta warp shuffle it calls no longer exists; its full-mask
// synchronising form is the same operation for the converged warps of those kernels.  Nothing else is touched.
#pragma once
#ifdef __CUDACC__
#define __shfl_xor(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#endif
