// ni_step_lean_bf16.cu -- explicit instantiations of the specialised step kernels (ni_step_lean.cuh) for bf16 state; one translation unit per
// storage type so that nvcc compiles them in parallel.
#include "ni_step_lean.cuh"

namespace ni {
template int launch_step_lean<__nv_bfloat16, __nv_bfloat16>(const NiStepDesc *, const void *, bool, cudaStream_t, bool *);
} // namespace ni
