// ni_step_lean_f32.cu -- explicit instantiations of the specialised step kernels (ni_step_lean.cuh) for fp32 state (fp32 / fp16 / bf16 model outputs); one translation unit per
// storage type so that nvcc compiles them in parallel.
#include "ni_step_lean.cuh"

namespace ni {
template int launch_step_lean<float, float>(const NiStepDesc *, const void *, bool, cudaStream_t, bool *);
template int launch_step_lean<float, __half>(const NiStepDesc *, const void *, bool, cudaStream_t, bool *);
template int launch_step_lean<float, __nv_bfloat16>(const NiStepDesc *, const void *, bool, cudaStream_t, bool *);
} // namespace ni
