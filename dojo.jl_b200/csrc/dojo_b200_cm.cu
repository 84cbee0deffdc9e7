// dojo_b200_cm.cu -- second compilation of the step / gradient kernels with the complete model set of the reference enabled
// (macro DJ_ANY_CONTACT, named after its first use):
//   * every contact model: NonlinearContact (contacts/nonlinear.jl) as in dojo_b200.cu, plus ImpactContact
//     (contacts/impact.jl) and LinearContact (contacts/linear.jl) from dojo_contact_orthant.cuh;
//   * translational springs / dampers / limits of the joints from dojo_joint_tra.cuh.
//
// Why a second translation unit instead of a run-time branch inside the one kernel: the Newton loop of dojo_step_kernel is ~20 k
// straight-line instructions at 255 registers and is sensitive to instruction-cache footprint and register allocation
// (profiles/README.md); the benchmarked kernels (all BASELINE models use NonlinearContact) must not change when a contact model is
// added.  The same headers are compiled here under another namespace with DJ_ANY_CONTACT; dojo_create picks this compilation only for
// mechanisms that contain an impact / linear contact or a joint with translational springs / dampers / limits.  The two compilations share the argument block layout (StepArgs, Plan and the
// plan tables are plain data), so the host code in dojo_b200.cu fills one StepArgs and launches whichever kernel the handle holds.
#include <cuda_runtime.h>

#define DJ_ANY_CONTACT 1
#define dj dj_cm
#include "dojo_step_kernel.cuh"
#undef dj

extern "C" __attribute__((visibility("hidden"))) const void* dojo_cm_step_kernel(int grad) {
  return grad ? (const void*)dj_cm::dojo_step_kernel<true> : (const void*)dj_cm::dojo_step_kernel<false>;
}
