// Instantiates the fused step kernel for one portfolio capacity: compile with -DMDG_CAP=1|2|4|8|16.
#include "mdg_step_kernel.cuh"

#ifndef MDG_CAP
#error "compile with -DMDG_CAP=<capacity>"
#endif
#define MDG_CAT2(a, b) a##b
#define MDG_CAT(a, b) MDG_CAT2(a, b)

namespace mdg {
int MDG_CAT(launch_step_cap, MDG_CAP)(const StepArgs& a, bool exact) { return launch_step<MDG_CAP>(a, exact); }
}  // namespace mdg
