// step_cdrq.cu — the rlr_step_kernel instantiations of one policy family (see rlr_step.cuh).
#include "rlr_step.cuh"

RLR_DEFINE_STEP_PICK(cdrq, RLR_POLICY_CDRQ)
