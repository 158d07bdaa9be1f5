// step_qroute_f32.cu — the rlr_step_kernel instantiations of Qroute with its table stored in fp32 (see rlr_step.cuh).
#include "rlr_step.cuh"

RLR_DEFINE_STEP_PICK_F32(qroute_f32, RLR_POLICY_QROUTE)
