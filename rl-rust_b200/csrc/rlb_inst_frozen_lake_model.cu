// rlb_inst_frozen_lake_model.cu — k_run with the Dyna model attached, for RLB_ENV_FROZEN_LAKE (see rlb_launch.h).
#include "rlb_launch_impl.cuh"
namespace rlb { RLB_INSTANTIATE_ENV_MODEL(RLB_ENV_FROZEN_LAKE) }
