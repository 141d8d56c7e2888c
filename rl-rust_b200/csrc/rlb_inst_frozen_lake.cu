// rlb_inst_frozen_lake.cu — kernel instantiations for RLB_ENV_FROZEN_LAKE (see rlb_launch.h).
#include "rlb_launch_impl.cuh"
namespace rlb { RLB_INSTANTIATE_ENV(RLB_ENV_FROZEN_LAKE) }
