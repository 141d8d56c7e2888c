// rlb_inst_cliff_walking.cu — kernel instantiations for RLB_ENV_CLIFF_WALKING (see rlb_launch.h).
#include "rlb_launch_impl.cuh"
namespace rlb { RLB_INSTANTIATE_ENV(RLB_ENV_CLIFF_WALKING) }
