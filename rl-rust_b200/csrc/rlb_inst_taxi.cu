// rlb_inst_taxi.cu — kernel instantiations for RLB_ENV_TAXI (see rlb_launch.h).
#include "rlb_launch_impl.cuh"
namespace rlb { RLB_INSTANTIATE_ENV(RLB_ENV_TAXI) }
