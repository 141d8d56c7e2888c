// rlb_inst_blackjack.cu — kernel instantiations for RLB_ENV_BLACKJACK (see rlb_launch.h).
#include "rlb_launch_impl.cuh"
namespace rlb { RLB_INSTANTIATE_ENV(RLB_ENV_BLACKJACK) }
