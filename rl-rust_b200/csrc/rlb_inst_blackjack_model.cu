// rlb_inst_blackjack_model.cu — k_run with the Dyna model attached, for RLB_ENV_BLACKJACK (see rlb_launch.h).
#include "rlb_launch_impl.cuh"
namespace rlb { RLB_INSTANTIATE_ENV_MODEL(RLB_ENV_BLACKJACK) }
