// rlb_launch.h — per-env launch entry points.  Each env's kernels are instantiated in its
// own translation unit (rlb_inst_<env>.cu) so the 16 (Real, Policy, Selector, Trace)
// variants of every kernel compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "rlb_device.cuh"

namespace rlb {

struct Variant {
    int real;     // rlb_real_kind
    int policy;   // rlb_policy_kind
    int sel;      // rlb_selector_kind
    int trace;    // 0 one-step, 1 eligibility traces
};

struct StepArgs {   // device pointers for the step-level kernels
    const uint32_t* obs = nullptr;
    const uint32_t* action = nullptr;
    const double* reward = nullptr;
    const uint8_t* term = nullptr;
    const uint32_t* obs2 = nullptr;
    const uint32_t* action2 = nullptr;
    const void* values = nullptr;
    const void* td_in = nullptr;
    void* real_out = nullptr;
    uint32_t* u32_out = nullptr;
    uint32_t* u32_out2 = nullptr;
    double* reward_out = nullptr;
    uint8_t* term_out = nullptr;
    uint8_t* kind_out = nullptr;
    uint8_t* not_ready_out = nullptr;
    uint32_t* any_not_ready = nullptr;   // the engine's flag word (FLAG_* bits, rlb_step_kernels.cuh)
    int which = 0;
};

enum StepOp {
    OP_ENV_CONSTRUCT, OP_ENV_RESET, OP_ENV_STEP, OP_GET_ACTION, OP_UPDATE, OP_POLICY_ROWS, OP_POLICY_UPDATE,
    OP_SELECTOR_GET_ACTION, OP_SELECTOR_PROBS, OP_AGENT_STEP
};

template <int ENV> cudaError_t launch_run(const Variant& v, const DevParams& p, int store, cudaStream_t stream);
// the same with the Dyna model attached (InternalModelAgent): HBM store only; instantiated in rlb_inst_<env>_model.cu
template <int ENV> cudaError_t launch_run_model(const Variant& v, const DevParams& p, cudaStream_t stream);
// bytes of dynamic shared memory one 1-warp CTA of the shared-memory / hybrid store needs (0: env not compiled for it)
template <int ENV> size_t smem_store_bytes(const Variant& v, int store, uint32_t rows, uint32_t S, uint32_t vmax);   // rows: table rows kept on chip
template <int ENV> cudaError_t launch_step(StepOp op, const Variant& v, const DevParams& p, const StepArgs& a, cudaStream_t stream);
template <int ENV> cudaError_t run_kernel_attributes(const Variant& v, int store, cudaFuncAttributes* attr);

}   // namespace rlb
