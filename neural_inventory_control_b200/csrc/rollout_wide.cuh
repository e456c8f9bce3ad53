// rollout_wide.cuh - fused rollout for wide policy nets / many stores (VanillaWarehouse, SymmetryAware).
#pragma once

#include "hdpo_internal.cuh"

namespace hdpo {
namespace wide {

bool supported(const HdpoRolloutDesc* d);
size_t workspace_bytes(const HdpoRolloutDesc* d);
int forward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st,
            const HdpoState* init, float* cost_b, float* report_b, float* reward_tb, double* totals,
            HdpoState* final_state, void* workspace, size_t workspace_bytes, void* stream);
int backward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st, float g_total,
             float g_report, float* grad_params, void* workspace, size_t workspace_bytes, void* stream);

}  // namespace wide
}  // namespace hdpo
