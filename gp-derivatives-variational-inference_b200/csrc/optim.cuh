// Fused multi-tensor Adam step (definitions in optim.cu).
#pragma once
#include "common.cuh"

namespace dsvgp {

constexpr int ADAM_MAX_TENSORS = 24, ADAM_MAX_GROUPS = 4;

// desc_host: ntensors x 8 int64 {param, grad, exp_avg, exp_avg_sq (device addresses), numel, group, tri_n, 0}
// group_host: ngroups x 8 double {lr, beta1, beta2, eps, weight_decay, step (1-based, after increment), 0, 0}
template <typename T>
int adam_step(int ntensors, const int64_t* desc_host, int ngroups, const double* group_host, cudaStream_t st);

}  // namespace dsvgp
