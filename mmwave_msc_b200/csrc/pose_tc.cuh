// Tensor-core (tcgen05 + TMEM + TMA) path for the dense contraction of the pose network.
#pragma once
#include "pose.cuh"

namespace mmw {

struct PoseTc {
    bool ready = false;
    int K = 0, H = 0, rows_cap = 0;
    void* impl = nullptr;
};

// Prepares the split-bf16 weight operand and the TMA descriptors.  Returns 0 on success.
int pose_tc_init(PoseTc* tc, const float* host_w_kh, int K, int H, int rows_cap, cudaStream_t st);
// out = BN(relu(A W + b)) for the first *n_rows rows; *n_launches = kernels launched.
int pose_tc_fc1(PoseTc* tc, const FcArgs& a, int max_rows, cudaStream_t st, int* n_launches);
void pose_tc_free(PoseTc* tc);
const char* pose_tc_error();

}  // namespace mmw
