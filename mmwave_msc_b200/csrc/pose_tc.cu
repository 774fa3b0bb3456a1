#include "pose_tc.cuh"

namespace mmw {
static const char* g_tc_err = "";
const char* pose_tc_error() { return g_tc_err; }
int pose_tc_init(PoseTc* tc, const float*, int K, int H, int rows_cap, cudaStream_t) {
    tc->ready = false; tc->K = K; tc->H = H; tc->rows_cap = rows_cap;
    return 0;
}
int pose_tc_fc1(PoseTc*, const FcArgs&, int, cudaStream_t, int*) { g_tc_err = "not built"; return -1; }
void pose_tc_free(PoseTc* tc) { tc->ready = false; }
}  // namespace mmw
