#include "k_ws.inc"

namespace ddch {
int launch_ws64(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt) {
    return launch_ws_j<64>(h, p, d_in, n_rows, st, step, jt);
}
}  // namespace ddch
