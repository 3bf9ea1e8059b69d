// placeholder: blocked DMMA sweep variant (filled in below)
#include "sweep_common.cuh"
bool sweep_blocked_supports(int, int) { return false; }
size_t sweep_blocked_ws_bytes(int, int, long long) { return 0; }
int sweep_blocked_launch(const SweepParams&, size_t, cudaStream_t) { return -18; }
