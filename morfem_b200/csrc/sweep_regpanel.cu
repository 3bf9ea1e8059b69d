// placeholder: register-panel sweep variant (filled in below)
#include "sweep_common.cuh"
bool sweep_regpanel_supports(int, int) { return false; }
int sweep_regpanel_launch(const SweepParams&, cudaStream_t) { return -18; }
