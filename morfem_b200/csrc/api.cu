// C-ABI glue: version / error reporting and the dispatcher of the batched reduced sweep.
#include <stdlib.h>
#include "sweep_common.cuh"

thread_local char g_mf_err[512] = "";
std::atomic<long long> g_mf_launches{0};

// variant entry points (defined in sweep_*.cu)
int sweep_generic_launch(const SweepParams& p, size_t ws_bytes, cudaStream_t stream);
size_t sweep_generic_ws_bytes(int r, int m, long long F);
int sweep_regpanel_launch(const SweepParams& p, cudaStream_t stream);
bool sweep_regpanel_supports(int r, int m);
int sweep_blocked_launch(const SweepParams& p, size_t ws_bytes, cudaStream_t stream);
size_t sweep_blocked_ws_bytes(int r, int m, long long F);
bool sweep_blocked_supports(int r, int m);
bool sweep_blocked_fits_smem(int r, int m);
bool sweep_stream_supports(int r, int m);
size_t sweep_stream_ws_bytes(int r, int m, long long F);
int sweep_stream_launch(const SweepParams& p, size_t ws_bytes, cudaStream_t stream);
bool sweep_left_supports_c128(int r, int m);
size_t sweep_left_ws_bytes_c128(int r, int m, long long F);
int sweep_left_launch_c128(const SweepParams& p, size_t ws_bytes, cudaStream_t stream);

extern "C" int mf_version(void) { return MF_VERSION; }
extern "C" const char* mf_last_error(void) { return g_mf_err; }
extern "C" int64_t mf_launch_count(void) { return (int64_t)g_mf_launches.load(std::memory_order_relaxed); }

// Variants: 1 generic (checker), 2 register panel (r <= 64), 3 blocked family = the kernel the library picks (matrix in shared
// memory up to r = 112, left-looking streamed LU above), 4 right-looking streamed LU (sweep_stream.cu, 113..512; round-1 kernel,
// kept as a cross-check and for measurements), 5 left-looking LU forced (any r <= 512).
static int pick_variant(int r, int m, int variant) {
    if (variant != 0) return variant;
    if (sweep_blocked_fits_smem(r, m) || sweep_left_supports_c128(r, m)) return 3;
    if (sweep_regpanel_supports(r, m)) return 2;
    return 1;
}

// Measured crossover (profiles/r02_sweep_crossover.md): the shared-memory kernel wins up to r = 64 (three 4-warp CTAs per SM), from
// r = 65 its one 8-warp CTA per SM loses to the left-looking kernel's four 4-warp CTAs (3.9 M against 2.2 M points/s at r = 80).
static bool blocked_family_uses_left(int r, int m) {
    if (getenv("MF_SWEEP_FORCE_LEFT") && sweep_left_supports_c128(r, m)) return true;
    if (getenv("MF_SWEEP_FORCE_SMEM") && sweep_blocked_fits_smem(r, m)) return false;
    return !sweep_blocked_fits_smem(r, m) || (r > 64 && sweep_left_supports_c128(r, m));
}

extern "C" int mf_sweep_variant_supported(int r, int m, int variant) {
    if (r <= 0 || r > 1024 || m <= 0 || m > MF_MAX_PORTS) return 0;
    if (variant == 0 || variant == 1) return 1;
    if (variant == 2) return sweep_regpanel_supports(r, m) ? 1 : 0;
    if (variant == 3) return (sweep_blocked_fits_smem(r, m) || sweep_left_supports_c128(r, m)) ? 1 : 0;
    if (variant == 4) return (!sweep_blocked_fits_smem(r, m) && sweep_stream_supports(r, m)) ? 1 : 0;
    if (variant == 5) return sweep_left_supports_c128(r, m) ? 1 : 0;
    return 0;
}

extern "C" size_t mf_sweep_ws_bytes(int r, int m, int64_t F, int variant) {
    if (r <= 0 || m <= 0 || F <= 0) return 256;
    const int v = pick_variant(r, m, variant);
    size_t need = 0;
    if (v == 1) need = sweep_generic_ws_bytes(r, m, F);
    else if (v == 3) need = blocked_family_uses_left(r, m) ? sweep_left_ws_bytes_c128(r, m, F) : 0;
    else if (v == 4) need = sweep_stream_ws_bytes(r, m, F);
    else if (v == 5) need = sweep_left_ws_bytes_c128(r, m, F);
    return need < 256 ? 256 : need;
}

extern "C" int mf_sweep_lu_gsm_c128(const mf_c128* A0, const mf_c128* A1, const mf_c128* A2, int64_t lda,
                                    const mf_c128* Br, int64_t ldb, int r, int m,
                                    const double* c0, const double* c1, const double* c2, const double* cb,
                                    const double* zscale, int64_t F,
                                    mf_c128* X, mf_c128* S, int* info, int variant,
                                    void* ws, size_t ws_bytes, void* stream) {
    if (!A0 && !A1 && !A2) MF_FAIL_ARG(1, "all three operators are NULL");
    if (lda < r) MF_FAIL_ARG(4, "lda < r");
    if (!Br || ldb < m) MF_FAIL_ARG(5, "Br is NULL or ldb < m");
    if (r <= 0 || r > 1024) MF_FAIL_ARG(7, "need 0 < r <= 1024");
    if (m <= 0 || m > MF_MAX_PORTS) MF_FAIL_ARG(8, "need 0 < m <= MF_MAX_PORTS");
    if (!c0 || !c1 || !c2) MF_FAIL_ARG(9, "coefficient arrays must not be NULL");
    if (!cb) MF_FAIL_ARG(12, "cb is NULL");
    if (S && !zscale) MF_FAIL_ARG(13, "zscale is NULL but S is requested");
    if (F < 0) MF_FAIL_ARG(14, "F < 0");
    if (!X && !S) MF_FAIL_ARG(15, "neither X nor S requested");
    if (variant < 0 || variant > 5) MF_FAIL_ARG(18, "variant must be 0..5");
    if (F == 0) return 0;
    SweepParams p;
    p.A0 = (const cplx*)A0; p.A1 = (const cplx*)A1; p.A2 = (const cplx*)A2; p.lda = lda;
    p.Br = (const cplx*)Br; p.ldb = ldb; p.r = r; p.m = m;
    p.c0 = c0; p.c1 = c1; p.c2 = c2; p.cb = cb; p.zs = zscale; p.F = F;
    p.X = (cplx*)X; p.S = (cplx*)S; p.info = info; p.ws = (cplx*)ws; p.ws_stride = 0;
    const int v = pick_variant(r, m, variant);
    cudaStream_t st = (cudaStream_t)stream;
    if (v == 2) {
        if (!sweep_regpanel_supports(r, m)) MF_FAIL_ARG(18, "register-panel variant does not support this (r, m)");
        return sweep_regpanel_launch(p, st);
    }
    if (v == 3) {
        if (!mf_sweep_variant_supported(r, m, 3)) MF_FAIL_ARG(18, "blocked variant does not support this (r, m)");
        if (blocked_family_uses_left(r, m)) return sweep_left_launch_c128(p, ws_bytes, st);
        return sweep_blocked_launch(p, ws_bytes, st);
    }
    if (v == 4) {
        if (!mf_sweep_variant_supported(r, m, 4)) MF_FAIL_ARG(18, "right-looking streamed variant does not support this (r, m)");
        return sweep_stream_launch(p, ws_bytes, st);
    }
    if (v == 5) {
        if (!sweep_left_supports_c128(r, m)) MF_FAIL_ARG(18, "left-looking variant does not support this (r, m)");
        return sweep_left_launch_c128(p, ws_bytes, st);
    }
    return sweep_generic_launch(p, ws_bytes, st);
}
