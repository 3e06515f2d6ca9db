// pd_api.cu -- C ABI of libpydisort_b200.so (include/pydisort_b200.h): configuration checks, workspace sizing,
// the prologue kernel, the solve entry points and the FP64 probe.  sm_100a only.
#include "pd_launch.h"

__global__ void __launch_bounds__(128) k_prologue(PdPrologue a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= a.B) return;
    SubWarp<32> g;
    int chk = pd_prologue_column(g, a, warp);
    chk = __reduce_or_sync(0xffffffffu, chk);
    if (chk && g.lane() == 0) atomicOr(a.checks, chk);
}

__global__ void __launch_bounds__(256) k_fp64_probe(double* sink, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double x = 0.999999, y = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

int pd_check_cfg(const pd_config* c) {
    if (!c) return -1;
    if (c->B < 1 || c->L < 1) return -2;
    if (c->NQuad < 2 || (c->NQuad & 1)) return -3;
    if (c->NLeg < 1 || c->NLeg > c->NQuad || c->NLeg > c->NLeg_all) return -4;
    if (c->NFourier < 1 || c->NFourier > c->NLeg) return -5;
    if (c->NBDRF < 0 || c->Nscoeffs < 0) return -6;
    if (c->NFb != 1 && c->NFb != c->NFourier) return -7;
    if (c->NQuad > PD_MAX_NQUAD) return -8;
    return 0;
}

extern "C" {

int pd_abi_version(void) { return PD_ABI_VERSION; }

size_t pd_workspace_bytes(const pd_config* cfg) {
    if (pd_check_cfg(cfg)) return 0;
    return PD_WS_HEAD + pd_plan_stage_b(cfg->B, cfg->NFourier, cfg->NQuad / 2, cfg->L, cfg->flags).bytes();
}

int pd_prologue(const pd_config* cfg, const double* tau, const double* omega, const double* leg_all, const double* f,
                const double* s_poly, const double* mu0, const double* I0, const double* phi0, const double* b_pos,
                const double* b_neg, const double* mu_nodes, int nt_requested, double* taus, double* omega_s,
                double* wleg, double* scale_tau, double* s_s, double* colp, double* bpos_s, double* bneg_s,
                double* pmu0, int32_t* checks, void* stream) {
    if (int e = pd_check_cfg(cfg)) return e;
    PdPrologue a;
    a.B = cfg->B; a.L = cfg->L; a.N = cfg->NQuad / 2; a.NLeg = cfg->NLeg; a.NLeg_all = cfg->NLeg_all;
    a.NF = cfg->NFourier; a.Ns = cfg->Nscoeffs; a.NFb = cfg->NFb; a.nt_requested = nt_requested;
    a.tau = tau; a.omega = omega; a.leg_all = leg_all; a.f = f; a.s_poly = s_poly; a.mu0 = mu0; a.I0 = I0;
    a.phi0 = phi0; a.b_pos = b_pos; a.b_neg = b_neg; a.mu_nodes = mu_nodes;
    a.taus = taus; a.omega_s = omega_s; a.wleg = wleg; a.scale_tau = scale_tau; a.s_s = s_s; a.colp = colp;
    a.bpos_s = bpos_s; a.bneg_s = bneg_s; a.pmu0 = pmu0; a.checks = checks;
    cudaError_t e = cudaMemsetAsync(checks, 0, sizeof(int32_t), pd_stream(stream));
    if (e != cudaSuccess) return (int)e;
    const int wpb = 4;
    k_prologue<<<(cfg->B + wpb - 1) / wpb, wpb * 32, 0, pd_stream(stream)>>>(a);
    return (int)cudaGetLastError();
}

int pd_solve_stages(const pd_config* cfg, int stages, const double* taus, const double* omega_s, const double* wleg,
                    const double* s_s, const double* colp, const double* bpos_s, const double* bneg_s,
                    const double* pmu0, const double* mu_nodes, const double* w_nodes, const double* ptab,
                    const double* bdrf_q, const double* bdrf_q0, void* workspace, size_t workspace_bytes, double* K,
                    double* G, double* Bv, double* dth, double* C, double* Uif, int32_t* status, void* stream) {
    if (int e = pd_check_cfg(cfg)) return e;
    const int N = cfg->NQuad / 2;
    const bool beam = (cfg->flags & PD_FLAG_BEAM) != 0, iso = (cfg->flags & PD_FLAG_ISO) != 0;
    if (stages & PD_STAGE_EIGEN) {
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t) * cfg->B, pd_stream(stream));
        if (e != cudaSuccess) return (int)e;
        PdStageA a;
        a.B = cfg->B; a.L = cfg->L; a.N = N; a.NLeg = cfg->NLeg; a.NF = cfg->NFourier; a.Ns = cfg->Nscoeffs;
        a.beam = beam; a.iso = iso; a.only_flagged = 0;
        a.omega_s = omega_s; a.wleg = wleg; a.s_s = s_s; a.colp = colp; a.pmu0 = pmu0; a.mu = mu_nodes; a.w = w_nodes;
        a.K = K; a.G = G; a.Bv = Bv; a.dth = dth; a.status = status;
        a.nflagged = (workspace && workspace_bytes >= PD_WS_HEAD) ? static_cast<int32_t*>(workspace) : nullptr;
        if (int rc = pd_launch_stage_a(a, cfg->flags, ptab, pd_stream(stream))) return rc;
    }
    if (stages & PD_STAGE_BC) {
        PdStageB sb;
        sb.B = cfg->B; sb.L = cfg->L; sb.N = N; sb.NF = cfg->NFourier; sb.Ns = cfg->Nscoeffs; sb.NBDRF = cfg->NBDRF;
        sb.NFb = cfg->NFb; sb.beam = beam; sb.iso = iso; sb.bdrf_percol = (cfg->flags & PD_FLAG_BDRF_PERCOL) != 0;
        sb.taus = taus; sb.colp = colp; sb.bpos = bpos_s; sb.bneg = bneg_s; sb.mu = mu_nodes; sb.w = w_nodes;
        sb.bdrf_q = bdrf_q; sb.bdrf_q0 = bdrf_q0; sb.K = K; sb.G = G; sb.RT = nullptr; sb.Bv = Bv; sb.dth = dth; sb.C = C;
        sb.Uif = Uif; sb.status = status;
        if (!workspace || workspace_bytes < PD_WS_HEAD) return -20;
        if (int rc = pd_launch_stage_b(sb, cfg->flags, static_cast<char*>(workspace) + PD_WS_HEAD, workspace_bytes - PD_WS_HEAD,
                                       pd_stream(stream)))
            return rc;
    }
    return 0;
}

int pd_solve(const pd_config* cfg, const double* taus, const double* omega_s, const double* wleg, const double* s_s,
             const double* colp, const double* bpos_s, const double* bneg_s, const double* pmu0,
             const double* mu_nodes, const double* w_nodes, const double* ptab, const double* bdrf_q,
             const double* bdrf_q0, void* workspace, size_t workspace_bytes, double* K, double* G, double* Bv,
             double* dth, double* C, double* Uif, int32_t* status, void* stream) {
    return pd_solve_stages(cfg, PD_STAGE_EIGEN | PD_STAGE_BC, taus, omega_s, wleg, s_s, colp, bpos_s, bneg_s, pmu0,
                           mu_nodes, w_nodes, ptab, bdrf_q, bdrf_q0, workspace, workspace_bytes, K, G, Bv, dth, C,
                           Uif, status, stream);
}

double pd_fp64_probe(double* sink, int iters, void* stream) {
    const int blocks = PD_NUM_SMS * 8, threads = 256;
    k_fp64_probe<<<blocks, threads, 0, pd_stream(stream)>>>(sink, iters);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return 2.0 * 8.0 * (double)iters * blocks * threads;
}

}  // extern "C"
