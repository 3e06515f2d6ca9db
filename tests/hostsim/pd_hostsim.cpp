// pd_hostsim.cpp -- TEST INFRASTRUCTURE ONLY (never built or loaded by the package).
//
// Compiles the very same device routines (pythonic_disort_b200/csrc/*.cuh) for
// the host with SerialGroup (one "lane") and exposes them under the C ABI of
// include/pydisort_b200.h, pointers being HOST pointers.  The CPU test-suite
// uses it to check the kernels' arithmetic and the Python host logic where no
// GPU exists; it says nothing about intra-warp synchronisation, which only the
// -m gpu tests (and compute-sanitizer on the GPU box) exercise.
#include <stdlib.h>
#include <string.h>

#include "../../pythonic_disort_b200/csrc/pd_eval.cuh"
#include "../../pythonic_disort_b200/csrc/pd_inputs.cuh"
#include "../../pythonic_disort_b200/csrc/pd_prologue.cuh"
#include "../../pythonic_disort_b200/csrc/pd_stage_a.cuh"
#include "../../pythonic_disort_b200/csrc/pd_stage_a_sym.cuh"
#include "../../pythonic_disort_b200/csrc/pd_stage_b.cuh"
#include "../../pythonic_disort_b200/csrc/pd_layer_ops.cuh"
#include "../../pythonic_disort_b200/csrc/pd_stage_b_add.cuh"
#include "../../pythonic_disort_b200/csrc/pd_stage_b_tps.cuh"

// same dispatch as pd_launch_stage_b: for N = 8 the layer operators R^, T^ are formed first, one item at a time
// (pd_layer_ops.cuh), and the sweep runs in its SPLIT form
template <int N>
static bool host_stage_b_add(const PdStageB& sb_in, int b, int m, double* hist) {
    SerialGroup g;
    PdStageB sb = sb_in;
    constexpr bool SPLIT = (N == 8);
    double* sm = (double*)malloc(sizeof(double) * PdStageBAdd<N>::SD);
    double* rt = nullptr;
    if (SPLIT) {
        const long sys = (long)b * sb.NF + m;
        constexpr int NP2 = N * (N + 1);
        rt = (double*)malloc(sizeof(double) * sb.L * NP2);
        double hD[N], park[N * (N + 1) / 2];
        for (int i = 0; i < N; ++i) hD[i] = 0.5 * sqrt(sb.w[i] * sb.mu[i]);
        for (int l = 0; l < sb.L; ++l) {
            const double* ts = sb.taus + (long)b * (sb.L + 1) + l;
            pd_layer_ops_item<N>(sb.G + pd_g_base(sys * sb.L + l, N), sb.K + (sys * sb.L + l) * N, ts[1] - ts[0], hD,
                                 rt + (long)l * NP2, park, 1);
        }
        sb.RT = rt - sys * sb.L * NP2;  // the sweep indexes RT by item
    }
    const bool ok = pd_stage_b_add<SerialGroup, N, SPLIT>(g, sb, b, m, sm, hist);
    free(sm);
    free(rt);
    return ok;
}

extern "C" {

int pd_abi_version(void) { return PD_ABI_VERSION; }
int pd_is_hostsim(void) { return 1; }
static long g_sym_done = 0, g_general = 0, g_bc_add = 0, g_bc_band = 0;
// which: 0 symmetric eigen items, 1 general eigen items, 2 systems solved by the adding stage B, 3 by the band solver
long pd_hostsim_count(int which) { return which == 0 ? g_sym_done : which == 1 ? g_general : which == 2 ? g_bc_add : g_bc_band; }
void pd_hostsim_reset(void) { g_sym_done = g_general = g_bc_add = g_bc_band = 0; }

static long add_hist_doubles(int N, int L) { return (long)L * (2 * N * N + 2 * N); }

size_t pd_workspace_bytes(const pd_config* cfg) {
    const int N = cfg->NQuad / 2;
    const long a = pd_stage_b_history_doubles(N, cfg->L), b = add_hist_doubles(N, cfg->L);
    return (size_t)(a > b ? a : b) * 8;
}

int pd_prologue(const pd_config* cfg, const double* tau, const double* omega, const double* leg_all, const double* f,
                const double* s_poly, const double* mu0, const double* I0, const double* phi0, const double* b_pos,
                const double* b_neg, const double* mu_nodes, int nt_requested, double* taus, double* omega_s,
                double* wleg, double* scale_tau, double* s_s, double* colp, double* bpos_s, double* bneg_s,
                double* pmu0, int32_t* checks, void*) {
    PdPrologue a;
    a.B = cfg->B; a.L = cfg->L; a.N = cfg->NQuad / 2; a.NLeg = cfg->NLeg; a.NLeg_all = cfg->NLeg_all;
    a.NF = cfg->NFourier; a.Ns = cfg->Nscoeffs; a.NFb = cfg->NFb; a.nt_requested = nt_requested;
    a.tau = tau; a.omega = omega; a.leg_all = leg_all; a.f = f; a.s_poly = s_poly; a.mu0 = mu0; a.I0 = I0;
    a.phi0 = phi0; a.b_pos = b_pos; a.b_neg = b_neg; a.mu_nodes = mu_nodes;
    a.taus = taus; a.omega_s = omega_s; a.wleg = wleg; a.scale_tau = scale_tau; a.s_s = s_s; a.colp = colp;
    a.bpos_s = bpos_s; a.bneg_s = bneg_s; a.pmu0 = pmu0; a.checks = checks;
    int chk = 0;
    SerialGroup g;
    for (int b = 0; b < cfg->B; ++b) chk |= pd_prologue_column(g, a, b);
    *checks = chk;
    return 0;
}

int pd_solve_stages(const pd_config* cfg, int stages, const double* taus, const double* omega_s, const double* wleg,
                    const double* s_s, const double* colp, const double* bpos_s, const double* bneg_s,
                    const double* pmu0, const double* mu_nodes, const double* w_nodes, const double* ptab,
                    const double* bdrf_q, const double* bdrf_q0, void* workspace, size_t, double* K, double* G,
                    double* Bv, double* dth, double* C, double* Uif, int32_t* status, void*) {
    const int N = cfg->NQuad / 2;
    if (stages & PD_STAGE_EIGEN) memset(status, 0, sizeof(int32_t) * cfg->B);
    PdStageA a;
    a.B = cfg->B; a.L = cfg->L; a.N = N; a.NLeg = cfg->NLeg; a.NF = cfg->NFourier; a.Ns = cfg->Nscoeffs;
    a.beam = (cfg->flags & PD_FLAG_BEAM) != 0; a.iso = (cfg->flags & PD_FLAG_ISO) != 0;
    a.omega_s = omega_s; a.wleg = wleg; a.s_s = s_s; a.colp = colp; a.pmu0 = pmu0; a.mu = mu_nodes; a.w = w_nodes;
    a.K = K; a.G = G; a.Bv = Bv; a.dth = dth; a.status = status; a.only_flagged = 0; a.nflagged = nullptr;
    SerialGroup g;
    double* sm = (double*)malloc(sizeof(double) * (pd_stage_a_item_doubles(N, cfg->NLeg) + 16));
    double* Q = (double*)malloc(sizeof(double) * cfg->NLeg * N);
    for (int m = 0; m < cfg->NFourier && (stages & PD_STAGE_EIGEN); ++m) {
        const int nm = cfg->NLeg - m;
        for (int idx = 0; idx < nm * N; ++idx)
            Q[idx] = ptab[((long)m * cfg->NLeg + m) * N + idx] * sqrt(w_nodes[idx % N] / mu_nodes[idx % N]);
        // same dispatch as pd_launch_stage_a: symmetric one-thread-per-item path for N = 4, 8 (unless
        // PD_FLAG_GENERIC_KERNELS), general Hessenberg-QR path otherwise and for flagged items
        const bool sym = (N == 4 || N == 8) && !(cfg->flags & PD_FLAG_GENERIC_KERNELS);
        const int NPk = N * (N + 1) / 2;
        double* QQ = (double*)malloc(sizeof(double) * (nm * NPk + 1));
        double* park = (double*)malloc(sizeof(double) * (NPk + 3 * N + 1));
        for (int t = 0; t < nm && sym; ++t) {
            int e = 0;
            for (int i = 0; i < N; ++i)
                for (int j = i; j < N; ++j) QQ[t * NPk + e++] = Q[t * N + i] * Q[t * N + j];
        }
        for (int b = 0; b < cfg->B; ++b)
            for (int l = 0; l < cfg->L; ++l) {
                bool done = false;
                if (sym && N == 4) done = pd_stage_a_sym_item<4>(a, b, m, l, QQ, Q, park, 1);
                if (sym && N == 8) done = pd_stage_a_sym_item<8>(a, b, m, l, QQ, Q, park, 1);
                if (done) ++g_sym_done;
                else ++g_general;
                if (!done) pd_stage_a_item(g, a, b, m, l, Q, sm);
            }
        free(QQ);
        free(park);
    }
    free(sm);
    free(Q);
    if (!(stages & PD_STAGE_BC)) return 0;
    PdStageB sb;
    sb.B = cfg->B; sb.L = cfg->L; sb.N = N; sb.NF = cfg->NFourier; sb.Ns = cfg->Nscoeffs; sb.NBDRF = cfg->NBDRF;
    sb.NFb = cfg->NFb; sb.beam = a.beam; sb.iso = a.iso; sb.bdrf_percol = (cfg->flags & PD_FLAG_BDRF_PERCOL) != 0;
    sb.taus = taus; sb.colp = colp; sb.bpos = bpos_s; sb.bneg = bneg_s; sb.mu = mu_nodes; sb.w = w_nodes;
    sb.bdrf_q = bdrf_q; sb.bdrf_q0 = bdrf_q0; sb.K = K; sb.G = G; sb.RT = nullptr; sb.Bv = Bv; sb.dth = dth; sb.C = C;
    sb.Uif = Uif; sb.status = status;
    double* smb = (double*)malloc(sizeof(double) * (pd_stage_b_doubles(N) + 16));
    // same dispatch as pd_launch_stage_b: interface-radiance elimination for N = 2, 4 (thread per system), 8, 16 (unless
    // PD_FLAG_GENERIC_KERNELS), pivoted band solver otherwise and for the systems the first one hands back
    const bool generic_b = (cfg->flags & PD_FLAG_GENERIC_KERNELS) != 0;
    for (int b = 0; b < cfg->B; ++b)
        for (int m = 0; m < cfg->NFourier; ++m) {
            bool ok = false;
            if (!generic_b) {
                if (N == 2) ok = pd_stage_b_tps<2>(sb, b, m, (double*)workspace, 1);  // one thread per system
                if (N == 4) ok = pd_stage_b_tps<4>(sb, b, m, (double*)workspace, 1);
                if (N == 8) ok = host_stage_b_add<8>(sb, b, m, (double*)workspace);
                if (N == 16) ok = host_stage_b_add<16>(sb, b, m, (double*)workspace);
            }
            if (ok) ++g_bc_add;
            else {
                ++g_bc_band;
                pd_stage_b_system(g, sb, b, m, smb, (double*)workspace);
            }
        }
    free(smb);
    return 0;
}

int pd_solve(const pd_config* cfg, const double* taus, const double* omega_s, const double* wleg, const double* s_s,
             const double* colp, const double* bpos_s, const double* bneg_s, const double* pmu0,
             const double* mu_nodes, const double* w_nodes, const double* ptab, const double* bdrf_q,
             const double* bdrf_q0, void* workspace, size_t wsb, double* K, double* G, double* Bv, double* dth, double* C,
             double* Uif, int32_t* status, void* stream) {
    return pd_solve_stages(cfg, PD_STAGE_EIGEN | PD_STAGE_BC, taus, omega_s, wleg, s_s, colp, bpos_s, bneg_s, pmu0,
                           mu_nodes, w_nodes, ptab, bdrf_q, bdrf_q0, workspace, wsb, K, G, Bv, dth, C, Uif, status, stream);
}

static PdEval make_eval(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti) {
    PdEval a;
    a.B = cfg->B; a.L = cfg->L; a.N = cfg->NQuad / 2; a.NF = cfg->NFourier; a.Ns = cfg->Nscoeffs;
    a.NLeg = cfg->NLeg; a.NLeg_all = cfg->NLeg_all;
    a.beam = (cfg->flags & PD_FLAG_BEAM) != 0; a.iso = (cfg->flags & PD_FLAG_ISO) != 0;
    a.st = *st; a.tau_q = tau_q; a.ntau = ntau; a.anti = anti;
    return a;
}

int pd_eval_flux(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti, double* Fup,
                 double* Fdn_diffuse, double* Fdn_direct, void*) {
    const PdEval a = make_eval(cfg, st, tau_q, ntau, anti);
    SerialGroup g;
    double* sm = (double*)malloc(sizeof(double) * 4 * a.N);
    for (int b = 0; b < a.B; ++b)
        for (int t = 0; t < ntau; ++t)  // as k_eval_flux: the one-thread interface routine first, the group routine otherwise
            if (!(a.st.Uif && !a.anti && pd_flux_point_interface<0>(a, b, t, Fup, Fdn_diffuse, Fdn_direct)))
                pd_flux_point(g, a, b, t, sm, Fup, Fdn_diffuse, Fdn_direct);
    free(sm);
    return 0;
}

int pd_eval_u0(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti, double* u0,
               double* recl, void*) {
    const PdEval a = make_eval(cfg, st, tau_q, ntau, anti);
    SerialGroup g;
    double* sm = (double*)malloc(sizeof(double) * 4 * a.N);
    for (int b = 0; b < a.B; ++b)
        for (int t = 0; t < ntau; ++t) pd_u0_point(g, a, b, t, sm, u0, recl);
    free(sm);
    return 0;
}

int pd_eval_u(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, const double* phi_q, int nphi,
              int anti, int nt, const double* omega, const double* f, const double* leg_all, const double* omega_s,
              const double* wleg, double* u, double* ulast, void*) {
    const PdEval a = make_eval(cfg, st, tau_q, ntau, anti);
    SerialGroup g;
    const int n = a.N, n2 = 2 * n, L = a.L;
    double* ev = (double*)malloc(sizeof(double) * (a.NF + 1) * n2);
    double* um = ev + n2;
    double* scr = (double*)malloc(sizeof(double) * (2 * n * L + 2 * a.NLeg_all + 4));
    double *Rpos = scr, *Rneg = scr + n * L, *imsc = scr + 2 * n * L, *imsv = imsc + a.NLeg_all, *rinv = imsv + 2;
    for (int i = 0; i <= a.NLeg_all; ++i) rinv[i] = (i > 0) ? 1.0 / (double)i : 0.0;
    PdNT p;
    p.omega = omega; p.f = f; p.leg_all = leg_all; p.omega_s = omega_s; p.wleg = wleg;
    for (int b = 0; b < a.B; ++b) {
        const double* cp = st->colp + (long)b * PD_NCOLP;
        const double resc = cp[PD_COL_RESCALE], phi0 = cp[PD_COL_PHI0];
        const int ntb = nt && cp[PD_COL_NT] != 0.0;
        if (ntb) {
            if (L > 1) pd_tms_scans(g, a, b, Rpos, Rneg);
            pd_ims_setup(g, a, p, b, imsc, imsv);
        }
        for (int t = 0; t < ntau; ++t) {
            const double tq = tau_q[(long)b * ntau + t];
            const int l = pd_locate(st->tau + (long)b * L, L, tq);
            const double ts = pd_scaled_tau(a, b, l, tq);
            pd_all_modes_point(g, a, b, l, pd_interface_level(a, b, l, tq), ts, ev, um);
            for (int i = 0; i < n2; ++i) {
                for (int q = 0; q < nphi; ++q) {
                    double v = resc * pd_azimuth_sum(um + i, n2, a.NF, phi0 - phi_q[q]);
                    if (ntb)
                        v += resc * pd_nt_value(a, p, b, i, l, tq, ts, phi_q[q], Rpos, Rneg, imsc, imsv,
                                                leg_all + ((long)b * L + l) * a.NLeg_all, rinv);
                    u[(((long)b * n2 + i) * ntau + t) * nphi + q] = v;
                }
                if (ulast) ulast[((long)b * n2 + i) * ntau + t] = um[(a.NF - 1) * n2 + i];
            }
        }
    }
    free(ev);
    free(scr);
    return 0;
}

int pd_interp_mu(int B, int n2, long M, int nmu, const double* wts, const double* u, double* out, void*) {
    if (B < 1 || n2 < 2 || (n2 & 1) || M < 1 || nmu < 1 || !wts || !u || !out) return -40;
    for (long b = 0; b < B; ++b)
        for (int o = 0; o < nmu; ++o)
            for (long m = 0; m < M; ++m) {
                double acc = 0.0;
                for (int i = 0; i < n2; ++i) acc = fma(wts[(long)o * n2 + i], u[(b * n2 + i) * M + m], acc);
                out[(b * nmu + o) * M + m] = acc;
            }
    return 0;
}

int pd_hg_moments(long n, int NA, const double* g, double* out, void*) {
    for (long i = 0; i < n; ++i) {
        for (int k = 0; k < NA; ++k) out[i * NA + k] = (k == 0) ? 1.0 : ((k == 1) ? g[i] : pow(g[i], (double)k));
    }
    return 0;
}

int pd_level_source(int B, int L, const double* tau, const double* lev, double* s_poly, void*) {
    for (long b = 0; b < B; ++b)
        for (int l = 0; l < L; ++l)
            pd_linear_segment(l == 0 ? 0.0 : tau[b * L + l - 1], lev[b * (L + 1) + l], tau[b * L + l], lev[b * (L + 1) + l + 1],
                              s_poly + (b * L + l) * 2);
    return 0;
}

int pd_planck_band(long n, const double* T, double wlo, double whi, const double* gl16, double* out, void*) {
    if (n < 1 || !T || !gl16 || !out) return -50;
    for (long i = 0; i < n; ++i) out[i] = pd_planck_band_value(T[i], wlo, whi, gl16);
    return 0;
}

int pd_s_poly_coeffs(int B, int L, const double* tau, const double* temper, double wlo, double whi, const double* gl16,
                     double* s_poly, void*) {
    if (B < 1 || L < 1 || !tau || !temper || !gl16 || !s_poly) return -50;
    for (long b = 0; b < B; ++b)
        for (int l = 0; l < L; ++l) {
            const double e0 = pd_planck_band_value(temper[b * (L + 1) + l], wlo, whi, gl16);
            const double e1 = pd_planck_band_value(temper[b * (L + 1) + l + 1], wlo, whi, gl16);
            pd_linear_segment(l == 0 ? 0.0 : tau[b * L + l - 1], e0, tau[b * L + l], e1, s_poly + (b * L + l) * 2);
        }
    return 0;
}

int pd_hapke_modes(int N, long M, int NF, int npanel, const double* gl16, const double* mu, const double* mup, double B0,
                   double HH, double W, double* out, void*) {
    if (N < 1 || M < 1 || NF < 1 || NF > 64 || npanel < 1 || 16 * npanel < NF || !gl16 || !mu || !mup || !out) return -60;
    for (int i = 0; i < N; ++i)
        for (long j = 0; j < M; ++j)
            pd_hapke_modes_point<64>(mu[i], mup[j], NF, npanel, gl16, B0, HH, W, out + (long)i * M + j, (long)N * M);
    return 0;
}

double pd_fp64_probe(double*, int, void*) { return -1.0; }

}  // extern "C"
