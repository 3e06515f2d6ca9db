/* pydisort_b200.h -- C ABI of libpydisort_b200.so (sm_100a CUDA kernels).
 *
 * Drop-in boundary for the PythonicDISORT solver hot path.  The reference has
 * no FFI: its seam is the Python function PythonicDISORT.pydisort
 * (src/PythonicDISORT/pydisort.py:13-29) and, below it, the private
 * _assemble_intensity_and_fluxes (src/PythonicDISORT/_assemble_intensity_and_fluxes.py:8-32).
 * Each entry point below names the reference code it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to FP64 (or int32 where said) owned by
 *    the caller; the library never allocates, frees or keeps state;
 *  - arrays are C-contiguous with the shapes given, B (columns) outermost;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *  - return value: 0 ok, <0 invalid argument, >0 cudaError_t of the launch.
 *    Numerical trouble (QR non-convergence, non-positive k^2, zero pivot) is
 *    reported per column in the `status` array (bit mask PD_ST_*), not thrown.
 *  - N = NQuad/2; stream order of every "2N" axis is [+mu_1..+mu_N, -mu_1..-mu_N]
 *    (pydisort.py:304-305).
 */
#ifndef PYDISORT_B200_H
#define PYDISORT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_ABI_VERSION 2
#define PD_MAX_NQUAD 92   /* largest NQuad: one system's elimination panel (size-generic stage B) must fit one CTA's
                           * 227 KB of shared memory; larger values are rejected by every entry point (-8)        */

/* pd_config.flags */
#define PD_FLAG_BEAM      (1 << 0) /* there_is_beam_source  (pydisort.py:215)            */
#define PD_FLAG_ISO       (1 << 1) /* there_is_iso_source   (pydisort.py:216)            */
#define PD_FLAG_DELTA_M   (1 << 2) /* np.any(f_arr > 0)     (pydisort.py:316)            */
#define PD_FLAG_BDRF_PERCOL (1 << 3) /* bdrf_q / bdrf_q0 carry a leading B axis          */
#define PD_FLAG_GENERIC_KERNELS (1 << 8) /* TEST BIT: run the size-generic kernels (Hessenberg-QR eigen stage, pivoted
                                          * band solver, per-output NT recurrences) where a production kernel
                                          * specialised for this NQuad / NLeg_all exists; results agree to rounding.
                                          * The only kernel-selection switch of the library (no environment variables). */

/* per-column status bits (pd_solve) */
#define PD_ST_QR_NOCONV   (1 << 0) /* shifted QR hit its iteration cap                   */
#define PD_ST_BAD_EIGEN   (1 << 1) /* k^2 <= 0 or non-finite: reference would give NaN   */
#define PD_ST_ZERO_PIVOT  (1 << 2) /* exactly singular pivot in an LU                    */

/* input-check bits (pd_prologue), one per ValueError of pydisort.py:223-288 */
#define PD_CHK_TAU_POS      (1 << 0)  /* :223 tau values cannot be non-positive          */
#define PD_CHK_THICK_POS    (1 << 1)  /* :225 layer thicknesses cannot be non-positive   */
#define PD_CHK_OMEGA_RANGE  (1 << 2)  /* :228 omega must be in [0, 1)                    */
#define PD_CHK_LEG_RANGE    (1 << 3)  /* :249 Legendre coefficients must be in (-1, 1)   */
#define PD_CHK_I0_NEG       (1 << 4)  /* :266 beam intensity cannot be negative          */
#define PD_CHK_MU0_RANGE    (1 << 5)  /* :269 mu0 must be in (0, 1]                      */
#define PD_CHK_PHI0_RANGE   (1 << 6)  /* :271 phi0 must be in [0, 2 pi)                  */
#define PD_CHK_F_RANGE      (1 << 7)  /* :287 f must be in [0, 1]                        */
#define PD_CHK_LEG0_FIXED   (1 << 8)  /* :246 warning: g_0 != 1 was corrected            */
#define PD_CHK_OMEGA_NEAR1  (1 << 9)  /* :340 warning: scaled omega > 1 - 1e-6           */
#define PD_CHK_LEG_NEAR1    (1 << 10) /* :342 warning: |scaled g_l| > 0.95               */
#define PD_CHK_MU0_AT_NODE  (1 << 11) /* :309 NT_cor and |mu_i - mu0| < 1e-8             */

typedef struct pd_config {
    int32_t B;         /* columns in this call                                             */
    int32_t L;         /* NLayers                                                          */
    int32_t NQuad;     /* streams, even, >= 2                                              */
    int32_t NLeg;      /* Legendre terms used by the solver, NFourier <= NLeg <= NQuad     */
    int32_t NLeg_all;  /* Legendre terms supplied (>= NLeg; the NT corrections use all)    */
    int32_t NFourier;  /* azimuthal modes solved                                           */
    int32_t NBDRF;     /* BDRF Fourier modes supplied (0 = black surface)                  */
    int32_t Nscoeffs;  /* thermal source polynomial coefficients per layer (0 = none)      */
    int32_t NFb;       /* leading size of b_pos / b_neg mode axis: 1 (mode 0 only) or NFourier */
    int32_t flags;     /* PD_FLAG_*                                                        */
} pd_config;

int pd_abi_version(void);

/* Scratch needed by pd_solve / pd_solve_stages for this configuration (bytes, device memory, 256-byte aligned).
 * Layout (private): a 256-byte head (counter of eigen items the symmetric kernels hand to the general one, so
 * that the fallback pass can return at once), per-system flags, and the history of the boundary-condition sweep
 * of every resident system (Q, Rup Q, q, Rup q + S per layer).  The library never allocates. */
size_t pd_workspace_bytes(const pd_config* cfg);

/* pydisort prologue: input checks, delta-M scaling, thermal-source rescaling
 * and affine transform, source rescale (pydisort.py:223-372), plus the
 * normalised associated Legendre functions at -mu0 for every column
 * (_solve_for_gen_and_part_sols.py:80,101-103).
 *
 * in : tau[B][L] omega[B][L] leg_all[B][L][NLeg_all] f[B][L] s_poly[B][L][Ns]
 *      mu0[B] I0[B] phi0[B] b_pos[B][NFb][N] b_neg[B][NFb][N]   (f, s_poly may be NULL)
 *      mu_nodes[N] (for the NT mu0-vs-node check; pass nt_requested != 0)
 * out: taus[B][L+1] omega_s[B][L] wleg[B][L][NLeg] scale_tau[B][L] s_s[B][L][Ns]
 *      colp[B][PD_NCOLP] (PD_COL_* below) bpos_s[B][NFb][N] bneg_s[B][NFb][N]
 *      pmu0[B][NFourier][NLeg]  checks[1] (OR of PD_CHK_* over all columns, int32) */
#define PD_NCOLP 8
#define PD_COL_MU0     0
#define PD_COL_I0      1  /* rescaled beam intensity                          */
#define PD_COL_RESCALE 2  /* rescale_factor (pydisort.py:351-370)             */
#define PD_COL_PHI0    3
#define PD_COL_I0_RAW  4
#define PD_COL_DM      5  /* 1.0 if this column is delta-M scaled (any f > 0), else 0.0 */
#define PD_COL_NT      6  /* 1.0 if the column-dependent part of the NT gate holds (pydisort.py:375) */
int pd_prologue(const pd_config* cfg,
                const double* tau, const double* omega, const double* leg_all, const double* f,
                const double* s_poly, const double* mu0, const double* I0, const double* phi0,
                const double* b_pos, const double* b_neg, const double* mu_nodes, int nt_requested,
                double* taus, double* omega_s, double* wleg, double* scale_tau, double* s_s,
                double* colp, double* bpos_s, double* bneg_s, double* pmu0, int32_t* checks,
                void* stream);

/* The solve: replaces _solve_for_gen_and_part_sols (whole file) and
 * _solve_for_coeffs (whole file) as orchestrated by
 * _assemble_intensity_and_fluxes.py:109-164.
 *
 * in : prologue outputs; mu_nodes[N] w_nodes[N];
 *      ptab[NFourier][NLeg][N]  normalised P~_l^m(mu_i), zero for l < m;
 *      bdrf_q[(B)][NBDRF][N][N] q^m(mu_i, mu_j); bdrf_q0[(B)][NBDRF][N] q^m(mu_i, mu0)
 * out: K  [B][NFourier][L][N]      positive eigenvalues k (K_collect = [-k, +k])
 *      G  ceil(B*NFourier*L / 32) * 32 items of 2*N*N doubles: the two distinct blocks of G_collect per
 *                                   (column, mode, layer) item = (b * NFourier + m) * L + l,
 *                                   block 0 = G[:N,:N] = G[N:,N:],  block 1 = G[:N,N:] = G[N:,:N];
 *                                   element e = block*N*N + row*N + col of an item lives at (doubles)
 *                                     item * 2*N*N + e                                          (NQuad != 16)
 *                                     (item/32) * 32*2*N*N + (e/4)*128 + (item%32)*4 + e%4     (NQuad == 16:
 *                                   groups of 32 items interleaved by 32-byte sectors, see pd_common.cuh)
 *      Bv [B][NFourier][L][2N]     beam particular solution (B_collect)
 *      dth[B][L][Ns][2N]           thermal particular solution, coefficient of tau*^q
 *      C  [B][NFourier][L][2N]     boundary-condition coefficients (GC_collect = G * C)
 *      Uif[B][L+1][NFourier][2N]   radiances u^m of the 2N streams at the L + 1 layer interfaces (0 = top), particular
 *                                   solutions included, not multiplied by rescale_factor: what the continuity rows of
 *                                   _solve_for_coeffs.py:184-205 equate on the two sides of an interface.  The
 *                                   evaluation entry points return these for query points that ARE interfaces instead
 *                                   of re-assembling G (C * exp) there (pd_state.Uif).  May be NULL (not written).
 *      status[B] int32             PD_ST_* bits */
int pd_solve(const pd_config* cfg,
             const double* taus, const double* omega_s, const double* wleg, const double* s_s,
             const double* colp, const double* bpos_s, const double* bneg_s, const double* pmu0,
             const double* mu_nodes, const double* w_nodes, const double* ptab,
             const double* bdrf_q, const double* bdrf_q0,
             void* workspace, size_t workspace_bytes,
             double* K, double* G, double* Bv, double* dth, double* C, double* Uif, int32_t* status,
             void* stream);

/* Same as pd_solve but runs only the stages selected in `stages`
 * (PD_STAGE_EIGEN: per-(column, mode, layer) eigen-decomposition and particular
 * solutions -> K, G, Bv, dth;  PD_STAGE_BC: per-(column, mode) boundary-condition
 * solve -> C, Uif).  Lets callers time or re-run the two kernels separately. */
#define PD_STAGE_EIGEN 1
#define PD_STAGE_BC    2
int pd_solve_stages(const pd_config* cfg, int stages,
             const double* taus, const double* omega_s, const double* wleg, const double* s_s,
             const double* colp, const double* bpos_s, const double* bneg_s, const double* pmu0,
             const double* mu_nodes, const double* w_nodes, const double* ptab,
             const double* bdrf_q, const double* bdrf_q0,
             void* workspace, size_t workspace_bytes,
             double* K, double* G, double* Bv, double* dth, double* C, double* Uif, int32_t* status,
             void* stream);

/* Output functions evaluated at ntau optical depths per column
 * (tau_q[B][ntau], unscaled tau as the user gives it; each value must lie in
 * [0, tau_L], checked by the caller).  `anti` != 0 selects the tau-antiderivative.
 * A query point equal to 0 or to one of the column's tau[l] (bit for bit) is a layer interface: with st->Uif set and
 * anti == 0 its mode radiances are read from Uif (the same quantity to rounding; G is not read for that point).
 *
 * pd_eval_flux : flux_up / flux_down, _assemble_intensity_and_fluxes.py:446-613
 *   out Fup[B][ntau] Fdn_diffuse[B][ntau] Fdn_direct[B][ntau]
 * pd_eval_u0   : u0 (+ actinic reclassification term), :334-433
 *   out u0[B][2N][ntau]  recl[B][ntau] (may be NULL)
 * pd_eval_u    : u, :170-329, summed over modes at nphi azimuths phi_q[nphi]
 *   out u[B][2N][ntau][nphi]; if nt != 0 the Nakajima-Tanaka TMS + IMS
 *   corrections (pydisort.py:409-694) are added (needs leg_all, f, omega, tau);
 *   ulast[B][2N][ntau] (may be NULL) receives the last Fourier mode for the
 *   Cauchy convergence estimate (:265-316). */
typedef struct pd_state {
    const double* tau;       /* [B][L]   unscaled layer bottoms                    */
    const double* taus;      /* [B][L+1]                                            */
    const double* scale_tau; /* [B][L]                                              */
    const double* colp;      /* [B][PD_NCOLP]                                       */
    const double* K;
    const double* G;
    const double* Bv;
    const double* dth;
    const double* C;
    const double* mu_nodes;
    const double* w_nodes;
    const double* Uif;       /* [B][L+1][NFourier][2N] from pd_solve, or NULL: every point is assembled from G, C   */
} pd_state;

int pd_eval_flux(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti,
                 double* Fup, double* Fdn_diffuse, double* Fdn_direct, void* stream);
int pd_eval_u0(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti,
               double* u0, double* recl, void* stream);
int pd_eval_u(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau,
              const double* phi_q, int nphi, int anti, int nt,
              const double* omega, const double* f, const double* leg_all,
              const double* omega_s, const double* wleg,
              double* u, double* ulast, void* stream);

/* Interpolation of u or u0 from the quadrature streams to user polar angles
 * (PythonicDISORT.subroutines.interpolate, subroutines.py:614-705: barycentric
 * interpolation, separately per hemisphere).  The interpolation is linear in the
 * stream values, so the caller passes it as a weight matrix wts[nmu][2N] (row o:
 * barycentric weights of the N streams of the hemisphere of mu_o, zeros for the
 * other hemisphere) and the kernel contracts the stream axis:
 *   out[b][o][m] = sum_i wts[o][i] * u[b][i][m],   m < M  (M = ntau * nphi, or ntau for u0). */
int pd_interp_mu(int B, int n2, long M, int nmu, const double* wts, const double* u, double* out, void* stream);

/* Thermal-source inputs on the device.
 * pd_planck_band   : PythonicDISORT.subroutines.blackbody_contrib_to_BCs (subroutines.py:354-377), i.e. the integral of
 *                    Planck(T, nu) (:322-350) over [wvnmlo, wvnmhi], for n temperatures T[n] -> out[n]  (W m^-2).
 * pd_s_poly_coeffs : generate_s_poly_coeffs (subroutines.py:413-454) for B columns: tau[B][L] (lower boundaries),
 *                    temper[B][L+1] (level temperatures, top to bottom) -> s_poly[B][L][2] (intercept, slope), the
 *                    `s_poly_coeffs` input of pydisort().
 * gl16[32]: the 16 Gauss-Legendre nodes on [-1, 1] followed by their weights (numpy.polynomial.legendre.leggauss(16)). */
int pd_planck_band(long n, const double* T, double wvnmlo, double wvnmhi, const double* gl16, double* out, void* stream);
int pd_s_poly_coeffs(int B, int L, const double* tau, const double* temper, double wvnmlo, double wvnmhi,
                     const double* gl16, double* s_poly, void* stream);

/* Inputs a caller describes by a few numbers per layer, expanded on the device so that the host -> device link carries
 * the description only ("callers either side of the path"; used by pydisort() / solve_ensemble() for the
 * HenyeyGreenstein and LevelSource input objects of pythonic_disort_b200.api):
 * pd_hg_moments   : phase-function moments of Henyey-Greenstein layers, the `Leg_coeffs_all` input the reference's users
 *                   build as g ** np.arange(NLeg_all) (pydisotest/4_test.py:27, 5_test.py:24):  g[n] -> out[n][NLeg_all],
 *                   out[i][k] = pow(g[i], k) (within 2 ulp of numpy's result), out[i][0] = 1.
 * pd_level_source : `s_poly_coeffs` of a thermal source that is linear in tau inside every layer, from its values at the
 *                   L + 1 levels -- the second half of generate_s_poly_coeffs (subroutines.py:440-454), for callers who
 *                   hold band emissions rather than temperatures:  tau[B][L], lev[B][L+1] -> s_poly[B][L][2]. */
int pd_hg_moments(long n, int NLeg_all, const double* g, double* out, void* stream);
int pd_level_source(int B, int L, const double* tau, const double* lev, double* s_poly, void* stream);

/* Surface input on the device: Fourier modes m < NF of the Hapke BDRF of DISORT's test problems
 * (pydisotest/6_test.py:11-24) in the relative azimuth, as the reference's users compute them with quad_vec
 * (pydisotest/6_test.py:193-201) before handing them to pydisort() / cache_BDRF_Fourier_modes (subroutines.py:490-570):
 *   out[m][i][j] = 1 / ((1 + delta_m0) pi) * int_0^2pi Hapke(mu[i], mup[j], dphi) cos(m dphi) d dphi,   i < N, j < M,
 * as 2 int_0^pi by npanel 16-point Gauss-Legendre panels (gl16[32]: nodes on [-1, 1], then weights; the integrand is even
 * about pi and analytic on [0, pi], also on the diagonal mu == mu' where the opposition surge has its cusp at pi).
 * mup = the quadrature nodes gives q^m(mu_i, mu_j); mup = the beam cosines of B columns gives the per-column
 * q^m(mu_i, mu0_b).  NF <= 64, 16 npanel >= NF. */
int pd_hapke_modes(int N, long M, int NF, int npanel, const double* gl16, const double* mu, const double* mup, double B0,
                   double HH, double W, double* out, void* stream);

/* FP64 FMA throughput probe (one launch of dependent-free DFMA chains); used
 * by bench.py to measure the FP64 roofline denominator on the box.
 * Returns the number of FLOPs the launch performs; time it with CUDA events. */
double pd_fp64_probe(double* sink, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYDISORT_B200_H */
