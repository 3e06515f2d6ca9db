// pd_inputs.cuh -- thermal-source input generation on the device:
//   band-integrated Planck emission      PythonicDISORT.subroutines.Planck / blackbody_contrib_to_BCs (subroutines.py:322-377)
//   linear-in-tau source coefficients    generate_s_poly_coeffs / linear_spline_coefficients        (subroutines.py:381-454)
// The reference integrates Planck(T, nu) over [WVNMLO, WVNMHI] with scipy.integrate.quad_vec (adaptive Gauss-Kronrod,
// relative tolerance 1e-8) once per call on the host; for a million-column ensemble that and the upload of the
// coefficients dominate the end-to-end time.  Here every temperature gets a fixed rule: with x = 100 h c nu / (k T) the
// integrand is (T/a)^4 x^3 / (e^x - 1), smooth, and 16-point Gauss-Legendre panels of width <= 2 in x integrate it to
// rounding; the range is cut where the tail is below 1e-19 of the integral.
#pragma once
#include "pd_common.cuh"

// gl: 16 Gauss-Legendre nodes on [-1, 1] followed by their 16 weights
PD_HD double pd_planck_band_value(double T, double wlo, double whi, const double* gl) {
    if (!(T > 0.0) || !(whi > wlo)) return 0.0;  // the reference returns 0 for T == 0
    const double h = 6.62607015e-34, c = 299792458.0, k = 1.380649e-23;  // scipy.constants (exact SI values)
    const double a = 100.0 * h * c / k;
    const double xlo = a * wlo / T, xhi = a * whi / T;
    const double cap = fmax(xlo + 45.0, 80.0);
    const double xend = xhi < cap ? xhi : cap;
    int nint = (int)ceil((xend - xlo) * 0.5);
    if (nint < 1) nint = 1;
    const double hw = 0.5 * (xend - xlo) / nint;
    double sum = 0.0;
    for (int s = 0; s < nint; ++s) {
        const double mid = xlo + (2 * s + 1) * hw;
        double part = 0.0;
        for (int j = 0; j < 16; ++j) {
            const double x = fma(hw, gl[j], mid);
            const double f = (x > 0.0) ? x * x * x / expm1(x) : 0.0;
            part = fma(gl[16 + j], f, part);
        }
        sum += part;
    }
    const double ta = T / a;
    return 2e8 * h * c * c * (ta * ta) * (ta * ta) * (sum * hw);
}

// (intercept, slope) of the linear segment through (x0, y0), (x1, y1)   (subroutines.py:406-409)
PD_HD void pd_linear_segment(double x0, double y0, double x1, double y1, double* out2) {
    const double slope = (y1 - y0) / (x1 - x0);
    out2[0] = y0 - slope * x0;
    out2[1] = slope;
}

// ---- Hapke surface BDRF ----
// The BDRF DISORT's test problems use (pydisotest/6_test.py:11-24) and the Fourier decomposition the reference's users
// write around it (pydisotest/6_test.py:193-201: quad_vec over delta-phi of BDRF * cos(m dphi), divided by (1 + d_m0) pi).
// tan(alpha / 2) is written as sqrt((1 - cos alpha) / (1 + cos alpha)): no arccos / tan pair.
PD_HD double pd_hapke(double mu, double mup, double cos_dphi, double B0, double HH, double W) {
    double ca = mu * mup - sqrt(1.0 - mu * mu) * sqrt(1.0 - mup * mup) * cos_dphi;
    ca = ca < -1.0 ? -1.0 : (ca > 1.0 ? 1.0 : ca);
    const double phase = 1.0 + 0.5 * ca;
    const double th = sqrt((1.0 - ca) / (1.0 + ca));  // +inf at exact forward... backscatter handled: ca = -1 -> surge 0
    const double surge = B0 * HH / (HH + th);
    const double gam = sqrt(1.0 - W);
    const double h0 = (1.0 + 2.0 * mup) / (1.0 + 2.0 * mup * gam), h = (1.0 + 2.0 * mu) / (1.0 + 2.0 * mu * gam);
    return W / 4.0 / (mu + mup) * ((1.0 + surge) * phase + h0 * h - 1.0);
}

// Fourier modes m < NF of the Hapke BDRF at (mu, mup):
//   out[m * stride] = 1 / ((1 + d_m0) pi) * int_0^2pi BDRF(mu, mup, dphi) cos(m dphi) d dphi
// The integrand is even about dphi = pi, and on [0, pi] it is analytic even when mu == mup (the opposition surge has
// a cusp AT pi: tan(alpha/2) ~ |cos(dphi/2)|), so the integral is 2 int_0^pi by `npanel` 16-point Gauss-Legendre
// panels (a periodic trapezoid rule would converge only like 1/n^2 on the diagonal entries).  cos(m dphi) by the
// Chebyshev recurrence in m.  NFMAX bounds the register accumulators.  gl: 16 nodes on [-1, 1], then 16 weights.
template <int NFMAX>
PD_HD void pd_hapke_modes_point(double mu, double mup, int NF, int npanel, const double* gl, double B0, double HH, double W,
                                double* out, long stride) {
    double acc[NFMAX];
#pragma unroll
    for (int m = 0; m < NFMAX; ++m) acc[m] = 0.0;
    const double hw = 0.5 * PD_PI / npanel;
    for (int s = 0; s < npanel; ++s) {
        const double mid = (2 * s + 1) * hw;
        for (int j = 0; j < 16; ++j) {
            const double c1 = cos(fma(hw, gl[j], mid));
            const double f = gl[16 + j] * pd_hapke(mu, mup, c1, B0, HH, W);
            double cm1 = 1.0, cm = c1;
            acc[0] += f;
            if (NFMAX > 1) acc[1] = fma(f, c1, acc[1]);
#pragma unroll
            for (int m = 2; m < NFMAX; ++m) {
                const double cn = fma(2.0 * c1, cm, -cm1);
                acc[m] = fma(f, cn, acc[m]);
                cm1 = cm;
                cm = cn;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < NFMAX; ++m)
        if (m < NF) out[m * stride] = acc[m] * (2.0 * hw) / ((m == 0 ? 2.0 : 1.0) * PD_PI);
}
