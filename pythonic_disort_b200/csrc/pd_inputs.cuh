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
