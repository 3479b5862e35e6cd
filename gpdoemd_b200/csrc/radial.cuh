// Radial profiles kappa(r) of the six stationary kernels (variance excluded) with first and second derivative:
// GPdoemd/kernels/stationary.py:33-99 on top of GPy's K_of_r / dK_dr.
#pragma once

namespace gpd {

// ---------------------------------------------------------------------------------------------
// radial profiles  kappa(r) (variance excluded), first and second derivative
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void radial_profile(int kernel, double r, double power, double& k, double& dk,
                                               double& d2k) {
    switch (kernel) {
        case 0: {  // RBF: GPy rbf.py, dK2_drdr = (r^2-1) k
            k = exp(-0.5 * r * r);
            dk = -r * k;
            d2k = (r * r - 1.0) * k;
        } break;
        case 1: {  // Exponential, stationary.py:46-47
            k = exp(-r);
            dk = -k;
            d2k = k;
        } break;
        case 2: {  // Matern32, stationary.py:57-59
            const double s3 = 1.7320508075688772;
            double e = exp(-s3 * r);
            k = (1.0 + s3 * r) * e;
            dk = -3.0 * r * e;
            d2k = 3.0 * (s3 * r - 1.0) * e;
        } break;
        case 3: {  // Matern52, stationary.py:69-71
            const double s5 = 2.23606797749979;
            double e = exp(-s5 * r);
            k = (1.0 + s5 * r + 5.0 / 3.0 * r * r) * e;
            dk = (10.0 / 3.0 * r - 5.0 * r - 5.0 * s5 / 3.0 * r * r) * e;
            double ar = s5 * r;
            d2k = 5.0 / 3.0 * (ar * ar - ar - 1.0) * e;
        } break;
        case 4: {  // Cosine, stationary.py:81-82
            k = cos(r);
            dk = -sin(r);
            d2k = -k;
        } break;
        default: {  // RatQuad, stationary.py:92-99
            double r2 = r * r;
            double lp = log1p(0.5 * r2);
            k = exp(-power * lp);
            double a = (power + 1.0) * lp;
            dk = -power * r * exp(-a);
            double dr1 = -power * exp(-a);
            double dr2 = power * (power + 1.0) * r2 * exp(-(a + lp));
            d2k = dr1 + dr2;
        } break;
    }
}

}  // namespace gpd
