// prox_math.cuh -- the pointwise projection of stepB, shared by the stand-alone kernel, K3 (foto_kernels.cu) and the
// TMA-staged K3 (prox_tma.cu).  Compiled with -fmad=false like every kernel of the library.
#pragma once
#include "common.cuh"

namespace foto {

// Projection of (alpha, beta1, beta2) onto K = {alpha + |beta|^2/2 <= 0}
// (benamou_brenier.py:123-147), with the algebraic forms SURVEY.md section 7 verified:
//   4/3 a^3 + 4 a^2 + 4 a + 4/3 = 4/3 (a+1)^3,  cos(atan2(b2,b1)) = b1/rho, sin = b2/rho,
//   pow(s, 1/3) = cbrt(s),  zh = c - (a+1)/(3c).
__device__ __forceinline__ void project_K(double a, double b1, double b2, double &qa, double &qb1, double &qb2)
{
    const double rho2 = b1 * b1 + b2 * b2;
    if (2.0 * a + rho2 <= 0.0) { qa = a; qb1 = b1; qb2 = b2; return; }
    // One rsqrt gives rho and the direction (cos, sin) = (b1, b2)/rho; atan2(0, 0) = 0 -> (1, 0).
    // K3 is fp64-instruction bound, not HBM bound, when written with sqrt + two divisions here and
    // cbrt + a division below (52 % of the HBM roofline at 1080x1920x16); rsqrt/rcbrt halve the count.
    double rho = 0.0, ct = 1.0, st = 0.0;
    if (rho2 > 0.0) {
        const double rinv = rsqrt(rho2);
        rho = rho2 * rinv; ct = b1 * rinv; st = b2 * rinv;
    }
    const double a1 = a + 1.0;
    const double cube = a1 * a1 * a1;
    double aH, rhoH;
    if (-32.0 * cube - 108.0 * rho2 < 0.0) {                 // single real root
        const double rad = (4.0 / 3.0) * cube + 4.5 * rho2;
        const double s = 0.35355339059327379 * rho + (1.0 / 6.0) * sqrt(rad);   // sqrt(2)/4
        const double rc = rcbrt(s);                           // 1 / c,  c = s^(1/3) = s * rc^2
        const double c = s * rc * rc;
        const double zh = c - a1 * (rc * (1.0 / 3.0));        // c - (a+1)/(3c)
        aH = -(zh * zh);
        rhoH = 1.4142135623730951 * zh;
    } else {                                                  // three real roots
        const double t = -a1;
        const double arg = 1.8371173070873836 * rho / (t * sqrt(t));            // (3/2)^(3/2)
        const double zh = 1.6329931618554521 * sqrt(t) * cos(acos(arg) / 3.0);  // 2 sqrt(2/3)
        aH = -0.5 * (zh * zh);
        rhoH = zh;
    }
    qa = aH; qb1 = rhoH * ct; qb2 = rhoH * st;
}

}  // namespace foto
