// Two-sided p-values of the pooled two-sample statistic, fp64, and the Fisher term.
//
// z branch (phet.py:428; statsmodels ztest -> 2*scipy.stats.norm.sf(|z|) -> cephes ndtr/erfc):
//   p = erfc(|z|/sqrt(2)), and exactly 0 when (|z|/sqrt(2))^2 > MAXLOG, which is cephes erfc's
//   underflow rule (p == 0 iff z^2/2 > 709.782712893384).  That rule is
//   observable in the scores because p == 0 terms are dropped from Fisher's sum (phet.py:450).
// t branch (phet.py:426; scipy ttest_ind equal_var=True -> 2*stdtr(df, -|t|)):
//   2*stdtr(df, -|t|) = I_x(df/2, 1/2), x = df/(df + t^2) (regularised incomplete beta),
//   evaluated with the modified-Lentz continued fraction; ln B(df/2, 1/2) comes from the host.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>

namespace phet {

constexpr double kSqrtH = 7.07106781186547524401E-1;   // sqrt(2)/2 (cephes SQRTH)
constexpr double kMaxLog = 7.09782712893383996843E2;   // cephes MAXLOG

__host__ __device__ __forceinline__ double p_two_sided_z(double z) {
    const double a = fabs(z) * kSqrtH;
    if (!(a == a)) return __builtin_nan("");
    if (a * a > kMaxLog) return 0.0;
    return erfc(a);
}

// Continued fraction for I_x(a, b) (Lentz).  Converges quickly for x < (a+1)/(a+b+2).
__host__ __device__ __forceinline__ double betacf(double a, double b, double x) {
    const double tiny = 1.0e-300;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int i = 1; i <= 300; ++i) {
        const double m2 = 2.0 * i;
        double aa = i * (b - i) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + i) * (qab + i) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 2.0e-16) break;
    }
    return h;
}

// p = 2*stdtr(df, -|t|) = I_x(df/2, 1/2) with x = df/(df+t^2), for df > 0.
// x and y = 1-x are both formed directly so neither suffers cancellation.
__host__ __device__ __forceinline__ double p_two_sided_t(double t, double df, double ln_beta) {
    if (!(t == t)) return __builtin_nan("");
    if (isinf(t)) return 0.0;
    const double t2 = t * t;
    const double x = df / (df + t2);
    const double y = t2 / (df + t2);
    if (y <= 0.0) return 1.0;
    if (x <= 0.0) return 0.0;
    const double a = 0.5 * df, b = 0.5;
    const double lnx = (y < 0.5) ? log1p(-y) : log(x);
    const double lny = (x < 0.5) ? log1p(-x) : log(y);
    const double front = exp(a * lnx + b * lny - ln_beta);
    if (x < (a + 1.0) / (a + b + 2.0)) return front * betacf(a, b, x) / a;
    return 1.0 - front * betacf(b, a, y) / b;
}

// -2 ln p with the reference's clean-up: NaN / inf p -> 0 (phet.py:429), p == 0 contributes 0
// (phet.py:449-450: -2*log(0) = +inf is replaced by 0).
__host__ __device__ __forceinline__ double fisher_term(double p) {
    if (!(p > 0.0) || isinf(p)) return 0.0;
    return -2.0 * log(p);
}

}  // namespace phet
