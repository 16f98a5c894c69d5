// Per-voxel least-squares solve and reliability (always evaluated in fp64). Internal header.
//
// 3D: calc_flow.py:337-340 (Sarrus determinant, adjugate solve with the additive eps
// regulariser) and :352-357 (smallest eigenvalue of the symmetric 3x3 window tensor).
// 2D: calc_flow.py:154-156 and :163-168.
#pragma once
#include "common.cuh"

namespace of3d {

constexpr double kEps = 2.220446049250313e-16;  // np.finfo(float).eps, calc_flow.py:155,338

struct Flow3 { double vx, vy, vz, rel; };
struct Flow2 { double vx, vy, rel; };

// Branch-free double-precision reciprocal and reciprocal square root: hardware seed (MUFU.RCP64H / RSQ64H,
// 20 mantissa bits, full exponent range) plus two Newton steps -> ~1 ulp.  No slow-path branch, so several
// voxels' dependency chains can be interleaved by the compiler (the IEEE division's fix-up branch prevents that).
// Special values: the Newton steps would turn the seed of 0 / denormal (inf) or of inf / huge (0) into NaN; outside
// [DBL_MIN, 1e300] the seed itself is returned (+-inf resp. +-0, what the reference's `**-1` gives up to the denormals).
__device__ __forceinline__ double rcp_fast(double x) {
    double s;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x));
    double r = fma(fma(-x, s, 1.0), s, s);
    r = fma(fma(-x, r, 1.0), r, r);
    const double ax = fabs(x);
    return (ax >= 2.2250738585072014e-308 && ax <= 1e300) ? r : s;
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Smallest eigenvalue of the symmetric matrix [[xx,xy,xz],[xy,yy,yz],[xz,yz,zz]].
// The reference runs LAPACK cgeev on complex64 (float32 accuracy); for a real symmetric matrix the
// eigenvalues are real and the lexicographic complex minimum is the smallest one, which is what this
// returns, in float64.
//
// Closed form (Smith 1961): with q = tr/3, p = sqrt(|A - qI|_F^2 / 6), B = (A - qI)/p, r = det(B)/2,
// the eigenvalues are q + 2p*t where t runs over the roots of the Chebyshev cubic 4t^3 - 3t = r; the
// smallest is t = cos(acos(r)/3 + 2pi/3) in [-1, -1/2].  fp64 acos/cos cost ~100 FP64-pipe instructions,
// so the root is found instead on the shifted cubic: with s = t + 1/2 in [-1/2, 0] and e = 1 - r in [0, 2],
//     s^2 (4s - 6) + e = 0 .
// The seed is the fp32 fixed-point iterate of s = -sqrt(e / (6 - 4s)) (contraction rate < 0.19, exact as
// e -> 0; relative error < 6e-3 after two rounds), polished by two Newton steps in fp32 and ONE in fp64 whose
// slope reciprocal only needs fp32 accuracy because Newton is self-correcting.  (Three fp64 Newton steps were the
// longest part of the per-voxel dependency chain; the solve is latency-bound inside strip_window_solve.)  In this form the convergence is
// quadratic in the RELATIVE error of s, also next to the double root (two equal smallest eigenvalues, e -> 0)
// where the trigonometric form loses half the digits.  The whole function is straight-line code (selects
// instead of branches).
__device__ __forceinline__ double min_eig_sym3(double xx, double xy, double xz, double yy, double yz, double zz) {
    const double q = (xx + yy + zz) * (1.0 / 3.0);
    const double a = xx - q, b = yy - q, c = zz - q;
    const double p1 = xy * xy + xz * xz + yz * yz;
    const double p2 = (a * a + b * b + c * c + 2.0 * p1) * (1.0 / 6.0);
    // scalar matrix (incl. all-zero): every eigenvalue is q.  Also taken below 1e-280, where rsqrt.approx.ftz would see a
    // denormal: the eigenvalues then differ from q by less than 1e-139
    const bool scalar = !(p2 > 1e-280);
    const double ip = rsqrt_fast(scalar ? 1.0 : p2);
    const double p = p2 * ip;
    const double ba = a * ip, bb = b * ip, bc = c * ip, bxy = xy * ip, bxz = xz * ip, byz = yz * ip;
    double r = 0.5 * (ba * (bb * bc - byz * byz) - bxy * (bxy * bc - byz * bxz) + bxz * (bxy * byz - bb * bxz));
    // (clamps by compare + select: fmin / fmax on doubles are eight instructions each for the sake of a NaN that the
    // finite window sums never produce)
    r = r > 1.0 ? 1.0 : r;
    r = r < -1.0 ? -1.0 : r;
    const double e = 1.0 - r;
    const float ef = fmaxf((float)e, 1e-30f);
    // The fp32 seed runs on the hardware approximations themselves (MUFU.RSQ / MUFU.RCP, arguments >= 6e-30 are normal):
    // rsqrtf() and __frcp_rn() wrap them in denormal scaling and an IEEE fix-up path behind a branch -- a third of the
    // instructions of this function and three breaks in its instruction schedule -- for last bits that the Newton steps
    // below do not need.
    float sf = -ef * rsqrt_approx(ef * 6.0f);        // -sqrt(e/6)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float d = 6.0f - 4.0f * sf;            // in [6, 8]
        sf = -ef * rsqrt_approx(ef * d);             // -sqrt(e/d)
    }
    // two Newton steps in fp32 (FP32 pipe, off the FP64 dependency chain): relative error ~1e-7
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float f = fmaf(sf * sf, fmaf(4.0f, sf, -6.0f), ef);
        const float fp = 12.0f * sf * (sf - 1.0f);   // > 0 for s < 0
        sf = fminf(-0.0f, sf - f * rcp_approx(fmaxf(fp, 1e-30f)));
    }
    double s = (double)sf;
    {   // one Newton step in fp64: quadratic convergence takes 1e-7 to ~1e-14
        const double f = s * s * (4.0 * s - 6.0) + e;
        const double fp = 12.0 * s * (s - 1.0);
        s -= f * (double)rcp_approx(fmaxf((float)fp, 1e-30f));
    }
    s = s > 0.0 ? 0.0 : s;
    s = s < -0.5 ? -0.5 : s;
    const double lam = (q - p) + 2.0 * p * s;
    return scalar ? q : lam;
}

// The same in float32, for the fp32 mode (whose reliability is float32, like the reference's: LAPACK cgeev on complex64):
// the fp64 version is a chain of ~70 dependent operations at DFMA latency and half of the packed fp32 window kernel's
// instructions.  Error ~ a few eps32 * lambda_max (the fp32 mode's bar is 1e-4 * lambda_max).
__device__ __forceinline__ float min_eig_sym3_f32(float xx, float xy, float xz, float yy, float yz, float zz) {
    const float q = (xx + yy + zz) * (1.0f / 3.0f);
    const float a = xx - q, b = yy - q, c = zz - q;
    const float p1 = xy * xy + xz * xz + yz * yz;
    const float p2 = (a * a + b * b + c * c + 2.0f * p1) * (1.0f / 6.0f);
    const bool scalar = !(p2 > 1e-30f);
    const float ip = rsqrt_approx(scalar ? 1.0f : p2);
    const float p = p2 * ip;
    const float ba = a * ip, bb = b * ip, bc = c * ip, bxy = xy * ip, bxz = xz * ip, byz = yz * ip;
    float r = 0.5f * (ba * (bb * bc - byz * byz) - bxy * (bxy * bc - byz * bxz) + bxz * (bxy * byz - bb * bxz));
    r = fminf(1.0f, fmaxf(-1.0f, r));
    const float ef = fmaxf(1.0f - r, 1e-30f);
    float sf = -ef * rsqrt_approx(ef * 6.0f);
#pragma unroll
    for (int it = 0; it < 2; ++it) sf = -ef * rsqrt_approx(ef * (6.0f - 4.0f * sf));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float f = fmaf(sf * sf, fmaf(4.0f, sf, -6.0f), ef);
        const float fp = 12.0f * sf * (sf - 1.0f);
        sf = fminf(-0.0f, sf - f * rcp_approx(fmaxf(fp, 1e-30f)));
    }
    sf = fmaxf(-0.5f, sf);
    const float lam = (q - p) + 2.0f * p * sf;
    return scalar ? q : lam;
}

// EXACT = true reproduces NumPy's evaluation order with individually rounded operations
// (no FMA contraction), so that given bit-identical window sums the flow is bit-identical.
template <bool EXACT, bool EIG32 = false>
__device__ __forceinline__ Flow3 solve3(double xx, double xy, double xz, double yy, double yz, double zz,
                                        double tx, double ty, double tz) {
    Flow3 o;
    if (EXACT) {
#define M(a, b) __dmul_rn(a, b)
#define A(a, b) __dadd_rn(a, b)
#define S(a, b) __dsub_rn(a, b)
        double det = M(M(xx, yy), zz);
        det = A(det, M(M(M(2.0, xy), xz), yz));
        det = S(det, M(yy, M(xz, xz)));
        det = S(det, M(zz, M(xy, xy)));
        det = S(det, M(xx, M(yz, yz)));
        const double ninv = -__drcp_rn(A(det, kEps));
        o.vx = M(ninv, A(A(M(S(M(yy, zz), M(yz, yz)), tx), M(S(M(xz, yz), M(xy, zz)), ty)), M(S(M(xy, yz), M(xz, yy)), tz)));
        o.vy = M(ninv, A(A(M(S(M(yz, xz), M(xy, zz)), tx), M(S(M(xx, zz), M(xz, xz)), ty)), M(S(M(xz, xy), M(xx, yz)), tz)));
        o.vz = M(ninv, A(A(M(S(M(xy, yz), M(yy, xz)), tx), M(S(M(xy, xz), M(xx, yz)), ty)), M(S(M(xx, yy), M(xy, xy)), tz)));
#undef M
#undef A
#undef S
    } else {
        const double cxx = yy * zz - yz * yz, cxy = xz * yz - xy * zz, cxz = xy * yz - xz * yy;
        const double cyy = xx * zz - xz * xz, cyz = xz * xy - xx * yz, czz = xx * yy - xy * xy;
        // Sarrus determinant written exactly as the reference does (cofactor expansion would
        // round differently where the tensor is near-singular)
        const double det = xx * yy * zz + 2.0 * xy * xz * yz - yy * xz * xz - zz * xy * xy - xx * yz * yz;
        const double ninv = -rcp_fast(det + kEps);
        o.vx = ninv * (cxx * tx + cxy * ty + cxz * tz);
        o.vy = ninv * (cxy * tx + cyy * ty + cyz * tz);
        o.vz = ninv * (cxz * tx + cyz * ty + czz * tz);
    }
    if (EIG32) o.rel = (double)min_eig_sym3_f32((float)xx, (float)xy, (float)xz, (float)yy, (float)yz, (float)zz);
    else o.rel = min_eig_sym3(xx, xy, xz, yy, yz, zz);
    return o;
}

// 2D: the discriminant is evaluated with individually rounded operations in both modes so
// that its sign (NaN vs tiny real, calc_flow.py:166) follows NumPy's.
template <bool EXACT>
__device__ __forceinline__ Flow2 solve2(double xx, double xy, double yy, double tx, double ty) {
    Flow2 o;
    if (!EXACT) {
        // Straight-line code: the IEEE reciprocal and square root each hide a fix-up call behind a branch (cf. rcp_fast).  The
        // determinant and the discriminant keep their individually rounded form, so the SIGN of the discriminant -- NaN or
        // tiny real, calc_flow.py:166 -- still follows NumPy's.
        const double det = __dsub_rn(__dmul_rn(xx, yy), __dmul_rn(xy, xy));
        const double inv = rcp_fast(__dadd_rn(det, kEps));
        o.vx = inv * (xy * ty - yy * tx);
        o.vy = inv * (xy * tx - xx * ty);
        const double tr = __dadd_rn(xx, yy);
        const double disc = __dsub_rn(__dmul_rn(tr, tr), __dmul_rn(4.0, det));
        const bool small = disc < 1e-200;                       // rsqrt.approx.ftz would flush a denormal: scale by 4^300
        const double ds = small ? disc * 0x1p600 : disc;
        double root = ds * rsqrt_fast(ds);
        root = ds > 0.0 ? (small ? root * 0x1p-300 : root) : 0.0;
        // np.minimum of the two roots is the one with the minus sign; a negative discriminant gives NaN like np.sqrt
        o.rel = disc < 0.0 ? __longlong_as_double(0x7ff8000000000000LL) : 0.5 * (tr - root);
        return o;
    }
    const double det = __dsub_rn(__dmul_rn(xx, yy), __dmul_rn(xy, xy));
    const double inv = __drcp_rn(__dadd_rn(det, kEps));
    o.vx = __dmul_rn(inv, __dadd_rn(__dmul_rn(yy, -tx), __dmul_rn(-xy, -ty)));
    o.vy = __dmul_rn(inv, __dadd_rn(__dmul_rn(-xy, -tx), __dmul_rn(xx, -ty)));
    const double tr = __dadd_rn(xx, yy);
    const double disc = __dsub_rn(__dmul_rn(tr, tr), __dmul_rn(4.0, det));
    const double root = sqrt(disc);  // NaN when disc rounds negative, like np.sqrt
    const double l1 = __dmul_rn(__dadd_rn(tr, root), 0.5);
    const double l2 = __dmul_rn(__dsub_rn(tr, root), 0.5);
    o.rel = (l1 != l1 || l2 != l2) ? (l1 + l2) : fmin(l1, l2);  // np.minimum propagates NaN
    return o;
}

}  // namespace of3d
