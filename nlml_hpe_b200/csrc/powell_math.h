// Derivative-free minimisation exactly as the reference runs it: scipy.optimize.minimize(method="Powell").
//
// The reference's shipped default fit (TD_Tester.Test, /root/reference/TD_Tester.py:191-194) and its cosine-curve trainer
// (TD_Trainer.optimize_for_matrix_using_grads, /root/reference/TD_Trainer.py:81-88) both hand their objective to scipy's
// modified Powell method.  scipy is a third-party dependency that is neither vendored under /root/reference nor pinned
// (it is not even listed in requirements.txt); the algorithm restated here is the published one of scipy 1.18.1
// (the version in the build container, the one tests/golden/powell_golden.npz was produced with):
//   scipy/optimize/_optimize.py  bracket()  Brent.optimize()  _linesearch_powell()  _minimize_powell()
// with the defaults the reference leaves in place: xtol = ftol = 1e-4, maxiter = maxfev = N * 1000, direc = I, unbounded
// line searches (Brent with tol = xtol * 100 and the default bracket (0, 1)).  Float64 throughout, statement for
// statement, so that the same sequence of evaluations is taken: tests/test_hostcheck.py pins this restatement to the
// reference's own results INCLUDING the number of function evaluations.
//
// __host__ __device__: the CUDA kernels (tucker_fit.cu) and the CPU test tier (tests/hostcheck) compile the same text.
#pragma once
#include "tucker_math.h"

namespace nlml {
namespace powell {

constexpr int kMaxN = 3 + 16;   // parameters: three angles + identity rank (<= 16), or 4 for a cosine row

// Every arithmetic operation of the search is rounded on its own, as CPython / numpy evaluate it: nvcc would otherwise
// contract a*b+c into a fused multiply-add (one rounding instead of two), the search would evaluate the objective at
// slightly different points and leave the reference's path.  (The host check build is compiled with -ffp-contract=off.)
NLML_HD double mul_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
NLML_HD double add_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
NLML_HD double sub_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
NLML_HD double div_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// Counts evaluations and enforces maxfev like scipy's _wrap_scalar_function_maxfun_validation: the call that would exceed
// the budget is not made, `aborted` is set and every caller unwinds (scipy raises _MaxFuncCallError).
template <class F>
struct Counted {
    F& f;
    int calls = 0, maxfun;
    bool aborted = false;
    NLML_HD Counted(F& f_, int maxfun_) : f(f_), maxfun(maxfun_) {}
    NLML_HD double operator()(const double* x) {
        if (calls >= maxfun) {
            aborted = true;
            return 0.0;
        }
        ++calls;
        return f(x);
    }
};

// f along a line: alpha -> f(p + alpha * xi)
template <class CF>
struct LineFn {
    CF& cf;
    const double* p;
    const double* xi;
    int n;
    NLML_HD double operator()(double alpha) {
        double y[kMaxN];
        for (int i = 0; i < n; ++i) y[i] = add_(p[i], mul_(alpha, xi[i]));
        return cf(y);
    }
};

NLML_HD double nan_value() {
    const double zero = 0.0;
    return zero / zero;
}

struct Bracket {
    double xa, xb, xc, fa, fb, fc;
    bool valid;
};

// scipy.optimize.bracket(func, xa=0.0, xb=1.0, grow_limit=110.0, maxiter=1000)
template <class L>
NLML_HD Bracket bracket(L& func, bool& aborted_flag, const bool& aborted) {
    const double gold = 1.618034, verysmall = 1e-21, grow_limit = 110.0;
    const int maxiter = 1000;
    Bracket r{};
    r.valid = false;
    double xa = 0.0, xb = 1.0;
    double fa = func(xa);
    if (aborted) { aborted_flag = true; return r; }
    double fb = func(xb);
    if (aborted) { aborted_flag = true; return r; }
    if (fa < fb) {
        double t = xa; xa = xb; xb = t;
        t = fa; fa = fb; fb = t;
    }
    double xc = add_(xb, mul_(gold, sub_(xb, xa)));
    double fc = func(xc);
    if (aborted) { aborted_flag = true; return r; }
    int iter = 0;
    while (fc < fb) {
        const double tmp1 = mul_(sub_(xb, xa), sub_(fb, fc));
        const double tmp2 = mul_(sub_(xb, xc), sub_(fb, fa));
        const double val = sub_(tmp2, tmp1);
        const double denom = fabs(val) < verysmall ? 2.0 * verysmall : mul_(2.0, val);
        double w = sub_(xb, div_(sub_(mul_(sub_(xb, xc), tmp2), mul_(sub_(xb, xa), tmp1)), denom));
        const double wlim = add_(xb, mul_(grow_limit, sub_(xc, xb)));
        if (iter > maxiter) break;   // scipy raises RuntimeError here; unreachable for the objectives of this library
        ++iter;
        double fw;
        if (mul_(sub_(w, xc), sub_(xb, w)) > 0.0) {
            fw = func(w);
            if (aborted) { aborted_flag = true; return r; }
            if (fw < fc) {
                xa = xb; xb = w; fa = fb; fb = fw;
                break;
            } else if (fw > fb) {
                xc = w; fc = fw;
                break;
            }
            w = add_(xc, mul_(gold, sub_(xc, xb)));
            fw = func(w);
            if (aborted) { aborted_flag = true; return r; }
        } else if (mul_(sub_(w, wlim), sub_(wlim, xc)) >= 0.0) {
            w = wlim;
            fw = func(w);
            if (aborted) { aborted_flag = true; return r; }
        } else if (mul_(sub_(w, wlim), sub_(xc, w)) > 0.0) {
            fw = func(w);
            if (aborted) { aborted_flag = true; return r; }
            if (fw < fc) {
                xb = xc; xc = w;
                w = add_(xc, mul_(gold, sub_(xc, xb)));
                fb = fc; fc = fw;
                fw = func(w);
                if (aborted) { aborted_flag = true; return r; }
            }
        } else {
            w = add_(xc, mul_(gold, sub_(xc, xb)));
            fw = func(w);
            if (aborted) { aborted_flag = true; return r; }
        }
        xa = xb; xb = xc; xc = w;
        fa = fb; fb = fc; fc = fw;
    }
    const bool cond1 = (fb < fc && fb <= fa) || (fb < fa && fb <= fc);
    const bool cond2 = (xa < xb && xb < xc) || (xc < xb && xb < xa);
    const bool cond3 = (xa - xa == 0.0) && (xb - xb == 0.0) && (xc - xc == 0.0);   // all finite
    r.xa = xa; r.xb = xb; r.xc = xc; r.fa = fa; r.fb = fb; r.fc = fc;
    r.valid = cond1 && cond2 && cond3;
    return r;
}

// _recover_from_bracket_error(_minimize_scalar_brent, myfunc, None, (), xtol=tol): Brent.optimize() on the default
// bracket; an invalid bracket returns the best of its three points (scipy intercepts BracketError that way).
template <class L>
NLML_HD void brent_minimize(L& func, double tol, bool& aborted_flag, const bool& aborted, double& xmin, double& fmin) {
    const double mintol = 1.0e-11, cg = 0.3819660;
    const int maxiter = 500;
    Bracket br = bracket(func, aborted_flag, aborted);
    if (aborted_flag) return;
    if (!br.valid) {
        const double xs[3] = {br.xa, br.xb, br.xc}, fs[3] = {br.fa, br.fb, br.fc};
        bool nan = false;
        for (int i = 0; i < 3; ++i) nan = nan || xs[i] != xs[i] || fs[i] != fs[i];
        if (nan) {
            xmin = fmin = nan_value();
            return;
        }
        int imin = 0;   // np.argmin: first minimum
        for (int i = 1; i < 3; ++i)
            if (fs[i] < fs[imin]) imin = i;
        xmin = xs[imin];
        fmin = fs[imin];
        return;
    }
    double x, w, v, fw, fv, fx, a, b;
    x = w = v = br.xb;
    fw = fv = fx = br.fb;
    if (br.xa < br.xc) { a = br.xa; b = br.xc; } else { a = br.xc; b = br.xa; }
    double deltax = 0.0, rat = 0.0;
    int iter = 0;
    while (iter < maxiter) {
        const double tol1 = add_(mul_(tol, fabs(x)), mintol);
        const double tol2 = mul_(2.0, tol1);
        const double xmid = mul_(0.5, add_(a, b));
        if (fabs(sub_(x, xmid)) < sub_(tol2, mul_(0.5, sub_(b, a)))) break;
        if (fabs(deltax) <= tol1) {
            deltax = (x >= xmid) ? sub_(a, x) : sub_(b, x);   // golden section step
            rat = mul_(cg, deltax);
        } else {                                     // parabolic step
            double tmp1 = mul_(sub_(x, w), sub_(fx, fv));
            double tmp2 = mul_(sub_(x, v), sub_(fx, fw));
            double p = sub_(mul_(sub_(x, v), tmp2), mul_(sub_(x, w), tmp1));
            tmp2 = mul_(2.0, sub_(tmp2, tmp1));
            if (tmp2 > 0.0) p = -p;
            tmp2 = fabs(tmp2);
            const double dx_temp = deltax;
            deltax = rat;
            if ((p > mul_(tmp2, sub_(a, x))) && (p < mul_(tmp2, sub_(b, x))) && (fabs(p) < fabs(mul_(mul_(0.5, tmp2), dx_temp)))) {
                rat = div_(p, tmp2);
                const double u = add_(x, rat);
                if (sub_(u, a) < tol2 || sub_(b, u) < tol2) rat = (sub_(xmid, x) >= 0) ? tol1 : -tol1;
            } else {
                deltax = (x >= xmid) ? sub_(a, x) : sub_(b, x);
                rat = mul_(cg, deltax);
            }
        }
        double u;
        if (fabs(rat) < tol1) u = (rat >= 0) ? add_(x, tol1) : sub_(x, tol1);
        else u = add_(x, rat);
        const double fu = func(u);
        if (aborted) { aborted_flag = true; return; }
        if (fu > fx) {
            if (u < x) a = u; else b = u;
            if ((fu <= fw) || (w == x)) {
                v = w; w = u; fv = fw; fw = fu;
            } else if ((fu <= fv) || (v == x) || (v == w)) {
                v = u; fv = fu;
            }
        } else {
            if (u >= x) a = x; else b = x;
            v = w; w = x; x = u;
            fv = fw; fw = fx; fx = fu;
        }
        ++iter;
    }
    xmin = x;
    fmin = fx;
}

// _linesearch_powell(func, p, xi, tol, fval=fval) without bounds: (fval, p, xi) <- (f(p + a xi), p + a xi, a xi)
template <class CF>
NLML_HD void linesearch(CF& cf, int n, double* p, double* xi, double tol, double& fval, bool& aborted_flag) {
    bool any = false;
    for (int i = 0; i < n; ++i) any = any || xi[i] != 0.0;
    if (!any) return;
    LineFn<CF> lf{cf, p, xi, n};
    double alpha = 0.0, fret = 0.0;
    brent_minimize(lf, tol, aborted_flag, cf.aborted, alpha, fret);
    if (aborted_flag) return;
    for (int i = 0; i < n; ++i) {
        xi[i] = mul_(alpha, xi[i]);
        p[i] = add_(p[i], xi[i]);
    }
    fval = fret;
}

struct Result {
    double fun;
    int nfev, nit, status;   // status: 0 converged, 1 maxfev, 2 maxiter, 3 nan
};

// _minimize_powell(func, x0): x holds x0 on entry and the solution on return; n <= kMaxN; direc is n x n scratch
template <class F>
NLML_HD Result minimize(F& f, int n, double* x, double* direc, double xtol = 1e-4, double ftol = 1e-4) {
    const int maxiter = n * 1000, maxfun = n * 1000;
    Counted<F> cf(f, maxfun);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) direc[i * n + j] = (i == j) ? 1.0 : 0.0;
    bool aborted = false;
    double fval = cf(x);
    double x1[kMaxN], direc1[kMaxN], x2[kMaxN];
    for (int i = 0; i < n; ++i) x1[i] = x[i];
    int iter = 0;
    while (true) {
        const double fx = fval;
        int bigind = 0;
        double delta = 0.0;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) direc1[j] = direc[i * n + j];
            const double fx2 = fval;
            linesearch(cf, n, x, direc1, xtol * 100, fval, aborted);
            if (aborted) break;
            if (sub_(fx2, fval) > delta) {
                delta = sub_(fx2, fval);
                bigind = i;
            }
        }
        if (aborted) break;
        ++iter;
        const double bnd = add_(mul_(ftol, add_(fabs(fx), fabs(fval))), 1e-20);
        if (mul_(2.0, sub_(fx, fval)) <= bnd) break;
        if (cf.calls >= maxfun) break;
        if (iter >= maxiter) break;
        if (fx != fx && fval != fval) break;
        // the extrapolated point
        for (int i = 0; i < n; ++i) {
            direc1[i] = sub_(x[i], x1[i]);
            x1[i] = x[i];
            x2[i] = add_(x[i], direc1[i]);
        }
        const double fx2 = cf(x2);
        if (cf.aborted) break;
        if (fx > fx2) {
            double t = mul_(2.0, sub_(add_(fx, fx2), mul_(2.0, fval)));
            double temp = sub_(sub_(fx, fval), delta);
            t = mul_(t, mul_(temp, temp));
            temp = sub_(fx, fx2);
            t = sub_(t, mul_(mul_(delta, temp), temp));
            if (t < 0.0) {
                linesearch(cf, n, x, direc1, xtol * 100, fval, aborted);
                if (aborted) break;
                bool any = false;
                for (int i = 0; i < n; ++i) any = any || direc1[i] != 0.0;
                if (any) {
                    for (int j = 0; j < n; ++j) {
                        direc[bigind * n + j] = direc[(n - 1) * n + j];
                        direc[(n - 1) * n + j] = direc1[j];
                    }
                }
            }
        }
    }
    Result r;
    r.fun = fval;
    r.nfev = cf.calls;
    r.nit = iter;
    bool xnan = false;
    for (int i = 0; i < n; ++i) xnan = xnan || x[i] != x[i];
    r.status = cf.calls >= maxfun ? 1 : (iter >= maxiter ? 2 : ((fval != fval || xnan) ? 3 : 0));
    return r;
}

// ---- objectives ----------------------------------------------------------------------------------------------------
// a cos(b w + c) + d with every operation rounded on its own, as numpy evaluates it (no fused multiply-add)
NLML_HD double cos_row(double a, double b, double c, double d, double w) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn(a, cos(__dadd_rn(__dmul_rn(b, w), c))), d);
#else
    volatile double arg = b * w;
    arg = arg + c;
    volatile double v = a * cos(arg);
    return v + d;
#endif
}

// TD_Tester.objective (/root/reference/TD_Tester.py:31-58) in float64 on the folded form:
//   f_a = float32(a cos(b w + c) + d) per factor row (:36-43), u float64;
//   0.5 ||x - W x1 u x2 f_y x3 f_p x4 f_r||^2 = 0.5 x.x - q.z + sum_ABCD S[A,B,C,D] UU_A YY_B PP_C RR_D
// q = W2 x and x.x accumulated in float64, S = the folded Gram tensor kept in float64.  Element r of q at q[r * qstride].
struct TuckerObjective {
    int ri, ry, rp, rr;
    const double* S;      // [nB*nC*nD][nA]
    const double* q;
    long long qstride;
    double half_xx;
    const double* rows_y; // [r][4] (a,b,c,d), float64 as the reference passes them (TD_Inference.py:43-45)
    const double* rows_p;
    const double* rows_r;
    NLML_HD double operator()(const double* p) const {
        double fy[16], fp[16], fr[16];
        for (int j = 0; j < ry; ++j) fy[j] = (double)(float)cos_row(rows_y[4 * j], rows_y[4 * j + 1], rows_y[4 * j + 2], rows_y[4 * j + 3], p[0]);
        for (int j = 0; j < rp; ++j) fp[j] = (double)(float)cos_row(rows_p[4 * j], rows_p[4 * j + 1], rows_p[4 * j + 2], rows_p[4 * j + 3], p[1]);
        for (int j = 0; j < rr; ++j) fr[j] = (double)(float)cos_row(rows_r[4 * j], rows_r[4 * j + 1], rows_r[4 * j + 2], rows_r[4 * j + 3], p[2]);
        const double* u = p + 3;
        // linear term
        double lin = 0.0;
        long long r = 0;
        for (int i = 0; i < ri; ++i)
            for (int j = 0; j < ry; ++j) {
                double sj = 0.0;
                for (int k = 0; k < rp; ++k) {
                    double sk = 0.0;
                    for (int l = 0; l < rr; ++l, ++r) sk = fma(q[r * qstride], fr[l], sk);
                    sj = fma(sk, fp[k], sj);
                }
                lin = fma(u[i] * fy[j], sj, lin);
            }
        // quadratic term on the folded Gram tensor
        const int nA = tri(ri);
        double UU[136];
        {
            int a = 0;
            for (int i = 0; i < ri; ++i)
                for (int j = i; j < ri; ++j) UU[a++] = u[i] * u[j];
        }
        double quad = 0.0;
        const double* row = S;
        for (int b0 = 0; b0 < ry; ++b0)
            for (int b1 = b0; b1 < ry; ++b1) {
                const double yy = fy[b0] * fy[b1];
                for (int c0 = 0; c0 < rp; ++c0)
                    for (int c1 = c0; c1 < rp; ++c1) {
                        const double yp = yy * (fp[c0] * fp[c1]);
                        for (int d0 = 0; d0 < rr; ++d0)
                            for (int d1 = d0; d1 < rr; ++d1, row += nA) {
                                double t = 0.0;
                                for (int a = 0; a < nA; ++a) t = fma(row[a], UU[a], t);
                                quad = fma(t, yp * (fr[d0] * fr[d1]), quad);
                            }
                    }
            }
        return half_xx - lin + quad;
    }
};

// numpy's pairwise summation of n contiguous float64 values (numpy/_core/src/umath/loops_utils.h.src, what np.sum runs):
// blocks of at most 128 elements are summed with eight strided accumulators, larger ranges are halved (the left half
// rounded down to a multiple of 8).  Reproduced operation for operation so that the objective below returns the
// reference's bits.
NLML_HD double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// The three factor vectors of TD_Tester.objective: float32(a cos(b w + c) + d) per row (:36-43), held as float64
struct TuckerFactors {
    double fy[16], fp[16], fr[16];
};
NLML_HD void tucker_factors(const double* p, int ry, int rp, int rr, const double* rows_y, const double* rows_p,
                            const double* rows_r, TuckerFactors& t) {
    for (int j = 0; j < ry; ++j) t.fy[j] = (double)(float)cos_row(rows_y[4 * j], rows_y[4 * j + 1], rows_y[4 * j + 2], rows_y[4 * j + 3], p[0]);
    for (int j = 0; j < rp; ++j) t.fp[j] = (double)(float)cos_row(rows_p[4 * j], rows_p[4 * j + 1], rows_p[4 * j + 2], rows_p[4 * j + 3], p[1]);
    for (int j = 0; j < rr; ++j) t.fr[j] = (double)(float)cos_row(rows_r[4 * j], rows_r[4 * j + 1], rows_r[4 * j + 2], rows_r[4 * j + 3], p[2]);
}
// (x[m] - x_hat[m])^2 for one feature m, x_hat accumulated as np.einsum('ijklm,i,j,k,l->m', W, u, f_y, f_p, f_r) does it
// (:46): float64, over (i,j,k,l) in storage order, each term ((((W u_i) f_y) f_p) f_r), every operation rounded on its own.
// wcol: W2[0][m], element r at wcol[r * wstride].
NLML_HD double residual_sq(const float* wcol, long long wstride, float xm, const double* u, const TuckerFactors& t, int ri,
                           int ry, int rp, int rr) {
    double acc = 0.0;
    long long r = 0;
    for (int i = 0; i < ri; ++i)
        for (int j = 0; j < ry; ++j)
            for (int k = 0; k < rp; ++k)
                for (int l = 0; l < rr; ++l, ++r) {
#if defined(__CUDA_ARCH__)
                    const double term = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn((double)wcol[r * wstride], u[i]), t.fy[j]), t.fp[k]), t.fr[l]);
                    acc = __dadd_rn(acc, term);
#else
                    volatile double term = (double)wcol[r * wstride] * u[i];
                    term = term * t.fy[j];
                    term = term * t.fp[k];
                    term = term * t.fr[l];
                    acc = acc + term;
#endif
                }
#if defined(__CUDA_ARCH__)
    const double d = __dsub_rn((double)xm, acc);
    return __dmul_rn(d, d);
#else
    volatile double d = (double)xm - acc;
    volatile double sq = d * d;
    return sq;
#endif
}

// TD_Tester.objective (:31-58) evaluated the reference's way, bit for bit: error = 0.5 * np.sum((x - x_hat) ** 2).
// Sequential form (host check build; the CUDA kernel evaluates the same operations cooperatively, one CTA per sample).
struct TuckerObjectiveExact {
    int ri, ry, rp, rr, F;
    const float* W2;      // [R][F]
    const float* x;       // [F]
    const double* rows_y;
    const double* rows_p;
    const double* rows_r;
    double* scratch;      // [F]
    NLML_HD double operator()(const double* p) const {
        TuckerFactors t;
        tucker_factors(p, ry, rp, rr, rows_y, rows_p, rows_r, t);
        for (int m = 0; m < F; ++m) scratch[m] = residual_sq(W2 + m, F, x[m], p + 3, t, ri, ry, rp, rr);
        return 0.5 * np_pairwise_sum(scratch, F);
    }
};

// TD_Trainer.objective (/root/reference/TD_Trainer.py:38-43): 0.5 * sum_i (U_ij - (a cos(b rad(w_i) + c) + d))^2
struct CosineObjective {
    const double* U;       // column j of the factor matrix: element i at U[i * stride]
    const double* w_rad;   // np.radians(w)
    int n, stride;
    NLML_HD double operator()(const double* p) const {
        double e = 0.0;
        for (int i = 0; i < n; ++i) {
            const double d = U[i * stride] - cos_row(p[0], p[1], p[2], p[3], w_rad[i]);
#if defined(__CUDA_ARCH__)
            e = __dadd_rn(e, __dmul_rn(d, d));
#else
            volatile double sq = d * d;
            e = e + sq;
#endif
        }
        return e / 2;
    }
};

// one entry of the folded Gram tensor in float64 (tucker_math.h fold_entry without the final cast)
NLML_HD double fold_entry_f64(const double* M, int ri, int ry, int rp, int rr, int A, int B, int C, int D) {
    const int R = ri * ry * rp * rr;
    int i0, i1, j0, j1, k0, k1, l0, l1;
    unpair(A, ri, &i0, &i1);
    unpair(B, ry, &j0, &j1);
    unpair(C, rp, &k0, &k1);
    unpair(D, rr, &l0, &l1);
    double acc = 0.0;
    for (int si = 0; si < (i0 == i1 ? 1 : 2); ++si)
        for (int sj = 0; sj < (j0 == j1 ? 1 : 2); ++sj)
            for (int sk = 0; sk < (k0 == k1 ? 1 : 2); ++sk)
                for (int sl = 0; sl < (l0 == l1 ? 1 : 2); ++sl) {
                    const int ia = si ? i1 : i0, ib = si ? i0 : i1;
                    const int ja = sj ? j1 : j0, jb = sj ? j0 : j1;
                    const int ka = sk ? k1 : k0, kb = sk ? k0 : k1;
                    const int la = sl ? l1 : l0, lb = sl ? l0 : l1;
                    const long long r = ((ia * ry + ja) * rp + ka) * rr + la;
                    const long long c = ((ib * ry + jb) * rp + kb) * rr + lb;
                    acc += M[r * R + c];
                }
    return 0.5 * acc;
}

// TD_Trainer.est_params_by_Uniform_Fourier (/root/reference/TD_Trainer.py:125-148): initial (a, b, c, d) of one column from
// the dominant non-zero frequency of its discrete Fourier transform.  w_rad: the (uniformly spaced) bins in radians.
// The real input's spectrum is conjugate-symmetric, so the first maximum of |X_k| over k >= 1 lies in k <= n/2.
NLML_HD void fourier_init(const double* U, int stride, const double* w_rad, int n, double* out4) {
    const double two_pi = 6.283185307179586476925286766559;
    const double dstep = w_rad[1] - w_rad[0];
    double best = -1.0, best_re = 0.0, best_im = 0.0;
    int best_k = 1;
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += U[i * stride];
    mean /= n;
    for (int k = 1; k <= n / 2; ++k) {
        double re = 0.0, im = 0.0;
        for (int i = 0; i < n; ++i) {
            const double ang = -two_pi * (double)((long long)k * i % n) / n;
            re += U[i * stride] * cos(ang);
            im += U[i * stride] * sin(ang);
        }
        const double mag = sqrt(re * re + im * im);
        if (mag > best) {
            best = mag; best_re = re; best_im = im; best_k = k;
        }
    }
    const double freq = (double)best_k / (n * dstep);   // np.fft.fftfreq(n, d)[k] for k <= (n-1)/2 (|.| for the Nyquist bin)
    out4[0] = 2.0 * best / n;
    out4[1] = two_pi * fabs(freq);
    out4[2] = atan2(best_im, best_re);
    out4[3] = mean;
}

}  // namespace powell
}  // namespace nlml
