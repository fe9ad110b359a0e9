// Host build of the per-sample Tucker-fit arithmetic (nlml_hpe_b200/csrc/tucker_math.h).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only test tier execute the very statements the CUDA
// kernels run (folded Gram tensor, gradient, clip, step) against the reference goldens.
// The product never loads this library.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../nlml_hpe_b200/csrc/tucker_math.h"
#include "../../nlml_hpe_b200/csrc/powell_math.h"

using namespace nlml;

extern "C" int hostcheck_tucker_fit_5333(const float* W2, int F, const double* rows_y, const double* rows_p,
                                         const double* rows_r, const float* X, int64_t N, int64_t ldx, int T,
                                         float lr, float clip, float* P /*[N][8]*/) {
    constexpr int RI = 5, RY = 3, RP = 3, RR = 3, R = RI * RY * RP * RR, NP = 3 + RI;
    constexpr int nA = tri(RI), NAP = (nA + 3) / 4 * 4, nBCD = tri(RY) * tri(RP) * tri(RR);
    std::vector<double> M((size_t)R * R);
    for (int r = 0; r < R; ++r)
        for (int c = r; c < R; ++c) M[(size_t)r * R + c] = M[(size_t)c * R + r] = gram_entry(W2, F, r, c);
    std::vector<float> S((size_t)nBCD * NAP, 0.f);
    for (int a = 0; a < nA; ++a)
        for (int b = 0; b < tri(RY); ++b)
            for (int c = 0; c < tri(RP); ++c)
                for (int d = 0; d < tri(RR); ++d)
                    S[(size_t)((b * tri(RP) + c) * tri(RR) + d) * NAP + a] = fold_entry(M.data(), RI, RY, RP, RR, a, b, c, d);
    float ry[12], rp[12], rr[12];
    for (int i = 0; i < 12; ++i) { ry[i] = (float)rows_y[i]; rp[i] = (float)rows_p[i]; rr[i] = (float)rows_r[i]; }
    for (int64_t s = 0; s < N; ++s) {
        float q[R];
        for (int r = 0; r < R; ++r) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(W2[(size_t)r * F + f], X[s * ldx + f], acc);
            q[r] = acc;
        }
        float p[1][NP] = {{0}};
        for (int it = 0; it < T; ++it) {
            float g[1][NP];
            float scr[tri(RY) * tri(RP)];
            tucker_gradient<RI, RY, RP, RR, NAP, 1>(p, S.data(), QStrided{q, 1, 0}, scr, 1, 0, ry, rp, rr, g);
            clip_and_step<NP>(p[0], g[0], lr, clip);
        }
        for (int i = 0; i < NP; ++i) P[s * NP + i] = p[0][i];
    }
    return 0;
}

// gradient only, at given parameter points (checked against autograd of the reference objective)
extern "C" int hostcheck_tucker_grad_5333(const float* W2, int F, const double* rows_y, const double* rows_p,
                                          const double* rows_r, const float* X, int64_t N, int64_t ldx,
                                          const float* Pin /*[N][8]*/, float* G /*[N][8]*/) {
    constexpr int RI = 5, RY = 3, RP = 3, RR = 3, R = RI * RY * RP * RR, NP = 3 + RI;
    constexpr int nA = tri(RI), NAP = (nA + 3) / 4 * 4, nBCD = tri(RY) * tri(RP) * tri(RR);
    std::vector<double> M((size_t)R * R);
    for (int r = 0; r < R; ++r)
        for (int c = r; c < R; ++c) M[(size_t)r * R + c] = M[(size_t)c * R + r] = gram_entry(W2, F, r, c);
    std::vector<float> S((size_t)nBCD * NAP, 0.f);
    for (int a = 0; a < nA; ++a)
        for (int b = 0; b < tri(RY); ++b)
            for (int c = 0; c < tri(RP); ++c)
                for (int d = 0; d < tri(RR); ++d)
                    S[(size_t)((b * tri(RP) + c) * tri(RR) + d) * NAP + a] = fold_entry(M.data(), RI, RY, RP, RR, a, b, c, d);
    float ry[12], rp[12], rr[12];
    for (int i = 0; i < 12; ++i) { ry[i] = (float)rows_y[i]; rp[i] = (float)rows_p[i]; rr[i] = (float)rows_r[i]; }
    for (int64_t s = 0; s < N; ++s) {
        float q[R];
        for (int r = 0; r < R; ++r) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(W2[(size_t)r * F + f], X[s * ldx + f], acc);
            q[r] = acc;
        }
        float scr[2][tri(RY) * tri(RP)];
        // two samples per call (the same sample twice) exercises the NS=2 path the GPU kernel uses
        float pin[2][NP], gout[2][NP];
        for (int i = 0; i < NP; ++i) pin[0][i] = pin[1][i] = Pin[s * NP + i];
        float q2[2][R];
        for (int r = 0; r < R; ++r) q2[0][r] = q2[1][r] = q[r];
        tucker_gradient<RI, RY, RP, RR, NAP, 2>(pin, S.data(), QStrided{&q2[0][0], 1, R}, &scr[0][0], 1, tri(RY) * tri(RP), ry, rp, rr, gout);
        for (int i = 0; i < NP; ++i) G[s * NP + i] = gout[1][i];
        for (int i = 0; i < NP; ++i) if (gout[0][i] != gout[1][i]) return 1;
    }
    return 0;
}

// sin/cos of the tensor-core kernel (tucker_math.h sincos_small), exposed so its accuracy can be pinned against libm.
extern "C" void hostcheck_sincos_small(const float* x, int64_t n, float* sn, float* cs) {
    for (int64_t i = 0; i < n; ++i) nlml::sincos_small(x[i], sn + i, cs + i);
}

namespace {
struct Prepared5333 {
    static constexpr int RI = 5, RY = 3, RP = 3, RR = 3, R = RI * RY * RP * RR, NP = 3 + RI;
    static constexpr int nA = tri(RI), NAP = (nA + 3) / 4 * 4, nBCD = tri(RY) * tri(RP) * tri(RR);
    std::vector<float> S;
    float ry[12], rp[12], rr[12];
    Prepared5333(const float* W2, int F, const double* rows_y, const double* rows_p, const double* rows_r)
        : S((size_t)nBCD * NAP, 0.f) {
        std::vector<double> M((size_t)R * R);
        for (int r = 0; r < R; ++r)
            for (int c = r; c < R; ++c) M[(size_t)r * R + c] = M[(size_t)c * R + r] = gram_entry(W2, F, r, c);
        for (int a = 0; a < nA; ++a)
            for (int b = 0; b < tri(RY); ++b)
                for (int c = 0; c < tri(RP); ++c)
                    for (int d = 0; d < tri(RR); ++d)
                        S[(size_t)((b * tri(RP) + c) * tri(RR) + d) * NAP + a] = fold_entry(M.data(), RI, RY, RP, RR, a, b, c, d);
        for (int i = 0; i < 12; ++i) { ry[i] = (float)rows_y[i]; rp[i] = (float)rows_p[i]; rr[i] = (float)rows_r[i]; }
    }
    static void project(const float* W2, int F, const float* x, float* q) {
        for (int r = 0; r < R; ++r) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(W2[(size_t)r * F + f], x[f], acc);
            q[r] = acc;
        }
    }
};
}  // namespace

// value (without 0.5 x.x), gradient and packed lower-triangular Hessian at given points (tucker_newton_eval)
extern "C" int hostcheck_tucker_newton_5333(const float* W2, int F, const double* rows_y, const double* rows_p,
                                            const double* rows_r, const float* X, int64_t N, int64_t ldx,
                                            const float* Pin /*[N][8]*/, float* L /*[N]*/, float* G /*[N][8]*/,
                                            float* H /*[N][36]*/) {
    using C = Prepared5333;
    C pre(W2, F, rows_y, rows_p, rows_r);
    for (int64_t s = 0; s < N; ++s) {
        float q[C::R], scr[3 * (tri(C::RY) + tri(C::RP))];
        C::project(W2, F, X + s * ldx, q);
        float p[C::NP], g[C::NP], h[C::NP * (C::NP + 1) / 2], l;
        for (int i = 0; i < C::NP; ++i) p[i] = Pin[s * C::NP + i];
        tucker_newton_eval<C::RI, C::RY, C::RP, C::RR, C::NAP>(p, pre.S.data(), q, 1, scr, 1, pre.ry, pre.rp, pre.rr, l, g, h);
        L[s] = l;
        for (int i = 0; i < C::NP; ++i) G[s * C::NP + i] = g[i];
        for (int i = 0; i < 36; ++i) H[s * 36 + i] = h[i];
    }
    return 0;
}

// converged fit (tucker_lm_solve) from p = 0
extern "C" int hostcheck_tucker_solve_5333(const float* W2, int F, const double* rows_y, const double* rows_p,
                                           const double* rows_r, const float* X, int64_t N, int64_t ldx,
                                           int max_evals, float* P /*[N][8]*/, int* evals /*[N]*/,
                                           const float* opts /*optional [8]: lambda0, down, up, cap, tol, floor, noise_step, lambda_min*/,
                                           float* Lout /*optional [N]*/) {
    using C = Prepared5333;
    C pre(W2, F, rows_y, rows_p, rows_r);
    LmOptions o = lm_default_options();
    if (max_evals > 0) o.max_evals = max_evals;
    if (opts) {
        o.lambda0 = opts[0]; o.lambda_down = opts[1]; o.lambda_up = opts[2]; o.angle_cap = opts[3];
        o.step_tol = opts[4]; o.diag_floor = opts[5]; o.noise_step = opts[6]; o.lambda_min = opts[7];
    }
    for (int64_t s = 0; s < N; ++s) {
        float q[C::R], scr[3 * (tri(C::RY) + tri(C::RP))];
        C::project(W2, F, X + s * ldx, q);
        float p[C::NP], l;
        evals[s] = tucker_lm_solve<C::RI, C::RY, C::RP, C::RR, C::NAP>(pre.S.data(), q, 1, scr, 1, pre.ry, pre.rp, pre.rr, o, p, l);
        for (int i = 0; i < C::NP; ++i) P[s * C::NP + i] = p[i];
        if (Lout) Lout[s] = l;
    }
    return 0;
}

// packed 8x8 Cholesky solve of the converged solver (tucker_math.h chol_solve): A lower triangle packed row-major,
// returns 1 when every pivot was positive
extern "C" int hostcheck_chol_solve8(const float* A /*[36]*/, const float* g /*[8]*/, float* d /*[8]*/) {
    float a[36], gg[8], dd[8];
    for (int i = 0; i < 36; ++i) a[i] = A[i];
    for (int i = 0; i < 8; ++i) gg[i] = g[i];
    const bool ok = nlml::chol_solve<8>(a, gg, dd);
    for (int i = 0; i < 8; ++i) d[i] = dd[i];
    return ok ? 1 : 0;
}


// ---- scipy-Powell restatement (powell_math.h) on the CPU ------------------------------------------------------------
// TD_Tester.Test's search: Powell from p = 0 over the float64 objective, run-time ranks.
extern "C" int hostcheck_powell_tucker(const float* W2, int ri, int ry, int rp, int rr, int F, const double* rows_y,
                                       const double* rows_p, const double* rows_r, const float* X, int64_t N, int64_t ldx,
                                       double* P /*[N][3+ri]*/, double* fun /*[N]*/, int* nfev /*[N]*/, int exact_mode) {
    const int R = ri * ry * rp * rr, nA = tri(ri), nB = tri(ry), nC = tri(rp), nD = tri(rr), NP = 3 + ri;
    std::vector<double> M((size_t)R * R);
    for (int r = 0; r < R; ++r)
        for (int c = r; c < R; ++c) M[(size_t)r * R + c] = M[(size_t)c * R + r] = gram_entry(W2, F, r, c);
    std::vector<double> S((size_t)nB * nC * nD * nA);
    for (int b = 0; b < nB; ++b)
        for (int c = 0; c < nC; ++c)
            for (int d = 0; d < nD; ++d)
                for (int a = 0; a < nA; ++a)
                    S[(size_t)((b * nC + c) * nD + d) * nA + a] = powell::fold_entry_f64(M.data(), ri, ry, rp, rr, a, b, c, d);
    std::vector<double> q(R), direc((size_t)NP * NP), scratch(F);
    for (int64_t s = 0; s < N; ++s) {
        if (exact_mode) {   // exact_mode: the reference's own evaluation order (what the CUDA kernel runs)
            powell::TuckerObjectiveExact obj{ri, ry, rp, rr, F, W2, X + s * ldx, rows_y, rows_p, rows_r, scratch.data()};
            double x[powell::kMaxN] = {0};
            const powell::Result res = powell::minimize(obj, NP, x, direc.data());
            for (int i = 0; i < NP; ++i) P[s * NP + i] = x[i];
            fun[s] = res.fun;
            nfev[s] = res.nfev;
            continue;
        }
        double xx = 0.0;
        for (int f = 0; f < F; ++f) xx += (double)X[s * ldx + f] * (double)X[s * ldx + f];
        for (int r = 0; r < R; ++r) {
            double acc = 0.0;
            for (int f = 0; f < F; ++f) acc += (double)W2[(size_t)r * F + f] * (double)X[s * ldx + f];
            q[r] = acc;
        }
        powell::TuckerObjective obj{ri, ry, rp, rr, S.data(), q.data(), 1, 0.5 * xx, rows_y, rows_p, rows_r};
        double x[powell::kMaxN] = {0};
        const powell::Result res = powell::minimize(obj, NP, x, direc.data());
        for (int i = 0; i < NP; ++i) P[s * NP + i] = x[i];
        fun[s] = res.fun;
        nfev[s] = res.nfev;
    }
    return 0;
}

// TD_Trainer.Train for one factor matrix: Fourier initial guess + Powell per column.  U [n_rows][n_cols] row-major.
extern "C" int hostcheck_cosine_fit(const double* U, int n_rows, int n_cols, const double* w_deg, double* init /*[n_cols][4]*/,
                                    double* out /*[n_cols][4]*/, double* fun, int* nfev) {
    std::vector<double> w(n_rows);
    for (int i = 0; i < n_rows; ++i) w[i] = w_deg[i] * (3.141592653589793238462643383279502884 / 180.0);   // np.radians
    for (int j = 0; j < n_cols; ++j) {
        powell::fourier_init(U + j, n_cols, w.data(), n_rows, init + 4 * j);
        powell::CosineObjective obj{U + j, w.data(), n_rows, n_cols};
        double x[powell::kMaxN], direc[16];
        for (int i = 0; i < 4; ++i) x[i] = init[4 * j + i];
        const powell::Result res = powell::minimize(obj, 4, x, direc);
        for (int i = 0; i < 4; ++i) out[4 * j + i] = x[i];
        fun[j] = res.fun;
        nfev[j] = res.nfev;
    }
    return 0;
}

// the exact objective alone, at given points of one sample
extern "C" int hostcheck_powell_objective(const float* W2, int ri, int ry, int rp, int rr, int F, const double* rows_y, const double* rows_p,
                                          const double* rows_r, const float* x, const double* pts, int npts, double* vals) {
    std::vector<double> scratch(F);
    powell::TuckerObjectiveExact obj{ri, ry, rp, rr, F, W2, x, rows_y, rows_p, rows_r, scratch.data()};
    for (int i = 0; i < npts; ++i) vals[i] = obj(pts + (size_t)i * (3 + ri));
    return 0;
}
