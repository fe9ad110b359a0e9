// Per-sample arithmetic of the fixed-iteration Tucker fit, shared by the CUDA kernels
// (tucker_fit.cu) and by the host-side check build (tests/hostcheck) so the exact same
// statements can be exercised on a CPU-only box.
//
// Reference recurrence: TD_Tester.optimize_with_sgd, /root/reference/TD_Tester.py:127-159
//   p = 0 (:130);  repeat T times:  g = d/dp objective_torch(p) (:139,:145);
//   g *= min(1, clip/(||g||_2 + 1e-6)) (:150);  p -= lr*g (:153-154)
// with objective_torch (:110-125)  L(p) = 0.5*|| x - W x1 u x2 c_y(w_y) x3 c_p(w_p) x4 c_r(w_r) ||^2
// and c_a(w)[j] = a_j cos(b_j w + c_j) + d_j (:105-107).
//
// Algebra used here (exact in real arithmetic):
//   z = u (x) c_y (x) c_p (x) c_r  in R^R,  W2 = W reshaped [R,F],  M = W2 W2^T,  q = W2 x
//   L = 0.5 x.x - q.z + 0.5 z^T M z
// and because z is a Kronecker product, z^T M z is a polynomial that only sees the symmetric
// second-order monomials  UU_A = u_i u_i' (i<=i'), YY_B, PP_C, RR_D:
//   0.5 z^T M z = sum_{A,B,C,D} S[A,B,C,D] UU_A YY_B PP_C RR_D
// S ("folded Gram tensor") has nA*nB*nC*nD entries (15*6*6*6 = 3240 at ranks 5,3,3,3, against
// 135*135 = 18225 for M).  One iteration needs both  T[BCD] = sum_A S[A,BCD] UU_A  and
// GU[A] = sum_BCD S[A,BCD] YY_B PP_C RR_D ; everything else is O(R).
#pragma once

#if defined(__CUDACC__)
#define NLML_HD __host__ __device__ __forceinline__
#else
#define NLML_HD inline
#include <cmath>
struct alignas(16) float4 { float x, y, z, w; };   // host check build only
#endif

namespace nlml {

NLML_HD constexpr int tri(int r) { return r * (r + 1) / 2; }

// index of the unordered pair (i<=j) among the tri(r) pairs, row-major over the upper triangle
NLML_HD constexpr int pair_index(int i, int j, int r) { return i * r - i * (i - 1) / 2 + (j - i); }

NLML_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
NLML_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b;
    return r;
#endif
}

// Cosine factor rows: rows[j*4 + {0,1,2,3}] = (a,b,c,d).  c[j] = a cos(bw+c)+d, dc[j] = -a b sin(bw+c).
template <int R>
NLML_HD void cos_features(float w, const float* rows, float* c, float* dc) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const float a = rows[4 * j + 0], b = rows[4 * j + 1], ph = rows[4 * j + 2], d = rows[4 * j + 3];
        float s, co;
        const float arg = b * w + ph;
#if defined(__CUDA_ARCH__)
        sincosf(arg, &s, &co);
#else
        s = sinf(arg);
        co = cosf(arg);
#endif
        c[j] = a * co + d;
        dc[j] = -(a * b) * s;
    }
}

// sin and cos of a moderate argument (|x| < ~1e4) without the large-argument slow path of sincosf: three-term
// Cody-Waite reduction by pi/2 and the classic single-precision minimax kernels on [-pi/4, pi/4] (~1 ulp).
// Used by the tensor-core kernel, where nine inlined sincosf slow paths would push the iteration body out of
// the instruction cache.  The factor arguments b*w + c stay within a few radians.
NLML_HD void sincos_small(float x, float* sn, float* cs) {
    const float kf = rintf(x * 0.636619772367581343f);   // x * 2/pi
    float r = fmaf(-kf, 1.5703125f, x);
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    const float r2 = r * r;
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, r2, -1.6666654611e-1f);
    const float s = fmaf(ps * r2, r, r);
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, r2, 4.166664568298827e-2f);
    const float c = fmaf(pc * r2, r2, fmaf(-0.5f, r2, 1.0f));
    const int q = (int)kf;
    const float a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

// The same rows through the special-function unit: sin.approx / cos.approx (one multiply by 1/2pi shared + one MUFU each)
// instead of ~26 instructions per sincos_small.  Absolute error <= 2^-21.4 + |x| 2^-24 (~1e-6 for the few radians the
// arguments b*w + c span) against ~6e-8: a systematic 1e-6 relative perturbation of the gradient moves the T = 3000
// result by ~1e-4 degrees (measured, tests/test_tucker_gpu.py) inside a 1e-2 degree budget.  Device only; used by the
// tensor-core iteration kernel, where nine sincos per thread were a quarter of the iteration's instructions.
#if defined(__CUDACC__)
template <int R>
__device__ __forceinline__ void cos_features_sfu(float w, const float* rows, float* c, float* dc) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const float a = rows[4 * j + 0], b = rows[4 * j + 1], ph = rows[4 * j + 2], d = rows[4 * j + 3];
        const float arg = fmaf(b, w, ph);
        c[j] = fmaf(a, __cosf(arg), d);
        dc[j] = -(a * b) * __sinf(arg);
    }
}
// values only (the caller does not differentiate along this angle)
template <int R>
__device__ __forceinline__ void cos_values_sfu(float w, const float* rows, float* c) {
#pragma unroll
    for (int j = 0; j < R; ++j) c[j] = fmaf(rows[4 * j + 0], __cosf(fmaf(rows[4 * j + 1], w, rows[4 * j + 2])), rows[4 * j + 3]);
}
#endif

template <int R>
NLML_HD void cos_features_fast(float w, const float* rows, float* c, float* dc) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const float a = rows[4 * j + 0], b = rows[4 * j + 1], ph = rows[4 * j + 2], d = rows[4 * j + 3];
        float s, co;
        sincos_small(b * w + ph, &s, &co);
        c[j] = a * co + d;
        dc[j] = -(a * b) * s;
    }
}

// symmetric second-order monomials v_i v_j (i<=j), packed with pair_index
template <int R>
NLML_HD void sym_products(const float* v, float* vv) {
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = i; j < R; ++j) vv[pair_index(i, j, R)] = v[i] * v[j];
}

// d/dv_m of sum_A G[A] vv_A  =  2 G[(m,m)] v_m + sum_{i != m} G[(min,max)] v_i
template <int R>
NLML_HD void sym_backprop(const float* G, const float* v, float* dv) {
#pragma unroll
    for (int m = 0; m < R; ++m) {
        float acc = 2.0f * G[pair_index(m, m, R)] * v[m];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (i == m) continue;
            const int lo = i < m ? i : m, hi = i < m ? m : i;
            acc = fmaf(G[pair_index(lo, hi, R)], v[i], acc);
        }
        dv[m] = acc;
    }
}

// Joint L2 clip of torch.nn.utils.clip_grad_norm_ (TD_Tester.py:150) followed by p -= lr*g (:153-154).
template <int NP>
NLML_HD void clip_and_step(float* p, float* g, float lr, float clip) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) ss = fmaf(g[i], g[i], ss);
    // coef = min(1, clip / (sqrt(ss) + 1e-6)) is exactly 1 whenever sqrt(ss) + 1e-6 <= clip; the test below is 1 %
    // inside that boundary, so the square root and the division are only evaluated while the clip is (nearly) active
    const float lim = clip - 1e-6f;
    float coef = 1.0f;
    if (!(lim > 0.f && ss < 0.98f * lim * lim)) {
        const float norm = sqrtf(ss);
        coef = clip / (norm + 1e-6f);
        coef = coef < 1.0f ? coef : 1.0f;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = sub_rn(p[i], mul_rn(lr, mul_rn(g[i], coef)));
}

// The same update with a short dependency chain, for the tensor-core iteration kernel (every warp of the CTA sits in
// this step at once, right after the CTA barrier, so its latency is exposed): pairwise sum of squares, rsqrt and a
// reciprocal from the special-function unit (2 ulp on the clip coefficient: a 1e-7 relative change of the step length).
#if defined(__CUDACC__)
template <int NP>
__device__ __forceinline__ void clip_and_step_fast(float* p, const float* g, float lr, float clip) {
    static_assert(NP == 8, "pairwise tree written for 8 parameters");
    const float s01 = fmaf(g[1], g[1], g[0] * g[0]), s23 = fmaf(g[3], g[3], g[2] * g[2]);
    const float s45 = fmaf(g[5], g[5], g[4] * g[4]), s67 = fmaf(g[7], g[7], g[6] * g[6]);
    const float ss = (s01 + s23) + (s45 + s67);
    const float lim = clip - 1e-6f;
    float coef = 1.0f;
    if (!(lim > 0.f && ss < 0.98f * lim * lim)) {
        const float norm = ss > 0.f ? ss * rsqrtf(ss) : 0.f;
        coef = fminf(__fdividef(clip, norm + 1e-6f), 1.0f);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = sub_rn(p[i], mul_rn(lr, mul_rn(g[i], coef)));
}
#endif

// q accessor for memory-resident q: element r of sample n at q[n*sample + r*stride] (a shared-memory column on
// the GPU, a plain array on the host).  The thread-per-sample kernel's other accessor reads q from tensor memory.
struct QStrided {
    const float* q;
    int stride, sample;
    template <int R0, int CNT>
    NLML_HD void load(int n, float (&v)[32]) const {
#pragma unroll
        for (int x = 0; x < CNT; ++x) v[x] = q[n * sample + (R0 + x) * stride];
    }
};

// q accessor for the vectorised layout [R/4][pitch] of float4 (element r of the thread's sample = component r%4 of
// q4[(r/4)*pitch]): eight 128-bit loads per 32 elements instead of 32 scalar ones.
struct QVec4 {
    const float4* q4;
    int pitch;
    template <int R0, int CNT>
    NLML_HD void load(int /*n*/, float (&v)[32]) const {
        static_assert(R0 % 4 == 0, "chunks start on a float4 boundary");
#pragma unroll
        for (int x4 = 0; x4 < (CNT + 3) / 4; ++x4) {
            const float4 t = q4[(R0 / 4 + x4) * pitch];
            v[4 * x4 + 0] = t.x; v[4 * x4 + 1] = t.y; v[4 * x4 + 2] = t.z; v[4 * x4 + 3] = t.w;
        }
    }
};

// Static walk over q in chunks of 32: chunk C covers r in [32C, min(32C+32, R)), r = i*JKL + jkl.
template <int RI, int JKL, class QA, int C>
struct QLoadChunk {
    static constexpr int R = RI * JKL, R0 = 32 * C, CNT = (R - R0) < 32 ? (R - R0) : 32;
    NLML_HD static void run(const QA& qa, int n, const float (&tq)[JKL], const float* u, float (&lin_u)[RI], float (&e)[JKL]) {
        if constexpr (R0 < R) {
            float v[32];
            qa.template load<R0, CNT>(n, v);
#pragma unroll
            for (int x = 0; x < CNT; ++x) {
                constexpr int dummy = 0; (void)dummy;
                const int r = R0 + x, i = r / JKL, jkl = r % JKL;   // compile-time after unrolling
                lin_u[i] = fmaf(v[x], tq[jkl], lin_u[i]);
                e[jkl] = fmaf(v[x], u[i], e[jkl]);
            }
            QLoadChunk<RI, JKL, QA, C + 1>::run(qa, n, tq, u, lin_u, e);
        }
    }
};

// Linear term -q.z : lin_u[i] = sum_jkl q[ijkl] t_jkl with t = cy (x) cp (x) cr, and the contractions of
// e[jkl] = sum_i u_i q[ijkl] with the other two angle factors (ey, ep, er).  q is consumed in its storage order
// (i slowest), 32 consecutive elements at a time.
template <int RI, int RY, int RP, int RR, class QA>
NLML_HD void linear_term(const QA& qa, int n, const float* cy, const float* cp, const float* cr, const float* u,
                         float (&lin_u)[RI], float (&ey)[RY], float (&ep)[RP], float (&er)[RR]) {
    constexpr int JKL = RY * RP * RR;
    float tq[JKL], e[JKL];
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
        for (int k = 0; k < RP; ++k)
#pragma unroll
            for (int l = 0; l < RR; ++l) {
                tq[(j * RP + k) * RR + l] = cy[j] * cp[k] * cr[l];
                e[(j * RP + k) * RR + l] = 0.f;
            }
#pragma unroll
    for (int i = 0; i < RI; ++i) lin_u[i] = 0.f;
    QLoadChunk<RI, JKL, QA, 0>::run(qa, n, tq, u, lin_u, e);
#pragma unroll
    for (int j = 0; j < RY; ++j) ey[j] = 0.f;
#pragma unroll
    for (int k = 0; k < RP; ++k) ep[k] = 0.f;
#pragma unroll
    for (int l = 0; l < RR; ++l) er[l] = 0.f;
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
        for (int k = 0; k < RP; ++k) {
            const float yk = cy[j] * cp[k];
            float e_jk = 0.f;  // sum_l e[jkl] cr_l
#pragma unroll
            for (int l = 0; l < RR; ++l) {
                const float ev = e[(j * RP + k) * RR + l];
                er[l] = fmaf(ev, yk, er[l]);
                e_jk = fmaf(ev, cr[l], e_jk);
            }
            ey[j] = fmaf(e_jk, cp[k], ey[j]);
            ep[k] = fmaf(e_jk, cy[j], ep[k]);
        }
}

// Chain rule from the derivatives w.r.t. the monomials (GU, GY, GP, GR) and the linear term to d/dp.
template <int RI, int RY, int RP, int RR>
NLML_HD void assemble_gradient(const float* GU, const float* GY, const float* GP, const float* GR, const float* u,
                               const float* cy, const float* cp, const float* cr, const float* dcy, const float* dcp,
                               const float* dcr, const float* lin_u, const float* ey, const float* ep, const float* er,
                               float* g) {
    float du[RI], dy[RY], dp[RP], dr[RR];
    sym_backprop<RI>(GU, u, du);
    sym_backprop<RY>(GY, cy, dy);
    sym_backprop<RP>(GP, cp, dp);
    sym_backprop<RR>(GR, cr, dr);
    float gy = 0.f, gp = 0.f, gr = 0.f;
#pragma unroll
    for (int j = 0; j < RY; ++j) gy = fmaf(dy[j] - ey[j], dcy[j], gy);
#pragma unroll
    for (int k = 0; k < RP; ++k) gp = fmaf(dp[k] - ep[k], dcp[k], gp);
#pragma unroll
    for (int l = 0; l < RR; ++l) gr = fmaf(dr[l] - er[l], dcr[l], gr);
    g[0] = gy;
    g[1] = gp;
    g[2] = gr;
#pragma unroll
    for (int i = 0; i < RI; ++i) g[3 + i] = du[i] - lin_u[i];
}

// One full gradient evaluation for NS samples held by one thread (fixed small ranks).
//   S : folded Gram tensor laid out [nB*nC*nD][NAP] (NAP = nA padded to a multiple of 4), read-only, identical
//       for all samples (shared-memory broadcast on the GPU).  Every element loaded from S feeds 2*NS FMAs,
//       which is what keeps the kernel off the shared-memory wavefront limit.
//   qa : accessor for q = W2 x; qa.load<R0,CNT>(n, v) fetches q[R0 .. R0+CNT) of sample n, 32 at a time.
//   scr : nB*nC floats of scratch per sample, element k of sample n at scr[n*ssample + k*sstride]
//         (a shared-memory column on the GPU).
//   rows_* : cosine rows of the three angle factors.
// Reads p[n][3+RI], writes g[n][3+RI].
template <int RI, int RY, int RP, int RR, int NAP, int NS, class QA>
NLML_HD void tucker_gradient(const float (&p)[NS][3 + RI], const float* __restrict__ S, const QA& qa,
                             float* scr, int sstride, int ssample, const float* rows_y,
                             const float* rows_p, const float* rows_r, float (&g)[NS][3 + RI]) {
    constexpr int nA = tri(RI), nB = tri(RY), nC = tri(RP), nD = tri(RR);
    float cy[NS][RY], dcy[NS][RY], cp[NS][RP], dcp[NS][RP], cr[NS][RR], dcr[NS][RR], u[NS][RI];
    float UU[NS][nA], YY[NS][nB], PP[NS][nC], RRv[NS][nD];
    float GU[NS][nA], GY[NS][nB], GP[NS][nC], GR[NS][nD];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        cos_features<RY>(p[n][0], rows_y, cy[n], dcy[n]);
        cos_features<RP>(p[n][1], rows_p, cp[n], dcp[n]);
        cos_features<RR>(p[n][2], rows_r, cr[n], dcr[n]);
#pragma unroll
        for (int i = 0; i < RI; ++i) u[n][i] = p[n][3 + i];
        sym_products<RI>(u[n], UU[n]);
        sym_products<RY>(cy[n], YY[n]);
        sym_products<RP>(cp[n], PP[n]);
        sym_products<RR>(cr[n], RRv[n]);
#pragma unroll
        for (int a = 0; a < nA; ++a) GU[n][a] = 0.f;
#pragma unroll
        for (int b = 0; b < nB; ++b) GY[n][b] = 0.f;
#pragma unroll
        for (int c = 0; c < nC; ++c) GP[n][c] = 0.f;
#pragma unroll
        for (int d = 0; d < nD; ++d) GR[n][d] = 0.f;
        // per-sample values indexed by the (b,c) loop counter go through the scratch column
#pragma unroll
        for (int b = 0; b < nB; ++b)
#pragma unroll
            for (int c = 0; c < nC; ++c) scr[n * ssample + (b * nC + c) * sstride] = YY[n][b] * PP[n][c];
    }

    // quadratic term: one pass over S feeds both contractions of all NS samples.  The (b,c) loop is a REAL
    // loop (36 trips at ranks 3,3) so its body stays inside the instruction cache.  The T dot product is split
    // in three chains to shorten the dependent-FMA latency.
#pragma unroll 1
    for (int bc = 0; bc < nB * nC; ++bc) {
        const float* __restrict__ rows = S + bc * (nD * NAP);
        float yp[NS], tr[NS];
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            yp[n] = scr[n * ssample + bc * sstride];
            tr[n] = 0.f;  // sum_d T[b,c,d] * RR_d
        }
#pragma unroll
        for (int d = 0; d < nD; ++d) {
            const float* __restrict__ row = rows + d * NAP;
            float ypr[NS], t0[NS], t1[NS], t2[NS];
#pragma unroll
            for (int n = 0; n < NS; ++n) {
                ypr[n] = yp[n] * RRv[n][d];
                t0[n] = t1[n] = t2[n] = 0.f;
            }
#pragma unroll
            for (int a = 0; a < nA; ++a) {
                const float sv = row[a];
#pragma unroll
                for (int n = 0; n < NS; ++n) {
                    if (a % 3 == 0) t0[n] = fmaf(sv, UU[n][a], t0[n]);
                    else if (a % 3 == 1) t1[n] = fmaf(sv, UU[n][a], t1[n]);
                    else t2[n] = fmaf(sv, UU[n][a], t2[n]);
                    GU[n][a] = fmaf(sv, ypr[n], GU[n][a]);
                }
            }
#pragma unroll
            for (int n = 0; n < NS; ++n) {
                const float t = (t0[n] + t1[n]) + t2[n];
                GR[n][d] = fmaf(t, yp[n], GR[n][d]);
                tr[n] = fmaf(t, RRv[n][d], tr[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) scr[n * ssample + bc * sstride] = tr[n];
    }

#pragma unroll
    for (int n = 0; n < NS; ++n) {
#pragma unroll
        for (int b = 0; b < nB; ++b)
#pragma unroll
            for (int c = 0; c < nC; ++c) {
                const float tr = scr[n * ssample + (b * nC + c) * sstride];
                GY[n][b] = fmaf(tr, PP[n][c], GY[n][b]);
                GP[n][c] = fmaf(tr, YY[n][b], GP[n][c]);
            }
        float lin_u[RI], ey[RY], ep[RP], er[RR];
        linear_term<RI, RY, RP, RR>(qa, n, cy[n], cp[n], cr[n], u[n], lin_u, ey, ep, er);
        assemble_gradient<RI, RY, RP, RR>(GU[n], GY[n], GP[n], GR[n], u[n], cy[n], cp[n], cr[n], dcy[n], dcp[n], dcr[n],
                                          lin_u, ey, ep, er, g[n]);
    }
}

// ---------------------------------------------------------------------------------------------
// Converged fit: damped Newton / Levenberg-Marquardt on the same folded form.
// What the reference ships by default is a scipy Powell search over this objective from p = 0
// (TD_Tester.py:164, :191-194); the solver below reaches the local minimum of the same basin with the exact
// Hessian.  Because L - 0.5 x.x = F1 + F2 is a polynomial in (u, c_y, c_p, c_r), every derivative along an angle
// is the same contraction with the monomials replaced by their derivatives along that angle:
//   YY -> Y1 = d(YY)/dw_y, Y2 = d2(YY)/dw_y^2, ...   (sym_products_d)
// so one pass over S yields value, gradient and Hessian.
// ---------------------------------------------------------------------------------------------
template <int R>
NLML_HD void cos_features2(float w, const float* rows, float* c, float* dc, float* d2c) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const float a = rows[4 * j + 0], b = rows[4 * j + 1], ph = rows[4 * j + 2], d = rows[4 * j + 3];
        float s, co;
        sincos_small(b * w + ph, &s, &co);
        c[j] = a * co + d;
        dc[j] = -(a * b) * s;
        d2c[j] = -(a * b * b) * co;
    }
}

// monomials v_i v_j (i<=j) of a vector that depends on one scalar, with their first and second derivatives
template <int R>
NLML_HD void sym_products_d(const float* v, const float* dv, const float* d2v, float* vv, float* vv1, float* vv2) {
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = i; j < R; ++j) {
            const int k = pair_index(i, j, R);
            vv[k] = v[i] * v[j];
            vv1[k] = fmaf(dv[i], v[j], v[i] * dv[j]);
            vv2[k] = fmaf(d2v[i], v[j], fmaf(2.0f * dv[i], dv[j], v[i] * d2v[j]));
        }
}

// packed lower triangle of a symmetric NP x NP matrix
NLML_HD constexpr int tri_index(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

// Value (without the constant 0.5 x.x), gradient and Hessian of the objective at p = (w_y, w_p, w_r, u).
//   scr : 3*(nB+nC) floats of per-sample scratch (element k at scr[k*sstride]).
//   q   : q = W2 x of this sample, element r at q[r*qstride].
template <int RI, int RY, int RP, int RR, int NAP>
NLML_HD void tucker_newton_eval(const float (&p)[3 + RI], const float* __restrict__ S, const float* __restrict__ q,
                                int qstride, float* scr, int sstride, const float* rows_y, const float* rows_p,
                                const float* rows_r, float& Lval, float (&g)[3 + RI],
                                float (&H)[(3 + RI) * (4 + RI) / 2]) {
    constexpr int nA = tri(RI), nB = tri(RY), nC = tri(RP), nD = tri(RR);
    float u[RI], UU[nA];
    float cy[RY], dcy[RY], d2cy[RY], cp[RP], dcp[RP], d2cp[RP], cr[RR], dcr[RR], d2cr[RR];
    float R0[nD], R1[nD], R2[nD];
    cos_features2<RY>(p[0], rows_y, cy, dcy, d2cy);
    cos_features2<RP>(p[1], rows_p, cp, dcp, d2cp);
    cos_features2<RR>(p[2], rows_r, cr, dcr, d2cr);
#pragma unroll
    for (int i = 0; i < RI; ++i) u[i] = p[3 + i];
    sym_products<RI>(u, UU);
    sym_products_d<RR>(cr, dcr, d2cr, R0, R1, R2);
    {   // values indexed by the (b,c) loop counters go through the scratch column
        float Y0[nB], Y1[nB], Y2[nB], P0[nC], P1[nC], P2[nC];
        sym_products_d<RY>(cy, dcy, d2cy, Y0, Y1, Y2);
        sym_products_d<RP>(cp, dcp, d2cp, P0, P1, P2);
#pragma unroll
        for (int b = 0; b < nB; ++b) {
            scr[(3 * b + 0) * sstride] = Y0[b];
            scr[(3 * b + 1) * sstride] = Y1[b];
            scr[(3 * b + 2) * sstride] = Y2[b];
        }
#pragma unroll
        for (int c = 0; c < nC; ++c) {
            scr[(3 * nB + 3 * c + 0) * sstride] = P0[c];
            scr[(3 * nB + 3 * c + 1) * sstride] = P1[c];
            scr[(3 * nB + 3 * c + 2) * sstride] = P2[c];
        }
    }

    // ---- quadratic term F2 = sum S UU YY PP RR and its derivatives ----
    float GU[nA], HY[nA], HP[nA], HR[nA];
#pragma unroll
    for (int a = 0; a < nA; ++a) GU[a] = HY[a] = HP[a] = HR[a] = 0.f;
    float F2 = 0.f, gy = 0.f, gp = 0.f, gr = 0.f, hyy = 0.f, hpp = 0.f, hrr = 0.f, hyp = 0.f, hyr = 0.f, hpr = 0.f;
#pragma unroll 1
    for (int b = 0; b < nB; ++b) {
        const float y0 = scr[(3 * b + 0) * sstride], y1 = scr[(3 * b + 1) * sstride], y2 = scr[(3 * b + 2) * sstride];
#pragma unroll 1
        for (int c = 0; c < nC; ++c) {
            const float p0 = scr[(3 * nB + 3 * c + 0) * sstride], p1 = scr[(3 * nB + 3 * c + 1) * sstride],
                        p2 = scr[(3 * nB + 3 * c + 2) * sstride];
            const float m00 = y0 * p0, m10 = y1 * p0, m01 = y0 * p1, m20 = y2 * p0, m02 = y0 * p2, m11 = y1 * p1;
            const float* __restrict__ rows = S + (b * nC + c) * (nD * NAP);
#pragma unroll
            for (int d = 0; d < nD; ++d) {
                const float* __restrict__ row = rows + d * NAP;
                const float w00 = m00 * R0[d], wy = m10 * R0[d], wp = m01 * R0[d], wr = m00 * R1[d];
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int a = 0; a < nA; ++a) {
                    const float sv = row[a];
                    if (a % 3 == 0) t0 = fmaf(sv, UU[a], t0);
                    else if (a % 3 == 1) t1 = fmaf(sv, UU[a], t1);
                    else t2 = fmaf(sv, UU[a], t2);
                    GU[a] = fmaf(sv, w00, GU[a]);
                    HY[a] = fmaf(sv, wy, HY[a]);
                    HP[a] = fmaf(sv, wp, HP[a]);
                    HR[a] = fmaf(sv, wr, HR[a]);
                }
                const float t = (t0 + t1) + t2;
                F2 = fmaf(t, w00, F2);
                gy = fmaf(t, wy, gy);
                gp = fmaf(t, wp, gp);
                gr = fmaf(t, wr, gr);
                hyy = fmaf(t, m20 * R0[d], hyy);
                hpp = fmaf(t, m02 * R0[d], hpp);
                hrr = fmaf(t, m00 * R2[d], hrr);
                hyp = fmaf(t, m11 * R0[d], hyp);
                hyr = fmaf(t, m10 * R1[d], hyr);
                hpr = fmaf(t, m01 * R1[d], hpr);
            }
        }
    }

    // ---- linear term F1 = -sum q[ijkl] u_i cy_j cp_k cr_l and its derivatives ----
    float kl0[RP * RR], klp[RP * RR], klr[RP * RR], klpp[RP * RR], klrr[RP * RR], klpr[RP * RR];
#pragma unroll
    for (int k = 0; k < RP; ++k)
#pragma unroll
        for (int l = 0; l < RR; ++l) {
            kl0[k * RR + l] = cp[k] * cr[l];
            klp[k * RR + l] = dcp[k] * cr[l];
            klr[k * RR + l] = cp[k] * dcr[l];
            klpp[k * RR + l] = d2cp[k] * cr[l];
            klrr[k * RR + l] = cp[k] * d2cr[l];
            klpr[k * RR + l] = dcp[k] * dcr[l];
        }
    float F1 = 0.f, A0[RI], Ay[RI], Ap[RI], Ar[RI];
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        float a0 = 0.f, ay = 0.f, ap = 0.f, ar = 0.f, byy = 0.f, bpp = 0.f, brr = 0.f, byp = 0.f, byr = 0.f, bpr = 0.f;
#pragma unroll
        for (int j = 0; j < RY; ++j) {
            float s0 = 0.f, sp = 0.f, sr = 0.f, spp = 0.f, srr = 0.f, spr = 0.f;
#pragma unroll
            for (int kl = 0; kl < RP * RR; ++kl) {
                const float qv = q[((i * RY + j) * RP * RR + kl) * qstride];
                s0 = fmaf(qv, kl0[kl], s0);
                sp = fmaf(qv, klp[kl], sp);
                sr = fmaf(qv, klr[kl], sr);
                spp = fmaf(qv, klpp[kl], spp);
                srr = fmaf(qv, klrr[kl], srr);
                spr = fmaf(qv, klpr[kl], spr);
            }
            a0 = fmaf(cy[j], s0, a0);
            ay = fmaf(dcy[j], s0, ay);
            ap = fmaf(cy[j], sp, ap);
            ar = fmaf(cy[j], sr, ar);
            byy = fmaf(d2cy[j], s0, byy);
            bpp = fmaf(cy[j], spp, bpp);
            brr = fmaf(cy[j], srr, brr);
            byp = fmaf(dcy[j], sp, byp);
            byr = fmaf(dcy[j], sr, byr);
            bpr = fmaf(cy[j], spr, bpr);
        }
        A0[i] = a0; Ay[i] = ay; Ap[i] = ap; Ar[i] = ar;
        F1 = fmaf(-u[i], a0, F1);
        gy = fmaf(-u[i], ay, gy);
        gp = fmaf(-u[i], ap, gp);
        gr = fmaf(-u[i], ar, gr);
        hyy = fmaf(-u[i], byy, hyy);
        hpp = fmaf(-u[i], bpp, hpp);
        hrr = fmaf(-u[i], brr, hrr);
        hyp = fmaf(-u[i], byp, hyp);
        hyr = fmaf(-u[i], byr, hyr);
        hpr = fmaf(-u[i], bpr, hpr);
    }

    // ---- assemble ----
    Lval = F1 + F2;
    g[0] = gy; g[1] = gp; g[2] = gr;
    H[tri_index(0, 0)] = hyy; H[tri_index(1, 1)] = hpp; H[tri_index(2, 2)] = hrr;
    H[tri_index(1, 0)] = hyp; H[tri_index(2, 0)] = hyr; H[tri_index(2, 1)] = hpr;
    float du[RI], duy[RI], dup[RI], dur[RI];
    sym_backprop<RI>(GU, u, du);
    sym_backprop<RI>(HY, u, duy);
    sym_backprop<RI>(HP, u, dup);
    sym_backprop<RI>(HR, u, dur);
#pragma unroll
    for (int m = 0; m < RI; ++m) {
        g[3 + m] = du[m] - A0[m];
        H[tri_index(3 + m, 0)] = duy[m] - Ay[m];
        H[tri_index(3 + m, 1)] = dup[m] - Ap[m];
        H[tri_index(3 + m, 2)] = dur[m] - Ar[m];
#pragma unroll
        for (int n = 0; n <= m; ++n)
            H[tri_index(3 + m, 3 + n)] = (m == n) ? 2.0f * GU[pair_index(m, m, RI)] : GU[pair_index(n, m, RI)];
    }
}

// Solve A d = g for a packed symmetric positive-definite A by Cholesky; false when a pivot is not positive.
template <int NP>
NLML_HD bool chol_solve(const float (&A)[NP * (NP + 1) / 2], const float (&g)[NP], float (&d)[NP]) {
    float Lm[NP * (NP + 1) / 2];
    bool ok = true;
#pragma unroll
    for (int r = 0; r < NP; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            float acc = A[tri_index(r, c)];
#pragma unroll
            for (int k = 0; k < c; ++k) acc = fmaf(-Lm[tri_index(r, k)], Lm[tri_index(c, k)], acc);
            if (c == r) {
                ok = ok && (acc > 0.f);
                Lm[tri_index(r, r)] = sqrtf(acc > 0.f ? acc : 1.0f);
            } else {
                Lm[tri_index(r, c)] = acc / Lm[tri_index(c, c)];
            }
        }
    }
    float y[NP];
#pragma unroll
    for (int r = 0; r < NP; ++r) {
        float acc = g[r];
#pragma unroll
        for (int k = 0; k < r; ++k) acc = fmaf(-Lm[tri_index(r, k)], y[k], acc);
        y[r] = acc / Lm[tri_index(r, r)];
    }
#pragma unroll
    for (int r = NP - 1; r >= 0; --r) {
        float acc = y[r];
#pragma unroll
        for (int k = r + 1; k < NP; ++k) acc = fmaf(-Lm[tri_index(k, r)], d[k], acc);
        d[r] = acc / Lm[tri_index(r, r)];
    }
    return ok;
}

struct LmOptions {
    int max_evals;      // objective/gradient/Hessian evaluations per sample (the unit of work)
    float lambda0;      // initial damping (Marquardt scaling: A = H + lambda * diag(|H_ii|))
    float lambda_down;  // damping divisor after an accepted, uncapped step
    float lambda_up;    // damping multiplier after a rejected step or a failed factorisation
    float angle_cap;    // largest angle change per step (rad): keeps the search in the basin reached from p = 0
    float step_tol;     // stop when max |delta p| falls below this
    float diag_floor;   // added to |H_ii| in the damping term (at p = 0 the angle block of H is exactly zero)
    float lambda_min;   // damping floor: a smaller floor buys nothing (the step is already Newton's) and costs one
                        // rejected evaluation per factor lambda_up when a step overshoots in the curved valley
    float noise_step;   // FP32 floor: on ill-conditioned samples the Newton step itself becomes rounding noise
                        // (up to ~1e-4 rad); stop after two accepted steps below this that no longer contract
};
NLML_HD LmOptions lm_default_options() { return LmOptions{64, 1e-3f, 5.f, 4.f, 0.15f, 2e-6f, 1.0f, 1e-5f, 3e-4f}; }

// From p = 0 (TD_Tester.py:164) to the local minimum.  Returns the number of evaluations used.
template <int RI, int RY, int RP, int RR, int NAP>
NLML_HD int tucker_lm_solve(const float* __restrict__ S, const float* __restrict__ q, int qstride, float* scr, int sstride,
                            const float* rows_y, const float* rows_p, const float* rows_r, const LmOptions& o,
                            float (&p)[3 + RI], float& Lfinal) {
    constexpr int NP = 3 + RI, NH = NP * (NP + 1) / 2;
    float L, g[NP], H[NH];
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = 0.f;
    tucker_newton_eval<RI, RY, RP, RR, NAP>(p, S, q, qstride, scr, sstride, rows_y, rows_p, rows_r, L, g, H);
    int evals = 1;
    float lam = o.lambda0;
    int guard = 4 * o.max_evals;   // failed factorisations do not consume evaluations
    int flat = 0;                  // consecutive small accepted steps that did not contract
    float prev_step = 1e30f;
    while (evals < o.max_evals && guard-- > 0) {
        float A[NH], d[NP];
#pragma unroll
        for (int i = 0; i < NH; ++i) A[i] = H[i];
#pragma unroll
        for (int i = 0; i < NP; ++i) A[tri_index(i, i)] += lam * (fabsf(H[tri_index(i, i)]) + o.diag_floor);
        if (!chol_solve<NP>(A, g, d)) {
            lam *= o.lambda_up;
            if (lam > 1e12f) break;
            continue;
        }
        float ma = fmaxf(fabsf(d[0]), fmaxf(fabsf(d[1]), fabsf(d[2])));
        const bool capped = ma > o.angle_cap;
        const float scale = capped ? o.angle_cap / ma : 1.0f;
        float step = 0.f, pn[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const float di = scale * d[i];
            step = fmaxf(step, fabsf(di));
            pn[i] = p[i] - di;
        }
        if (step < o.step_tol && lam < 1.0f) {   // converged: a (nearly) undamped correction below the tolerance; take it unevaluated
#pragma unroll
            for (int i = 0; i < NP; ++i) p[i] = pn[i];
            break;
        }
        float Ln, gn[NP], Hn[NH];
        tucker_newton_eval<RI, RY, RP, RR, NAP>(pn, S, q, qstride, scr, sstride, rows_y, rows_p, rows_r, Ln, gn, Hn);
        ++evals;
        // FP32 resolution of L (|F1 + F2| ~ 0.5 |x_hat|^2): differences below it carry no information
        const float resolution = 2e-6f * (1.0f + fabsf(L));
        if (Ln <= L + resolution) {
            flat = (step < o.noise_step && step > 0.5f * prev_step && lam < 1.0f) ? flat + 1 : 0;
            prev_step = step;
            const bool stalled = flat >= 2;
#pragma unroll
            for (int i = 0; i < NP; ++i) { p[i] = pn[i]; g[i] = gn[i]; }
#pragma unroll
            for (int i = 0; i < NH; ++i) H[i] = Hn[i];
            L = Ln;
            if (stalled) break;
            if (!capped) lam = fmaxf(lam / o.lambda_down, o.lambda_min);
        } else {
            lam *= o.lambda_up;
            if (lam > 1e12f) break;
        }
    }
    Lfinal = L;
    return evals;
}

// ---- one-time constant preparation (per Tucker core), identical on host and device ----

// One entry of M = W2 W2^T in double.
NLML_HD double gram_entry(const float* W2, int F, int r, int c) {
    double acc = 0.0;
    const float* a = W2 + (long long)r * F;
    const float* b = W2 + (long long)c * F;
    for (int f = 0; f < F; ++f) acc += (double)a[f] * (double)b[f];
    return acc;
}

// unordered pair index -> (i,j), i<=j
NLML_HD void unpair(int idx, int r, int* i, int* j) {
    int ii = 0;
    while (idx >= r - ii) {
        idx -= r - ii;
        ++ii;
    }
    *i = ii;
    *j = ii + idx;
}

// One entry of the folded Gram tensor:  S[A,B,C,D] = 0.5 * sum over the (up to 16) ordered index
// tuples in the class of M[(i,j,k,l),(i',j',k',l')].
NLML_HD float fold_entry(const double* M, int ri, int ry, int rp, int rr, int A, int B, int C, int D) {
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
    return (float)(0.5 * acc);
}

}  // namespace nlml
