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
    const float norm = sqrtf(ss);
    float coef = clip / (norm + 1e-6f);
    coef = coef < 1.0f ? coef : 1.0f;
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = sub_rn(p[i], mul_rn(lr, mul_rn(g[i], coef)));
}

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
