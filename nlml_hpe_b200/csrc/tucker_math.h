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

// One full gradient evaluation at p for fixed small ranks, thread-per-sample form.
//   S : folded Gram tensor laid out [nB*nC*nD][NAP] (NAP = nA padded to a multiple of 4), read-only,
//       identical for all samples (shared-memory broadcast on the GPU).
//   q : this sample's q = W2 x, element r at q[r*qstride].
//   scr : nB*nC floats of per-sample scratch, element k at scr[k*sstride] (a shared-memory column on the GPU).
//   rows_* : cosine rows of the three angle factors.
// Writes g[3+RI].
template <int RI, int RY, int RP, int RR, int NAP>
NLML_HD void tucker_gradient(const float* p, const float* __restrict__ S, const float* q, int qstride,
                             float* scr, int sstride,
                             const float* rows_y, const float* rows_p, const float* rows_r, float* g) {
    constexpr int nA = tri(RI), nB = tri(RY), nC = tri(RP), nD = tri(RR);
    float cy[RY], dcy[RY], cp[RP], dcp[RP], cr[RR], dcr[RR], u[RI];
    cos_features<RY>(p[0], rows_y, cy, dcy);
    cos_features<RP>(p[1], rows_p, cp, dcp);
    cos_features<RR>(p[2], rows_r, cr, dcr);
#pragma unroll
    for (int i = 0; i < RI; ++i) u[i] = p[3 + i];

    float UU[nA], YY[nB], PP[nC], RRv[nD];
    sym_products<RI>(u, UU);
    sym_products<RY>(cy, YY);
    sym_products<RP>(cp, PP);
    sym_products<RR>(cr, RRv);

    float GU[nA], GY[nB], GP[nC], GR[nD];
#pragma unroll
    for (int a = 0; a < nA; ++a) GU[a] = 0.f;
#pragma unroll
    for (int b = 0; b < nB; ++b) GY[b] = 0.f;
#pragma unroll
    for (int c = 0; c < nC; ++c) GP[c] = 0.f;
#pragma unroll
    for (int d = 0; d < nD; ++d) GR[d] = 0.f;

    // quadratic term: one pass over S feeds both contractions.  The (b,c) loop is a REAL loop (36 trips at
    // ranks 3,3) so its body stays inside the instruction cache; the per-thread values indexed by the loop
    // counter (YY_b*PP_c in, sum_d T[b,c,d]*RR_d out) go through the caller's scratch column instead of
    // registers.  The T dot product is split in three chains to shorten the dependent-FMA latency.
#pragma unroll
    for (int b = 0; b < nB; ++b)
#pragma unroll
        for (int c = 0; c < nC; ++c) scr[(b * nC + c) * sstride] = YY[b] * PP[c];
#pragma unroll 1
    for (int bc = 0; bc < nB * nC; ++bc) {
        const float yp = scr[bc * sstride];
        const float* __restrict__ rows = S + bc * (nD * NAP);
        float tr = 0.f;  // sum_d T[b,c,d] * RR_d
#pragma unroll
        for (int d = 0; d < nD; ++d) {
            const float* __restrict__ row = rows + d * NAP;
            const float ypr = yp * RRv[d];
            float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int a = 0; a < nA; ++a) {
                const float s = row[a];
                if (a % 3 == 0) t0 = fmaf(s, UU[a], t0);
                else if (a % 3 == 1) t1 = fmaf(s, UU[a], t1);
                else t2 = fmaf(s, UU[a], t2);
                GU[a] = fmaf(s, ypr, GU[a]);
            }
            const float t = (t0 + t1) + t2;
            GR[d] = fmaf(t, yp, GR[d]);
            tr = fmaf(t, RRv[d], tr);
        }
        scr[bc * sstride] = tr;
    }
#pragma unroll
    for (int b = 0; b < nB; ++b)
#pragma unroll
        for (int c = 0; c < nC; ++c) {
            const float tr = scr[(b * nC + c) * sstride];
            GY[b] = fmaf(tr, PP[c], GY[b]);
            GP[c] = fmaf(tr, YY[b], GP[c]);
        }
    float du[RI], dy[RY], dp[RP], dr[RR];
    sym_backprop<RI>(GU, u, du);
    sym_backprop<RY>(GY, cy, dy);
    sym_backprop<RP>(GP, cp, dp);
    sym_backprop<RR>(GR, cr, dr);

    // linear term -q.z : d/du_i = -sum_jkl q[ijkl] cy_j cp_k cr_l ; e[jkl] = sum_i u_i q[ijkl]
    float ey[RY], ep[RP], er[RR];
#pragma unroll
    for (int j = 0; j < RY; ++j) ey[j] = 0.f;
#pragma unroll
    for (int k = 0; k < RP; ++k) ep[k] = 0.f;
#pragma unroll
    for (int l = 0; l < RR; ++l) er[l] = 0.f;
    float lin_u[RI];
#pragma unroll
    for (int i = 0; i < RI; ++i) lin_u[i] = 0.f;
#pragma unroll
    for (int j = 0; j < RY; ++j) {
#pragma unroll
        for (int k = 0; k < RP; ++k) {
            const float yk = cy[j] * cp[k];
            float e_jk = 0.f;  // sum_l e[jkl] cr_l
#pragma unroll
            for (int l = 0; l < RR; ++l) {
                const float t_jkl = yk * cr[l];
                float e = 0.f;  // sum_i u_i q[ijkl]
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    const float qv = q[(((i * RY + j) * RP + k) * RR + l) * qstride];
                    lin_u[i] = fmaf(qv, t_jkl, lin_u[i]);
                    e = fmaf(qv, u[i], e);
                }
                er[l] = fmaf(e, yk, er[l]);
                e_jk = fmaf(e, cr[l], e_jk);
            }
            ey[j] = fmaf(e_jk, cp[k], ey[j]);
            ep[k] = fmaf(e_jk, cy[j], ep[k]);
        }
    }
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
