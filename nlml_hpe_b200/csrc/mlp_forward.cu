// Encoder + yaw/pitch/roll heads forward for sm_100a.
//
// Replaces CombinedAnglePredictionModel.forward (/root/reference/NLML_HPE_Model_Builder.py:115-126):
//   encoder  1404-1024-512-256-128-64-9   ReLU x4, Tanh, none          (:33-53)
//   split    latent[:,0:3] | [:,3:6] | [:,6:9]                           (:55-68, :118-120)
//   heads    3 x (3-128-256-128-64-1)      ReLU x4, none                 (:76-92)
// Every layer is y = act(x W^T + b) with nn.Linear's weight[out][in] layout.
//
// This file holds the FP32 CUDA-core chain (exact-precision path, also the on-device reference the
// tensor-core chain in mlp_tc.cu is validated against): one tiled GEMM kernel with bias + activation
// fused in the epilogue, launched once per layer; the three heads ride blockIdx.z of one launch per
// head layer.  The batch is processed in chunks whose activations stay L2-resident.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "mlp_tc.cuh"

namespace nlml {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct LinearArgs {
    // up to 3 independent problems of identical shape (grid.z): the three heads
    const float* X[3];
    const float* W[3];
    const float* B[3];
    float* Y[3];        // FP32 output, or null when only the split planes are wanted
    __half* Yhi[3];     // optional FP16 hi/lo planes of the output (input format of a tensor-core layer)
    __half* Ylo[3];
    long long ldx, ldy;
    long long N;  // rows (samples)
    int in, out;
    int act;
    int vec_x;  // X rows 16B aligned (ldx % 4 == 0, base aligned) and in % 4 == 0
    int vec_w;  // W rows 16B aligned (in % 4 == 0)
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) {   // NaN-propagating, as torch.relu (see relu_nan in mlp_tc.cuh)
        float r;
        asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
        return r;
    }
    if (act == ACT_TANH) return tanhf(v);  // accurate tanhf, as torch.tanh on f32 (Model_Builder.py:48)
    return v;
}

__device__ __forceinline__ float4 ld4_guard(const float* __restrict__ row, int k, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && k + 3 < K) {
        v = __ldg(reinterpret_cast<const float4*>(row + k));
    } else {
        if (k + 0 < K) v.x = __ldg(row + k + 0);
        if (k + 1 < K) v.y = __ldg(row + k + 1);
        if (k + 2 < K) v.z = __ldg(row + k + 2);
        if (k + 3 < K) v.w = __ldg(row + k + 3);
    }
    return v;
}

// 128 x BN output tile per CTA, 256 threads, each thread an 8 x (BN/16) register tile split in
// 4-wide halves so every shared-memory read is a conflict-free 128-bit load.
template <int BN>
__global__ void __launch_bounds__(256) linear_simt_kernel(const __grid_constant__ LinearArgs a) {
    constexpr int BM = 128, BK = 16, TN = BN / 16;  // TN in {4, 8}
    constexpr int ASTR = BM + 4, BSTR = BN + 4;
    __shared__ __align__(16) float As[2][BK][ASTR];
    __shared__ __align__(16) float Bs[2][BK][BSTR];

    const int z = blockIdx.z;
    const float* __restrict__ X = a.X[z];
    const float* __restrict__ W = a.W[z];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = a.in;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 xa[2], wb[BN / 64];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256, r = idx >> 2, c4 = idx & 3;
            const long long row = m0 + r;
            xa[i] = row < a.N ? ld4_guard(X + row * a.ldx, k0 + 4 * c4, K, a.vec_x != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
            const int idx = tid + i * 256, r = idx >> 2, c4 = idx & 3;
            const int col = n0 + r;
            wb[i] = col < a.out ? ld4_guard(W + (long long)col * K, k0 + 4 * c4, K, a.vec_w != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256, r = idx >> 2, c4 = idx & 3;
            As[buf][4 * c4 + 0][r] = xa[i].x;
            As[buf][4 * c4 + 1][r] = xa[i].y;
            As[buf][4 * c4 + 2][r] = xa[i].z;
            As[buf][4 * c4 + 3][r] = xa[i].w;
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
            const int idx = tid + i * 256, r = idx >> 2, c4 = idx & 3;
            Bs[buf][4 * c4 + 0][r] = wb[i].x;
            Bs[buf][4 * c4 + 1][r] = wb[i].y;
            Bs[buf][4 * c4 + 2][r] = wb[i].z;
            Bs[buf][4 * c4 + 3][r] = wb[i].w;
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        const bool more = k0 + BK < K;
        if (more) gload(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + h * 64]);
                bv[4 * h + 0] = b.x; bv[4 * h + 1] = b.y; bv[4 * h + 2] = b.z; bv[4 * h + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    const float* __restrict__ Bv = a.B[z];
    float* __restrict__ Y = a.Y[z];
    __half* __restrict__ Yhi = a.Yhi[z];
    __half* __restrict__ Ylo = a.Ylo[z];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long row = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (row >= a.N) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
            if (col < a.out) {
                const float v = apply_act(acc[i][j] + __ldg(Bv + col), a.act);
                if (Y) Y[row * a.ldy + col] = v;
                if (Yhi) {
                    const __half h = __float2half_rn(v);
                    Yhi[row * a.ldy + col] = h;
                    Ylo[row * a.ldy + col] = __float2half_rn(v - __half2float(h));
                }
            }
        }
    }
}

}  // namespace nlml

// =============================================================================================
// C ABI
// =============================================================================================
using namespace nlml;

namespace {
constexpr int kEnc = NLML_MLP_ENCODER_LAYERS, kHead = NLML_MLP_HEAD_LAYERS, kNumT = NLML_MLP_NUM_TENSORS;

struct Workspace {
    float* f32[2] = {nullptr, nullptr};   // FP32 activations, ping-pong by layer parity
    __half* hi[2] = {nullptr, nullptr};   // FP16 hi/lo planes, ping-pong by layer parity
    __half* lo[2] = {nullptr, nullptr};
    float* lat = nullptr;                 // [rows][latent]
    int64_t rows = 0;                     // capacity in samples (<= plan chunk); grown on demand
};
}  // namespace

// Accuracy / kernel-selection constants.  These are COMPILE-TIME: the product library reads no environment variable
// (a stray variable must not be able to change the angles).  Development builds override them with -D.
#ifndef NLML_MLP_TC_GROUP
#define NLML_MLP_TC_GROUP 1          // k-blocks accumulated in TMEM per promotion.  Max error vs the reference over 32768
#endif                               // samples: 1 -> 5.3e-4 deg, 2 -> 8.3e-4 deg (+3 % speed), never promote -> 3.7e-3 deg
#ifndef NLML_MLP_SHORTK_GROUP2
#define NLML_MLP_SHORTK_GROUP2 256   // layers with K <= this sum two k-blocks per TMEM accumulator (16.5 -> 16.2 ms per 1M
#endif                               // samples; max error over 32 k samples 5.8e-4 -> 6.0e-4 degrees); 0 turns it off
#ifndef NLML_MLP_TC_NECK
#define NLML_MLP_TC_NECK 1           // encoder.8 on the tensor cores with the narrow layers in its epilogue (0: neck_kernel)
#endif
#ifndef NLML_MLP_TC_TAIL
#define NLML_MLP_TC_TAIL 1           // heads' last two layers on the tensor cores (0: CUDA-core head_tail_kernel)
#endif
#ifndef NLML_MLP_TWO_CTA
#define NLML_MLP_TWO_CTA 1           // 256-wide layers: cta_group::2 kernel (0: 1-CTA MMAs with W multicast)
#endif
#ifndef NLML_MLP_CHUNK_WAVES
#define NLML_MLP_CHUNK_WAVES 8       // samples per pass = waves x num_sms x 128 (4 waves 16.9 ms / 1M, 8 waves 16.5, 16 waves 16.5)
#endif
#ifndef NLML_MLP_LEAD_KB
#define NLML_MLP_LEAD_KB 0           // long reductions (>= 8 k-blocks): the first this-many k-blocks of a tile summed in pairs, so
#endif                               // the MMAs run further ahead of the epilogue warps' store phase.  Measured with 4: encoder.0
                                     // 39.2 k -> 37.7 k cycles per tile (the layer is shared-memory-bandwidth bound either way),
                                     // whole chain +1 %, error 6.0e-4 -> 6.3e-4 deg on synthetic features and 1.02e-3 deg on the
                                     // landmark fixtures: over the 1e-3 budget, so it stays off
#ifndef NLML_MLP_DEV_LANES
#define NLML_MLP_DEV_LANES 2         // device-buffer batches above one chunk: chunks alternate between this many internal
#endif                               // streams (forked from / joined to the caller's), so one chunk's HBM-bound operand split
                                     // and epilogue-bound head layers overlap the other's tensor-bound encoder layers
namespace {
constexpr int kDevLanes = NLML_MLP_DEV_LANES, kLeadKb = NLML_MLP_LEAD_KB;
constexpr int kTcGroup = NLML_MLP_TC_GROUP, kShortKGroup2 = NLML_MLP_SHORTK_GROUP2;
constexpr bool kTcNeck = NLML_MLP_TC_NECK != 0, kTcTail = NLML_MLP_TC_TAIL != 0, kTwoCta = NLML_MLP_TWO_CTA != 0;
}  // namespace

struct nlml_mlp_plan {
    int device = 0;
    int out_dims[kNumT];
    int in_dims[kNumT];
    float* W[kNumT] = {};
    float* B[kNumT] = {};
    // tensor-core form of the eligible layers
    bool tc[kNumT] = {};
    __half* Whi[kNumT] = {};
    __half* Wlo[kNumT] = {};
    float inv_scale[kNumT] = {};
    int Kp[kNumT] = {};
    CUtensorMap wmap_hi[kNumT], wmap_lo[kNumT];
    float* Wt[kNumT] = {};   // transposed FP32 copies [in][out] for the fused narrow-layer kernels
    int path = 0;  // 0 = tensor-core chain where eligible, 1 = FP32 CUDA-core chain everywhere
    int input_size = 0, latent = 0, head_in = 0;
    int64_t chunk = 8 * 148 * 128;  // samples per pass: 8 x 148 M-tiles = whole waves of the persistent GEMMs; the per-launch
                                    // fixed costs of the 9 kernels amortise over it (final kernels, per 1M samples:
                                    // 4 waves 16.9 ms, 8 waves 16.5 ms, 16 waves 16.5 ms; workspaces 2.6 GB at 8)
    size_t f32_width[2] = {0, 0}, plane_width[2] = {0, 0};
    Workspace ws_dev;        // device-buffer entry points (ordered across caller streams by ws_dev_done)
    Workspace ws_lane;       // second lane of the device path (batches above one chunk)
    cudaStream_t lane[2] = {nullptr, nullptr};
    cudaEvent_t lane_fork = nullptr, lane_join[2] = {nullptr, nullptr};
    Workspace ws_host[2];    // the host pipeline's two slots (its own internal streams); never touched by the device path
    cudaEvent_t ws_dev_done = nullptr;   // recorded after the last device-path call's kernels: a later call on ANOTHER stream
                                         // waits for it before reusing ws_dev (two unordered streams must not share the buffers)
    int num_sms = 148;
    int64_t launches = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    float* x_dev[2] = {nullptr, nullptr};
    float* y_dev[2] = {nullptr, nullptr};
    float* y_stage[2] = {nullptr, nullptr};   // pinned result staging (see nlml_mlp_forward_host_f32)
    int64_t host_rows = 0;   // capacity of x_dev / y_dev
};

namespace {

inline int enc_t(int li) { return li; }
inline int head_t(int h, int li) { return kEnc + h * kHead + li; }
inline int act_of(int t) {
    if (t < kEnc) return t < 4 ? ACT_RELU : (t == 4 ? ACT_TANH : ACT_NONE);
    return ((t - kEnc) % kHead) < kHead - 1 ? ACT_RELU : ACT_NONE;
}
inline bool use_tc(const nlml_mlp_plan* pl, int t) { return pl->path == 0 && pl->tc[t]; }

void free_workspace(Workspace& w) {
    for (int i = 0; i < 2; ++i) { cudaFree(w.f32[i]); cudaFree(w.hi[i]); cudaFree(w.lo[i]); }
    cudaFree(w.lat);
    w = Workspace();
}
// Activation buffers for up to `rows` samples per pass (whole 128-row tiles).  Small batches (the reference's
// batch-1 calls) get a small workspace; it grows, never shrinks.  Growing synchronises the device first because
// earlier launches on any stream may still use the old buffers.
int ensure_workspace(nlml_mlp_plan* pl, Workspace& w, int64_t rows) {
    rows = std::min<int64_t>(pl->chunk, ceil_div(rows, 128) * 128);
    if (w.rows >= rows) return 0;
    if (w.rows) NLML_CUDA(cudaDeviceSynchronize());
    free_workspace(w);
    for (int i = 0; i < 2; ++i) {
        NLML_CUDA(cudaMalloc(&w.f32[i], sizeof(float) * rows * pl->f32_width[i]));
        if (pl->plane_width[i]) {
            NLML_CUDA(cudaMalloc(&w.hi[i], sizeof(__half) * rows * pl->plane_width[i]));
            NLML_CUDA(cudaMalloc(&w.lo[i], sizeof(__half) * rows * pl->plane_width[i]));
            // pad columns of the input planes are written by split_planes_kernel; everything else is fully overwritten
        }
    }
    NLML_CUDA(cudaMalloc(&w.lat, sizeof(float) * rows * pl->latent));
    w.rows = rows;
    return 0;
}

int launch_simt(nlml_mlp_plan* pl, LinearArgs& a, int nz, cudaStream_t st) {
    auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    a.vec_w = (a.in % 4 == 0);
    a.vec_x = (a.in % 4 == 0) && (a.ldx % 4 == 0);
    for (int z = 0; z < nz; ++z) a.vec_x = a.vec_x && aligned(a.X[z]);
    if (a.out > 64) {
        dim3 grid((unsigned)ceil_div(a.N, 128), (unsigned)ceil_div(a.out, 128), (unsigned)nz);
        linear_simt_kernel<128><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid((unsigned)ceil_div(a.N, 128), (unsigned)ceil_div(a.out, 64), (unsigned)nz);
        linear_simt_kernel<64><<<grid, 256, 0, st>>>(a);
    }
    NLML_CUDA(cudaGetLastError());
    pl->launches += 1;
    return 0;
}

// k-blocks summed in one TMEM accumulator before promotion.  The long reductions promote after every k-block (the
// truncating accumulate of the tensor core biases long chains); layers with K <= short_k_group2 (256: the heads and
// encoder layers 3-4, two or four k-blocks in all) sum two: the chain stays 24 MMAs long.
inline int group_for(const nlml_mlp_plan* pl, int Kp) { return Kp <= kShortKGroup2 ? std::max(kTcGroup, 2) : kTcGroup; }

#ifdef NLML_MLP_TIMING
// development build only: every linear_tc2_kernel launch writes its per-warp phase cycles into the next slice of this buffer
static float* g_mlp_timing_buf = nullptr;
static int g_mlp_timing_slices = 0, g_mlp_timing_next = 0;
#endif

// one tensor-core layer for `nz` problems of identical shape (1 = encoder layer, 3 = the heads):
// A planes [n][Kp] -> planes and/or FP32
// dot_t: when non-null, tensor ids of the FOLLOWING single-output layers, fused into the epilogue as a dot product
// (the heads' last two layers: 128 -> 64 relu on the tensor cores, 64 -> 1 in the epilogue); results go to Ydot[row*3+z]
int launch_tc(nlml_mlp_plan* pl, const int* t, int nz, const __half* const* Ahi, const __half* const* Alo, int64_t n,
              __half* const* Yhi, __half* const* Ylo, float* const* Yf32, cudaStream_t st, const int* dot_t = nullptr,
              float* Ydot = nullptr) {
    const int out = pl->out_dims[t[0]], Kp = pl->Kp[t[0]];
    tc::TcMaps maps;
    tc::LinearTcArgs a{};
    a.N = n; a.out = out; a.Kp = Kp; a.act = act_of(t[0]); a.problems = nz; a.ldy = out; a.group = group_for(pl, Kp);
    a.lead_kb = Kp / tc::BK >= 8 ? kLeadKb : 0;
    a.Ydot = Ydot; a.ldd = 3;
    for (int z = 0; z < nz && dot_t; ++z) { a.dot_w[z] = pl->W[dot_t[z]]; a.dot_b[z] = pl->B[dot_t[z]]; }
    for (int z = 0; z < nz; ++z) {
        if (int rc = tc::make_plane_map(&maps.a_hi[z], Ahi[z], n, Kp, Kp, tc::BM)) return rc;
        if (int rc = tc::make_plane_map(&maps.a_lo[z], Alo[z], n, Kp, Kp, tc::BM)) return rc;
        maps.w_hi[z] = pl->wmap_hi[t[z]];
        maps.w_lo[z] = pl->wmap_lo[t[z]];
        a.inv_scale[z] = pl->inv_scale[t[z]]; a.bias[z] = pl->B[t[z]];
        a.Yhi[z] = Yhi[z]; a.Ylo[z] = Ylo[z]; a.Yf32[z] = Yf32[z];
    }
    // plane-only outputs of the 128- and 256-wide tiles leave through TMA (32-row x 64-column boxes per epilogue warp)
    a.y_tma = out % 128 == 0;
    for (int z = 0; z < nz; ++z) a.y_tma = a.y_tma && Yhi[z] && Ylo[z] && !Yf32[z];
    for (int z = 0; z < tc::kMaxProblems; ++z) {
        const int zz = z < nz ? z : 0;
        if (a.y_tma) {
            if (int rc = tc::make_plane_map(&maps.y_hi[z], Yhi[zz], n, out, out, 32)) return rc;
            if (int rc = tc::make_plane_map(&maps.y_lo[z], Ylo[zz], n, out, out, 32)) return rc;
        } else {
            maps.y_hi[z] = maps.a_hi[0]; maps.y_lo[z] = maps.a_lo[0];   // never dereferenced
        }
    }
    for (int z = nz; z < tc::kMaxProblems; ++z) {
        maps.a_hi[z] = maps.a_hi[0]; maps.a_lo[z] = maps.a_lo[0]; maps.w_hi[z] = maps.w_hi[0]; maps.w_lo[z] = maps.w_lo[0];
    }
    const int64_t tiles_m = ceil_div(n, tc::BM);
    if (out % 256 == 0) {
        // clusters of 2 CTAs share each W tile by TMA multicast (W maps of 256-wide layers have 128-row boxes)
        const int64_t work = ceil_div(tiles_m, 2) * (out / 256) * nz;
        const unsigned grid = 2 * (unsigned)std::min<int64_t>(work, pl->num_sms / 2);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::kThreads);
        cfg.dynamicSmemBytes = kTwoCta ? tc::Cfg2::SMEM_BYTES : tc::Cfg<256>::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
#ifdef NLML_MLP_TIMING
        a.timing = (g_mlp_timing_buf && g_mlp_timing_next < g_mlp_timing_slices)
                       ? g_mlp_timing_buf + (size_t)(g_mlp_timing_next++) * pl->num_sms * tc::kEpilogueWarps * 4 : nullptr;
#endif
        if (kTwoCta) NLML_CUDA(cudaLaunchKernelEx(&cfg, tc::linear_tc2_kernel, maps, a));
        else NLML_CUDA(cudaLaunchKernelEx(&cfg, tc::linear_tc_kernel<256, 2>, maps, a));
    } else if (out == 64) {
        const unsigned grid = (unsigned)std::min<int64_t>(tiles_m * nz, pl->num_sms);
        tc::linear_tc_kernel<64, 1><<<grid, tc::kThreads, tc::Cfg<64>::SMEM_BYTES, st>>>(maps, a);
    } else {
        const unsigned grid = (unsigned)std::min<int64_t>(tiles_m * (out / 128) * nz, pl->num_sms);
        tc::linear_tc_kernel<128, 1><<<grid, tc::kThreads, tc::Cfg<128>::SMEM_BYTES, st>>>(maps, a);
    }
    NLML_CUDA(cudaGetLastError());
    pl->launches += 1;
    return 0;
}

// the reference architecture's narrow layers have dedicated fused kernels
constexpr int kNeckIn = 128, kNeckMid = 64, kNeckLat = 9, kHeadIn = 3, kHeadW = 128, kTailMid = 64;
inline bool neck_fusable(const nlml_mlp_plan* pl) {
    return pl->in_dims[4] == kNeckIn && pl->out_dims[4] == kNeckMid && pl->out_dims[5] == kNeckLat &&
           pl->head_in == kHeadIn && pl->out_dims[head_t(0, 0)] == kHeadW;
}
inline bool tail_fusable(const nlml_mlp_plan* pl) {
    return pl->in_dims[head_t(0, 3)] == kHeadW && pl->out_dims[head_t(0, 3)] == kTailMid;
}
// heads' last two layers on the tensor cores: 128 -> 64 relu as a BN = 64 tile, 64 -> 1 as a dot product in its epilogue
// encoder.8 (128 -> 64 tanh) on the tensor cores with encoder.10 and the heads' model.0 in its epilogue
inline bool neck_on_tc(const nlml_mlp_plan* pl) {
    return pl->path == 0 && kTcNeck && neck_fusable(pl) && pl->tc[enc_t(3)] && pl->tc[enc_t(4)] && pl->tc[head_t(0, 1)];
}
inline bool tail_on_tc(const nlml_mlp_plan* pl) {
    return pl->path == 0 && kTcTail && pl->tc[head_t(0, 2)] && pl->tc[head_t(0, 3)];
}
constexpr size_t kNeckSmem = sizeof(float) * (kNeckIn * kNeckMid + tc::kRowsPerBlock * (kNeckIn + 4) + kNeckLat * kNeckMid +
                                              3 * kHeadW * kHeadIn + kNeckMid + kNeckLat + 3 * kHeadW);
constexpr size_t kTailSmem = sizeof(float) * (kHeadW * kTailMid + tc::kRowsPerBlock * (kHeadW + 4) + 2 * kTailMid + 4);

// one chunk (n <= pl->chunk samples) through the whole chain
int forward_chunk(nlml_mlp_plan* pl, const float* X, int64_t n, int64_t ldx, float* YPR, float* LAT_out, Workspace& w,
                  cudaStream_t st, int pre = 0) {
    if (n == 0) return 0;
    if (pre && !use_tc(pl, enc_t(0)))
        return set_error(NLML_E_UNSUPPORTED, "IPD normalisation is fused into the tensor-core path's operand split (path 0)");
    const bool fuse_neck = pl->path == 0 && neck_fusable(pl), fuse_tail = pl->path == 0 && tail_fusable(pl);
    // ---- encoder ----
    const float* cur_f32 = X;
    long long cur_ld = ldx;
    const __half *cur_hi = nullptr, *cur_lo = nullptr;
    if (use_tc(pl, enc_t(0))) {
        const int K = pl->in_dims[0], Kp = pl->Kp[0];
        const int vec_ok = (K % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
        const long long total = n * (Kp / 4);
        const unsigned grid = (unsigned)std::min<long long>(ceil_div(total, 256), (long long)pl->num_sms * 16);
        if (pre) tc::split_planes_kernel<true><<<grid, 256, 0, st>>>(X, n, ldx, K, Kp, vec_ok, w.hi[1], w.lo[1]);
        else tc::split_planes_kernel<false><<<grid, 256, 0, st>>>(X, n, ldx, K, Kp, vec_ok, w.hi[1], w.lo[1]);
        NLML_CUDA(cudaGetLastError());
        pl->launches += 1;
        cur_hi = w.hi[1]; cur_lo = w.lo[1];
    }
    const int enc_layers = fuse_neck ? 4 : kEnc;   // the neck kernel takes encoder.8 / encoder.10 / model.0
    for (int li = 0; li < enc_layers; ++li) {
        const int t = enc_t(li), par = li & 1;
        const bool last = li == kEnc - 1;
        const bool next_tc = !last && (!(fuse_neck && li == 3) || neck_on_tc(pl)) && use_tc(pl, enc_t(li + 1));
        float* yf = last ? (LAT_out ? LAT_out : w.lat) : (next_tc ? nullptr : w.f32[par]);
        __half* yh = next_tc ? w.hi[par] : nullptr;
        __half* yl = next_tc ? w.lo[par] : nullptr;
        if (use_tc(pl, t)) {
            if (int rc = launch_tc(pl, &t, 1, &cur_hi, &cur_lo, n, &yh, &yl, &yf, st)) return rc;
        } else {
            LinearArgs a{};
            a.X[0] = cur_f32; a.W[0] = pl->W[t]; a.B[0] = pl->B[t]; a.Y[0] = yf; a.Yhi[0] = yh; a.Ylo[0] = yl;
            a.ldx = cur_ld; a.ldy = pl->out_dims[t]; a.N = n; a.in = pl->in_dims[t]; a.out = pl->out_dims[t];
            a.act = act_of(t);
            if (int rc = launch_simt(pl, a, 1, st)) return rc;
        }
        cur_f32 = yf; cur_ld = pl->out_dims[t]; cur_hi = yh; cur_lo = yl;
    }
    // ---- heads: activations of head h live at offset h * chunk * width of the ping-pong buffers ----
    const float* hx[3];
    const __half *hh[3] = {nullptr, nullptr, nullptr}, *hl[3] = {nullptr, nullptr, nullptr};
    long long hld = pl->latent;
    int first_head_layer = 0;
    if (fuse_neck && neck_on_tc(pl)) {
        const int t4 = enc_t(4);
        tc::TcMaps maps;
        tc::LinearTcArgs a{};
        a.N = n; a.out = pl->out_dims[t4]; a.Kp = pl->Kp[t4]; a.act = act_of(t4); a.problems = 1; a.ldy = a.out; a.group = group_for(pl, a.Kp);
        if (int rc = tc::make_plane_map(&maps.a_hi[0], cur_hi, n, a.Kp, a.Kp, tc::BM)) return rc;
        if (int rc = tc::make_plane_map(&maps.a_lo[0], cur_lo, n, a.Kp, a.Kp, tc::BM)) return rc;
        maps.w_hi[0] = pl->wmap_hi[t4]; maps.w_lo[0] = pl->wmap_lo[t4];
        for (int z = 1; z < tc::kMaxProblems; ++z) { maps.a_hi[z] = maps.a_hi[0]; maps.a_lo[z] = maps.a_lo[0]; maps.w_hi[z] = maps.w_hi[0]; maps.w_lo[z] = maps.w_lo[0]; }
        a.y_tma = 1;   // the heads' first-layer planes leave through TMA (32-row x 64-column boxes, maps per head)
        for (int h = 0; h < 3; ++h) {
            const size_t off = (size_t)h * w.rows * kHeadW;
            if (int rc = tc::make_plane_map(&maps.y_hi[h], w.hi[0] + off, n, kHeadW, kHeadW, 32)) return rc;
            if (int rc = tc::make_plane_map(&maps.y_lo[h], w.lo[0] + off, n, kHeadW, kHeadW, 32)) return rc;
        }
        a.inv_scale[0] = pl->inv_scale[t4]; a.bias[0] = pl->B[t4];
        a.neck = 1; a.neck_w5 = pl->W[5]; a.neck_b5 = pl->B[5];
        a.neck_lat = LAT_out ? LAT_out : w.lat;
        for (int h = 0; h < 3; ++h) {
            const size_t off = (size_t)h * w.rows * kHeadW;
            a.neck_wh[h] = pl->W[head_t(h, 0)]; a.neck_bh[h] = pl->B[head_t(h, 0)];
            a.neck_hi[h] = w.hi[0] + off; a.neck_lo[h] = w.lo[0] + off;
            hx[h] = nullptr; hh[h] = a.neck_hi[h]; hl[h] = a.neck_lo[h];
        }
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, tc::BM), pl->num_sms);
        tc::linear_tc_kernel<64, 1><<<grid, tc::kThreads, tc::Cfg<64>::SMEM_BYTES, st>>>(maps, a);
        NLML_CUDA(cudaGetLastError());
        pl->launches += 1;
        hld = kHeadW;
        first_head_layer = 1;
    } else if (fuse_neck) {
        tc::NeckArgs a{};
        a.X = cur_f32; a.N = n;
        a.W4t = pl->Wt[4]; a.B4 = pl->B[4]; a.W5 = pl->W[5]; a.B5 = pl->B[5];
        a.LAT = LAT_out ? LAT_out : w.lat;
        const bool next_tc = use_tc(pl, head_t(0, 1));
        for (int h = 0; h < 3; ++h) {
            const size_t off = (size_t)h * w.rows * kHeadW;
            a.Wh[h] = pl->W[head_t(h, 0)]; a.Bh[h] = pl->B[head_t(h, 0)];
            a.Hhi[h] = (!LAT_out && next_tc) ? w.hi[0] + off : nullptr;
            a.Hlo[h] = (!LAT_out && next_tc) ? w.lo[0] + off : nullptr;
            a.Hf32[h] = (!LAT_out && !next_tc) ? w.f32[0] + off : nullptr;
            hx[h] = a.Hf32[h]; hh[h] = a.Hhi[h]; hl[h] = a.Hlo[h];
        }
        tc::neck_kernel<kNeckIn, kNeckMid, kNeckLat, kHeadIn, kHeadW><<<(unsigned)ceil_div(n, tc::kRowsPerBlock), tc::kNarrowThreads, kNeckSmem, st>>>(a);
        NLML_CUDA(cudaGetLastError());
        pl->launches += 1;
        hld = kHeadW;
        first_head_layer = 1;
    } else {
        for (int h = 0; h < 3; ++h) hx[h] = w.lat + h * pl->head_in;
    }
    if (LAT_out) return 0;
    for (int li = first_head_layer; li < kHead; ++li) {
        const int t0 = head_t(0, li), par = li & 1;
        const int out = pl->out_dims[t0];
        const bool last = li == kHead - 1;
        if (li == 3 && tail_on_tc(pl)) {
            const int ts[3] = {head_t(0, 3), head_t(1, 3), head_t(2, 3)}, ds[3] = {head_t(0, 4), head_t(1, 4), head_t(2, 4)};
            __half* none_h[3] = {nullptr, nullptr, nullptr};
            float* none_f[3] = {nullptr, nullptr, nullptr};
            if (int rc = launch_tc(pl, ts, 3, hh, hl, n, none_h, none_h, none_f, st, ds, YPR)) return rc;
            break;
        }
        if (fuse_tail && li == 3) {
            tc::HeadTailArgs a{};
            a.N = n; a.YPR = YPR;
            for (int h = 0; h < 3; ++h) {
                a.X[h] = hx[h];
                a.W3t[h] = pl->Wt[head_t(h, 3)]; a.B3[h] = pl->B[head_t(h, 3)];
                a.W4[h] = pl->W[head_t(h, 4)]; a.B4[h] = pl->B[head_t(h, 4)];
            }
            tc::head_tail_kernel<kHeadW, kTailMid><<<dim3((unsigned)ceil_div(n, tc::kRowsPerBlock), 3), tc::kNarrowThreads, kTailSmem, st>>>(a);
            NLML_CUDA(cudaGetLastError());
            pl->launches += 1;
            break;
        }
        const bool next_tc = !last && (!(fuse_tail && li == 2) || tail_on_tc(pl)) && use_tc(pl, head_t(0, li + 1));
        float* yf[3]; __half* yh[3]; __half* yl[3];
        for (int h = 0; h < 3; ++h) {
            const size_t off = (size_t)h * w.rows * out;
            yf[h] = last ? YPR + h : (next_tc ? nullptr : w.f32[par] + off);
            yh[h] = next_tc ? w.hi[par] + off : nullptr;
            yl[h] = next_tc ? w.lo[par] + off : nullptr;
        }
        if (use_tc(pl, t0)) {
            const int ts[3] = {head_t(0, li), head_t(1, li), head_t(2, li)};
            if (int rc = launch_tc(pl, ts, 3, hh, hl, n, yh, yl, yf, st)) return rc;
        } else {
            LinearArgs a{};
            for (int h = 0; h < 3; ++h) {
                const int t = head_t(h, li);
                a.X[h] = hx[h]; a.W[h] = pl->W[t]; a.B[h] = pl->B[t]; a.Y[h] = yf[h]; a.Yhi[h] = yh[h]; a.Ylo[h] = yl[h];
            }
            a.ldx = hld; a.ldy = last ? 3 : out; a.N = n; a.in = pl->in_dims[t0]; a.out = out; a.act = act_of(t0);
            if (int rc = launch_simt(pl, a, 3, st)) return rc;
        }
        for (int h = 0; h < 3; ++h) { hx[h] = yf[h]; hh[h] = yh[h]; hl[h] = yl[h]; }
        hld = out;
    }
    return 0;
}

int forward_device(nlml_mlp_plan* pl, const float* X, int64_t N, int64_t ldx, float* YPR, float* LAT, cudaStream_t st,
                   int pre = 0) {
    if (N == 0) return 0;
    if (int rc = ensure_workspace(pl, pl->ws_dev, N)) return rc;
    if (!pl->ws_dev_done) NLML_CUDA(cudaEventCreateWithFlags(&pl->ws_dev_done, cudaEventDisableTiming));
    else NLML_CUDA(cudaStreamWaitEvent(st, pl->ws_dev_done, 0));   // the previous call may have run on another stream
    const bool lanes = kDevLanes == 2 && N > pl->chunk;
    if (lanes) {
        if (int rc = ensure_workspace(pl, pl->ws_lane, N - pl->chunk)) return rc;
        if (!pl->lane_fork) {
            for (int i = 0; i < 2; ++i) {
                NLML_CUDA(cudaStreamCreateWithFlags(&pl->lane[i], cudaStreamNonBlocking));
                NLML_CUDA(cudaEventCreateWithFlags(&pl->lane_join[i], cudaEventDisableTiming));
            }
            NLML_CUDA(cudaEventCreateWithFlags(&pl->lane_fork, cudaEventDisableTiming));
        }
        NLML_CUDA(cudaEventRecord(pl->lane_fork, st));
        for (int i = 0; i < 2; ++i) NLML_CUDA(cudaStreamWaitEvent(pl->lane[i], pl->lane_fork, 0));
    }
    int slot = 0, rc = 0;
    for (int64_t s0 = 0; s0 < N && !rc; s0 += pl->chunk, slot ^= 1) {
        const int64_t n = std::min<int64_t>(pl->chunk, N - s0);
        rc = forward_chunk(pl, X + s0 * ldx, n, ldx, YPR ? YPR + s0 * 3 : nullptr, LAT ? LAT + s0 * pl->latent : nullptr,
                           lanes && slot ? pl->ws_lane : pl->ws_dev, lanes ? pl->lane[slot] : st, pre);
    }
    if (lanes)   // joined on the error path too: the caller's stream must not run ahead of launched chunks
        for (int i = 0; i < 2; ++i) {
            NLML_CUDA(cudaEventRecord(pl->lane_join[i], pl->lane[i]));
            NLML_CUDA(cudaStreamWaitEvent(st, pl->lane_join[i], 0));
        }
    NLML_CUDA(cudaEventRecord(pl->ws_dev_done, st));
    return rc;
}

// W[out][in] FP32 -> power-of-two-scaled FP16 hi/lo planes [out][Kp]
int prepare_tc_layer(nlml_mlp_plan* pl, int t, const float* Wh) {
    const int out = pl->out_dims[t], in = pl->in_dims[t], Kp = (in + tc::BK - 1) / tc::BK * tc::BK;
    float maxabs = 0.f;
    for (size_t i = 0; i < (size_t)out * in; ++i) maxabs = std::max(maxabs, std::fabs(Wh[i]));
    // scale so the largest weight sits in [256, 512): the lo plane then stays in FP16's normal range
    int e = 0;
    if (maxabs > 0.f && std::isfinite(maxabs)) e = 8 - (int)std::floor(std::log2(maxabs));
    e = std::max(-14, std::min(e, 24));
    const float scale = std::ldexp(1.0f, e);
    std::vector<__half> hi((size_t)out * Kp, __float2half_rn(0.f)), lo((size_t)out * Kp, __float2half_rn(0.f));
    for (int o = 0; o < out; ++o)
        for (int k = 0; k < in; ++k) {
            const float v = Wh[(size_t)o * in + k] * scale;
            const __half h = __float2half_rn(v);
            hi[(size_t)o * Kp + k] = h;
            lo[(size_t)o * Kp + k] = __float2half_rn(v - __half2float(h));
        }
    const size_t bytes = sizeof(__half) * (size_t)out * Kp;
    NLML_CUDA(cudaMalloc(&pl->Whi[t], bytes));
    NLML_CUDA(cudaMalloc(&pl->Wlo[t], bytes));
    NLML_CUDA(cudaMemcpy(pl->Whi[t], hi.data(), bytes, cudaMemcpyHostToDevice));
    NLML_CUDA(cudaMemcpy(pl->Wlo[t], lo.data(), bytes, cudaMemcpyHostToDevice));
    pl->Kp[t] = Kp;
    pl->inv_scale[t] = std::ldexp(1.0f, -e);
    const int box_rows = std::min(out, 128);   // 256-wide tiles are loaded as two 128-row halves, one per CTA of the cluster
    if (int rc = tc::make_plane_map(&pl->wmap_hi[t], pl->Whi[t], out, Kp, Kp, box_rows)) return rc;
    if (int rc = tc::make_plane_map(&pl->wmap_lo[t], pl->Wlo[t], out, Kp, Kp, box_rows)) return rc;
    return 0;
}

}  // namespace

extern "C" int nlml_mlp_plan_create(const float* const* weights, const float* const* biases, const int* out_dims,
                                    const int* in_dims, int device, nlml_mlp_plan** plan_out) {
    if (!weights || !biases || !out_dims || !in_dims || !plan_out) return set_error(NLML_E_INVALID, "null pointer argument");
    for (int t = 0; t < kNumT; ++t) {
        if (!weights[t] || !biases[t]) return set_error(NLML_E_INVALID, "null tensor %d", t);
        if (out_dims[t] < 1 || in_dims[t] < 1) return set_error(NLML_E_INVALID, "tensor %d has non-positive shape", t);
    }
    for (int li = 1; li < kEnc; ++li)
        if (in_dims[li] != out_dims[li - 1]) return set_error(NLML_E_INVALID, "encoder layer %d input %d != previous output %d", li, in_dims[li], out_dims[li - 1]);
    const int latent = out_dims[kEnc - 1];
    const int head_in = in_dims[kEnc];
    if (latent != 3 * head_in) return set_error(NLML_E_INVALID, "latent width %d != 3 x head input %d", latent, head_in);
    for (int h = 0; h < 3; ++h)
        for (int li = 0; li < kHead; ++li) {
            const int t = head_t(h, li), t0 = head_t(0, li);
            if (out_dims[t] != out_dims[t0] || in_dims[t] != in_dims[t0]) return set_error(NLML_E_INVALID, "heads must share one architecture");
            if (li > 0 && in_dims[t] != out_dims[t - 1]) return set_error(NLML_E_INVALID, "head layer %d shape mismatch", li);
        }
    if (out_dims[kNumT - 1] != 1) return set_error(NLML_E_INVALID, "head output width must be 1");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(NLML_E_NO_DEVICE, "cudaSetDevice(%d) failed", device);

    // owned until the end of this function: every early error return below frees the plan and its device buffers
    struct PlanOwner {
        nlml_mlp_plan* p;
        ~PlanOwner() { if (p) nlml_mlp_plan_destroy(p); }
    } owner{new nlml_mlp_plan()};
    nlml_mlp_plan* pl = owner.p;
    pl->device = device;
    pl->input_size = in_dims[0];
    pl->latent = latent;
    pl->head_in = head_in;
    cudaDeviceProp prop;
    NLML_CUDA(cudaGetDeviceProperties(&prop, device));
    pl->num_sms = prop.multiProcessorCount;
    for (int t = 0; t < kNumT; ++t) {
        pl->out_dims[t] = out_dims[t];
        pl->in_dims[t] = in_dims[t];
        const size_t wb = sizeof(float) * (size_t)out_dims[t] * in_dims[t];
        NLML_CUDA(cudaMalloc(&pl->W[t], wb));
        NLML_CUDA(cudaMalloc(&pl->B[t], sizeof(float) * out_dims[t]));
        NLML_CUDA(cudaMemcpy(pl->W[t], weights[t], wb, cudaMemcpyHostToDevice));
        NLML_CUDA(cudaMemcpy(pl->B[t], biases[t], sizeof(float) * out_dims[t], cudaMemcpyHostToDevice));
    }
    // transposed copies [in][out] of the layers the fused narrow-layer kernels broadcast from shared memory
    for (int t = 0; t < kNumT; ++t) {
        const bool neck_l = t == 4, tail_l = t >= kEnc && (t - kEnc) % kHead == 3;
        if (!neck_l && !tail_l) continue;
        const int o = out_dims[t], in = in_dims[t];
        std::vector<float> tr((size_t)o * in);
        for (int j = 0; j < o; ++j)
            for (int k = 0; k < in; ++k) tr[(size_t)k * o + j] = weights[t][(size_t)j * in + k];
        NLML_CUDA(cudaMalloc(&pl->Wt[t], sizeof(float) * tr.size()));
        NLML_CUDA(cudaMemcpy(pl->Wt[t], tr.data(), sizeof(float) * tr.size(), cudaMemcpyHostToDevice));
    }
    // tensor-core eligibility: a real dense contraction (wide output, deep reduction) whose operand planes
    // line up with the 64-element k-blocks; the first encoder layer gets its planes from split_planes_kernel
    // (which pads), every other layer from its producer's epilogue (so in must already be a multiple of 64)
    for (int t = 0; t < kNumT; ++t) {
        const bool first = t == 0;
        pl->tc[t] = out_dims[t] % 128 == 0 && in_dims[t] >= 64 && (first || in_dims[t] % 64 == 0);
    }
    // encoder.8 (128 -> 64) rides the tensor cores as a 64-wide tile when the fused neck epilogue applies
    if (neck_fusable(pl)) pl->tc[enc_t(4)] = true;
    // the heads' 128 -> 64 layer rides the tensor cores as a 64-wide tile when its successor is the single-output layer
    if (tail_fusable(pl) && out_dims[head_t(0, 4)] == 1 && in_dims[head_t(0, 4)] == kTailMid) pl->tc[head_t(0, 3)] = true;
    for (int h = 1; h < 3; ++h)
        for (int li = 0; li < kHead; ++li) pl->tc[head_t(h, li)] = pl->tc[head_t(0, li)];
    for (int t = 0; t < kNumT; ++t)
        if (pl->tc[t])
            if (int rc = prepare_tc_layer(pl, t, weights[t])) return rc;
    NLML_CUDA(cudaFuncSetAttribute(tc::linear_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::Cfg2::SMEM_BYTES));
    NLML_CUDA(cudaFuncSetAttribute(tc::linear_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::Cfg<64>::SMEM_BYTES));
    pl->chunk = (int64_t)pl->num_sms * 128 * NLML_MLP_CHUNK_WAVES;
    NLML_CUDA(cudaFuncSetAttribute(tc::linear_tc_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::Cfg<256>::SMEM_BYTES));
    NLML_CUDA(cudaFuncSetAttribute(tc::linear_tc_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::Cfg<128>::SMEM_BYTES));
    NLML_CUDA(cudaFuncSetAttribute(tc::neck_kernel<kNeckIn, kNeckMid, kNeckLat, kHeadIn, kHeadW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNeckSmem));
    NLML_CUDA(cudaFuncSetAttribute(tc::head_tail_kernel<kHeadW, kTailMid>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmem));
    // workspace widths (elements per sample) by layer parity, sized for either path
    for (int li = 0; li < kEnc - 1; ++li) {
        pl->f32_width[li & 1] = std::max<size_t>(pl->f32_width[li & 1], out_dims[li]);
        if (pl->tc[enc_t(li + 1)]) pl->plane_width[li & 1] = std::max<size_t>(pl->plane_width[li & 1], out_dims[li]);
    }
    if (pl->tc[0]) pl->plane_width[1] = std::max<size_t>(pl->plane_width[1], pl->Kp[0]);
    for (int li = 0; li < kHead - 1; ++li) {
        const size_t w3 = 3 * (size_t)out_dims[head_t(0, li)];
        pl->f32_width[li & 1] = std::max(pl->f32_width[li & 1], w3);
        if (pl->tc[head_t(0, li + 1)]) pl->plane_width[li & 1] = std::max(pl->plane_width[li & 1], w3);
    }
    for (int i = 0; i < 2; ++i) pl->f32_width[i] = std::max<size_t>(pl->f32_width[i], 4);
    owner.p = nullptr;
    *plan_out = pl;
    return 0;
}

extern "C" void nlml_mlp_plan_destroy(nlml_mlp_plan* pl) {
    if (!pl) return;
    DeviceGuard guard(pl->device);
    for (int t = 0; t < kNumT; ++t) {
        cudaFree(pl->W[t]); cudaFree(pl->B[t]); cudaFree(pl->Whi[t]); cudaFree(pl->Wlo[t]); cudaFree(pl->Wt[t]);
    }
    for (int i = 0; i < 2; ++i) {
        if (pl->streams[i]) cudaStreamDestroy(pl->streams[i]);
        cudaFree(pl->x_dev[i]);
        cudaFree(pl->y_dev[i]);
        if (pl->y_stage[i]) cudaFreeHost(pl->y_stage[i]);
        free_workspace(pl->ws_host[i]);
    }
    free_workspace(pl->ws_dev);
    free_workspace(pl->ws_lane);
    for (int i = 0; i < 2; ++i) {
        if (pl->lane[i]) cudaStreamDestroy(pl->lane[i]);
        if (pl->lane_join[i]) cudaEventDestroy(pl->lane_join[i]);
    }
    if (pl->lane_fork) cudaEventDestroy(pl->lane_fork);
    if (pl->ws_dev_done) cudaEventDestroy(pl->ws_dev_done);
    delete pl;
}

extern "C" int nlml_mlp_set_path(nlml_mlp_plan* pl, int path) {
    if (!pl || path < 0 || path > 1) return set_error(NLML_E_INVALID, "path must be 0 (tensor-core chain) or 1 (FP32 CUDA-core chain)");
    pl->path = path;
    return 0;
}

extern "C" int nlml_mlp_forward_f32(nlml_mlp_plan* pl, const float* X_dev, int64_t N, int64_t ldx, float* YPR_out_dev,
                                    void* stream) {
    if (!pl || (N > 0 && (!X_dev || !YPR_out_dev))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes: N=%lld ldx=%lld (input_size=%d)", (long long)N, (long long)ldx, pl->input_size);
    DeviceGuard guard(pl->device);
    return forward_device(pl, X_dev, N, ldx, YPR_out_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int nlml_mlp_forward_landmarks_f32(nlml_mlp_plan* pl, const float* LM_dev, int64_t N, int64_t ldx,
                                              float* YPR_out_dev, void* stream) {
    if (!pl || (N > 0 && (!LM_dev || !YPR_out_dev))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes: N=%lld ldx=%lld (input_size=%d)", (long long)N, (long long)ldx, pl->input_size);
    if (pl->input_size % 3 != 0 || pl->input_size < 3 * 264)
        return set_error(NLML_E_INVALID, "IPD normalisation needs x,y,z triples up to landmark 263 (input_size=%d)", pl->input_size);
    DeviceGuard guard(pl->device);
    return forward_device(pl, LM_dev, N, ldx, YPR_out_dev, nullptr, (cudaStream_t)stream, 1);
}

extern "C" int nlml_pose_postprocess_f64(const float* YPR_dev, int64_t N, int decimals, double ema_alpha,
                                         double* DEG_out_dev, void* stream) {
    if (N > 0 && (!YPR_dev || !DEG_out_dev)) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || decimals < 0 || decimals > 15 || !(ema_alpha < 1.0))
        return set_error(NLML_E_INVALID, "bad arguments: N=%lld decimals=%d (0..15) ema_alpha=%g (< 1; <= 0 disables)", (long long)N, decimals, ema_alpha);
    if (N == 0) return 0;
    double scale = 1.0;
    for (int i = 0; i < decimals; ++i) scale *= 10.0;   // exact in double for decimals <= 22, as numpy's power-of-ten table
    const double alpha = ema_alpha > 0.0 ? ema_alpha : 0.0;
    const unsigned grid = alpha > 0.0 ? 1u : (unsigned)std::min<int64_t>(ceil_div(3 * N, 128), 148 * 8);
    tc::pose_post_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(YPR_dev, N, scale, alpha, 1.0 - alpha, DEG_out_dev);
    NLML_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int nlml_mlp_latent_f32(nlml_mlp_plan* pl, const float* X_dev, int64_t N, int64_t ldx, float* LAT_out_dev,
                                   void* stream) {
    if (!pl || (N > 0 && (!X_dev || !LAT_out_dev))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes");
    DeviceGuard guard(pl->device);
    return forward_device(pl, X_dev, N, ldx, nullptr, LAT_out_dev, (cudaStream_t)stream);
}

namespace {
int forward_host(nlml_mlp_plan* pl, const float* X_host, int64_t N, int64_t ldx, float* YPR_out_host, int pre);
}

extern "C" int nlml_mlp_forward_host_f32(nlml_mlp_plan* pl, const float* X_host, int64_t N, int64_t ldx,
                                         float* YPR_out_host) {
    if (!pl || (N > 0 && (!X_host || !YPR_out_host))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes");
    return forward_host(pl, X_host, N, ldx, YPR_out_host, 0);
}

extern "C" int nlml_mlp_forward_landmarks_host_f32(nlml_mlp_plan* pl, const float* LM_host, int64_t N, int64_t ldx,
                                                   float* YPR_out_host) {
    if (!pl || (N > 0 && (!LM_host || !YPR_out_host))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes");
    if (pl->input_size % 3 != 0 || pl->input_size < 3 * 264)
        return set_error(NLML_E_INVALID, "IPD normalisation needs x,y,z triples up to landmark 263 (input_size=%d)", pl->input_size);
    return forward_host(pl, LM_host, N, ldx, YPR_out_host, 1);
}

namespace {
int forward_host(nlml_mlp_plan* pl, const float* X_host, int64_t N, int64_t ldx, float* YPR_out_host, int pre) {
    DeviceGuard guard(pl->device);
    const int F = pl->input_size;
    if (!pl->streams[0])
        for (int i = 0; i < 2; ++i) NLML_CUDA(cudaStreamCreateWithFlags(&pl->streams[i], cudaStreamNonBlocking));
    const int64_t want = std::min<int64_t>(pl->chunk, ceil_div(std::max<int64_t>(N, 1), 128) * 128);
    if (pl->host_rows < want) {
        if (pl->host_rows) NLML_CUDA(cudaDeviceSynchronize());
        pl->host_rows = 0;   // nothing below is usable until every allocation has succeeded
        for (int i = 0; i < 2; ++i) {
            cudaFree(pl->x_dev[i]); cudaFree(pl->y_dev[i]);
            if (pl->y_stage[i]) cudaFreeHost(pl->y_stage[i]);
            pl->x_dev[i] = pl->y_dev[i] = pl->y_stage[i] = nullptr;
        }
        for (int i = 0; i < 2; ++i) {
            NLML_CUDA(cudaMallocHost(&pl->y_stage[i], sizeof(float) * want * 3));
            NLML_CUDA(cudaMalloc(&pl->x_dev[i], sizeof(float) * want * F));
            NLML_CUDA(cudaMalloc(&pl->y_dev[i], sizeof(float) * want * 3));
        }
        pl->host_rows = want;
    }
    for (int i = 0; i < 2; ++i)
        if (int rc = ensure_workspace(pl, pl->ws_host[i], N)) return rc;
    // results go to pinned staging: a D2H copy into pageable user memory would block the host thread until the
    // chunk's kernels are done, so the next chunk's H2D copy could not overlap them
    struct Pending { int64_t s0 = 0, n = 0; } pending[2];
    auto drain = [&](int slot) -> int {
        if (!pending[slot].n) return 0;
        NLML_CUDA(cudaStreamSynchronize(pl->streams[slot]));
        std::memcpy(YPR_out_host + pending[slot].s0 * 3, pl->y_stage[slot], sizeof(float) * 3 * pending[slot].n);
        pending[slot].n = 0;
        return 0;
    };
    int slot = 0;
    for (int64_t s0 = 0; s0 < N; s0 += pl->chunk, slot ^= 1) {
        const int64_t n = std::min<int64_t>(pl->chunk, N - s0);
        cudaStream_t st = pl->streams[slot];
        if (int rc = drain(slot)) return rc;
        if (ldx == F)   // contiguous rows: one linear DMA instead of a pitched copy
            NLML_CUDA(cudaMemcpyAsync(pl->x_dev[slot], X_host + s0 * ldx, sizeof(float) * F * n, cudaMemcpyHostToDevice, st));
        else
            NLML_CUDA(cudaMemcpy2DAsync(pl->x_dev[slot], sizeof(float) * F, X_host + s0 * ldx, sizeof(float) * ldx,
                                        sizeof(float) * F, (size_t)n, cudaMemcpyHostToDevice, st));
        if (int rc = forward_chunk(pl, pl->x_dev[slot], n, F, pl->y_dev[slot], nullptr, pl->ws_host[slot], st, pre)) return rc;
        NLML_CUDA(cudaMemcpyAsync(pl->y_stage[slot], pl->y_dev[slot], sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
        pending[slot].s0 = s0;
        pending[slot].n = n;
    }
    if (int rc = drain(slot)) return rc;
    if (int rc = drain(slot ^ 1)) return rc;
    return 0;
}
}  // namespace

extern "C" int64_t nlml_mlp_launch_count(const nlml_mlp_plan* pl) { return pl ? pl->launches : 0; }

#ifdef NLML_MLP_TIMING
extern "C" void nlml_debug_mlp_timing(float* dev_buf, int slices) { g_mlp_timing_buf = dev_buf; g_mlp_timing_slices = slices; g_mlp_timing_next = 0; }
#endif
