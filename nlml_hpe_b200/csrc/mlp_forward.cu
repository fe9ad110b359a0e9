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
#include <vector>

#include "common.cuh"

namespace nlml {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct LinearArgs {
    // up to 3 independent problems of identical shape (grid.z): the three heads
    const float* X[3];
    const float* W[3];
    const float* B[3];
    float* Y[3];
    long long ldx, ldy;
    long long N;  // rows (samples)
    int in, out;
    int act;
    int vec_x;  // X rows 16B aligned (ldx % 4 == 0, base aligned) and in % 4 == 0
    int vec_w;  // W rows 16B aligned (in % 4 == 0)
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
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
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long row = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (row >= a.N) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
            if (col < a.out) Y[row * a.ldy + col] = apply_act(acc[i][j] + __ldg(Bv + col), a.act);
        }
    }
}

}  // namespace nlml

// =============================================================================================
// C ABI
// =============================================================================================
using namespace nlml;

struct nlml_mlp_plan {
    int device = 0;
    int out_dims[NLML_MLP_NUM_TENSORS];
    int in_dims[NLML_MLP_NUM_TENSORS];
    float* W[NLML_MLP_NUM_TENSORS] = {};
    float* B[NLML_MLP_NUM_TENSORS] = {};
    int input_size = 0, latent = 0, head_in = 0;
    int64_t chunk = 16384;  // samples per pass; activations of a chunk stay L2-resident
    float* bufA = nullptr;  // [chunk][max even-layer width]
    float* bufB = nullptr;  // [chunk][max odd-layer width]
    float* lat = nullptr;   // [chunk][latent]
    size_t widthA = 0, widthB = 0;
    int64_t launches = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    float* x_dev[2] = {nullptr, nullptr};
    float* y_dev[2] = {nullptr, nullptr};
    // second workspace set so the two host-path streams do not share activations
    float* bufA2 = nullptr;
    float* bufB2 = nullptr;
    float* lat2 = nullptr;
};

namespace {

int launch_linear(nlml_mlp_plan* pl, LinearArgs& a, int nz, cudaStream_t st) {
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

// one chunk (n <= pl->chunk samples) through the whole chain
int forward_chunk(nlml_mlp_plan* pl, const float* X, int64_t n, int64_t ldx, float* YPR, float* LAT_out,
                  float* bufA, float* bufB, float* lat, cudaStream_t st) {
    const float* cur = X;
    long long ld = ldx;
    for (int li = 0; li < NLML_MLP_ENCODER_LAYERS; ++li) {
        LinearArgs a{};
        const bool last = li == NLML_MLP_ENCODER_LAYERS - 1;
        float* dst = last ? (LAT_out ? LAT_out : lat) : (li % 2 == 0 ? bufA : bufB);
        a.X[0] = cur; a.W[0] = pl->W[li]; a.B[0] = pl->B[li]; a.Y[0] = dst;
        a.ldx = ld; a.ldy = pl->out_dims[li]; a.N = n; a.in = pl->in_dims[li]; a.out = pl->out_dims[li];
        a.act = li < 4 ? ACT_RELU : (li == 4 ? ACT_TANH : ACT_NONE);
        if (int rc = launch_linear(pl, a, 1, st)) return rc;
        cur = dst;
        ld = pl->out_dims[li];
    }
    if (LAT_out) return 0;
    // heads: three problems per launch; activations for head h at offset h*chunk*width
    const float* hx[3];
    long long hld = pl->latent;
    for (int h = 0; h < 3; ++h) hx[h] = lat + h * pl->head_in;
    for (int li = 0; li < NLML_MLP_HEAD_LAYERS; ++li) {
        LinearArgs a{};
        const bool last = li == NLML_MLP_HEAD_LAYERS - 1;
        float* base = (li % 2 == 0) ? bufA : bufB;
        const int t0 = NLML_MLP_ENCODER_LAYERS + li;
        const int out = pl->out_dims[t0];
        for (int h = 0; h < 3; ++h) {
            const int t = NLML_MLP_ENCODER_LAYERS + h * NLML_MLP_HEAD_LAYERS + li;
            a.X[h] = hx[h]; a.W[h] = pl->W[t]; a.B[h] = pl->B[t];
            a.Y[h] = last ? YPR + h : base + (size_t)h * pl->chunk * out;
        }
        a.ldx = hld; a.ldy = last ? 3 : out; a.N = n; a.in = pl->in_dims[t0]; a.out = out;
        a.act = last ? ACT_NONE : ACT_RELU;
        if (int rc = launch_linear(pl, a, 3, st)) return rc;
        for (int h = 0; h < 3; ++h) hx[h] = a.Y[h];
        hld = out;
    }
    return 0;
}

int forward_device(nlml_mlp_plan* pl, const float* X, int64_t N, int64_t ldx, float* YPR, float* LAT, cudaStream_t st) {
    for (int64_t s0 = 0; s0 < N; s0 += pl->chunk) {
        const int64_t n = std::min<int64_t>(pl->chunk, N - s0);
        if (int rc = forward_chunk(pl, X + s0 * ldx, n, ldx, YPR ? YPR + s0 * 3 : nullptr,
                                   LAT ? LAT + s0 * pl->latent : nullptr, pl->bufA, pl->bufB, pl->lat, st))
            return rc;
    }
    return 0;
}

}  // namespace

extern "C" int nlml_mlp_plan_create(const float* const* weights, const float* const* biases, const int* out_dims,
                                    const int* in_dims, int device, nlml_mlp_plan** plan_out) {
    if (!weights || !biases || !out_dims || !in_dims || !plan_out) return set_error(NLML_E_INVALID, "null pointer argument");
    for (int t = 0; t < NLML_MLP_NUM_TENSORS; ++t) {
        if (!weights[t] || !biases[t]) return set_error(NLML_E_INVALID, "null tensor %d", t);
        if (out_dims[t] < 1 || in_dims[t] < 1) return set_error(NLML_E_INVALID, "tensor %d has non-positive shape", t);
    }
    for (int li = 1; li < NLML_MLP_ENCODER_LAYERS; ++li)
        if (in_dims[li] != out_dims[li - 1]) return set_error(NLML_E_INVALID, "encoder layer %d input %d != previous output %d", li, in_dims[li], out_dims[li - 1]);
    const int latent = out_dims[NLML_MLP_ENCODER_LAYERS - 1];
    const int head_in = in_dims[NLML_MLP_ENCODER_LAYERS];
    if (latent != 3 * head_in) return set_error(NLML_E_INVALID, "latent width %d != 3 x head input %d", latent, head_in);
    for (int h = 0; h < 3; ++h)
        for (int li = 0; li < NLML_MLP_HEAD_LAYERS; ++li) {
            const int t = NLML_MLP_ENCODER_LAYERS + h * NLML_MLP_HEAD_LAYERS + li;
            const int t0 = NLML_MLP_ENCODER_LAYERS + li;
            if (out_dims[t] != out_dims[t0] || in_dims[t] != in_dims[t0]) return set_error(NLML_E_INVALID, "heads must share one architecture");
            if (li > 0 && in_dims[t] != out_dims[t - 1]) return set_error(NLML_E_INVALID, "head layer %d shape mismatch", li);
        }
    if (out_dims[NLML_MLP_NUM_TENSORS - 1] != 1) return set_error(NLML_E_INVALID, "head output width must be 1");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(NLML_E_NO_DEVICE, "cudaSetDevice(%d) failed", device);

    auto* pl = new nlml_mlp_plan();
    pl->device = device;
    pl->input_size = in_dims[0];
    pl->latent = latent;
    pl->head_in = head_in;
    for (int t = 0; t < NLML_MLP_NUM_TENSORS; ++t) {
        pl->out_dims[t] = out_dims[t];
        pl->in_dims[t] = in_dims[t];
        const size_t wb = sizeof(float) * (size_t)out_dims[t] * in_dims[t];
        NLML_CUDA(cudaMalloc(&pl->W[t], wb));
        NLML_CUDA(cudaMalloc(&pl->B[t], sizeof(float) * out_dims[t]));
        NLML_CUDA(cudaMemcpy(pl->W[t], weights[t], wb, cudaMemcpyHostToDevice));
        NLML_CUDA(cudaMemcpy(pl->B[t], biases[t], sizeof(float) * out_dims[t], cudaMemcpyHostToDevice));
    }
    for (int li = 0; li < NLML_MLP_ENCODER_LAYERS - 1; ++li) {
        size_t& w = (li % 2 == 0) ? pl->widthA : pl->widthB;
        w = std::max<size_t>(w, out_dims[li]);
    }
    for (int li = 0; li < NLML_MLP_HEAD_LAYERS - 1; ++li) {
        size_t& w = (li % 2 == 0) ? pl->widthA : pl->widthB;
        w = std::max<size_t>(w, 3 * (size_t)out_dims[NLML_MLP_ENCODER_LAYERS + li]);
    }
    NLML_CUDA(cudaMalloc(&pl->bufA, sizeof(float) * pl->chunk * pl->widthA));
    NLML_CUDA(cudaMalloc(&pl->bufB, sizeof(float) * pl->chunk * pl->widthB));
    NLML_CUDA(cudaMalloc(&pl->lat, sizeof(float) * pl->chunk * latent));
    *plan_out = pl;
    return 0;
}

extern "C" void nlml_mlp_plan_destroy(nlml_mlp_plan* pl) {
    if (!pl) return;
    DeviceGuard guard(pl->device);
    for (int t = 0; t < NLML_MLP_NUM_TENSORS; ++t) {
        cudaFree(pl->W[t]);
        cudaFree(pl->B[t]);
    }
    for (int i = 0; i < 2; ++i) {
        if (pl->streams[i]) cudaStreamDestroy(pl->streams[i]);
        cudaFree(pl->x_dev[i]);
        cudaFree(pl->y_dev[i]);
    }
    cudaFree(pl->bufA); cudaFree(pl->bufB); cudaFree(pl->lat);
    cudaFree(pl->bufA2); cudaFree(pl->bufB2); cudaFree(pl->lat2);
    delete pl;
}

extern "C" int nlml_mlp_forward_f32(nlml_mlp_plan* pl, const float* X_dev, int64_t N, int64_t ldx, float* YPR_out_dev,
                                    void* stream) {
    if (!pl || (N > 0 && (!X_dev || !YPR_out_dev))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes: N=%lld ldx=%lld (input_size=%d)", (long long)N, (long long)ldx, pl->input_size);
    DeviceGuard guard(pl->device);
    return forward_device(pl, X_dev, N, ldx, YPR_out_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int nlml_mlp_latent_f32(nlml_mlp_plan* pl, const float* X_dev, int64_t N, int64_t ldx, float* LAT_out_dev,
                                   void* stream) {
    if (!pl || (N > 0 && (!X_dev || !LAT_out_dev))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes");
    DeviceGuard guard(pl->device);
    return forward_device(pl, X_dev, N, ldx, nullptr, LAT_out_dev, (cudaStream_t)stream);
}

extern "C" int nlml_mlp_forward_host_f32(nlml_mlp_plan* pl, const float* X_host, int64_t N, int64_t ldx,
                                         float* YPR_out_host) {
    if (!pl || (N > 0 && (!X_host || !YPR_out_host))) return set_error(NLML_E_INVALID, "null pointer argument");
    if (N < 0 || ldx < pl->input_size) return set_error(NLML_E_INVALID, "bad sizes");
    DeviceGuard guard(pl->device);
    const int F = pl->input_size;
    if (!pl->streams[0]) {
        for (int i = 0; i < 2; ++i) {
            NLML_CUDA(cudaStreamCreateWithFlags(&pl->streams[i], cudaStreamNonBlocking));
            NLML_CUDA(cudaMalloc(&pl->x_dev[i], sizeof(float) * pl->chunk * F));
            NLML_CUDA(cudaMalloc(&pl->y_dev[i], sizeof(float) * pl->chunk * 3));
        }
        NLML_CUDA(cudaMalloc(&pl->bufA2, sizeof(float) * pl->chunk * pl->widthA));
        NLML_CUDA(cudaMalloc(&pl->bufB2, sizeof(float) * pl->chunk * pl->widthB));
        NLML_CUDA(cudaMalloc(&pl->lat2, sizeof(float) * pl->chunk * pl->latent));
    }
    int slot = 0;
    for (int64_t s0 = 0; s0 < N; s0 += pl->chunk, slot ^= 1) {
        const int64_t n = std::min<int64_t>(pl->chunk, N - s0);
        cudaStream_t st = pl->streams[slot];
        NLML_CUDA(cudaMemcpy2DAsync(pl->x_dev[slot], sizeof(float) * F, X_host + s0 * ldx, sizeof(float) * ldx,
                                    sizeof(float) * F, (size_t)n, cudaMemcpyHostToDevice, st));
        if (int rc = forward_chunk(pl, pl->x_dev[slot], n, F, pl->y_dev[slot], nullptr, slot ? pl->bufA2 : pl->bufA,
                                   slot ? pl->bufB2 : pl->bufB, slot ? pl->lat2 : pl->lat, st))
            return rc;
        NLML_CUDA(cudaMemcpyAsync(YPR_out_host + s0 * 3, pl->y_dev[slot], sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    }
    NLML_CUDA(cudaStreamSynchronize(pl->streams[0]));
    NLML_CUDA(cudaStreamSynchronize(pl->streams[1]));
    return 0;
}

extern "C" int64_t nlml_mlp_launch_count(const nlml_mlp_plan* pl) { return pl ? pl->launches : 0; }
