// fp32 SIMT GEMMs with fused epilogues -- the strict-parity (1e-5) path for the dense layers of the
// encoder (K7, K9-K11) and their backward.  64x64x16 tiles, 256 threads, 4x4 register micro-tiles,
// register-prefetch double buffering, 128-bit shared-memory reads.
#include "common.cuh"

#define BM 64
#define BN 64
#define BK 16
#define PAD 4

__device__ __forceinline__ float4 load4_guard(const float* __restrict__ base, long long row, long long ld, int col,
                                              int rows, int cols, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) {
        const float* p = base + row * ld + col;
        if (vec_ok && col + 3 < cols) return __ldg(reinterpret_cast<const float4*>(p));
        if (col < cols) v.x = __ldg(p);
        if (col + 1 < cols) v.y = __ldg(p + 1);
        if (col + 2 < cols) v.z = __ldg(p + 2);
        if (col + 3 < cols) v.w = __ldg(p + 3);
    }
    return v;
}

struct EpilogueArgs {
    const float* bias;
    float* pre_act;
    int act;
    const float* mul_gelu_grad_of;
    float p_drop, inv_keep;
    uint64_t seed;
    uint32_t site;
    const float* residual;
};

__device__ __forceinline__ float apply_epilogue(float v, long long m, int n, int N, const EpilogueArgs& e) {
    if (e.bias) v += __ldg(e.bias + n);
    const long long idx = m * N + n;
    if (e.pre_act) e.pre_act[idx] = v;
    if (e.act == ASME_ACT_GELU) v = gelu_erf(v);
    if (e.mul_gelu_grad_of) v *= gelu_erf_grad(__ldg(e.mul_gelu_grad_of + idx));
    if (e.p_drop > 0.f) v *= dropout_scale(asme_seed(e.seed), e.site, (uint64_t)idx, e.p_drop, e.inv_keep);
    if (e.residual) v += __ldg(e.residual + idx);
    return v;
}

// C[M,N] = epi(A[M,K] * op(B)),  TRANS_B: B is (N,K) row-major; else B is (K,N) row-major
template <bool TRANS_B>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                   float* __restrict__ C, int M, int N, int K, EpilogueArgs epi,
                                                   bool a_vec, bool b_vec, const int32_t* __restrict__ m_live) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    M = asme_live_rows(M, m_live);
    if (m0 >= M) return;

    // loader coordinates
    const int a_row = tid / 4, a_k = (tid % 4) * 4;          // A tile: 64 rows x 16 k
    const int bt_row = tid / 4, bt_k = (tid % 4) * 4;        // TRANS_B: 64 n-rows x 16 k
    const int bn_k = tid / 16, bn_col = (tid % 16) * 4;      // !TRANS_B: 16 k-rows x 64 n

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra, rb;
    auto fetch = [&](int k0) {
        ra = load4_guard(A, m0 + a_row, K, k0 + a_k, M, K, a_vec);
        if (TRANS_B) rb = load4_guard(B, n0 + bt_row, K, k0 + bt_k, N, K, b_vec);
        else rb = load4_guard(B, k0 + bn_k, N, n0 + bn_col, K, N, b_vec);
    };
    auto stash = [&]() {
        As[a_k + 0][a_row] = ra.x; As[a_k + 1][a_row] = ra.y; As[a_k + 2][a_row] = ra.z; As[a_k + 3][a_row] = ra.w;
        if (TRANS_B) {
            Bs[bt_k + 0][bt_row] = rb.x; Bs[bt_k + 1][bt_row] = rb.y; Bs[bt_k + 2][bt_row] = rb.z; Bs[bt_k + 3][bt_row] = rb.w;
        } else {
            *reinterpret_cast<float4*>(&Bs[bn_k][bn_col]) = rb;
        }
    };

    fetch(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        __syncthreads();
        stash();
        __syncthreads();
        if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) C[m * N + n] = apply_epilogue(acc[i][j], m, n, N, epi);
        }
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

extern "C" int asme_b200_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_b,
                              const asme_gemm_epilogue* epi, asme_stream_t stream) {
    ASME_REQUIRE(A && B && C, "gemm: null argument");
    ASME_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    if (M == 0) return ASME_OK;
    EpilogueArgs e = {nullptr, nullptr, ASME_ACT_NONE, nullptr, 0.f, 1.f, 0, 0, nullptr};
    if (epi) {
        ASME_REQUIRE(epi->p_drop >= 0.f && epi->p_drop < 1.f, "gemm: dropout p=%f out of range", epi->p_drop);
        e.bias = epi->bias; e.pre_act = epi->pre_act; e.act = epi->act; e.mul_gelu_grad_of = epi->mul_gelu_grad_of;
        e.p_drop = epi->p_drop; e.inv_keep = epi->p_drop > 0.f ? 1.0f / (1.0f - epi->p_drop) : 1.0f;
        e.seed = epi->seed; e.site = epi->site; e.residual = epi->residual;
    }
    const bool a_vec = (K % 4 == 0) && aligned16(A);
    const bool b_vec = trans_b ? ((K % 4 == 0) && aligned16(B)) : ((N % 4 == 0) && aligned16(B));
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
    if (trans_b) gemm_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, M, N, K, e, a_vec, b_vec, epi ? epi->m_live : nullptr);
    else gemm_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, M, N, K, e, a_vec, b_vec, epi ? epi->m_live : nullptr);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dW[N,K] = dY[M,N]^T X[M,K], split over M (grid.z) into ws[splits][N][K], then a
// deterministic reduction over splits.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, int M, int N,
                                                    int K, int rows_per_split, float* __restrict__ partial, bool y_vec,
                                                    bool x_vec, const int32_t* __restrict__ m_live) {
    __shared__ __align__(16) float Ys[BK][BM + PAD];
    __shared__ __align__(16) float Xs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int n0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
    M = asme_live_rows(M, m_live);
    const int m_begin = blockIdx.z * rows_per_split;
    const int m_end = min(M, m_begin + rows_per_split);
    const int l_row = tid / 16, l_col = (tid % 16) * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ry, rx;
    auto fetch = [&](int m0) {
        ry = load4_guard(dY, m0 + l_row, N, n0 + l_col, m_end, N, y_vec);
        rx = load4_guard(X, m0 + l_row, K, k0 + l_col, m_end, K, x_vec);
    };
    if (m_begin < m_end) fetch(m_begin);
    for (int m0 = m_begin; m0 < m_end; m0 += BK) {
        __syncthreads();
        *reinterpret_cast<float4*>(&Ys[l_row][l_col]) = ry;
        *reinterpret_cast<float4*>(&Xs[l_row][l_col]) = rx;
        __syncthreads();
        if (m0 + BK < m_end) fetch(m0 + BK);
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            const float4 a = *reinterpret_cast<const float4*>(&Ys[mm][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Xs[mm][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    float* out = partial + (size_t)blockIdx.z * N * K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) out[(size_t)n * K + k] = acc[i][j];
        }
    }
}

__global__ void split_reduce_kernel(const float* __restrict__ partial, int splits, long long n, float* __restrict__ out,
                                    int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < splits; ++p) s += partial[(size_t)p * n + i];
    out[i] = accumulate ? out[i] + s : s;
}

static int wgrad_splits(int M, int N, int K) {
    const int tiles = ceil_div(N, BM) * ceil_div(K, BN);
    int splits = ceil_div(2 * ASME_NUM_SMS, tiles);
    const int max_splits = ceil_div(M, 4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    return splits;
}

extern "C" size_t asme_b200_colsum_workspace_bytes(int M, int N);
extern "C" int asme_b200_colsum_accumulate(const float* x, int M, int N, float* out, void* ws, size_t ws_bytes,
                                           asme_stream_t stream);

extern "C" size_t asme_b200_gemm_wgrad_workspace_bytes(int M, int N, int K) {
    const size_t a = (size_t)wgrad_splits(M, N, K) * N * K * sizeof(float);
    const size_t b = asme_b200_colsum_workspace_bytes(M, N);
    return a > b ? a : b;
}

extern "C" int asme_b200_colsum_accumulate_live(const float* x, int M, int N, float* out, void* ws, size_t ws_bytes,
                                                const int32_t* m_live, asme_stream_t stream);
extern "C" int asme_b200_gemm_wgrad(const float* dY, const float* X, int M, int N, int K, float* dW, float* dbias,
                                    int accumulate, void* ws, size_t ws_bytes, const int32_t* m_live, asme_stream_t stream) {
    ASME_REQUIRE(dY && X && dW, "gemm_wgrad: null argument");
    ASME_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_wgrad: bad shape M=%d N=%d K=%d", M, N, K);
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        if (!accumulate) {
            ASME_CUDA_OK(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st));
            if (dbias) ASME_CUDA_OK(cudaMemsetAsync(dbias, 0, (size_t)N * sizeof(float), st));
        }
        return ASME_OK;
    }
    const int splits = wgrad_splits(M, N, K);
    if (ws_bytes < asme_b200_gemm_wgrad_workspace_bytes(M, N, K)) {
        asme_set_error("gemm_wgrad: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    int rows_per_split = ceil_div(M, splits);
    rows_per_split = ceil_div(rows_per_split, BK) * BK;
    const bool y_vec = (N % 4 == 0) && aligned16(dY);
    const bool x_vec = (K % 4 == 0) && aligned16(X);
    dim3 grid(ceil_div(N, BM), ceil_div(K, BN), splits);
    wgrad_kernel<<<grid, 256, 0, st>>>(dY, X, M, N, K, rows_per_split, (float*)ws, y_vec, x_vec, m_live);
    ASME_LAUNCH_OK();
    const long long n = (long long)N * K;
    split_reduce_kernel<<<ceil_div(n, 256), 256, 0, st>>>((const float*)ws, splits, n, dW, accumulate);
    ASME_LAUNCH_OK();
    if (dbias) {
        if (!accumulate) ASME_CUDA_OK(cudaMemsetAsync(dbias, 0, (size_t)N * sizeof(float), st));
        return asme_b200_colsum_accumulate_live(dY, M, N, dbias, ws, ws_bytes, m_live, stream);
    }
    return ASME_OK;
}
