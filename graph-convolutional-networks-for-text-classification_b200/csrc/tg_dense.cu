// tg_dense.cu — the skinny dense products and reductions around the hidden layer (sm_100a, fp32 CUDA cores).
//
//   tg_dense_nn_f32    S2 = H1 * W2          reference layer.py:102 (dense branch of th.spmm in layer 2)
//   tg_hidden_bwd_f32  dH1 = dS2 * W2^T, dZ1 = dH1 * [H1>0] * scale, dW2 = H1^T * dS2, db1 = colsum(dZ1)
//                      (autograd of layer.py:102,182,185,110 — SURVEY §2.2 rows B4..B7, one pass over H1)
//   tg_colsum_f32      db2 = colsum(dZ2)     
//   tg_reduce_sum_f32  loss = sum(row_loss)  (trainer.py:358-359 mean, fixed order)
//   tg_relu_dropout_bwd_f32  dZ = dH * [H>0] * scale   (threshold_backward + mask mul)
//
// All of them read an [N x H] operand once from HBM; the arithmetic (2*N*H*C flop per product) is done on the
// fp32 FMA pipes because the reference's parity budget (1e-5 relative) rules out TF32 tensor-core inputs.
// Every cross-block reduction is two-stage with a fixed order: no float atomics.
#include "tg_common.cuh"

namespace tg {

// ------------------------------------------------------------------------------------------------------------
// C[n x c] = A[n x h] * W[h x c]     (c <= 32 per pass, wider c loops column blocks on the host side)
// Thread t owns RT rows {t, t + 256, ...} of the block and all CP = 4*NC4 columns: RT*CP accumulators in registers.
// Each thread streams its own rows with 128-bit loads (8 k-steps per 128-byte line, the line is consumed while it
// is L1 resident), the next k-block's loads are issued before the current block's FMAs (register double buffer);
// W (<= 32 KB) is staged once per CTA in shared memory and read at a warp-uniform address (broadcast).
// ------------------------------------------------------------------------------------------------------------
constexpr int kNnThreads = 256;
constexpr int kNnRT = 2;
constexpr int kNnRows = kNnThreads * kNnRT;

template <int NC4, bool VEC_A>
__global__ void __launch_bounds__(kNnThreads) dense_nn_kernel(const float* __restrict__ A, int64_t lda,
                                                              const float* __restrict__ W, int64_t ldw,
                                                              float* __restrict__ C, int64_t ldc, int64_t n,
                                                              int h, int c) {
    constexpr int CP = NC4 * 4;
    extern __shared__ __align__(16) float Ws[];  // [h4][CP], h4 = h rounded up to 4, zero padded
    const int t = threadIdx.x;
    const int h4 = (h + 3) & ~3;
    for (int idx = t; idx < h4 * CP; idx += kNnThreads) {
        const int kk = idx / CP, j = idx % CP;
        Ws[idx] = (kk < h && j < c) ? __ldg(W + (int64_t)kk * ldw + j) : 0.f;
    }
    __syncthreads();
    const int64_t row0 = (int64_t)blockIdx.x * kNnRows + t;
    const float* ap[kNnRT];
    bool live[kNnRT];
#pragma unroll
    for (int r = 0; r < kNnRT; ++r) {
        const int64_t row = row0 + (int64_t)r * kNnThreads;
        live[r] = row < n;
        ap[r] = A + (live[r] ? row : 0) * lda;
    }
    float acc[kNnRT][CP];
#pragma unroll
    for (int r = 0; r < kNnRT; ++r)
#pragma unroll
        for (int j = 0; j < CP; ++j) acc[r][j] = 0.f;

    auto load4 = [&](int r, int k) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!live[r]) return v;
        if (VEC_A && k + 3 < h) return __ldg(reinterpret_cast<const float4*>(ap[r] + k));
        if (k + 0 < h) v.x = __ldg(ap[r] + k + 0);
        if (k + 1 < h) v.y = __ldg(ap[r] + k + 1);
        if (k + 2 < h) v.z = __ldg(ap[r] + k + 2);
        if (k + 3 < h) v.w = __ldg(ap[r] + k + 3);
        return v;
    };
    constexpr int KB = 2;  // float4 k-steps per register buffer
    float4 cur[kNnRT][KB], nxt[kNnRT][KB];
#pragma unroll
    for (int r = 0; r < kNnRT; ++r)
#pragma unroll
        for (int u = 0; u < KB; ++u) cur[r][u] = load4(r, u * 4);
    for (int k0 = 0; k0 < h4; k0 += 4 * KB) {
#pragma unroll
        for (int r = 0; r < kNnRT; ++r)
#pragma unroll
            for (int u = 0; u < KB; ++u) nxt[r][u] = load4(r, k0 + 4 * KB + u * 4);
#pragma unroll
        for (int u = 0; u < KB; ++u) {
            const int kb = k0 + u * 4;
            if (kb < h4) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int q = 0; q < NC4; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(Ws + (kb + kk) * CP + q * 4);
#pragma unroll
                        for (int r = 0; r < kNnRT; ++r) {
                            const float av = kk == 0 ? cur[r][u].x : kk == 1 ? cur[r][u].y : kk == 2 ? cur[r][u].z : cur[r][u].w;
                            acc[r][q * 4 + 0] = fmaf(av, w.x, acc[r][q * 4 + 0]);
                            acc[r][q * 4 + 1] = fmaf(av, w.y, acc[r][q * 4 + 1]);
                            acc[r][q * 4 + 2] = fmaf(av, w.z, acc[r][q * 4 + 2]);
                            acc[r][q * 4 + 3] = fmaf(av, w.w, acc[r][q * 4 + 3]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kNnRT; ++r)
#pragma unroll
            for (int u = 0; u < KB; ++u) cur[r][u] = nxt[r][u];
    }
    const bool c_vec = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) && (c % 4 == 0);
#pragma unroll
    for (int r = 0; r < kNnRT; ++r) {
        if (!live[r]) continue;
        float* dst = C + (row0 + (int64_t)r * kNnThreads) * ldc;
        if (c_vec) {
#pragma unroll
            for (int q = 0; q < NC4; ++q)
                if (q * 4 < c) st_f4(dst + q * 4, make_float4(acc[r][q * 4], acc[r][q * 4 + 1], acc[r][q * 4 + 2], acc[r][q * 4 + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < CP; ++j)
                if (j < c) dst[j] = acc[r][j];
        }
    }
}

template <int NC4>
static int launch_dense_nn(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                           int64_t n, int h, int c, cudaStream_t st) {
    const int h4 = (h + 3) & ~3;
    const size_t smem = (size_t)h4 * NC4 * 4 * sizeof(float);
    TG_REQUIRE(smem <= 200 * 1024, TG_ERR_UNSUPPORTED, "inner dimension %d too large for the skinny product", h);
    const bool vec_a = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
    const unsigned grid = (unsigned)ceil_div64(n, kNnRows);
    if (vec_a) {
        TG_CUDA(cudaFuncSetAttribute(dense_nn_kernel<NC4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_nn_kernel<NC4, true><<<grid, kNnThreads, smem, st>>>(A, lda, W, ldw, C, ldc, n, h, c);
    } else {
        TG_CUDA(cudaFuncSetAttribute(dense_nn_kernel<NC4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_nn_kernel<NC4, false><<<grid, kNnThreads, smem, st>>>(A, lda, W, ldw, C, ldc, n, h, c);
    }
    TG_LAUNCH_CHECK();
    return TG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Fused hidden-layer backward.  One thread per hidden unit j: W2[j,:] and the dW2[j,:] accumulators live in
// registers; rows are streamed, the dS2 row is read from a small shared tile at a warp-uniform address.
// ------------------------------------------------------------------------------------------------------------
constexpr int kHbTile = 64;        // rows per staged dS2 tile
constexpr int kHbBatch = 8;        // H1 rows in flight per thread
constexpr int kHbMaxGrid = 3 * kNumSM;  // 80 registers x 256 threads: three CTAs per SM

template <int NC4, int TB>
__global__ void __launch_bounds__(TB, (TB == 256 ? 3 : 1)) hidden_bwd_kernel(const float* __restrict__ H1, int64_t ldh,
                                                          const float* __restrict__ dS2, int64_t ldd,
                                                          const float* __restrict__ W2, int64_t ldw, float scale,
                                                          float* __restrict__ dZ1, int64_t ldz,
                                                          float* __restrict__ partials, int64_t n, int h, int c,
                                                          int64_t rows_per_block) {
    constexpr int CP = NC4 * 4;
    __shared__ __align__(16) float ds[kHbTile * CP];
    const int j = threadIdx.x;
    const bool live = j < h;
    float w[CP], gw[CP];
#pragma unroll
    for (int q = 0; q < CP; ++q) {
        w[q] = (live && q < c) ? __ldg(W2 + (int64_t)j * ldw + q) : 0.f;
        gw[q] = 0.f;
    }
    float gb = 0.f;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(n, r_begin + rows_per_block);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += kHbTile) {
        const int tile = (int)min((int64_t)kHbTile, r_end - r0);
        __syncthreads();
        for (int idx = j; idx < kHbTile * CP; idx += blockDim.x) {
            const int rr = idx / CP, q = idx % CP;
            ds[idx] = (rr < tile && q < c) ? __ldg(dS2 + (r0 + rr) * ldd + q) : 0.f;
        }
        __syncthreads();
        if (!live) continue;
        // running pointers (one 64-bit add per row) instead of re-deriving row*ld+j for every access
        const float* hp = H1 + r0 * ldh + j;
        float* zp = dZ1 + r0 * ldz + j;
        for (int rb = 0; rb < tile; rb += kHbBatch) {
            float a[kHbBatch];
#pragma unroll
            for (int u = 0; u < kHbBatch; ++u) a[u] = (rb + u < tile) ? __ldg(hp + (int64_t)u * ldh) : 0.f;
            const float* dsr = ds + rb * CP;
            // all kHbBatch rows are computed unconditionally (rows past the tile read zeros), so the compiler can
            // interleave their independent FMA chains; dh is kept as four partial sums to shorten the dependent chain
#pragma unroll
            for (int u = 0; u < kHbBatch; ++u) {
                float dh0 = 0.f, dh1 = 0.f, dh2 = 0.f, dh3 = 0.f;
#pragma unroll
                for (int q4 = 0; q4 < NC4; ++q4) {
                    const float4 d = *reinterpret_cast<const float4*>(dsr + u * CP + q4 * 4);
                    dh0 = fmaf(d.x, w[q4 * 4 + 0], dh0);
                    dh1 = fmaf(d.y, w[q4 * 4 + 1], dh1);
                    dh2 = fmaf(d.z, w[q4 * 4 + 2], dh2);
                    dh3 = fmaf(d.w, w[q4 * 4 + 3], dh3);
                    gw[q4 * 4 + 0] = fmaf(a[u], d.x, gw[q4 * 4 + 0]);
                    gw[q4 * 4 + 1] = fmaf(a[u], d.y, gw[q4 * 4 + 1]);
                    gw[q4 * 4 + 2] = fmaf(a[u], d.z, gw[q4 * 4 + 2]);
                    gw[q4 * 4 + 3] = fmaf(a[u], d.w, gw[q4 * 4 + 3]);
                }
                const float dz = (a[u] > 0.f) ? ((dh0 + dh1) + (dh2 + dh3)) * scale : 0.f;
                if (rb + u < tile) zp[(int64_t)u * ldz] = dz;
                gb += dz;
            }
            hp += (int64_t)kHbBatch * ldh;
            zp += (int64_t)kHbBatch * ldz;
        }
    }
    if (live) {
        float* out = partials + (int64_t)blockIdx.x * h * (c + 1);
#pragma unroll
        for (int q = 0; q < CP; ++q)
            if (q < c) out[(int64_t)j * c + q] = gw[q];
        out[(int64_t)h * c + j] = gb;
    }
}

// out[e] = sum_{b < n_blocks} partials[b][e]  in block order
__global__ void sum_partials_kernel(const float* __restrict__ partials, int64_t n_blocks, int64_t n_elem,
                                    float* __restrict__ out_a, int64_t split, float* __restrict__ out_b) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    float s = 0.f;
    for (int64_t b = 0; b < n_blocks; ++b) s += partials[b * n_elem + e];
    if (e < split) out_a[e] = s;
    else out_b[e - split] = s;
}

template <int NC4>
static int launch_hidden_bwd(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                             float scale, float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n,
                             int h, int c, cudaStream_t st) {
    int64_t grid = ceil_div64(n, 2 * kHbTile);
    if (grid > kHbMaxGrid) grid = kHbMaxGrid;
    if (grid < 1) grid = 1;
    int64_t rpb = ceil_div64(n, grid);
    rpb = ceil_div64(rpb, kHbTile) * kHbTile;
    grid = ceil_div64(n > 0 ? n : 1, rpb);
    const int threads = ((h + 31) / 32) * 32;
    // the register budget follows the block size: W2[j,:] and dW2[j,:] (2*CP floats) stay in registers for h <= 512
    if (threads <= 256)
        hidden_bwd_kernel<NC4, 256><<<(unsigned)grid, threads, 0, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz,
                                                                        partials, n, h, c, rpb);
    else if (threads <= 512)
        hidden_bwd_kernel<NC4, 512><<<(unsigned)grid, threads, 0, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz,
                                                                        partials, n, h, c, rpb);
    else
        hidden_bwd_kernel<NC4, 1024><<<(unsigned)grid, threads, 0, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz,
                                                                         partials, n, h, c, rpb);
    TG_LAUNCH_CHECK();
    const int64_t n_elem = (int64_t)h * (c + 1);
    sum_partials_kernel<<<(unsigned)ceil_div64(n_elem, 256), 256, 0, st>>>(partials, grid, n_elem, dW2, (int64_t)h * c, db1);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// column sums / scalar sum: per-block partial in fixed thread order, then block-ordered final pass
// ------------------------------------------------------------------------------------------------------------
constexpr int kCsMaxGrid = 4 * kNumSM;

__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, int64_t ldx, int64_t n, int c,
                                                             int cw, int64_t rows_per_block, float* __restrict__ partials) {
    __shared__ float sh[256];
    const int t = threadIdx.x;
    const int col = t % cw, rl = t / cw, rstep = 256 / cw;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(n, r_begin + rows_per_block);
    for (int c0 = 0; c0 < c; c0 += cw) {
        float s = 0.f;
        if (c0 + col < c) {
            // 8 independent loads in flight, added in a fixed order
            int64_t r = r_begin + rl;
            for (; r + 7 * (int64_t)rstep < r_end; r += 8 * (int64_t)rstep) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(X + (r + (int64_t)u * rstep) * ldx + c0 + col);
#pragma unroll
                for (int u = 0; u < 8; ++u) s += v[u];
            }
            for (; r < r_end; r += rstep) s += __ldg(X + r * ldx + c0 + col);
        }
        sh[t] = s;
        __syncthreads();
        if (rl == 0 && c0 + col < c) {
            float tot = 0.f;
            for (int k = 0; k < rstep; ++k) tot += sh[k * cw + col];
            partials[(int64_t)blockIdx.x * c + c0 + col] = tot;
        }
        __syncthreads();
    }
}

static int colsum_grid(int64_t n, int64_t* rows_per_block) {
    int64_t grid = ceil_div64(n > 0 ? n : 1, 512);
    if (grid > kCsMaxGrid) grid = kCsMaxGrid;
    const int64_t rpb = ceil_div64(n > 0 ? n : 1, grid);
    *rows_per_block = rpb;
    return (int)ceil_div64(n > 0 ? n : 1, rpb);
}

__global__ void relu_dropout_bwd_kernel(const float* __restrict__ H, int64_t ldh, const float* __restrict__ dH,
                                        int64_t lddh, float scale, float* __restrict__ dZ, int64_t ldz, int64_t n,
                                        int f) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * (int64_t)f) return;
    const int64_t r = idx / f;
    const int j = (int)(idx % f);
    const float hv = H[r * ldh + j];
    dZ[r * ldz + j] = hv > 0.f ? dH[r * lddh + j] * scale : 0.f;
}

}  // namespace tg

extern "C" {

int tg_dense_nn_f32(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t n,
                    int32_t h, int32_t c, void* stream) {
    using namespace tg;
    TG_REQUIRE(A && W && C, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n >= 0 && h > 0 && c > 0 && lda >= h && ldw >= c && ldc >= c, TG_ERR_INVALID_ARG, "bad shape");
    if (n == 0) return TG_OK;
    cudaStream_t st = as_stream(stream);
    for (int c0 = 0; c0 < c; c0 += 32) {
        const int cb = (c - c0 < 32) ? (c - c0) : 32;
        const int nc4 = (cb + 3) / 4;
        int rc;
        switch (nc4) {
            case 1: rc = launch_dense_nn<1>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            case 2: rc = launch_dense_nn<2>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            case 3: rc = launch_dense_nn<3>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            case 4: rc = launch_dense_nn<4>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            case 5: rc = launch_dense_nn<5>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            case 6: rc = launch_dense_nn<6>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            case 7: rc = launch_dense_nn<7>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
            default: rc = launch_dense_nn<8>(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st); break;
        }
        if (rc != TG_OK) return rc;
    }
    return TG_OK;
}

int64_t tg_hidden_bwd_scratch_floats(int64_t n, int32_t h, int32_t c) {
    (void)n;
    return (int64_t)tg::kHbMaxGrid * h * (c + 1);
}

int tg_hidden_bwd_f32(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                      float scale, float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n,
                      int32_t h, int32_t c, void* stream) {
    using namespace tg;
    TG_REQUIRE(H1 && dS2 && W2 && dZ1 && dW2 && db1 && partials, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n >= 0 && h > 0 && c > 0 && ldh >= h && ldd >= c && ldw >= c && ldz >= h, TG_ERR_INVALID_ARG, "bad shape");
    TG_REQUIRE(h <= 1024, TG_ERR_UNSUPPORTED, "hidden width %d > 1024 not supported by the fused backward", h);
    TG_REQUIRE(c <= 32, TG_ERR_UNSUPPORTED, "n_class %d > 32 not supported by the fused backward", c);
    cudaStream_t st = as_stream(stream);
    const int nc4 = (c + 3) / 4;
    switch (nc4) {
        case 1: return launch_hidden_bwd<1>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        case 2: return launch_hidden_bwd<2>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        case 3: return launch_hidden_bwd<3>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        case 4: return launch_hidden_bwd<4>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        case 5: return launch_hidden_bwd<5>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        case 6: return launch_hidden_bwd<6>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        case 7: return launch_hidden_bwd<7>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
        default: return launch_hidden_bwd<8>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, st);
    }
}

int64_t tg_colsum_scratch_floats(int64_t n, int32_t c) {
    (void)n;
    return (int64_t)tg::kCsMaxGrid * c;
}

int tg_colsum_f32(const float* X, int64_t ldx, int64_t n, int32_t c, float* scratch, float* out, void* stream) {
    using namespace tg;
    TG_REQUIRE(X && scratch && out, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n >= 0 && c > 0 && ldx >= c, TG_ERR_INVALID_ARG, "bad shape");
    cudaStream_t st = as_stream(stream);
    int cw = 1;
    while (cw < c && cw < 256) cw <<= 1;
    int64_t rpb = 0;
    const int grid = colsum_grid(n, &rpb);
    colsum_partial_kernel<<<grid, 256, 0, st>>>(X, ldx, n, c, cw, rpb, scratch);
    TG_LAUNCH_CHECK();
    sum_partials_kernel<<<(unsigned)ceil_div64(c, 256), 256, 0, st>>>(scratch, grid, c, out, c, nullptr);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

int64_t tg_reduce_scratch_floats(int64_t n) {
    (void)n;
    return tg::kCsMaxGrid;
}

int tg_reduce_sum_f32(const float* x, int64_t n, float* scratch, float* out, void* stream) {
    using namespace tg;
    TG_REQUIRE(x && scratch && out, TG_ERR_INVALID_ARG, "null pointer");
    cudaStream_t st = as_stream(stream);
    // a length-n vector is an [n x 1] matrix: per-block sums in fixed thread order, then block order
    int64_t rpb = 0;
    const int grid = colsum_grid(n, &rpb);
    colsum_partial_kernel<<<grid, 256, 0, st>>>(x, 1, n, 1, 1, rpb, scratch);
    TG_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 256, 0, st>>>(scratch, grid, 1, out, 1, nullptr);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

int tg_relu_dropout_bwd_f32(const float* H, int64_t ldh, const float* dH, int64_t lddh, float scale, float* dZ,
                            int64_t ldz, int64_t n, int32_t f, void* stream) {
    using namespace tg;
    TG_REQUIRE(H && dH && dZ, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n >= 0 && f > 0 && ldh >= f && lddh >= f && ldz >= f, TG_ERR_INVALID_ARG, "bad shape");
    const int64_t tot = n * (int64_t)f;
    if (tot == 0) return TG_OK;
    relu_dropout_bwd_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, as_stream(stream)>>>(H, ldh, dH, lddh, scale, dZ,
                                                                                          ldz, n, f);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

}  // extern "C"
