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
#include <math.h>
#include <stdlib.h>

#include <type_traits>

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

// ------------------------------------------------------------------------------------------------------------
// Tensor-core variant of the skinny product:  C[n x c] = A[n x h] * W[h x c]  with 3xTF32 split accumulation.
// ncu shows the FFMA version is not bandwidth bound (DRAM 29 % of peak, L1/issue bound, profiles/r01_dense_a_full.md),
// so per the north star the product moves to the tensor cores.  fp32 parity needs more than TF32's 10-bit mantissa:
// every operand is split  x = hi + lo  (hi = tf32(x), lo = tf32(x - hi)) and  hi*hi + hi*lo + lo*hi  is accumulated in
// fp32 (relative error ~2^-21).  mma.sync.m16n8k8 (legacy HMMA path) is enough to make this product HBM-bound; the
// operand A is streamed from global memory straight into fragment layout: one k-step of a row is one 32-byte sector.
//   warp tile: 32 rows (two m16 tiles) x 24 columns (three n8 tiles, columns >= c are zero), k in steps of 8
//   W hi/lo are staged once per CTA in shared memory as [h8][24] (row stride 24 floats: conflict-free fragment reads)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// hi / lo split in four integer / float instructions (cvt.rna.tf32.f32 compiles to a nine-instruction sequence and this kernel
// is issue bound): hi = x rounded to 10 mantissa bits (add half an ulp to the bit pattern, clear the low 13 bits — ties
// away from zero like cvt.rna; an operand at the very top of the fp32 range would carry into the exponent, H1 / dS2 never
// are), lo = x - hi exactly, truncated to tf32 (|error| < 2^-21 |x|, the size of the dropped lo*lo term).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kMmaWarps = 8;
constexpr int kMmaRowsPerWarp = 32;
constexpr int kMmaNP = 24;   // padded column count (three n8 tiles)
constexpr int kMmaWS = 26;   // row stride of the staged W (floats): 4*26 = 8 (mod 32) makes the fragment reads conflict free

// A is read with one 128-bit load per row per 16 k: lane (g, t) holds A[row][k16 + 4t .. 4t+3].  The k index inside an
// MMA is only a summation index, so the two k8 steps of a 16-block use the permutation  slot t -> k = 4t + 2s,
// slot t+4 -> k = 4t + 2s + 1  (s = 0, 1) on BOTH operands: the 16 bytes a lane loaded are exactly its two fragments.
__global__ void __launch_bounds__(kMmaWarps * 32) dense_nn_mma_kernel(const float* __restrict__ A, int64_t lda,
                                                                    const float* __restrict__ W, int64_t ldw,
                                                                    float* __restrict__ C, int64_t ldc, int64_t n, int h,
                                                                    int c) {
    extern __shared__ __align__(16) float wsm[];  // Whi [h16][26] | Wlo [h16][26]
    const int h16 = (h + 15) & ~15;
    float* Whi = wsm;
    float* Wlo = wsm + (size_t)h16 * kMmaWS;
    for (int idx = threadIdx.x; idx < h16 * kMmaNP; idx += blockDim.x) {
        const int k = idx / kMmaNP, j = idx % kMmaNP;
        const float w = (k < h && j < c) ? __ldg(W + (int64_t)k * ldw + j) : 0.f;
        const float hi = __uint_as_float(to_tf32(w));
        Whi[k * kMmaWS + j] = hi;
        Wlo[k * kMmaWS + j] = __uint_as_float(to_tf32(w - hi));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // persistent CTAs: W hi/lo are split and staged once per CTA, then the CTA walks its row blocks
    for (int64_t blk = blockIdx.x; blk * (kMmaWarps * kMmaRowsPerWarp) < n; blk += gridDim.x) {
    const int64_t row0 = (blk * kMmaWarps + warp) * kMmaRowsPerWarp;
    if (row0 >= n) continue;
    // rows of the two m16 tiles this lane touches: row0 + {g, g+8, g+16, g+24}
    const float* ap[4];
    bool ok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = row0 + g + 8 * r;
        ok[r] = row < n;
        ap[r] = A + (ok[r] ? row : 0) * lda + 4 * t;
    }
    float acc[2][3][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;

    auto load_a = [&](int k16, float4 (&v)[4]) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int k = k16 + 4 * t;
            if (ok[r] && k + 3 < h) {
                v[r] = __ldg(reinterpret_cast<const float4*>(ap[r] + k16));
            } else {
                v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[r]) {
                    if (k + 0 < h) v[r].x = __ldg(ap[r] + k16 + 0);
                    if (k + 1 < h) v[r].y = __ldg(ap[r] + k16 + 1);
                    if (k + 2 < h) v[r].z = __ldg(ap[r] + k16 + 2);
                }
            }
        }
    };
    float4 cur[4], nxt[4];
    load_a(0, cur);
    for (int k16 = 0; k16 < h16; k16 += 16) {
        load_a(k16 + 16, nxt);  // next block travels while this one is multiplied (past h it loads zeros)
#pragma unroll
        for (int sstep = 0; sstep < 2; ++sstep) {
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                // fragment order a0:(g, slot t) a1:(g+8, slot t) a2:(g, slot t+4) a3:(g+8, slot t+4)
                const float4 lo_row = cur[2 * m], hi_row = cur[2 * m + 1];
                const float x[4] = {sstep ? lo_row.z : lo_row.x, sstep ? hi_row.z : hi_row.x,
                                    sstep ? lo_row.w : lo_row.y, sstep ? hi_row.w : hi_row.y};
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(x[e], ahi[m][e], alo[m][e]);   // four integer / float instructions per element
            }
            const int kb = k16 + 4 * t + 2 * sstep;  // actual k of slot t; slot t+4 is kb + 1
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint32_t bhi[2], blo[2];
                bhi[0] = __float_as_uint(Whi[kb * kMmaWS + 8 * j + g]);
                bhi[1] = __float_as_uint(Whi[(kb + 1) * kMmaWS + 8 * j + g]);
                blo[0] = __float_as_uint(Wlo[kb * kMmaWS + 8 * j + g]);
                blo[1] = __float_as_uint(Wlo[(kb + 1) * kMmaWS + 8 * j + g]);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma_tf32(acc[m][j], alo[m], bhi);  // small terms first
                    mma_tf32(acc[m][j], ahi[m], blo);
                    mma_tf32(acc[m][j], ahi[m], bhi);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) cur[r] = nxt[r];
    }
    // c0:(g, 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int64_t row = row0 + 16 * m + g + 8 * hh;
            if (row >= n) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int col = 8 * j + 2 * t;
                if (col < c) C[row * ldc + col] = acc[m][j][2 * hh];
                if (col + 1 < c) C[row * ldc + col + 1] = acc[m][j][2 * hh + 1];
            }
        }
    }  // row blocks of this CTA
}

static size_t dense_mma_smem(int h) { return (size_t)2 * ((h + 15) & ~15) * kMmaWS * sizeof(float); }

static int launch_dense_nn_mma(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t n,
                               int h, int c, cudaStream_t st) {
    const size_t smem = dense_mma_smem(h);
    TG_CUDA(cudaFuncSetAttribute(dense_nn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = ceil_div64(n, kMmaWarps * kMmaRowsPerWarp);
    if (grid > 2 * kNumSM) grid = 2 * kNumSM;  // two CTAs per SM (registers), each staging W once
    dense_nn_mma_kernel<<<(unsigned)grid, kMmaWarps * 32, smem, st>>>(A, lda, W, ldw, C, ldc, n, h, c);
    TG_LAUNCH_CHECK();
    return TG_OK;
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
constexpr int kHbTile = 32;        // rows per staged tile (one bit per row in the per-thread activity mask)
constexpr int kHbStages = 3;       // cp.async pipeline depth (tiles in flight: kHbStages - 1)
constexpr int kHbMaxGrid = 4 * kNumSM;  // partial rows of the workspace (sparse kernel: 2 CTAs per SM, dense kernel: up to 4)

__device__ __forceinline__ void hb_cp16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hb_cp4(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

// H1 and dS2 tiles travel through a cp.async pipeline in shared memory.  The FMAs exploit the sparsity of H1: after
// ReLU and dropout (p = 0.5) about three quarters of its entries are exactly zero, and a zero entry contributes nothing
// to dW2 and has dZ1 = 0.  Every thread (one hidden unit j) therefore builds a 32-bit mask of the rows of the tile
// where H1[row, j] > 0 and walks only those rows; the loop runs for the warp's maximum population count (~13 of 32), the
// finished lanes ride along with a = 0.  dZ1 is written in place over the staged H1 tile (entries that were zero are
// already the correct dZ1 = 0) and the tile goes back to HBM with coalesced 128-bit stores.
// History: the dense variants (register-loaded, register double-buffered, cp.async staged, two units per thread) all
// ran in 0.97 ms at C3 — 320 M three-register FFMAs at ~0.3 per cycle per scheduler, an issue / register-port bound
// (profiles/r01_hidden_a_full.md) — so the only way down was to issue fewer of them.
template <int NC4, int TB>
__global__ void __launch_bounds__(TB, (TB == 256 ? 2 : 1)) hidden_bwd_kernel(const float* __restrict__ H1, int64_t ldh,
                                                          const float* __restrict__ dS2, int64_t ldd,
                                                          const float* __restrict__ W2, int64_t ldw, float scale,
                                                          float* __restrict__ dZ1, int64_t ldz,
                                                          float* __restrict__ partials, int64_t n, int h, int c,
                                                          int64_t rows_per_block, int vec_ok, int out_vec, int64_t n_count) {
    // rows >= n_count get their dZ1 but do not count into dW2 / db1 (replicated rows of a document-sharded graph)
    constexpr int CP = NC4 * 4;
    extern __shared__ __align__(16) float hb_smem[];
    const int hp4 = (h + 3) & ~3;                     // row stride of the staged H1 tile (floats)
    float* As = hb_smem;                              // [kHbStages][kHbTile][hp4]
    float* Ds = hb_smem + (size_t)kHbStages * kHbTile * hp4;  // [kHbStages][kHbTile][CP]
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
    const int n_tiles = (int)((r_end - r_begin + kHbTile - 1) / kHbTile);

    auto issue = [&](int t) {
        if (t < n_tiles) {
            const int st = t % kHbStages;
            const int64_t r0 = r_begin + (int64_t)t * kHbTile;
            float* as = As + (size_t)st * kHbTile * hp4;
            float* dsm = Ds + (size_t)st * kHbTile * CP;
            if (vec_ok) {
                const int per_row = hp4 / 4;
                for (int idx = j; idx < kHbTile * per_row; idx += blockDim.x) {
                    const int rr = idx / per_row, q = idx % per_row;
                    const bool ok = r0 + rr < r_end && q * 4 < h;
                    hb_cp16(as + rr * hp4 + q * 4, ok ? H1 + (r0 + rr) * ldh + q * 4 : H1, ok);
                }
            } else {
                for (int idx = j; idx < kHbTile * hp4; idx += blockDim.x) {
                    const int rr = idx / hp4, q = idx % hp4;
                    const bool ok = r0 + rr < r_end && q < h;
                    hb_cp4(as + rr * hp4 + q, ok ? H1 + (r0 + rr) * ldh + q : H1, ok);
                }
            }
            for (int idx = j; idx < kHbTile * CP; idx += blockDim.x) {
                const int rr = idx / CP, q = idx % CP;
                const bool ok = r0 + rr < r_end && q < c;
                hb_cp4(dsm + idx, ok ? dS2 + (r0 + rr) * ldd + q : dS2, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // (an empty group keeps the wait arithmetic uniform)
    };

#pragma unroll
    for (int t = 0; t < kHbStages - 1; ++t) issue(t);
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + kHbStages - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(kHbStages - 1) : "memory");
        __syncthreads();
        const int st = t % kHbStages;
        const int64_t r0 = r_begin + (int64_t)t * kHbTile;
        const int tile = (int)min((int64_t)kHbTile, r_end - r0);
        const float* dsm = Ds + (size_t)st * kHbTile * CP;
        float* acol = As + (size_t)st * kHbTile * hp4 + j;  // column j of the staged tile, row stride hp4
        unsigned mask = 0;
        if (live) {
#pragma unroll
            for (int u = 0; u < kHbTile; ++u) mask |= (acol[u * hp4] > 0.f ? 1u : 0u) << u;
        }
        const int iters = __reduce_max_sync(0xffffffffu, __popc(mask));
        for (int it = 0; it < iters; ++it) {
            const bool act = mask != 0;
            const int u = act ? (__ffs(mask) - 1) : 0;
            mask &= mask - 1;  // (0 stays 0)
            const float a = act ? acol[u * hp4] : 0.f;
            const bool counted = r0 + u < n_count;
            const float* dr = dsm + u * CP;  // lane-private row of the dS2 tile
            // packed fp32 FMAs (FFMA2): the products of two adjacent classes per instruction — half the issue slots of the
            // arithmetic that bounds this kernel; every lane of a pair is an IEEE fma, the sums keep their order
            float2 dh01 = make_float2(0.f, 0.f), dh23 = make_float2(0.f, 0.f);
            const float ac = counted ? a : 0.f;
            const float2 aa = make_float2(ac, ac);
#pragma unroll
            for (int q4 = 0; q4 < NC4; ++q4) {
                const float4 d = *reinterpret_cast<const float4*>(dr + q4 * 4);
                dh01 = __ffma2_rn(make_float2(d.x, d.y), make_float2(w[q4 * 4 + 0], w[q4 * 4 + 1]), dh01);
                dh23 = __ffma2_rn(make_float2(d.z, d.w), make_float2(w[q4 * 4 + 2], w[q4 * 4 + 3]), dh23);
                const float2 g01 = __ffma2_rn(aa, make_float2(d.x, d.y), make_float2(gw[q4 * 4 + 0], gw[q4 * 4 + 1]));
                const float2 g23 = __ffma2_rn(aa, make_float2(d.z, d.w), make_float2(gw[q4 * 4 + 2], gw[q4 * 4 + 3]));
                gw[q4 * 4 + 0] = g01.x; gw[q4 * 4 + 1] = g01.y;
                gw[q4 * 4 + 2] = g23.x; gw[q4 * 4 + 3] = g23.y;
            }
            const float dh0 = dh01.x, dh1 = dh01.y, dh2 = dh23.x, dh3 = dh23.y;
            if (act) {
                const float dz = ((dh0 + dh1) + (dh2 + dh3)) * scale;
                acol[u * hp4] = dz;  // in place: the tile turns into dZ1
                if (counted) gb += dz;
            }
        }
        __syncthreads();
        // coalesced write-back of the tile (rows past the range were zero filled and are skipped)
        {
            float* tile_s = As + (size_t)st * kHbTile * hp4;
            if (out_vec) {
                const int per_row = hp4 / 4;
                for (int idx = j; idx < tile * per_row; idx += blockDim.x) {
                    const int rr = idx / per_row, q = idx % per_row;
                    if (q * 4 < h)
                        *reinterpret_cast<float4*>(dZ1 + (r0 + rr) * ldz + q * 4) = *reinterpret_cast<const float4*>(tile_s + rr * hp4 + q * 4);
                }
            } else {
                for (int idx = j; idx < tile * hp4; idx += blockDim.x) {
                    const int rr = idx / hp4, q = idx % hp4;
                    if (q < h) dZ1[(r0 + rr) * ldz + q] = tile_s[rr * hp4 + q];
                }
            }
        }
        __syncthreads();  // the stage is refilled kHbStages - 1 iterations later
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
// Dense variant with packed FMAs (FFMA2).  The sparse-walk kernel above is bound by shared memory, not by arithmetic: its
// lanes walk DIFFERENT rows, so every step reads 32 lane-private rows of the dS2 tile (20 wavefronts), and the warp runs for
// the maximum population count of its lanes (13.5 of 32 rows at 25 % density).  Here every lane of a warp works on the SAME
// row: the dS2 row is five broadcast LDS.128 (5 wavefronts), H1 is read straight from global memory (one coalesced 128-byte
// line per warp and row, no staging, no write-back pass), dZ1 is written the same way, and all 40 multiply-adds of a
// (row, unit) pair are issued as 20 FFMA2 — zeros of H1 cost arithmetic again, but the FMA pipe has room for it now that
// one instruction carries two of them.  Same summation order per element as the sparse kernel (zeros add exactly 0).
constexpr int kHdTile = 64;   // rows of dS2 per staged tile
constexpr int kHdBatch = 16;  // rows whose H1 values a thread holds in flight
template <int NC4, bool MULTI>
__global__ void __launch_bounds__(256, 2) hidden_bwd_dense_kernel(const float* __restrict__ H1, int64_t ldh,
                                                                  const float* __restrict__ dS2, int64_t ldd,
                                                                  const float* __restrict__ W2, int64_t ldw, float scale,
                                                                  float* __restrict__ dZ1, int64_t ldz,
                                                                  float* __restrict__ partials, int64_t n, int h, int c,
                                                                  int64_t rows_per_block, int acc_in_arg, int final_pass_arg, int64_t n_count) {
    // MULTI = false is the single-pass kernel of the benchmark path: the two pass flags are compile-time constants there
    const int acc_in = MULTI ? acc_in_arg : 0, final_pass = MULTI ? final_pass_arg : 1;
    // Class counts above 32 run this kernel once per block of <= 32 classes: a pass that is not the last one stores the raw
    // sum dS2[:, block] * W2[:, block]^T (added to the previous passes' sum when acc_in is set) in dZ1; the last pass adds
    // its own block, then applies the ReLU / dropout mask and the scale.  acc_in = 0, final_pass = 1: the single-pass case.
    constexpr int CP = NC4 * 4;
    __shared__ __align__(16) float Ds[2][kHdTile][CP];
    const int j = threadIdx.x;
    const bool live = j < h;
    const int jc = live ? j : 0;
    float2 w2[CP / 2], gw2[CP / 2];
#pragma unroll
    for (int q = 0; q < CP / 2; ++q) {
        w2[q].x = (live && 2 * q < c) ? __ldg(W2 + (int64_t)j * ldw + 2 * q) : 0.f;
        w2[q].y = (live && 2 * q + 1 < c) ? __ldg(W2 + (int64_t)j * ldw + 2 * q + 1) : 0.f;
        gw2[q] = make_float2(0.f, 0.f);
    }
    float gb = 0.f;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(n, r_begin + rows_per_block);
    const int n_tiles = (int)((r_end - r_begin + kHdTile - 1) / kHdTile);
    auto issue = [&](int t) {
        if (t < n_tiles) {
            const int64_t r0 = r_begin + (int64_t)t * kHdTile;
            float* dsm = &Ds[t & 1][0][0];
            for (int idx = j; idx < kHdTile * CP; idx += blockDim.x) {
                const int rr = idx / CP, q = idx % CP;
                const bool ok = r0 + rr < r_end && q < c;
                hb_cp4(dsm + idx, ok ? dS2 + (r0 + rr) * ldd + q : dS2, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    // row pointers advance by the leading dimensions (no 64-bit multiply per row); tiles that lie wholly inside the block's
    // range (all but the last) take a path without per-row bounds checks
    const float* hp = H1 + r_begin * ldh + jc;   // next row whose H1 value is still to be loaded
    float* zp = dZ1 + r_begin * ldz + jc;        // next row to be written
    int64_t hr = r_begin;                        // row index hp points at
    float hv[kHdBatch];
#pragma unroll
    for (int u = 0; u < kHdBatch; ++u) {
        hv[u] = (hr < r_end) ? __ldg(hp) : 0.f;
        hp += ldh;
        ++hr;
    }
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int64_t r0 = r_begin + (int64_t)t * kHdTile;
        const bool full = r0 + kHdTile + kHdBatch <= r_end;   // this tile and the look-ahead batch are in range
#pragma unroll 1
        for (int b = 0; b < kHdTile; b += kHdBatch) {
            float hn[kHdBatch];
            if (full) {
#pragma unroll
                for (int u = 0; u < kHdBatch; ++u) {  // the next batch of H1 values: in flight while this batch is computed
                    hn[u] = __ldg(hp);
                    hp += ldh;
                }
                hr += kHdBatch;
            } else {
#pragma unroll
                for (int u = 0; u < kHdBatch; ++u) {
                    hn[u] = (hr < r_end) ? __ldg(hp) : 0.f;
                    hp += ldh;
                    ++hr;
                }
            }
            const float4* dr = reinterpret_cast<const float4*>(&Ds[t & 1][b][0]);  // warp-uniform: broadcast reads
            // rows past n_count give their dZ1 but stay out of dW2 / db1.  They are the last rows of the matrix: the test is made
            // once per batch (block-uniform), and the batches that lie wholly below n_count — all but one or two of the whole
            // launch — run the loop without any per-row test (a per-row test cost 11 % on this issue-bound kernel).
            const int64_t cnt_rows = n_count - (r0 + b);
            auto batch = [&](auto all_counted_tag) {
                constexpr bool kAll = decltype(all_counted_tag)::value;
#pragma unroll
                for (int u = 0; u < kHdBatch; ++u) {
                    const float a = hv[u];
                    const bool counted = kAll || u < cnt_rows;
                    float2 dh01 = make_float2(0.f, 0.f), dh23 = make_float2(0.f, 0.f);
                    const float ac = counted ? a : 0.f;
                    const float2 aa = make_float2(ac, ac);
#pragma unroll
                    for (int q4 = 0; q4 < NC4; ++q4) {
                        const float4 d = dr[u * NC4 + q4];
                        dh01 = __ffma2_rn(make_float2(d.x, d.y), w2[2 * q4], dh01);
                        dh23 = __ffma2_rn(make_float2(d.z, d.w), w2[2 * q4 + 1], dh23);
                        gw2[2 * q4] = __ffma2_rn(aa, make_float2(d.x, d.y), gw2[2 * q4]);
                        gw2[2 * q4 + 1] = __ffma2_rn(aa, make_float2(d.z, d.w), gw2[2 * q4 + 1]);
                    }
                    const bool in_range = live && (full || r0 + b + u < r_end);
                    float dh = (dh01.x + dh01.y) + (dh23.x + dh23.y);
                    if (acc_in && in_range) dh += *zp;  // (the same thread wrote it in the previous pass)
                    const float dz = final_pass ? ((a > 0.f) ? dh * scale : 0.f) : dh;
                    if (counted) gb += dz;
                    if (in_range) *zp = dz;
                    zp += ldz;
                }
            };
            if (cnt_rows >= kHdBatch) batch(std::true_type{});
            else batch(std::false_type{});
#pragma unroll
            for (int u = 0; u < kHdBatch; ++u) hv[u] = hn[u];
        }
        __syncthreads();  // the tile buffer is refilled by the next iteration's issue
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
        float* out = partials + (int64_t)blockIdx.x * h * (c + 1);
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) {
            if (2 * q < c) out[(int64_t)j * c + 2 * q] = gw2[q].x;
            if (2 * q + 1 < c) out[(int64_t)j * c + 2 * q + 1] = gw2[q].y;
        }
        out[(int64_t)h * c + j] = gb;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core variant of the fused hidden-layer backward (c <= 24 classes, h <= 256 hidden units, 16-byte aligned rows).
// The FFMA2 kernel above sits on the FMA pipe (10.2 G multiply-adds at C3, 0.66 ms); the two products of the backward
// are contractions, so they move to the tensor cores with the same 3xTF32 split as the forward product
// (hi*hi + hi*lo + lo*hi in fp32):
//   P1  dH1^T [j x rows] = W2 [j x c] * dS2^T [c x rows]      (M = hidden units, N = 8 rows, K = classes in steps of 8)
//   P2  dW2   [j x c]   += H1^T [j x rows] * dS2 [rows x c]   (M = hidden units, N = classes in tiles of 8, K = 8 rows)
// A warp owns 16 MT hidden units (MT = 1 or 2 m16 tiles) and walks the rows in groups of 8; the 16 / MT warps of a CTA share
// the rows, so the dS2 tile is staged (already split into hi / lo) once per CTA.  Index choices that make every H1 element travel
// exactly once and every dZ1 element leave as part of one 64- / 128-bit store per lane and row:
//   * m-slot g + 8 hh of m-tile mt is hidden unit j0 + 2 mt + hh with j0 = 16 MT warp + 2 MT g — a lane's 2 MT units are
//     adjacent in memory (the m index only names rows of W2 / dW2, any bijection works);
//   * k-slot t of P2 is row 2t of the group, slot t + 4 row 2t + 1 (k is a summation index) — P2's A fragment
//     {(g,t),(g+8,t),(g,t+4),(g+8,t+4)} is then the SAME set of elements as P1's C fragment {(g,2t),(g,2t+1),(g+8,2t),
//     (g+8,2t+1)}: the H1 values a lane loaded for the mask of its dH1 outputs are its operand of P2.
// W2 hi / lo fragments live in registers for the whole kernel; dW2 accumulates in the MMA accumulators for one 64-row tile
// and is then added to fp32 master registers with ordinary round-to-nearest adds (the tensor core's accumulation of a long
// chain is not round-to-nearest).  Per-CTA partial dW2 / db1 go through sum_partials_kernel like the other variants:
// deterministic.  Staged dS2 rows have a stride of 28 floats: both fragment patterns (row g, class t) and (row 2t, class g)
// are then bank-conflict free.  Measured mma.sync rate on this part: 917 m16n8k8 TF32 MMAs per microsecond and SM
// (tools/micro/mma_sync_rate.cu); the 36 M MMAs of the C3 shape are 0.27 ms of tensor pipe, the 2.1 GB 0.32 ms of HBM.
// ------------------------------------------------------------------------------------------------------------
constexpr int kHmTile = 128;     // rows per staged dS2 tile (two runs of 8 groups of 8 rows): one CTA barrier per tile
constexpr int kHmStride = 28;    // floats per staged dS2 row
constexpr int kHmDepth = 8;      // 8-row groups of H1 a warp keeps in flight (ring slot = group of the 64-row run)
// MT = m16 tiles (16 hidden units each) per warp: 1 -> 16 warps per CTA, 128 registers; 2 -> 8 warps, the dS2 fragments of a
// group feed twice the MMAs.  Row stride of a warp's private H1 ring: 16 MT units + pad, conflict-free fragment reads.
template <int MT> struct HmCfg {
    static constexpr int kThreads = 512 / MT;
    static constexpr int kHStride = MT == 1 ? 20 : 36;
    static constexpr size_t kSmem = (size_t)(4 * kHmTile * kHmStride + 2 * kHmTile * 24 + (kThreads / 32) * kHmDepth * 8 * kHStride) * sizeof(float);
};

__device__ __forceinline__ void mma_tf32_nv(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    // not volatile: a pure function of its operands, the scheduler may interleave independent chains
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int KS, int MT>   // KS: k8 steps over the classes = n8 tiles of dW2: ceil(c / 8)
__global__ void __launch_bounds__(HmCfg<MT>::kThreads, 1) hidden_bwd_mma_kernel(const float* __restrict__ H1, int64_t ldh,
                                                                                const float* __restrict__ dS2, int64_t ldd,
                                                                                const float* __restrict__ W2, int64_t ldw, float scale,
                                                                                float* __restrict__ dZ1, int64_t ldz,
                                                                                float* __restrict__ partials, int64_t n, int h, int c,
                                                                                int64_t rows_per_block, int64_t n_count) {
    constexpr int CPK = 8 * KS;                     // padded class count
    constexpr int kThreads = HmCfg<MT>::kThreads;
    constexpr int kHS = HmCfg<MT>::kHStride;
    constexpr int NU = 2 * MT;                      // hidden units per lane (adjacent in memory)
    extern __shared__ __align__(16) float hm_smem[];
    float (*Dhi)[kHmTile * kHmStride] = reinterpret_cast<float (*)[kHmTile * kHmStride]>(hm_smem);
    float (*Dlo)[kHmTile * kHmStride] = reinterpret_cast<float (*)[kHmTile * kHmStride]>(hm_smem + 2 * kHmTile * kHmStride);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int j0 = 16 * MT * warp + NU * g;         // m-slot g + 8 hh of m-tile mt is unit j0 + 2 mt + hh
    // the warp's private H1 ring: kHmDepth groups x 8 rows x 16 MT units.  A group is MT coalesced cp.async of 16 bytes per lane
    // (lane l: rows l / (4 MT) + (8 / MT) i, units 4 (l % (4 MT)) ..), read back as (row 2t, units NU g ..) after a __syncwarp
    float* hs_w = hm_smem + 4 * kHmTile * kHmStride + 2 * kHmTile * 24 + warp * (kHmDepth * 8 * kHS);
    constexpr int kCpl = 4 * MT;                    // 16-byte pieces per row of the warp's slice
    float* hs_cp = hs_w + (lane / kCpl) * kHS + 4 * (lane % kCpl);
    const float* hs = hs_w + (2 * t) * kHS + NU * g;
    float* draw = hm_smem + 4 * kHmTile * kHmStride;   // raw dS2 tiles [2][kHmTile * CPK]
    const bool jok = j0 < h;                                           // h is a multiple of 4: units are inside or outside in fours
    const bool cp_ok = 16 * MT * warp + 4 * (lane % kCpl) < h;         // the four units this lane copies

    // P1's A operand: W2 fragments  a0:(slot g, k t) a1:(slot g+8, k t) a2:(slot g, k t+4) a3:(slot g+8, k t+4)
    uint32_t whi[MT][KS][4], wlo[MT][KS][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + 2 * mt + (e & 1), kk = 8 * ks + t + 4 * (e >> 1);
                const float w = (jok && kk < c) ? __ldg(W2 + (int64_t)j * ldw + kk) : 0.f;
                whi[mt][ks][e] = to_tf32(w);
                wlo[mt][ks][e] = to_tf32(w - __uint_as_float(whi[mt][ks][e]));
            }
    float gwm[MT][KS][4];   // fp32 master copy of the lane's dW2 elements
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < KS; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) gwm[mt][nt][e] = 0.f;
    float gb[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) gb[u] = 0.f;

    // everything below is relative to the block's first row, in 32-bit arithmetic (the launcher checks that a block's rows
    // times the leading dimensions fit): 64-bit row bookkeeping cost enough registers to spill under the 128-register cap,
    // and the reloads showed up as 14 % long-scoreboard stalls
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int nrows = (int)(min(n, r_begin + rows_per_block) - r_begin);                       // rows of this block
    const int ncnt = (int)max((int64_t)0, min(n_count - r_begin, (int64_t)nrows));             // ... that count into dW2 / db1
    const int n_tiles = (nrows + kHmTile - 1) / kHmTile;
    const float* __restrict__ H1b = H1 + r_begin * ldh;
    const float* __restrict__ dSb = dS2 + r_begin * ldd;
    float* __restrict__ dZb = dZ1 + r_begin * ldz;
    const int ldh_i = (int)ldh, ldz_i = (int)ldz, ldd_i = (int)ldd;

    // dS2 staging.  Tile 0: global -> registers -> hi / lo in shared memory.  Later tiles travel as raw fp32 with cp.async,
    // issued two tiles ahead inside the commit group of an H1 group (a register prefetch ended up next to its consumer and
    // exposed a full DRAM latency per tile); every thread splits the elements it copied itself at the end of the tile
    // before their use, so the hand-over needs no barrier of its own.  Thread -> (class q = tid % 32, rows tid / 32 + kSRows i).
    constexpr int kSRows = kThreads / 32;            // rows one pass of the CTA covers
    constexpr int kSPass = kHmTile / kSRows;         // passes per tile
    const int sq = threadIdx.x & 31, sr = threadIdx.x >> 5;
    const bool s_act = sq < CPK;
    auto stage_tile0 = [&]() {
        if (s_act) {
#pragma unroll
            for (int i = 0; i < kSPass; ++i) {
                const int rr = sr + kSRows * i;
                const float v = (rr < nrows && sq < c) ? __ldg(dSb + rr * ldd_i + sq) : 0.f;
                uint32_t hi, lo;
                split_tf32(v, hi, lo);
                Dhi[0][rr * kHmStride + sq] = __uint_as_float(hi);
                Dlo[0][rr * kHmStride + sq] = __uint_as_float(lo);
            }
        }
    };
    auto stage_issue = [&](int tile) {   // raw tile -> draw[tile & 1]  (zero-filled past the range / the class count)
        if (s_act) {
            float* dst = draw + (tile & 1) * (kHmTile * CPK) + sr * CPK + sq;
            const int row0 = tile * kHmTile + sr;
#pragma unroll
            for (int i = 0; i < kSPass; ++i) {
                const bool ok = row0 + kSRows * i < nrows && sq < c;
                hb_cp4(dst + kSRows * i * CPK, ok ? dSb + (row0 + kSRows * i) * ldd_i + sq : dSb, ok);
            }
        }
    };
    auto stage_split = [&](int tile) {   // draw[tile & 1] -> hi / lo [tile & 1]
        if (s_act) {
            const float* src = draw + (tile & 1) * (kHmTile * CPK) + sr * CPK + sq;
            float* ohi = Dhi[tile & 1] + sr * kHmStride + sq;
            float* olo = Dlo[tile & 1] + sr * kHmStride + sq;
#pragma unroll
            for (int i = 0; i < kSPass; ++i) {
                uint32_t hi, lo;
                split_tf32(src[kSRows * i * CPK], hi, lo);
                ohi[kSRows * i * kHmStride] = __uint_as_float(hi);
                olo[kSRows * i * kHmStride] = __uint_as_float(lo);
            }
        }
    };
    stage_tile0();
    stage_issue(1);
    asm volatile("cp.async.commit_group;" ::: "memory");

    // H1 ring: the next kHmDepth groups travel global -> shared memory asynchronously (a register ring left the distance
    // between a load and its use to the instruction scheduler, which under the 128-register cap moved the loads next to
    // their consumers: ncu showed 40 % of the samples on long-scoreboard stalls)
    constexpr int kCpRows = 32 / kCpl;               // rows one cp.async instruction of the warp covers
    int hoff = (lane / kCpl) * ldh_i + 16 * MT * warp + 4 * (lane % kCpl);   // this lane's first piece of the next group to load
    int hrow = lane / kCpl;
    int zoff = (2 * t) * ldz_i + j0;                                          // row 2t of the next group to store
    auto load_group = [&](int slot, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        if (cp_ok) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
                hb_cp16(hs_cp + slot * (8 * kHS) + i * kCpRows * kHS, H1b + hoff + i * kCpRows * ldh_i, kFull || hrow + i * kCpRows < nrows);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        hoff += 8 * ldh_i;
        hrow += 8;
    };
#pragma unroll
    for (int d = 0; d < kHmDepth; ++d) load_group(d, std::false_type{});
    __syncthreads();

    for (int tile = 0; tile < n_tiles; ++tile) {
        const int buf = tile & 1;
        stage_issue(tile + 2);   // joins the commit group of this tile's first H1 refill: landed by the next tile's first wait
        const int r0 = tile * kHmTile;
        float gw[MT][KS][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < KS; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) gw[mt][nt][e] = 0.f;
        const float* dhi_t = Dhi[buf];
        const float* dlo_t = Dlo[buf];
        auto run_tile = [&](auto full_tag) {
            constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll 1
            for (int run = 0; run < kHmTile / 64; ++run) {
            const float* dhi = dhi_t + run * (64 * kHmStride);
            const float* dlo = dlo_t + run * (64 * kHmStride);
#pragma unroll
            for (int grp = 0; grp < kHmDepth; ++grp) {
                asm volatile("cp.async.wait_group %0;" ::"n"(kHmDepth - 1) : "memory");   // this group has landed
                __syncwarp();
                float xa[NU], xb[NU];   // rows 2t and 2t + 1 of the group, the lane's NU units
#pragma unroll
                for (int u = 0; u < NU; ++u) xa[u] = xb[u] = 0.f;
                if (jok) {
                    if (MT == 1) {
                        const float2 a2 = *reinterpret_cast<const float2*>(hs + grp * (8 * kHS));
                        const float2 b2 = *reinterpret_cast<const float2*>(hs + grp * (8 * kHS) + kHS);
                        xa[0] = a2.x; xa[1] = a2.y; xb[0] = b2.x; xb[1] = b2.y;
                    } else {
                        const float4 a4 = *reinterpret_cast<const float4*>(hs + grp * (8 * kHS));
                        const float4 b4 = *reinterpret_cast<const float4*>(hs + grp * (8 * kHS) + kHS);
                        xa[0] = a4.x; xa[1] = a4.y; xa[NU - 2] = a4.z; xa[NU - 1] = a4.w;
                        xb[0] = b4.x; xb[1] = b4.y; xb[NU - 2] = b4.z; xb[NU - 1] = b4.w;
                    }
                }
                __syncwarp();                // every lane has read the slot ...
                load_group(grp, full_tag);   // ... before it is refilled with the group kHmDepth ahead
                // ---- P1: pre-activation gradient of the lane's NU units x 2 rows
                float dh[MT][4], dhs[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) dh[mt][e] = dhs[mt][e] = 0.f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int o = (8 * grp + g) * kHmStride + 8 * ks + t;
                    const uint32_t bhi[2] = {__float_as_uint(dhi[o]), __float_as_uint(dhi[o + 4])};
                    const uint32_t blo[2] = {__float_as_uint(dlo[o]), __float_as_uint(dlo[o + 4])};
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32_nv(dhs[mt], wlo[mt][ks], bhi);   // small terms in their own chain
                        mma_tf32_nv(dhs[mt], whi[mt][ks], blo);
                        mma_tf32_nv(dh[mt], whi[mt][ks], bhi);
                    }
                }
                // c0:(unit j0 + 2 mt, row 2t) c1:(same unit, row 2t+1) c2:(unit j0 + 2 mt + 1, row 2t) c3:(.., row 2t+1)
                float za[NU], zb[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int mt = u >> 1, hh = u & 1;
                    za[u] = (xa[u] > 0.f) ? (dh[mt][2 * hh] + dhs[mt][2 * hh]) * scale : 0.f;
                    zb[u] = (xb[u] > 0.f) ? (dh[mt][2 * hh + 1] + dhs[mt][2 * hh + 1]) * scale : 0.f;
                }
                float* zp = dZb + zoff;
                bool cnt_a = true, cnt_b = true, st_a = jok, st_b = jok;
                if (!kFull) {
                    const int row0 = r0 + 64 * run + 8 * grp + 2 * t;
                    st_a = jok && row0 < nrows;
                    st_b = jok && row0 + 1 < nrows;
                    cnt_a = row0 < ncnt;
                    cnt_b = row0 + 1 < ncnt;
                }
                if (MT == 1) {
                    if (st_a) *reinterpret_cast<float2*>(zp) = make_float2(za[0], za[1]);
                    if (st_b) *reinterpret_cast<float2*>(zp + ldz_i) = make_float2(zb[0], zb[1]);
                } else {
                    if (st_a) *reinterpret_cast<float4*>(zp) = make_float4(za[0], za[1], za[NU - 2], za[NU - 1]);
                    if (st_b) *reinterpret_cast<float4*>(zp + ldz_i) = make_float4(zb[0], zb[1], zb[NU - 2], zb[NU - 1]);
                }
                zoff += 8 * ldz_i;
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    if (kFull) {
                        gb[u] += za[u] + zb[u];
                    } else {
                        gb[u] += (cnt_a ? za[u] : 0.f) + (cnt_b ? zb[u] : 0.f);
                        if (!cnt_a) xa[u] = 0.f;
                        if (!cnt_b) xb[u] = 0.f;
                    }
                }
                // ---- P2: dW2 += H1^T dS2 over the group's 8 rows;  A fragment order (j, 2t) (j + 1, 2t) (j, 2t+1) (j + 1, 2t+1)
                uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    split_tf32(xa[2 * mt], ahi[mt][0], alo[mt][0]);
                    split_tf32(xa[2 * mt + 1], ahi[mt][1], alo[mt][1]);
                    split_tf32(xb[2 * mt], ahi[mt][2], alo[mt][2]);
                    split_tf32(xb[2 * mt + 1], ahi[mt][3], alo[mt][3]);
                }
#pragma unroll
                for (int nt = 0; nt < KS; ++nt) {
                    const int o = (8 * grp + 2 * t) * kHmStride + 8 * nt + g;
                    const uint32_t bhi[2] = {__float_as_uint(dhi[o]), __float_as_uint(dhi[o + kHmStride])};
                    const uint32_t blo[2] = {__float_as_uint(dlo[o]), __float_as_uint(dlo[o + kHmStride])};
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32_nv(gw[mt][nt], alo[mt], bhi);
                        mma_tf32_nv(gw[mt][nt], ahi[mt], blo);
                        mma_tf32_nv(gw[mt][nt], ahi[mt], bhi);
                    }
                }
            }
            }
        };
        // every row of this tile and of the look-ahead groups is inside the block's range and counted: no per-row tests
        if (r0 + kHmTile + 8 * kHmDepth <= nrows && r0 + kHmTile <= ncnt) run_tile(std::true_type{});
        else run_tile(std::false_type{});
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < KS; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) gwm[mt][nt][e] += gw[mt][nt][e];
        if (tile + 1 < n_tiles) stage_split(tile + 1);   // (the last readers of that hi / lo buffer passed the previous barrier)
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // dW2 fragment  c0:(unit j0 + 2 mt, class 8nt+2t) c1:(same unit, 8nt+2t+1) c2:(unit j0 + 2 mt + 1, 8nt+2t) c3:(.., 8nt+2t+1)
    float* out = partials + (int64_t)blockIdx.x * h * (c + 1);
    if (jok) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < KS; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = j0 + 2 * mt + (e >> 1), q = 8 * nt + 2 * t + (e & 1);
                    if (q < c) out[(int64_t)j * c + q] = gwm[mt][nt][e];
                }
    }
    // db1: the four lanes of a quad hold different rows of the same units
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        gb[u] += __shfl_xor_sync(0xffffffffu, gb[u], 1);
        gb[u] += __shfl_xor_sync(0xffffffffu, gb[u], 2);
    }
    if (jok && t == 0) {
#pragma unroll
        for (int u = 0; u < NU; ++u) out[(int64_t)h * c + j0 + u] = gb[u];
    }
}

// Fixed-order sum of per-block partial rows.  Eight lanes per element take the blocks b = l, l + 8, ... (each in
// ascending order, four loads in flight) and their eight sums meet in a fixed shuffle tree: deterministic, and the
// latency of a serial walk over hundreds of L2-resident rows no longer sits on the step's critical path.
__global__ void sum_partials_kernel(const float* __restrict__ partials, int64_t n_blocks, int64_t n_elem,
                                    float* __restrict__ out_a, int64_t split, float* __restrict__ out_b, int a_cols, int64_t a_ld) {
    // elements [0, split) form a matrix with a_cols columns that is written with leading dimension a_ld (a column block of
    // dW2); elements [split, n_elem) go to out_b (db1) unless it is null
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t >> 3;
    const int l = (int)(t & 7);
    float s = 0.f;
    if (e < n_elem) {
        int64_t b = l;
        for (; b + 24 < n_blocks; b += 32) {
            const float v0 = partials[b * n_elem + e], v1 = partials[(b + 8) * n_elem + e];
            const float v2 = partials[(b + 16) * n_elem + e], v3 = partials[(b + 24) * n_elem + e];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; b < n_blocks; b += 8) s += partials[b * n_elem + e];
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (e >= n_elem || l != 0) return;
    if (e < split) out_a[(e / a_cols) * a_ld + (e % a_cols)] = s;
    else if (out_b) out_b[e - split] = s;
}

static bool dense_hidden_enabled() {
    static const bool on = [] {
        const char* v = getenv("TG_HIDDEN_DENSE");
        return !(v && *v) || atoi(v) != 0;
    }();
    return on;
}
static bool mma_hidden_enabled() {   // TG_HIDDEN_MMA=0 selects the CUDA-core kernels (read once)
    static const bool on = [] {
        const char* v = getenv("TG_HIDDEN_MMA");
        return !(v && *v) || atoi(v) != 0;
    }();
    return on;
}

// tensor-core kernel: one CTA per SM, every CTA a contiguous range of rows (a multiple of the 128-row tile)
template <int KS, int MT>
static int launch_hidden_bwd_mma_t(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw, float scale,
                                   float* dZ1, int64_t ldz, float* partials, int64_t n, int h, int c, int64_t rpb, int64_t n_count,
                                   unsigned grid, cudaStream_t st) {
    TG_CUDA(cudaFuncSetAttribute(hidden_bwd_mma_kernel<KS, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HmCfg<MT>::kSmem));
    hidden_bwd_mma_kernel<KS, MT><<<grid, HmCfg<MT>::kThreads, HmCfg<MT>::kSmem, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h,
                                                                                         c, rpb, n_count);
    return TG_OK;
}
// m16 tiles per warp.  Two (8 warps x 32 units, 255 registers): the dS2 fragments of a group feed twice the MMAs and the
// per-group bookkeeping is paid once per 32 units — 0.47 ms against 0.54 ms at C3; one (16 warps x 16 units) keeps more warps
// busy when the hidden layer is narrow.  TG_HIDDEN_MMA_MT = 1 / 2 forces a variant (read once).
static int hidden_mma_tiles(int h) {
    static const int forced = [] {
        const char* v = getenv("TG_HIDDEN_MMA_MT");
        const int x = (v && *v) ? atoi(v) : 0;
        return (x == 1 || x == 2) ? x : 0;
    }();
    if (forced) return forced;
    return h > 128 ? 2 : 1;
}
static int launch_hidden_bwd_mma(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                                 float scale, float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n,
                                 int h, int c, int64_t n_count, cudaStream_t st) {
    int64_t grid = kNumSM;
    int64_t rpb = ceil_div64(n > 0 ? n : 1, grid);
    rpb = ceil_div64(rpb, kHmTile) * kHmTile;
    grid = ceil_div64(n > 0 ? n : 1, rpb);
    const int ks = (c + 7) / 8;
    int rc;
    if (hidden_mma_tiles(h) == 2) {
        rc = ks == 1 ? launch_hidden_bwd_mma_t<1, 2>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, n_count, (unsigned)grid, st)
           : ks == 2 ? launch_hidden_bwd_mma_t<2, 2>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, n_count, (unsigned)grid, st)
                     : launch_hidden_bwd_mma_t<3, 2>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, n_count, (unsigned)grid, st);
    } else {
        rc = ks == 1 ? launch_hidden_bwd_mma_t<1, 1>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, n_count, (unsigned)grid, st)
           : ks == 2 ? launch_hidden_bwd_mma_t<2, 1>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, n_count, (unsigned)grid, st)
                     : launch_hidden_bwd_mma_t<3, 1>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, n_count, (unsigned)grid, st);
    }
    if (rc != TG_OK) return rc;
    TG_LAUNCH_CHECK();
    const int64_t n_elem = (int64_t)h * (c + 1);
    sum_partials_kernel<<<(unsigned)ceil_div64(n_elem * 8, 256), 256, 0, st>>>(partials, grid, n_elem, dW2, (int64_t)h * c, db1, c, c);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

template <int NC4>
static int launch_hidden_bwd(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                             float scale, float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n,
                             int h, int c, int64_t n_count, cudaStream_t st) {
    if (h <= 256 && dense_hidden_enabled()) {
        // dense FFMA2 kernel: three CTAs of 256 threads per SM
        int64_t grid = 2 * kNumSM;
        int64_t rpb = ceil_div64(n > 0 ? n : 1, grid);
        rpb = ceil_div64(rpb, kHdTile) * kHdTile;
        grid = ceil_div64(n > 0 ? n : 1, rpb);
        const int threads = ((h + 31) / 32) * 32;
        hidden_bwd_dense_kernel<NC4, false><<<(unsigned)grid, threads, 0, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, h, c, rpb, 0, 1, n_count);
        TG_LAUNCH_CHECK();
        const int64_t n_elem = (int64_t)h * (c + 1);
        sum_partials_kernel<<<(unsigned)ceil_div64(n_elem * 8, 256), 256, 0, st>>>(partials, grid, n_elem, dW2, (int64_t)h * c, db1, c, c);
        TG_LAUNCH_CHECK();
        return TG_OK;
    }
    int64_t grid = ceil_div64(n, 2 * kHbTile);
    if (grid > 2 * kNumSM) grid = 2 * kNumSM;  // the sparse-walk kernel stages 96 KB of tiles per CTA: two CTAs per SM
    if (grid < 1) grid = 1;
    int64_t rpb = ceil_div64(n, grid);
    rpb = ceil_div64(rpb, kHbTile) * kHbTile;
    grid = ceil_div64(n > 0 ? n : 1, rpb);
    const int threads = ((h + 31) / 32) * 32;
    const int hp4 = (h + 3) & ~3;
    const size_t smem = (size_t)kHbStages * kHbTile * (hp4 + NC4 * 4) * sizeof(float);
    const int vec_ok = (ldh % 4 == 0) && ((reinterpret_cast<uintptr_t>(H1) & 15u) == 0);
    const int out_vec = (ldz % 4 == 0) && (h % 4 == 0) && ((reinterpret_cast<uintptr_t>(dZ1) & 15u) == 0);
    // the register budget follows the block size: W2[j,:] and dW2[j,:] (2*CP floats) stay in registers for h <= 512
    if (threads <= 256) {
        TG_CUDA(cudaFuncSetAttribute(hidden_bwd_kernel<NC4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hidden_bwd_kernel<NC4, 256><<<(unsigned)grid, threads, smem, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz,
                                                                           partials, n, h, c, rpb, vec_ok, out_vec, n_count);
    } else if (threads <= 512) {
        TG_CUDA(cudaFuncSetAttribute(hidden_bwd_kernel<NC4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hidden_bwd_kernel<NC4, 512><<<(unsigned)grid, threads, smem, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz,
                                                                           partials, n, h, c, rpb, vec_ok, out_vec, n_count);
    } else {
        TG_CUDA(cudaFuncSetAttribute(hidden_bwd_kernel<NC4, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hidden_bwd_kernel<NC4, 1024><<<(unsigned)grid, threads, smem, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz,
                                                                            partials, n, h, c, rpb, vec_ok, out_vec, n_count);
    }
    TG_LAUNCH_CHECK();
    const int64_t n_elem = (int64_t)h * (c + 1);
    sum_partials_kernel<<<(unsigned)ceil_div64(n_elem * 8, 256), 256, 0, st>>>(partials, grid, n_elem, dW2, (int64_t)h * c, db1, c, c);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

// Any class count / hidden width: blocks of <= 256 hidden units x blocks of <= 32 classes on the dense kernel.  Within a
// hidden block the class blocks run in order with the running pre-activation sum kept in dZ1 (see the kernel); dW2 is
// produced block by block, db1 by the last class block.  (R52: 52 classes -> two passes; reference data/text_dataset/R52.txt.)
template <int NC4>
static int launch_hidden_block(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                               float scale, float* dZ1, int64_t ldz, float* dW2, int64_t ldg, float* db1, float* partials,
                               int64_t n, int hb, int cb, int acc_in, int final_pass, int64_t n_count, cudaStream_t st) {
    int64_t grid = 2 * kNumSM;
    int64_t rpb = ceil_div64(n > 0 ? n : 1, grid);
    rpb = ceil_div64(rpb, kHdTile) * kHdTile;
    grid = ceil_div64(n > 0 ? n : 1, rpb);
    const int threads = ((hb + 31) / 32) * 32;
    hidden_bwd_dense_kernel<NC4, true><<<(unsigned)grid, threads, 0, st>>>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, partials, n, hb, cb, rpb,
                                                                     acc_in, final_pass, n_count);
    TG_LAUNCH_CHECK();
    const int64_t n_elem = (int64_t)hb * (cb + 1);
    sum_partials_kernel<<<(unsigned)ceil_div64(n_elem * 8, 256), 256, 0, st>>>(partials, grid, n_elem, dW2, (int64_t)hb * cb,
                                                                               final_pass ? db1 : nullptr, cb, ldg);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

static int hidden_bwd_blocked(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw, float scale,
                              float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n, int h, int c,
                              int64_t n_count, cudaStream_t st) {
    for (int j0 = 0; j0 < h; j0 += 256) {
        const int hb = (h - j0 < 256) ? (h - j0) : 256;
        for (int c0 = 0; c0 < c; c0 += 32) {
            const int cb = (c - c0 < 32) ? (c - c0) : 32;
            const int acc_in = c0 > 0, final_pass = (c0 + 32 >= c);
            const float *h1 = H1 + j0, *ds = dS2 + c0, *w2 = W2 + (int64_t)j0 * ldw + c0;
            float *dz = dZ1 + j0, *dw = dW2 + (int64_t)j0 * c + c0, *db = db1 + j0;
            int rc;
            switch ((cb + 3) / 4) {
                case 1: rc = launch_hidden_block<1>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                case 2: rc = launch_hidden_block<2>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                case 3: rc = launch_hidden_block<3>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                case 4: rc = launch_hidden_block<4>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                case 5: rc = launch_hidden_block<5>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                case 6: rc = launch_hidden_block<6>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                case 7: rc = launch_hidden_block<7>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
                default: rc = launch_hidden_block<8>(h1, ldh, ds, ldd, w2, ldw, scale, dz, ldz, dw, c, db, partials, n, hb, cb, acc_in, final_pass, n_count, st); break;
            }
            if (rc != TG_OK) return rc;
        }
    }
    return TG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// plain fp32 GEMM for a DENSE layer-1 feature matrix (reference layer.py:102 when infeatn is dense; its autograd
// transpose product).  Not the reference's mode (its features are sparse) and not on the benchmark path: a
// straightforward shared-memory tiled kernel (64 x 64 tile, 4 x 4 per thread, fp32 FMA), deterministic.
//   NN: C[M x N] = A[M x K] * B[K x N]
//   TN: C[M x N] = A[K x M]^T * B[K x N], K split into `splits` ranges whose partial products are summed in range order
// ------------------------------------------------------------------------------------------------------------
constexpr int kGT = 64, kGK = 16;
template <bool TA>
__global__ void __launch_bounds__(256) gemm_tile_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                        int64_t ldb, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                        int64_t K, int64_t k_per_split, int64_t c_split_stride) {
    __shared__ float As[kGK][kGT + 1], Bs[kGK][kGT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * kGT, n0 = (int64_t)blockIdx.x * kGT;
    const int64_t kb = (int64_t)blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
    float acc[4][4] = {};
    for (int64_t k0 = kb; k0 < ke; k0 += kGK) {
        for (int idx = threadIdx.x; idx < kGK * kGT; idx += 256) {
            int kk, mm;
            if (TA) { mm = idx % kGT; kk = idx / kGT; } else { kk = idx % kGK; mm = idx / kGK; }   // coalesced along the contiguous axis
            const int64_t k = k0 + kk, m = m0 + mm;
            As[kk][mm] = (k < ke && m < M) ? (TA ? A[k * lda + m] : A[m * lda + k]) : 0.f;
        }
        for (int idx = threadIdx.x; idx < kGK * kGT; idx += 256) {
            const int nn = idx % kGT, kk = idx / kGT;
            const int64_t k = k0 + kk, nn_g = n0 + nn;
            Bs[kk][nn] = (k < ke && nn_g < N) ? B[k * ldb + nn_g] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(a[i], b[jn], acc[i][jn]);
        }
        __syncthreads();
    }
    float* Cz = C + (int64_t)blockIdx.z * c_split_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
            const int64_t m = m0 + ty * 4 + i, nn = n0 + tx * 4 + jn;
            if (m < M && nn < N) Cz[m * ldc + nn] = acc[i][jn];
        }
}

// C[e] = sum over splits (in order) of partials[s][e]
__global__ void gemm_sum_splits_kernel(const float* __restrict__ partials, int splits, int64_t M, int64_t N, float* __restrict__ C, int64_t ldc) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * N) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partials[(int64_t)z * M * N + e];
    C[(e / N) * ldc + (e % N)] = s;
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
    {
        // tensor-core path (3xTF32) for column blocks of up to 24; TG_DENSE_MMA=0 selects the FFMA kernel
        static const bool mma_ok = !(getenv("TG_DENSE_MMA") && atoi(getenv("TG_DENSE_MMA")) == 0);
        const bool a_vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
        if (mma_ok && a_vec && dense_mma_smem(h) <= 200 * 1024) {
            for (int c0 = 0; c0 < c; c0 += kMmaNP) {
                const int cb = (c - c0 < kMmaNP) ? (c - c0) : kMmaNP;
                const int rc = launch_dense_nn_mma(A, lda, W + c0, ldw, C + c0, ldc, n, h, cb, st);
                if (rc != TG_OK) return rc;
            }
            return TG_OK;
        }
    }
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

int tg_hidden_bwd_rows_f32(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                           float scale, float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n,
                           int32_t h, int32_t c, int64_t n_count, void* stream) {
    using namespace tg;
    TG_REQUIRE(H1 && dS2 && W2 && dZ1 && dW2 && db1 && partials, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n >= 0 && h > 0 && c > 0 && ldh >= h && ldd >= c && ldw >= c && ldz >= h, TG_ERR_INVALID_ARG, "bad shape");
    TG_REQUIRE(n_count >= 0 && n_count <= n, TG_ERR_INVALID_ARG, "n_count must be in [0, n]");
    cudaStream_t st = as_stream(stream);
    if (c > 32 || h > 1024) return hidden_bwd_blocked(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
    {
        const bool aligned = (ldh % 4 == 0) && (ldz % 2 == 0) && (h % 4 == 0) &&
                             ((reinterpret_cast<uintptr_t>(H1) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(dZ1) & 7u) == 0);
        const int64_t blk_rows = ceil_div64(n > 0 ? n : 1, kNumSM) + kHmTile;   // (32-bit offsets inside a block of rows)
        const bool fits = blk_rows * ldh < (int64_t)1 << 30 && blk_rows * ldz < (int64_t)1 << 30 && blk_rows * ldd < (int64_t)1 << 30;
        if (c <= 24 && h <= 256 && aligned && fits && mma_hidden_enabled())
            return launch_hidden_bwd_mma(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
    }
    const int nc4 = (c + 3) / 4;
    switch (nc4) {
        case 1: return launch_hidden_bwd<1>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        case 2: return launch_hidden_bwd<2>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        case 3: return launch_hidden_bwd<3>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        case 4: return launch_hidden_bwd<4>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        case 5: return launch_hidden_bwd<5>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        case 6: return launch_hidden_bwd<6>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        case 7: return launch_hidden_bwd<7>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
        default: return launch_hidden_bwd<8>(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n_count, st);
    }
}

int tg_hidden_bwd_f32(const float* H1, int64_t ldh, const float* dS2, int64_t ldd, const float* W2, int64_t ldw,
                      float scale, float* dZ1, int64_t ldz, float* dW2, float* db1, float* partials, int64_t n,
                      int32_t h, int32_t c, void* stream) {
    return tg_hidden_bwd_rows_f32(H1, ldh, dS2, ldd, W2, ldw, scale, dZ1, ldz, dW2, db1, partials, n, h, c, n, stream);
}

int64_t tg_gemm_scratch_floats(int32_t trans_a, int64_t m, int64_t n, int64_t k) {
    if (!trans_a) return 0;
    int64_t splits = tg::ceil_div64(k, 8192);
    if (splits > 512) splits = 512;
    return splits > 1 ? splits * m * n : 0;
}

int tg_gemm_f32(int32_t trans_a, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t m,
                int64_t n, int64_t k, float* scratch, void* stream) {
    using namespace tg;
    TG_REQUIRE(A && B && C, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(m >= 0 && n >= 0 && k >= 0 && ldb >= n && ldc >= n && lda >= (trans_a ? m : k), TG_ERR_INVALID_ARG, "bad shape");
    if (m == 0 || n == 0) return TG_OK;
    cudaStream_t st = as_stream(stream);
    const dim3 tiles((unsigned)ceil_div64(n, kGT), (unsigned)ceil_div64(m, kGT), 1);
    TG_REQUIRE(tiles.y <= 65535, TG_ERR_UNSUPPORTED, "too many row tiles for one launch (m = %lld)", (long long)m);
    if (!trans_a) {
        gemm_tile_kernel<false><<<tiles, 256, 0, st>>>(A, lda, B, ldb, C, ldc, m, n, k, k > 0 ? k : 1, 0);
        TG_LAUNCH_CHECK();
        return TG_OK;
    }
    // the reduction runs over the long axis (the nodes): split it, one partial product per range, summed in range order
    int64_t splits = ceil_div64(k, 8192);
    if (splits > 512) splits = 512;
    if (splits <= 1) {
        gemm_tile_kernel<true><<<tiles, 256, 0, st>>>(A, lda, B, ldb, C, ldc, m, n, k, k > 0 ? k : 1, 0);
        TG_LAUNCH_CHECK();
        return TG_OK;
    }
    TG_REQUIRE(scratch, TG_ERR_WORKSPACE, "tg_gemm_f32 (transposed) needs tg_gemm_scratch_floats() floats of scratch");
    const int64_t kps = ceil_div64(ceil_div64(k, splits), kGK) * kGK;
    const dim3 grid(tiles.x, tiles.y, (unsigned)ceil_div64(k, kps));
    gemm_tile_kernel<true><<<grid, 256, 0, st>>>(A, lda, B, ldb, scratch, n, m, n, k, kps, m * n);
    TG_LAUNCH_CHECK();
    gemm_sum_splits_kernel<<<(unsigned)ceil_div64(m * n, 256), 256, 0, st>>>(scratch, (int)grid.z, m, n, C, ldc);
    TG_LAUNCH_CHECK();
    return TG_OK;
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
    sum_partials_kernel<<<(unsigned)ceil_div64((int64_t)c * 8, 256), 256, 0, st>>>(scratch, grid, c, out, c, nullptr, c, c);
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
    sum_partials_kernel<<<1, 256, 0, st>>>(scratch, grid, 1, out, 1, nullptr, 1, 1);
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

// ------------------------------------------------------------------------------------------------------------
// Adam step (torch.optim.Adam semantics, no amsgrad): one streaming pass, 128-bit accesses, grid = 8 CTAs per SM
// ------------------------------------------------------------------------------------------------------------
namespace tg {
// om1 = 1 - beta1, om2 = 1 - beta2 are formed in double on the host (1.f - 0.999f is 1.3e-5 away from 0.001f)
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float lr_t, float om1, float b2, float om2,
                                          float inv_bc2_sqrt, float eps, float wd) {
    g = fmaf(wd, p, g);
    m = fmaf(om1, g - m, m);
    v = fmaf(om2, g * g, b2 * v);
    const float denom = fmaf(sqrtf(v), inv_bc2_sqrt, eps);
    p -= lr_t * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr_t, float om1, float b2, float om2,
                                                   float inv_bc2_sqrt, float eps, float wd, int vec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (; i < n4; i += stride) {
            float4 p4 = reinterpret_cast<float4*>(param)[i];
            const float4 g4 = ldg_f4_stream(grad + i * 4);
            float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
            adam_elem(p4.x, g4.x, m4.x, v4.x, lr_t, om1, b2, om2, inv_bc2_sqrt, eps, wd);
            adam_elem(p4.y, g4.y, m4.y, v4.y, lr_t, om1, b2, om2, inv_bc2_sqrt, eps, wd);
            adam_elem(p4.z, g4.z, m4.z, v4.z, lr_t, om1, b2, om2, inv_bc2_sqrt, eps, wd);
            adam_elem(p4.w, g4.w, m4.w, v4.w, lr_t, om1, b2, om2, inv_bc2_sqrt, eps, wd);
            reinterpret_cast<float4*>(param)[i] = p4;
            reinterpret_cast<float4*>(m)[i] = m4;
            reinterpret_cast<float4*>(v)[i] = v4;
        }
        i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // tail
    }
    for (; i < n; i += stride) {
        float p1 = param[i], m1 = m[i], v1 = v[i];
        adam_elem(p1, grad[i], m1, v1, lr_t, om1, b2, om2, inv_bc2_sqrt, eps, wd);
        param[i] = p1; m[i] = m1; v[i] = v1;
    }
}
}  // namespace tg

extern "C" int tg_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                           double beta2, double eps, double weight_decay, int64_t step, void* stream) {
    using namespace tg;
    TG_REQUIRE(param && grad && exp_avg && exp_avg_sq, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n >= 0 && step >= 1, TG_ERR_INVALID_ARG, "n must be >= 0 and step >= 1");
    if (n == 0) return TG_OK;
    // bias corrections in double on the host, like torch.optim.Adam
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float lr_t = (float)(lr / bc1);
    const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int vec = al16(param) && al16(grad) && al16(exp_avg) && al16(exp_avg_sq);
    int64_t grid = ceil_div64(vec ? (n + 3) / 4 : n, 256);
    if (grid > 8 * kNumSM) grid = 8 * kNumSM;
    adam_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr_t, (float)(1.0 - beta1), (float)beta2,
                                                              (float)(1.0 - beta2), inv_bc2_sqrt, (float)eps, (float)weight_decay, vec);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// per-class tp / fp / fn counts of argmax(logits) against the row labels (integer atomics: order independent)
// ------------------------------------------------------------------------------------------------------------
namespace tg {
constexpr int kCcMaxClass = 1024;
__global__ void __launch_bounds__(256) class_counts_kernel(const float* __restrict__ logits, int64_t ldl,
                                                           const int32_t* __restrict__ row_label, int64_t n, int c,
                                                           int32_t* __restrict__ counts) {
    extern __shared__ int32_t cc_sh[];  // [3 * c]
    for (int i = threadIdx.x; i < 3 * c; i += blockDim.x) cc_sh[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const int y = __ldg(row_label + r);
        if (y < 0) continue;
        const float* z = logits + r * ldl;
        float best = __ldg(z);
        int pred = 0;
        for (int k = 1; k < c; ++k) {
            const float v = __ldg(z + k);
            if (v > best) { best = v; pred = k; }  // strict: the first maximum wins, like th.max
        }
        if (pred == y) {
            atomicAdd(&cc_sh[y], 1);
        } else {
            atomicAdd(&cc_sh[c + pred], 1);
            if (y < c) atomicAdd(&cc_sh[2 * c + y], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * c; i += blockDim.x)
        if (cc_sh[i]) atomicAdd(counts + i, cc_sh[i]);
}
}  // namespace tg

extern "C" int tg_class_counts_i32(const float* logits, int64_t ldl, const int32_t* row_label, int64_t n, int32_t c,
                                   int32_t* counts, void* stream) {
    using namespace tg;
    TG_REQUIRE(logits && row_label && counts, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(c >= 1 && c <= kCcMaxClass && ldl >= c && n >= 0, TG_ERR_INVALID_ARG, "need 1 <= c <= %d, ldl >= c", kCcMaxClass);
    cudaStream_t st = as_stream(stream);
    TG_CUDA(cudaMemsetAsync(counts, 0, (size_t)3 * c * sizeof(int32_t), st));
    if (n == 0) return TG_OK;
    int64_t grid = ceil_div64(n, 256);
    if (grid > 4 * kNumSM) grid = 4 * kNumSM;
    class_counts_kernel<<<(unsigned)grid, 256, (size_t)3 * c * sizeof(int32_t), st>>>(logits, ldl, row_label, n, c, counts);
    TG_LAUNCH_CHECK();
    return TG_OK;
}
