// tg_epilogue.cuh — register tiles and the fused epilogues shared by the SpMM kernels (tg_spmm.cu, tg_stream.cu).
//
//   EpiStore : Y = dropout(relu(acc + bias))            reference layer.py:109-110, :182, :185
//   EpiLoss  : logits = acc + bias; log-softmax; masked cross-entropy and its gradient   trainer.py:358-359
#pragma once
#include "tg_common.cuh"

namespace tg {

// ---- VEC-generic register tiles ----------------------------------------------------------------------------
template <int VEC>
struct Chunk {
    float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Chunk<VEC> chunk_zero() {
    Chunk<VEC> c;
#pragma unroll
    for (int k = 0; k < VEC; ++k) c.v[k] = 0.f;
    return c;
}

template <int VEC>
__device__ __forceinline__ Chunk<VEC> chunk_ldg(const float* p);
template <>
__device__ __forceinline__ Chunk<4> chunk_ldg<4>(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Chunk<4> c;
    c.v[0] = t.x; c.v[1] = t.y; c.v[2] = t.z; c.v[3] = t.w;
    return c;
}
template <>
__device__ __forceinline__ Chunk<1> chunk_ldg<1>(const float* p) {
    Chunk<1> c;
    c.v[0] = __ldg(p);
    return c;
}

// partial rows are written by other SMs in this launch: read them at L2 (.cg), never through L1
template <int VEC>
__device__ __forceinline__ Chunk<VEC> chunk_ldcg(const float* p);
template <>
__device__ __forceinline__ Chunk<4> chunk_ldcg<4>(const float* p) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    Chunk<4> c;
    c.v[0] = t.x; c.v[1] = t.y; c.v[2] = t.z; c.v[3] = t.w;
    return c;
}
template <>
__device__ __forceinline__ Chunk<1> chunk_ldcg<1>(const float* p) {
    Chunk<1> c;
    c.v[0] = __ldcg(p);
    return c;
}

template <int VEC>
__device__ __forceinline__ void chunk_st(float* p, const Chunk<VEC>& c);
template <>
__device__ __forceinline__ void chunk_st<4>(float* p, const Chunk<4>& c) {
    *reinterpret_cast<float4*>(p) = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
}
template <>
__device__ __forceinline__ void chunk_st<1>(float* p, const Chunk<1>& c) {
    *p = c.v[0];
}

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
    if (G == 32) return 0xffffffffu;
    return ((1u << G) - 1u) << ((lane / G) * G);
}


// ---- epilogues ------------------------------------------------------------------------------------------------
// Elementwise: Y = dropout(relu(acc + bias))   (each stage optional)
struct EpiStore {
    float* Y;
    int64_t ldy;
    const float* bias;
    int relu;
    int drop_mode;  // 0 none, 1 Philox counter RNG, 2 explicit keep mask
    const uint8_t* keep_mask;
    uint32_t keep_thr;
    float scale;
    uint64_t seed, offset;
    int32_t n_feat;
    int64_t raw_row_begin;  // rows >= raw_row_begin are stored as plain sums (document-sharded mode), INT64_MAX = none
    const float* out_scale; // optional DEVICE scalar multiplied into every output element (autograd upstream gradient)
    const unsigned long long* offset_dev;  // optional DEVICE counter added to `offset` (lets a captured CUDA graph draw a fresh mask per replay)

    template <int VEC, int G, int CPL>
    __device__ __forceinline__ void apply(int64_t row, int gl, unsigned gmask, int n_chunks,
                                          Chunk<VEC> (&acc)[CPL]) const {
        if (row >= raw_row_begin) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int chunk = gl + i * G;
                if (chunk < n_chunks) chunk_st<VEC>(Y + row * ldy + chunk * VEC, acc[i]);
            }
            return;
        }
        if (!relu && drop_mode == 0) {
            // linear epilogue (the class-sized products of layer 2: optional bias, optional upstream scale): decided once per
            // row instead of testing every stage for every chunk
            const float gsc = out_scale ? __ldg(out_scale) : 1.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int chunk = gl + i * G;
                if (chunk >= n_chunks) continue;
                const int col0 = chunk * VEC;
                Chunk<VEC> y = acc[i];
                if (bias) {
                    const Chunk<VEC> bb = chunk_ldg<VEC>(bias + col0);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) y.v[k] += bb.v[k];
                }
                if (out_scale) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) y.v[k] *= gsc;
                }
                chunk_st<VEC>(Y + row * ldy + col0, y);
            }
            return;
        }
        Philox4 rnd = Philox4{0, 0, 0, 0};
        int rnd_cidx = -1;  // exact-half mode: the 128-column block `rnd` was drawn for
        const uint64_t offset = (drop_mode == 1 && offset_dev) ? this->offset + __ldg(offset_dev) : this->offset;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int chunk = gl + i * G;
            if (chunk >= n_chunks) continue;
            const int col0 = chunk * VEC;
            Chunk<VEC> y = acc[i];
            if (bias) {
                const Chunk<VEC> bb = chunk_ldg<VEC>(bias + col0);
#pragma unroll
                for (int k = 0; k < VEC; ++k) y.v[k] += bb.v[k];
            }
            if (relu) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) y.v[k] = fmaxf(y.v[k], 0.f);
            }
            if (out_scale) {
                const float g = __ldg(out_scale);
#pragma unroll
                for (int k = 0; k < VEC; ++k) y.v[k] *= g;
            }
            if (drop_mode == 1 && keep_thr == kDropoutHalfThr) {
                // exact-half mode (tg_common.cuh): one bit per element, one Philox call per 128 columns
                if (VEC == 4) {
                    const int q = chunk;
                    if ((q >> 5) != rnd_cidx) {
                        rnd_cidx = q >> 5;
                        rnd = dropout_philox_half(row, (uint32_t)rnd_cidx, seed, offset);
                    }
                    const uint32_t w = dropout_half_word(rnd, q) >> (4 * (q & 7));
#pragma unroll
                    for (int k = 0; k < VEC; ++k) y.v[k] = ((w >> k) & 1u) ? y.v[k] * scale : 0.f;
                } else {
                    if ((col0 >> 7) != rnd_cidx) {
                        rnd_cidx = col0 >> 7;
                        rnd = dropout_philox_half(row, (uint32_t)rnd_cidx, seed, offset);
                    }
                    const uint32_t w = dropout_half_word(rnd, col0 >> 2) >> (col0 & 31);
                    y.v[0] = (w & 1u) ? y.v[0] * scale : 0.f;
                }
            } else if (drop_mode == 1) {
                if (VEC == 4) {
                    // chunk q: Philox call (q % 8, q / 16), half (q / 8) % 2 (see tg_common.cuh); with G == 8 the
                    // chunks (q, q + 8) of consecutive i share one call
                    const int q = chunk;
                    const bool reuse = (G == 8) && (i & 1);
                    if (!reuse) rnd = dropout_philox(row, (uint32_t)(q & 7), (uint32_t)(q >> 4), seed, offset);
                    uint32_t u[4];
                    dropout_u16x4(rnd, (q >> 3) & 1, u);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) y.v[k] = (u[k] < keep_thr) ? y.v[k] * scale : 0.f;
                } else {
                    const int q = col0 >> 2;
                    const Philox4 r1 = dropout_philox(row, (uint32_t)(q & 7), (uint32_t)(q >> 4), seed, offset);
                    uint32_t u[4];
                    dropout_u16x4(r1, (q >> 3) & 1, u);
                    y.v[0] = (u[col0 & 3] < keep_thr) ? y.v[0] * scale : 0.f;
                }
            } else if (drop_mode == 2) {
                const uint8_t* mp = keep_mask + row * (int64_t)n_feat + col0;
                if (VEC == 4) {
                    const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(mp));
#pragma unroll
                    for (int k = 0; k < VEC; ++k) y.v[k] = ((m >> (8 * k)) & 0xffu) ? y.v[k] * scale : 0.f;
                } else {
                    y.v[0] = __ldg(mp) ? y.v[0] * scale : 0.f;
                }
            }
            chunk_st<VEC>(Y + row * ldy + col0, y);
        }
    }
};

// Row-wise: logits = acc + bias; log-softmax; masked cross-entropy and its gradient.
struct EpiLoss {
    const float* bias;
    const int32_t* row_label;
    float inv_count;
    float* logits;  // optional
    int64_t ldl;
    float* dZ;  // optional
    int64_t ldd;
    float* row_loss;
    int32_t n_class;

    template <int VEC, int G, int CPL>
    __device__ __forceinline__ void apply(int64_t row, int gl, unsigned gmask, int n_chunks,
                                          Chunk<VEC> (&acc)[CPL]) const {
        const int y = __ldg(row_label + row);
        if (y < 0 && !logits) {
            // a row without a label (validation / test documents, topic rows) has no loss and a zero gradient: nothing to
            // exponentiate (the whole group agrees on y: it is a property of the row)
            if (gl == 0) row_loss[row] = 0.f;
            if (dZ) {
                Chunk<VEC> z;
#pragma unroll
                for (int k = 0; k < VEC; ++k) z.v[k] = 0.f;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int chunk = gl + i * G;
                    if (chunk < n_chunks) chunk_st<VEC>(dZ + row * ldd + chunk * VEC, z);
                }
            }
            return;
        }
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int chunk = gl + i * G;
            if (chunk >= n_chunks) continue;
            const int col0 = chunk * VEC;
            if (bias) {
                const Chunk<VEC> bb = chunk_ldg<VEC>(bias + col0);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[i].v[k] += bb.v[k];
            }
            if (logits) chunk_st<VEC>(logits + row * ldl + col0, acc[i]);
#pragma unroll
            for (int k = 0; k < VEC; ++k) m = fmaxf(m, acc[i].v[k]);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(gmask, m, o, G));
        // e = exp(z - max) is kept: the softmax of the gradient is e / sum(e), one exponential per element instead of two
        Chunk<VEC> e[CPL];
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int chunk = gl + i * G;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                e[i].v[k] = (chunk < n_chunks) ? expf(acc[i].v[k] - m) : 0.f;
                se += e[i].v[k];
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) se += __shfl_xor_sync(gmask, se, o, G);
        const float lse = m + logf(se);
        // z_y lives in chunk y/VEC, owned by lane (y/VEC) % G, register (y/VEC) / G
        float zy = 0.f;
        if (y >= 0) {
            const int ychunk = y / VEC, yk = y % VEC;
            float mine = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i)
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    if (gl + i * G == ychunk && k == yk) mine = acc[i].v[k];
            zy = __shfl_sync(gmask, mine, ychunk % G, G);
        }
        if (gl == 0) row_loss[row] = (y >= 0) ? (lse - zy) * inv_count : 0.f;
        if (dZ) {
            const float inv_se = 1.f / se;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int chunk = gl + i * G;
                if (chunk >= n_chunks) continue;
                const int col0 = chunk * VEC;
                Chunk<VEC> g;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float sm = e[i].v[k] * inv_se;
                    g.v[k] = (y >= 0) ? (sm - ((col0 + k) == y ? 1.f : 0.f)) * inv_count : 0.f;
                }
                chunk_st<VEC>(dZ + row * ldd + col0, g);
            }
        }
    }
};


// shape dispatch: G lanes per row and CPL chunks per lane so that G*CPL >= n_chunks
#define TG_SHAPE_SWITCH(VEC, nc, LAUNCH)                                   \
    do {                                                                   \
        if ((nc) <= 2) return LAUNCH(VEC, 2, 1);                           \
        if ((nc) <= 4) return LAUNCH(VEC, 4, 1);                           \
        if ((nc) <= 8) return LAUNCH(VEC, 8, 1);                           \
        if ((nc) <= 16) return LAUNCH(VEC, 16, 1);                         \
        if ((nc) <= 32) return LAUNCH(VEC, 32, 1);                         \
        if ((nc) <= 64) return LAUNCH(VEC, 32, 2);                         \
        if ((nc) <= 96) return LAUNCH(VEC, 32, 3);                         \
        if ((nc) <= 128) return LAUNCH(VEC, 32, 4);                        \
        if ((nc) <= 192) return LAUNCH(VEC, 32, 6);                        \
        if ((nc) <= 256) return LAUNCH(VEC, 32, 8);                        \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace tg
