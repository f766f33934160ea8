// tg_finish.cuh — finishing kernel of the streaming SpMM kernels: hub row = fixed-order sum of the per-CTA partial rows
// of its virtual slots, then the fused epilogue.  Deterministic (no float atomics).
#pragma once
#include "tg_epilogue.cuh"

namespace tg {

// ---- finishing kernel: hub row k = sum over CTA groups (fixed order) + epilogue --------------------------------------
template <int VEC, int G, int CPL, class Epi>
__global__ void __launch_bounds__(256) stream_finish_kernel(const float* __restrict__ partials, int64_t ldp, int n_groups,
                                                            int Kh, int Kv, const int32_t* __restrict__ vmap,
                                                            const int32_t* __restrict__ vcnt,
                                                            const int32_t* __restrict__ hub_rows, int n_chunks,
                                                            const Epi epi) {
    // One block row-group per hub row; the 8 warps of a block take the CTA partials g = w, w+8, ... (each warp in
    // ascending order), deposit their sums in shared memory and warp 0 adds the 8 deposits in warp order: a fixed tree.
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) float fin_s[];  // [8 warps][GPW groups][G*CPL*VEC]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1);
    const int gw = lane / G;
    const unsigned gmask = group_mask<G>(lane);
    const int k = blockIdx.x * GPW + gw;
    Chunk<VEC> acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = chunk_zero<VEC>();
    if (k < Kh) {
        const int nv = __ldg(vcnt + k);
        for (int j = 0; j < nv; ++j) {
            const int v = __ldg(vmap + k * 8 + j);
            for (int g = warp; g < n_groups; g += 8) {
                const float* src = partials + ((int64_t)g * Kv + v) * ldp;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int chunk = gl + i * G;
                    if (chunk < n_chunks) {
                        const Chunk<VEC> t = chunk_ldg<VEC>(src + (int64_t)chunk * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[i].v[e] += t.v[e];
                    }
                }
            }
        }
    }
    constexpr int ROWF = G * CPL * VEC;
    float* mine = fin_s + ((size_t)warp * GPW + gw) * ROWF;
#pragma unroll
    for (int i = 0; i < CPL; ++i)
#pragma unroll
        for (int e = 0; e < VEC; ++e) mine[(gl + i * G) * VEC + e] = acc[i].v[e];
    __syncthreads();
    if (warp != 0 || k >= Kh) return;
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = chunk_zero<VEC>();
    for (int w = 0; w < 8; ++w) {
        const float* src = fin_s + ((size_t)w * GPW + gw) * ROWF;
#pragma unroll
        for (int i = 0; i < CPL; ++i)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[i].v[e] += src[(gl + i * G) * VEC + e];
    }
    epi.template apply<VEC, G, CPL>((int64_t)__ldg(hub_rows + k), gl, gmask, n_chunks, acc);
}

struct FinishArgs {
    const float* partials;
    int64_t ldp;
    int n_groups, Kh, Kv;
    const int32_t *vmap, *vcnt, *hub_rows;
    int n_chunks;
};

template <int VEC, int G, int CPL, class Epi>
static int launch_finish(const FinishArgs& f, const Epi& epi, cudaStream_t st) {
    constexpr int GPW = 32 / G;
    const size_t smem = (size_t)8 * GPW * G * CPL * VEC * sizeof(float);
    stream_finish_kernel<VEC, G, CPL, Epi><<<(unsigned)ceil_div64(f.Kh, GPW), 256, smem, st>>>(
        f.partials, f.ldp, f.n_groups, f.Kh, f.Kv, f.vmap, f.vcnt, f.hub_rows, f.n_chunks, epi);
    TG_LAUNCH_CHECK();
    return TG_OK;
}


template <class Epi>
static int finish_run(const FinishArgs& f, const Epi& epi, cudaStream_t st) {
#define TG_LAUNCH_FIN(V, G, C) launch_finish<V, G, C>(f, epi, st)
    TG_SHAPE_SWITCH(4, f.n_chunks, TG_LAUNCH_FIN);
#undef TG_LAUNCH_FIN
    set_error("n_feat too wide for the streaming finish kernel");
    return TG_ERR_UNSUPPORTED;
}

}  // namespace tg
