// tg_spmm.cu — skew-aware deterministic CSR SpMM  Y = A * B  with fused epilogues (sm_100a).
//
// Replaces `th.spmm(adj, support)` (reference layer.py:106) and its autograd `sparse.t().mm(grad)`
// (SURVEY §3.3), plus the elementwise tail it feeds: bias add (layer.py:109-110), relu (layer.py:182),
// dropout (layer.py:185) for layer 1, and bias + log-softmax + masked cross-entropy (trainer.py:358-359)
// for layer 2.
//
// Work decomposition (tg_plan, tg_csr.cu):
//   * short rows (<= hub_threshold stored entries; the millions of document rows with 5-20 entries):
//     one lane GROUP of G lanes per row, G*CPL*VEC >= F, so that a row of B is fetched with 128-bit loads
//     and each lane owns CPL vector chunks of the output row.  Index/value pairs are fetched coalesced by
//     the group and broadcast with warp shuffles; U gathered rows are kept in flight per lane.
//   * hub rows (topic rows with 10^3..10^5 entries): split into fixed segments; one group per segment
//     writes a partial row; the LAST group to arrive for a row (integer ticket, no float atomics) adds the
//     partials in segment order and applies the epilogue.  The order of every floating-point addition is
//     fixed by the plan, so results are bitwise reproducible run to run.
// Hub segments occupy the first blocks of the grid so the long work starts first.
#include "tg_epilogue.cuh"
#include <type_traits>

#include "tg_roles.cuh"

namespace tg {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = 4;  // gathered rows in flight per lane

struct SpmmArgs {
    const int32_t* __restrict__ rowptr;
    const int32_t* __restrict__ colidx;
    const float* __restrict__ vals;
    const float* __restrict__ B;
    int64_t ldb;
    int64_t n_rows;
    int32_t n_feat;
    int32_t n_chunks;  // ceil(n_feat / VEC)
    int32_t hub_threshold;
    // split rows
    const int32_t* __restrict__ hub_rows;
    const int32_t* __restrict__ hub_seg_ptr;
    const int32_t* __restrict__ seg_hub;
    const int32_t* __restrict__ seg_begin;
    const int32_t* __restrict__ seg_end;
    const int32_t* __restrict__ seg_order;
    uint32_t* tickets;
    float* partials;
    int64_t ldp;
    int32_t n_seg;
    int32_t n_hub_blocks;
};

// ---- the gather/accumulate core: acc += sum_{p in [s,e)} vals[p] * B[colidx[p], my chunks] -------------------
template <int VEC, int G, int CPL>
__device__ __forceinline__ void accumulate_range(const SpmmArgs& a, int s, int e, int gl, unsigned gmask,
                                                 Chunk<VEC> (&acc)[CPL]) {
    for (int base = s; base < e; base += G) {
        const int p = base + gl;
        int c = 0;
        float v = 0.f;
        if (p < e) {
            c = __ldg(a.colidx + p);
            v = __ldg(a.vals + p);
        }
        const int cnt = min(G, e - base);
        for (int j = 0; j < cnt; j += kUnroll) {
            int cj[kUnroll];
            float vj[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                cj[u] = __shfl_sync(gmask, c, j + u, G);
                vj[u] = __shfl_sync(gmask, v, j + u, G);
            }
            Chunk<VEC> b[kUnroll][CPL];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const float* bp = a.B + (int64_t)cj[u] * a.ldb;
                const bool live = (j + u) < cnt;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int chunk = gl + i * G;
                    b[u][i] = (live && chunk < a.n_chunks) ? chunk_ldg<VEC>(bp + (int64_t)chunk * VEC)
                                                           : chunk_zero<VEC>();
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int i = 0; i < CPL; ++i)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[i].v[k] = fmaf(vj[u], b[u][i].v[k], acc[i].v[k]);
        }
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------
template <int VEC, int G, int CPL, class Epi>
__global__ void __launch_bounds__(kThreads) spmm_kernel(const SpmmArgs a, const Epi epi) {
    constexpr int GPW = 32 / G;            // groups per warp
    constexpr int GPB = kWarps * GPW;      // groups per block
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1);
    const int grp = warp * GPW + lane / G;
    const unsigned gmask = group_mask<G>(lane);

    Chunk<VEC> acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = chunk_zero<VEC>();

    if ((int)blockIdx.x < a.n_hub_blocks) {
        // ---- split-row role: one WARP per segment ---------------------------------------------------------
        // The 32/G lane groups of the warp take the segment's entries round-robin and their partial sums are combined
        // with a fixed butterfly of shuffles (warp-cooperative segmented reduction): for narrow rows (G = 2..16) a
        // segment is no longer a serial chain of 256 dependent gathers by a handful of lanes.
        const int slot = blockIdx.x * kWarps + warp;
        if (slot >= a.n_seg) return;
        const int seg = a.seg_order ? __ldg(a.seg_order + slot) : slot;  // executed in column order, stored/added in segment order
        const int hub = __ldg(a.seg_hub + seg);
        if (G == 32) {
            accumulate_range<VEC, G, CPL>(a, __ldg(a.seg_begin + seg), __ldg(a.seg_end + seg), gl, gmask, acc);
        } else {
            const int sb = __ldg(a.seg_begin + seg), se = __ldg(a.seg_end + seg);
            const int sub = lane / G;
            for (int p = sb + sub; p < se; p += GPW * kUnroll) {
                int c[kUnroll];
                float v[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int q = p + u * GPW;
                    c[u] = (q < se) ? __ldg(a.colidx + q) : -1;
                    v[u] = (q < se) ? __ldg(a.vals + q) : 0.f;
                }
                Chunk<VEC> b[kUnroll][CPL];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const int chunk = gl + i * G;
                        b[u][i] = (c[u] >= 0 && chunk < a.n_chunks)
                                      ? chunk_ldg<VEC>(a.B + (int64_t)c[u] * a.ldb + (int64_t)chunk * VEC)
                                      : chunk_zero<VEC>();
                    }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                    for (int i = 0; i < CPL; ++i)
#pragma unroll
                        for (int k = 0; k < VEC; ++k) acc[i].v[k] = fmaf(v[u], b[u][i].v[k], acc[i].v[k]);
            }
#pragma unroll
            for (int off = G; off < 32; off <<= 1)
#pragma unroll
                for (int i = 0; i < CPL; ++i)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[i].v[k] += __shfl_xor_sync(0xffffffffu, acc[i].v[k], off);
            if (sub != 0) return;  // group 0 of the warp carries the segment's partial row from here on
        }
        const int s0 = __ldg(a.hub_seg_ptr + hub), s1 = __ldg(a.hub_seg_ptr + hub + 1);
        float* my = a.partials + (int64_t)seg * a.ldp;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int chunk = gl + i * G;
            if (chunk < a.n_chunks) chunk_st<VEC>(my + (int64_t)chunk * VEC, acc[i]);
        }
        __threadfence();  // partial row visible device-wide before the ticket is taken
        __syncwarp(gmask);
        unsigned t = 0;
        if (gl == 0) t = atomicAdd(a.tickets + hub, 1u);
        t = __shfl_sync(gmask, t, 0, G);
        if (t != (unsigned)(s1 - s0 - 1)) return;
        // ---- last arriver: fixed-order reduction of the row's partials + epilogue --------------------------
        __threadfence();
#pragma unroll
        for (int i = 0; i < CPL; ++i) acc[i] = chunk_zero<VEC>();
        int sg = s0;
        for (; sg + 4 <= s1; sg += 4) {
            Chunk<VEC> t4[4][CPL];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int chunk = gl + i * G;
                    t4[u][i] = (chunk < a.n_chunks)
                                   ? chunk_ldcg<VEC>(a.partials + (int64_t)(sg + u) * a.ldp + (int64_t)chunk * VEC)
                                   : chunk_zero<VEC>();
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < CPL; ++i)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[i].v[k] += t4[u][i].v[k];
        }
        for (; sg < s1; ++sg) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int chunk = gl + i * G;
                if (chunk < a.n_chunks) {
                    const Chunk<VEC> tt = chunk_ldcg<VEC>(a.partials + (int64_t)sg * a.ldp + (int64_t)chunk * VEC);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[i].v[k] += tt.v[k];
                }
            }
        }
        if (gl == 0) a.tickets[hub] = 0;  // re-arm for the next launch
        epi.template apply<VEC, G, CPL>((int64_t)__ldg(a.hub_rows + hub), gl, gmask, a.n_chunks, acc);
        return;
    }

    // ---- short-row role: one group per row ----------------------------------------------------------------
    const int64_t row = ((int64_t)blockIdx.x - a.n_hub_blocks) * GPB + grp;
    if (row >= a.n_rows) return;
    const int s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
    if (e - s > a.hub_threshold) return;  // produced by the split-row role
    accumulate_range<VEC, G, CPL>(a, s, e, gl, gmask, acc);
    epi.template apply<VEC, G, CPL>(row, gl, gmask, a.n_chunks, acc);
}

// ---- host-side dispatch ------------------------------------------------------------------------------------------
template <int VEC, int G, int CPL, class Epi>
static int launch_cfg(SpmmArgs a, const Epi& epi, cudaStream_t st) {
    constexpr int GPB = kWarps * (32 / G);
    a.n_hub_blocks = (int32_t)ceil_div64(a.n_seg, kWarps);  // one warp per split-row segment
    const int64_t row_blocks = ceil_div64(a.n_rows, GPB);
    const int64_t grid = a.n_hub_blocks + row_blocks;
    if (grid == 0) return TG_OK;
    TG_REQUIRE(grid < (int64_t)INT32_MAX, TG_ERR_OVERFLOW, "grid too large");
    spmm_kernel<VEC, G, CPL, Epi><<<(unsigned)grid, kThreads, 0, st>>>(a, epi);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

template <int VEC, class Epi>
static int launch_vec(SpmmArgs a, const Epi& epi, cudaStream_t st) {
#define TG_LAUNCH_SPMM(V, G, C) launch_cfg<V, G, C>(a, epi, st)
    TG_SHAPE_SWITCH(VEC, a.n_chunks, TG_LAUNCH_SPMM);
#undef TG_LAUNCH_SPMM
    set_error("n_feat=%d too wide for one pass (max %d)", a.n_feat, 256 * VEC);
    return TG_ERR_UNSUPPORTED;
}

// ---- stand-alone row-wise loss on existing logits (same epilogue code as the fused layer-2 kernel) ------------
template <int VEC, int G, int CPL>
__global__ void __launch_bounds__(kThreads) rowwise_loss_kernel(const float* __restrict__ Z, int64_t ldz,
                                                                int64_t n_rows, int n_chunks, const EpiLoss epi) {
    constexpr int GPW = 32 / G;
    constexpr int GPB = kWarps * GPW;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = group_mask<G>(lane);
    const int64_t row = (int64_t)blockIdx.x * GPB + (threadIdx.x >> 5) * GPW + lane / G;
    if (row >= n_rows) return;
    Chunk<VEC> acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int chunk = gl + i * G;
        acc[i] = (chunk < n_chunks) ? chunk_ldg<VEC>(Z + row * ldz + (int64_t)chunk * VEC) : chunk_zero<VEC>();
    }
    epi.template apply<VEC, G, CPL>(row, gl, gmask, n_chunks, acc);
}

template <int VEC, int G, int CPL>
static int launch_rowwise_loss(const float* Z, int64_t ldz, int64_t n_rows, int n_chunks, const EpiLoss& epi,
                               cudaStream_t st) {
    constexpr int GPB = kWarps * (32 / G);
    if (n_rows == 0) return TG_OK;
    rowwise_loss_kernel<VEC, G, CPL><<<(unsigned)ceil_div64(n_rows, GPB), kThreads, 0, st>>>(Z, ldz, n_rows, n_chunks, epi);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

template <int VEC>
static int rowwise_loss_vec(const float* Z, int64_t ldz, int64_t n_rows, int n_chunks, const EpiLoss& epi,
                            cudaStream_t st) {
#define TG_LAUNCH_RL(V, G, C) launch_rowwise_loss<V, G, C>(Z, ldz, n_rows, n_chunks, epi, st)
    TG_SHAPE_SWITCH(VEC, n_chunks, TG_LAUNCH_RL);
#undef TG_LAUNCH_RL
    set_error("n_class=%d too wide", epi.n_class);
    return TG_ERR_UNSUPPORTED;
}

template <class Epi> struct EpiTraits { static constexpr bool whole_row = false; };
template <> struct EpiTraits<EpiLoss> { static constexpr bool whole_row = true; };  // log-softmax needs the whole row in one lane group

template <class Epi>
static int run_spmm(const tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                    const float* B, int64_t ldb, int32_t n_feat, bool out_vec4_ok, const Epi& epi, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
    TG_REQUIRE(pl && rowptr && B, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(pl->nnz == 0 || (colidx && vals), TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n_feat > 0, TG_ERR_INVALID_ARG, "n_feat must be positive");
    TG_REQUIRE(ldb >= n_feat, TG_ERR_INVALID_ARG, "ldb < n_feat");
    const size_t need = tg_plan_workspace_bytes(pl, n_feat);
    TG_REQUIRE((pl->n_seg == 0 && !pl->r2_ok) || (workspace && workspace_bytes >= need), TG_ERR_WORKSPACE,
               "workspace %zu B < required %zu B", workspace_bytes, need);
    if (out_vec4_ok) {
        // role-specialised streaming kernels (tg_roles2.cu) when the plan carries their sub-plan and the operands qualify
        StreamCall sc{rowptr, vals, B, ldb, n_feat, workspace, workspace_bytes};
        if constexpr (std::is_same<Epi, EpiStore>::value) {
            // rectangular operands (sparse feature matrix x weight, and the transpose product): one role each
            // (the resident-table kernel has no Philox path: a dropout epilogue on X * W takes the gather kernel)
            if (roles2_rect_applicable(pl, sc) && !(pl->r2_rect == 1 && epi.drop_mode == 1)) return roles2_rect_run(pl, sc, epi, st);
        }
        if (roles2_applicable(pl, sc, EpiTraits<Epi>::whole_row)) return roles2_run(pl, sc, epi, st);
        if (roles2_narrow_applicable(pl, sc)) return roles2_narrow_run(pl, sc, epi, st);
    }
    SpmmArgs a;
    a.rowptr = rowptr; a.colidx = colidx; a.vals = vals; a.B = B; a.ldb = ldb;
    a.n_rows = pl->n_rows; a.n_feat = n_feat; a.hub_threshold = pl->hub_threshold;
    a.hub_rows = pl->hub_rows; a.hub_seg_ptr = pl->hub_seg_ptr; a.seg_hub = pl->seg_hub;
    a.seg_begin = pl->seg_begin; a.seg_end = pl->seg_end; a.tickets = pl->tickets;
    // column order pays off when B exceeds L2; narrow operands (<= 32 columns) are L2 resident and prefer storage order
    a.seg_order = (n_feat > 32) ? pl->seg_order : nullptr;
    // partial rows start 16 B aligned inside the caller's workspace
    a.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 15u) & ~(uintptr_t)15u);
    a.ldp = (int64_t)((n_feat + 3) / 4) * 4;
    a.n_seg = pl->n_seg; a.n_hub_blocks = 0;
    const bool vec4 = (n_feat % 4 == 0) && (ldb % 4 == 0) && aligned16(B) && out_vec4_ok;
    if (vec4) {
        a.n_chunks = n_feat / 4;
        return launch_vec<4>(a, epi, st);
    }
    a.n_chunks = n_feat;
    return launch_vec<1>(a, epi, st);
}

// ---- keep-mask materialisation (tests) ------------------------------------------------------------------------
__global__ void keep_mask_kernel(uint8_t* __restrict__ out, int64_t n_rows, int32_t n_feat, uint32_t thr,
                                 uint64_t seed, uint64_t offset) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * (int64_t)n_feat) return;
    const int64_t row = idx / n_feat;
    const int col = (int)(idx % n_feat);
    const int q = col >> 2;
    if (thr == kDropoutHalfThr) {  // exact-half mode: one bit per element
        const Philox4 r = dropout_philox_half(row, (uint32_t)(col >> 7), seed, offset);
        out[idx] = (dropout_half_word(r, q) >> (col & 31)) & 1u;
        return;
    }
    const Philox4 r = dropout_philox(row, (uint32_t)(q & 7), (uint32_t)(q >> 4), seed, offset);
    uint32_t u[4];
    dropout_u16x4(r, (q >> 3) & 1, u);
    out[idx] = (u[col & 3] < thr) ? 1 : 0;
}

}  // namespace tg

extern "C" {

int tg_spmm_f32(const tg_plan* plan, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                const float* B, int64_t ldb, float* Y, int64_t ldy, int32_t n_feat, const float* bias,
                const float* out_scale, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace tg;
    TG_REQUIRE(Y, TG_ERR_INVALID_ARG, "null output");
    TG_REQUIRE(ldy >= n_feat, TG_ERR_INVALID_ARG, "ldy < n_feat");
    EpiStore epi{};
    epi.Y = Y; epi.ldy = ldy; epi.bias = bias; epi.relu = 0; epi.drop_mode = 0; epi.keep_mask = nullptr;
    epi.keep_thr = 0; epi.scale = 1.f; epi.seed = 0; epi.offset = 0; epi.n_feat = n_feat;
    epi.raw_row_begin = INT64_MAX; epi.out_scale = out_scale; epi.offset_dev = nullptr;
    const bool ok4 = (ldy % 4 == 0) && aligned16(Y) && (!bias || aligned16(bias));
    return run_spmm(plan, rowptr, colidx, vals, B, ldb, n_feat, ok4, epi, workspace, workspace_bytes,
                    as_stream(stream));
}

int tg_gc1_fwd_f32(const tg_plan* plan, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                   const float* S, int64_t lds, const float* bias, float* H1, int64_t ldh, int32_t n_feat, float p,
                   int32_t training, const uint8_t* keep_mask, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                   int64_t raw_row_begin, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace tg;
    TG_REQUIRE(H1, TG_ERR_INVALID_ARG, "null output");
    TG_REQUIRE(ldh >= n_feat, TG_ERR_INVALID_ARG, "ldh < n_feat");
    TG_REQUIRE(p >= 0.f && p <= 1.f, TG_ERR_INVALID_ARG, "dropout p must be in [0,1]");  // (p = 1 drops everything, like torch.dropout)
    EpiStore epi{};
    epi.Y = H1; epi.ldy = ldh; epi.bias = bias; epi.relu = 1; epi.n_feat = n_feat;
    epi.seed = seed; epi.offset = offset; epi.keep_mask = keep_mask;
    epi.raw_row_begin = raw_row_begin < 0 ? INT64_MAX : raw_row_begin;
    epi.out_scale = nullptr;
    epi.offset_dev = reinterpret_cast<const unsigned long long*>(offset_dev);
    const bool drop = training && p > 0.f;
    epi.drop_mode = !drop ? 0 : (keep_mask ? 2 : 1);
    epi.keep_thr = dropout_keep_threshold(p);
    epi.scale = drop ? (p < 1.f ? 1.f / (1.f - p) : 0.f) : 1.f;
    const bool ok4 = (ldh % 4 == 0) && aligned16(H1) && (!bias || aligned16(bias)) &&
                     (epi.drop_mode != 2 || (reinterpret_cast<uintptr_t>(keep_mask) & 3u) == 0);
    return run_spmm(plan, rowptr, colidx, vals, S, lds, n_feat, ok4, epi, workspace, workspace_bytes,
                    as_stream(stream));
}

int tg_dropout_keep_mask(uint8_t* keep_mask, int64_t n_rows, int32_t n_feat, float p, uint64_t seed,
                         uint64_t offset, void* stream) {
    using namespace tg;
    TG_REQUIRE(keep_mask && n_rows >= 0 && n_feat > 0, TG_ERR_INVALID_ARG, "bad argument");
    const int64_t n = n_rows * (int64_t)n_feat;
    if (n == 0) return TG_OK;
    keep_mask_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(
        keep_mask, n_rows, n_feat, dropout_keep_threshold(p), seed, offset);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

int tg_gc2_loss_fwd_f32(const tg_plan* plan, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                        const float* S2, int64_t lds, const float* bias, const int32_t* row_label, float inv_count,
                        float* logits, int64_t ldl, float* dZ2, int64_t ldd, float* row_loss, int32_t n_class,
                        void* workspace, size_t workspace_bytes, void* stream) {
    using namespace tg;
    TG_REQUIRE(row_label && row_loss, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE((!logits || ldl >= n_class) && (!dZ2 || ldd >= n_class), TG_ERR_INVALID_ARG, "leading dim < n_class");
    EpiLoss epi{};
    epi.bias = bias; epi.row_label = row_label; epi.inv_count = inv_count; epi.logits = logits; epi.ldl = ldl;
    epi.dZ = dZ2; epi.ldd = ldd; epi.row_loss = row_loss; epi.n_class = n_class;
    const bool ok4 = (!logits || ((ldl % 4 == 0) && aligned16(logits))) && (!dZ2 || ((ldd % 4 == 0) && aligned16(dZ2))) &&
                     (!bias || aligned16(bias));
    return run_spmm(plan, rowptr, colidx, vals, S2, lds, n_class, ok4, epi, workspace, workspace_bytes,
                    as_stream(stream));
}

int tg_masked_ce_f32(const float* logits, int64_t ldl, const int32_t* row_label, float inv_count, float* dZ,
                     int64_t ldd, float* row_loss, int64_t n_rows, int32_t n_class, void* stream) {
    using namespace tg;
    TG_REQUIRE(logits && row_label && row_loss, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n_rows >= 0 && n_class > 0 && ldl >= n_class && (!dZ || ldd >= n_class), TG_ERR_INVALID_ARG, "bad shape");
    EpiLoss epi{};
    epi.bias = nullptr; epi.row_label = row_label; epi.inv_count = inv_count; epi.logits = nullptr; epi.ldl = 0;
    epi.dZ = dZ; epi.ldd = ldd; epi.row_loss = row_loss; epi.n_class = n_class;
    const bool vec4 = (n_class % 4 == 0) && (ldl % 4 == 0) && aligned16(logits) && (!dZ || ((ldd % 4 == 0) && aligned16(dZ)));
    if (vec4) return rowwise_loss_vec<4>(logits, ldl, n_rows, n_class / 4, epi, as_stream(stream));
    return rowwise_loss_vec<1>(logits, ldl, n_rows, n_class, epi, as_stream(stream));
}

}  // extern "C"
