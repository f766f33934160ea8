// tg_stream.cu — column-chunk STREAMING SpMM for graphs with a compact hub set (document-topic-topic graphs).
//
// Same product as tg_spmm.cu (reference layer.py:106 and its autograd transpose product), different data flow.
// The gather formulation re-reads every row of B once per stored entry that references it: a topic row with
// 30 000 entries gathers 30 000 document rows, so the 1M-document graph moves 9.5 GB through HBM for 2.2 GB of
// algorithmic traffic (profiles/r01_v1_spmm_full.md).  Here B is streamed ONCE in chunks of T consecutive nodes
// through shared memory, and both uses of a chunk happen while it is resident:
//
//   (i)  short rows r of the chunk (documents):   Y[r,:]  = sum_j A[r,j] * B[j,:]
//        - hub columns j are served from a shared-memory copy of the hub rows of B (loaded once per CTA),
//        - columns inside the chunk (the self loop) from the staged chunk, anything else from L2/global;
//   (ii) hub rows k (topics):                    acc[k,:] += sum_{j in chunk} A[k,j] * B[j,:]
//        - the hub rows' entries are kept in a chunk-major copy (plan time) so that segment (chunk, k) is one
//          contiguous run that is staged next to the chunk; accumulators live in REGISTERS for the whole kernel.
//
// Feature columns are sliced (FT = 4*GW columns per CTA) so that the hub rows of B fit in shared memory; a CTA is
// persistent over a contiguous range of chunks; per-CTA hub partials are written once and added in CTA order by a
// small finishing kernel that also applies the epilogue.  Every floating-point addition has a fixed order: the
// result is bitwise reproducible (no float atomics).  Data movement global->shared uses cp.async double buffering.
#include <cub/cub.cuh>
#include <stdlib.h>

#include <vector>

#include "tg_stream.cuh"

namespace tg {

constexpr int kSThreads = 512;
constexpr int kGW = 8;                    // lanes per row group: a quarter warp reads one 128-byte line of a row slice
constexpr int kNG = kSThreads / kGW;      // row groups per CTA
constexpr int kMaxHubStream = 512;
constexpr size_t kSmemBudget = 227 * 1024;

struct StreamArgs {
    const int32_t* __restrict__ rowptr;
    const int32_t* __restrict__ rsplit;  // [n] first hub-column entry of every row (entries are reordered: others | hubs)
    const int2* __restrict__ dent;       // [nnz]     {hub slot or column id, val bits}, per row: non-hub columns first
    const int2* __restrict__ hent;       // [hub_nnz] {column local to chunk, val bits}, chunk-major / hub-minor
    const int32_t* __restrict__ htab;    // [n_chunks][Kh+1]
    const int4* __restrict__ cdesc;      // [n_chunks]
    const int32_t* __restrict__ hub_rows;
    const float* __restrict__ B;
    int64_t ldb;
    int64_t n;
    int64_t nnz, hub_nnz;
    int32_t n_chunks4;  // n_feat / 4
    int32_t T, n_chunks, Kh, hub_threshold;
    int32_t cap_doc, cap_hub;
    int32_t n_slices, n_groups;
    float* partials;
    int64_t ldp;
};

// ---- cp.async helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct StageView {
    float* Bs;       // [T][FT]
    int32_t* rp;     // [T+1]
    int32_t* rs;     // [T]
    int2* dent;      // [cap_doc]
    int2* hent;      // [cap_hub]
    int32_t* htab;   // [Kh+1]
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t stage_bytes(int T, int FT, int cap_doc, int cap_hub, int Kh) {
    return align16((size_t)T * FT * 4) + align16((size_t)(T + 1) * 4) + align16((size_t)T * 4) +
           align16((size_t)cap_doc * 8) + align16((size_t)cap_hub * 8) + align16((size_t)(Kh + 1) * 4);
}

__device__ __forceinline__ StageView stage_view(unsigned char* base, int T, int FT, int cap_doc, int cap_hub) {
    StageView v;
    v.Bs = reinterpret_cast<float*>(base);
    base += align16((size_t)T * FT * 4);
    v.rp = reinterpret_cast<int32_t*>(base);
    base += align16((size_t)(T + 1) * 4);
    v.rs = reinterpret_cast<int32_t*>(base);
    base += align16((size_t)T * 4);
    v.dent = reinterpret_cast<int2*>(base);
    base += align16((size_t)cap_doc * 8);
    v.hent = reinterpret_cast<int2*>(base);
    base += align16((size_t)cap_hub * 8);
    v.htab = reinterpret_cast<int32_t*>(base);
    return v;
}

// acc[i] += v * row[gl*4 + i*32 .. +4)   — `row` is a shared-memory row slice of FT = 32*CPL floats
template <int CPL>
__device__ __forceinline__ void fma_row_smem(Chunk<4> (&acc)[CPL], float v, const float* row, int gl) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const float4 b = *reinterpret_cast<const float4*>(row + gl * 4 + i * 32);
        acc[i].v[0] = fmaf(v, b.x, acc[i].v[0]);
        acc[i].v[1] = fmaf(v, b.y, acc[i].v[1]);
        acc[i].v[2] = fmaf(v, b.z, acc[i].v[2]);
        acc[i].v[3] = fmaf(v, b.w, acc[i].v[3]);
    }
}

// ---- the streaming kernel ----------------------------------------------------------------------------------------------
// CTA = 64 row groups of 8 lanes; a lane owns CPL float4 chunks (q0 + 8*i) of the CTA's 32*CPL-column slice.
template <int CPL, int KPG, class Epi>
__global__ void __launch_bounds__(kSThreads, 1) stream_spmm_kernel(const StreamArgs a, const Epi epi) {
    constexpr int FT = 32 * CPL;
    constexpr int QS = 8 * CPL;  // float4 chunks per slice
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int gl = lane & (kGW - 1);
    const int grp = tid / kGW;
    const unsigned gmask = group_mask<kGW>(lane);
    const int slice = blockIdx.x % a.n_slices;
    const int cg = blockIdx.x / a.n_slices;
    const int q0 = slice * QS + gl;  // this lane's first float4 column within the full row (the others are q0 + 8*i)
    const int c_begin = (int)((int64_t)cg * a.n_chunks / a.n_groups);
    const int c_end = (int)((int64_t)(cg + 1) * a.n_chunks / a.n_groups);

    float* BH = reinterpret_cast<float*>(smem_raw);
    const size_t bh_bytes = align16((size_t)a.Kh * FT * 4);
    const size_t st_bytes = stage_bytes(a.T, FT, a.cap_doc, a.cap_hub, a.Kh);
    auto stage_at = [&](int buf) {
        return stage_view(smem_raw + bh_bytes + (size_t)buf * st_bytes, a.T, FT, a.cap_doc, a.cap_hub);
    };

    auto issue_stage = [&](const StageView& sv, int c, const int4 d) {
        const int64_t c0 = (int64_t)c * a.T;
        // chunk rows of B, this CTA's column slice: QS 16-byte pieces per row
        for (int idx = tid; idx < a.T * QS; idx += kSThreads) {
            const int r = idx / QS, q = idx % QS;
            const int64_t row = c0 + r;
            const bool ok = row < a.n && (slice * QS + q) < a.n_chunks4;
            const float* src = ok ? a.B + row * a.ldb + (int64_t)(slice * QS + q) * 4 : a.B;
            cp_async16(sv.Bs + r * FT + q * 4, src, ok);
        }
        for (int idx = tid; idx <= a.T; idx += kSThreads) {
            const bool ok = c0 + idx <= a.n;
            cp_async4(sv.rp + idx, ok ? a.rowptr + c0 + idx : a.rowptr, ok);
        }
        for (int idx = tid; idx < a.T; idx += kSThreads) {
            const bool ok = c0 + idx < a.n;
            cp_async4(sv.rs + idx, ok ? a.rsplit + c0 + idx : a.rsplit, ok);
        }
        // entry windows start at an even entry so that every copy is one aligned 16-byte pair
        const int64_t d0 = d.x & ~1, h0 = d.z & ~1;
        for (int idx = tid; idx < a.cap_doc / 2; idx += kSThreads) {
            const int64_t p = d0 + 2 * idx;
            const bool ok = p < d.y;
            cp_async16(sv.dent + 2 * idx, ok ? a.dent + p : a.dent, ok);
        }
        for (int idx = tid; idx < a.cap_hub / 2; idx += kSThreads) {
            const int64_t p = h0 + 2 * idx;
            const bool ok = p < d.w;
            cp_async16(sv.hent + 2 * idx, ok ? a.hent + p : a.hent, ok);
        }
        for (int idx = tid; idx <= a.Kh; idx += kSThreads)
            cp_async4(sv.htab + idx, a.htab + (int64_t)c * (a.Kh + 1) + idx, true);
    };

    Chunk<4> hacc[KPG][CPL];
#pragma unroll
    for (int kk = 0; kk < KPG; ++kk)
#pragma unroll
        for (int i = 0; i < CPL; ++i) hacc[kk][i] = chunk_zero<4>();

    if (c_begin < c_end) {
        // prologue: hub rows of B (slice) + first stage
        for (int idx = tid; idx < a.Kh * QS; idx += kSThreads) {
            const int k = idx / QS, q = idx % QS;
            const bool ok = (slice * QS + q) < a.n_chunks4;
            const float* src = ok ? a.B + (int64_t)__ldg(a.hub_rows + k) * a.ldb + (int64_t)(slice * QS + q) * 4 : a.B;
            cp_async16(BH + k * FT + q * 4, src, ok);
        }
        int4 d_cur = __ldg(a.cdesc + c_begin);
        issue_stage(stage_at(0), c_begin, d_cur);
        cp_async_commit();
        int4 d_next = (c_begin + 1 < c_end) ? __ldg(a.cdesc + c_begin + 1) : make_int4(0, 0, 0, 0);

        for (int c = c_begin; c < c_end; ++c) {
            const int buf = (c - c_begin) & 1;
            if (c + 1 < c_end) {
                issue_stage(stage_at(buf ^ 1), c + 1, d_next);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const int4 d_new = (c + 2 < c_end) ? __ldg(a.cdesc + c + 2) : make_int4(0, 0, 0, 0);
            const StageView sv = stage_at(buf);
            const int c0 = c * a.T;
            const int dwin0 = d_cur.x & ~1, hwin0 = d_cur.z & ~1;
            const int dwin_end = dwin0 + a.cap_doc, hwin_end = hwin0 + a.cap_hub;

            // ---- (i) short rows of this chunk --------------------------------------------------------------------
            for (int lr = grp; lr < a.T; lr += kNG) {
                const int64_t row = (int64_t)c0 + lr;
                if (row >= a.n) break;
                const int s = sv.rp[lr], e = sv.rp[lr + 1];
                if (e - s > a.hub_threshold) continue;  // hub rows are produced by (ii)
                const int m = sv.rs[lr];
                Chunk<4> acc[CPL];
#pragma unroll
                for (int i = 0; i < CPL; ++i) acc[i] = chunk_zero<4>();
                if (e <= dwin_end) {
                    // fast path: every entry of the row sits in the staged window (group-uniform broadcast reads)
                    const int2* ent = sv.dent - dwin0;
                    for (int p = s; p < m; ++p) {  // non-hub columns (the self loop): staged chunk or L2
                        const int2 en = ent[p];
                        const float v = __int_as_float(en.y);
                        const unsigned lc = (unsigned)(en.x - c0);
                        if (lc < (unsigned)a.T) {
                            fma_row_smem<CPL>(acc, v, sv.Bs + lc * FT, gl);
                        } else {
#pragma unroll
                            for (int i = 0; i < CPL; ++i)
                                if (q0 + 8 * i < a.n_chunks4) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.B + (int64_t)en.x * a.ldb + (int64_t)(q0 + 8 * i) * 4));
                                    acc[i].v[0] = fmaf(v, b.x, acc[i].v[0]); acc[i].v[1] = fmaf(v, b.y, acc[i].v[1]);
                                    acc[i].v[2] = fmaf(v, b.z, acc[i].v[2]); acc[i].v[3] = fmaf(v, b.w, acc[i].v[3]);
                                }
                        }
                    }
#pragma unroll 4
                    for (int p = m; p < e; ++p) {  // hub columns: resident hub rows of B
                        const int2 en = ent[p];
                        fma_row_smem<CPL>(acc, __int_as_float(en.y), BH + en.x * FT, gl);
                    }
                } else {
                    // slow path (rows whose entries overflow the staged window): entries from L2
                    for (int p = s; p < e; ++p) {
                        const int2 en = (p < dwin_end) ? sv.dent[p - dwin0] : __ldg(a.dent + p);
                        const float v = __int_as_float(en.y);
                        if (p >= m) {
                            fma_row_smem<CPL>(acc, v, BH + en.x * FT, gl);
                        } else if ((unsigned)(en.x - c0) < (unsigned)a.T) {
                            fma_row_smem<CPL>(acc, v, sv.Bs + (en.x - c0) * FT, gl);
                        } else {
#pragma unroll
                            for (int i = 0; i < CPL; ++i)
                                if (q0 + 8 * i < a.n_chunks4) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.B + (int64_t)en.x * a.ldb + (int64_t)(q0 + 8 * i) * 4));
                                    acc[i].v[0] = fmaf(v, b.x, acc[i].v[0]); acc[i].v[1] = fmaf(v, b.y, acc[i].v[1]);
                                    acc[i].v[2] = fmaf(v, b.z, acc[i].v[2]); acc[i].v[3] = fmaf(v, b.w, acc[i].v[3]);
                                }
                        }
                    }
                }
                epi.template apply<4, kGW, CPL>(row, q0, gmask, a.n_chunks4, acc);
            }

            // ---- (ii) hub rows: accumulate the chunk's contribution in registers ----------------------------------
#pragma unroll
            for (int kk = 0; kk < KPG; ++kk) {
                const int k = grp + kNG * kk;
                if (k < a.Kh) {
                    const int h0 = sv.htab[k], h1 = sv.htab[k + 1];
                    if (h1 <= hwin_end) {
                        const int2* ent = sv.hent - hwin0;
#pragma unroll 4
                        for (int q = h0; q < h1; ++q) {
                            const int2 en = ent[q];
                            fma_row_smem<CPL>(hacc[kk], __int_as_float(en.y), sv.Bs + en.x * FT, gl);
                        }
                    } else {
                        for (int q = h0; q < h1; ++q) {
                            const int2 en = (q < hwin_end) ? sv.hent[q - hwin0] : __ldg(a.hent + q);
                            fma_row_smem<CPL>(hacc[kk], __int_as_float(en.y), sv.Bs + en.x * FT, gl);
                        }
                    }
                }
            }
            __syncthreads();  // everyone is done with this stage before it is refilled two iterations later
            d_cur = d_next;
            d_next = d_new;
        }
    }

    // per-CTA hub partials (zeros when the CTA had no chunk)
#pragma unroll
    for (int kk = 0; kk < KPG; ++kk) {
        const int k = grp + kNG * kk;
        if (k < a.Kh) {
#pragma unroll
            for (int i = 0; i < CPL; ++i)
                if (q0 + 8 * i < a.n_chunks4)
                    chunk_st<4>(a.partials + ((int64_t)cg * a.Kh + k) * a.ldp + (int64_t)(q0 + 8 * i) * 4, hacc[kk][i]);
        }
    }
}

// ---- finishing kernel: hub row k = sum over CTA groups (fixed order) + epilogue --------------------------------------
template <int VEC, int G, int CPL, class Epi>
__global__ void __launch_bounds__(256) stream_finish_kernel(const float* __restrict__ partials, int64_t ldp, int n_groups,
                                                            int Kh, const int32_t* __restrict__ hub_rows, int n_chunks,
                                                            const Epi epi) {
    constexpr int GPW = 32 / G;
    constexpr int GPB = 8 * GPW;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = group_mask<G>(lane);
    const int k = blockIdx.x * GPB + (threadIdx.x >> 5) * GPW + lane / G;
    if (k >= Kh) return;
    Chunk<VEC> acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = chunk_zero<VEC>();
    for (int g = 0; g < n_groups; ++g) {
        const float* src = partials + ((int64_t)g * Kh + k) * ldp;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int chunk = gl + i * G;
            if (chunk < n_chunks) {
                const Chunk<VEC> t = chunk_ldg<VEC>(src + (int64_t)chunk * VEC);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[i].v[v] += t.v[v];
            }
        }
    }
    epi.template apply<VEC, G, CPL>((int64_t)__ldg(hub_rows + k), gl, gmask, n_chunks, acc);
}

template <int VEC, int G, int CPL, class Epi>
static int launch_finish(const float* partials, int64_t ldp, int n_groups, int Kh, const int32_t* hub_rows, int n_chunks,
                         const Epi& epi, cudaStream_t st) {
    constexpr int GPB = 8 * (32 / G);
    stream_finish_kernel<VEC, G, CPL, Epi><<<(unsigned)ceil_div64(Kh, GPB), 256, 0, st>>>(partials, ldp, n_groups, Kh,
                                                                                          hub_rows, n_chunks, epi);
    TG_LAUNCH_CHECK();
    return TG_OK;
}

template <class Epi>
static int finish_dispatch(const float* partials, int64_t ldp, int n_groups, int Kh, const int32_t* hub_rows,
                           int n_chunks, const Epi& epi, cudaStream_t st) {
#define TG_LAUNCH_FIN(V, G, C) launch_finish<V, G, C>(partials, ldp, n_groups, Kh, hub_rows, n_chunks, epi, st)
    TG_SHAPE_SWITCH(4, n_chunks, TG_LAUNCH_FIN);
#undef TG_LAUNCH_FIN
    set_error("n_feat too wide for the streaming finish kernel");
    return TG_ERR_UNSUPPORTED;
}

// ---- launch ------------------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static size_t stream_smem_bytes(const tg_plan* pl, int CPL) {
    const int FT = 32 * CPL;
    return align16((size_t)pl->n_hub * FT * 4) + 2 * stage_bytes(pl->chunk_rows, FT, pl->cap_doc, pl->cap_hub, pl->n_hub);
}

// chunks per lane: 2 (64-column slices) when the row is wide enough and the hub rows fit, else 1 (32-column slices)
static int pick_cpl(const tg_plan* pl, int n_feat, bool whole_row) {
    const int pref = env_int("TG_STREAM_CPL", 2);
    for (int CPL = (n_feat > 32 && pref >= 2) ? 2 : 1; CPL >= 1; --CPL) {
        if (whole_row && n_feat > 32 * CPL) continue;
        if (stream_smem_bytes(pl, CPL) > kSmemBudget) continue;
        return CPL;
    }
    return 0;
}

bool stream_applicable(const tg_plan* pl, const StreamCall& c, bool out_vec4_ok, bool whole_row) {
    if (!pl || !pl->stream_ok) return false;
    if (!(c.n_feat % 4 == 0 && c.ldb % 4 == 0 && aligned16(c.B) && out_vec4_ok)) return false;
    if (c.n_feat > 1024) return false;
    return pick_cpl(pl, c.n_feat, whole_row) != 0;
}

size_t stream_workspace_bytes(const tg_plan* pl, int32_t n_feat) {
    if (!pl || !pl->stream_ok) return 0;
    const size_t ld = (size_t)((n_feat + 3) / 4) * 4;
    return (size_t)kNumSM * pl->n_hub * ld * sizeof(float) + 16;
}

template <int CPL, int KPG, class Epi>
static int launch_stream(const tg_plan* pl, const StreamCall& c, StreamArgs a, const Epi& epi, cudaStream_t st) {
    const size_t smem = stream_smem_bytes(pl, CPL);
    TG_CUDA(cudaFuncSetAttribute(stream_spmm_kernel<CPL, KPG, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a.n_slices = (int)ceil_div64(a.n_chunks4, 8 * CPL);
    int groups = kNumSM / a.n_slices;
    if (groups < 1) groups = 1;
    if (groups > a.n_chunks) groups = a.n_chunks;
    a.n_groups = groups;
    const size_t need = (size_t)groups * pl->n_hub * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    stream_spmm_kernel<CPL, KPG, Epi><<<(unsigned)(groups * a.n_slices), kSThreads, smem, st>>>(a, epi);
    TG_LAUNCH_CHECK();
    return finish_dispatch(a.partials, a.ldp, groups, pl->n_hub, pl->hub_rows, a.n_chunks4, epi, st);
}

template <class Epi>
static int run_stream(const tg_plan* pl, const StreamCall& c, const Epi& epi, bool whole_row, cudaStream_t st) {
    const int CPL = pick_cpl(pl, c.n_feat, whole_row);
    TG_REQUIRE(CPL != 0, TG_ERR_UNSUPPORTED, "streaming kernel not applicable");
    StreamArgs a;
    a.rowptr = c.rowptr; a.rsplit = pl->rsplit; a.dent = reinterpret_cast<const int2*>(pl->colidx2);
    a.hent = reinterpret_cast<const int2*>(pl->hcol); a.htab = pl->htab; a.cdesc = pl->cdesc;
    a.hub_rows = pl->hub_rows; a.B = c.B; a.ldb = c.ldb; a.n = pl->n_rows; a.nnz = pl->nnz; a.hub_nnz = pl->hub_nnz;
    a.n_chunks4 = c.n_feat / 4; a.T = pl->chunk_rows; a.n_chunks = pl->n_chunks; a.Kh = pl->n_hub;
    a.hub_threshold = pl->hub_threshold; a.cap_doc = pl->cap_doc; a.cap_hub = pl->cap_hub;
    a.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(c.workspace) + 15u) & ~(uintptr_t)15u);
    a.ldp = (int64_t)((c.n_feat + 3) / 4) * 4;
    a.n_slices = a.n_groups = 0;
    const int kpg = (int)ceil_div64(pl->n_hub, kNG);
#define TG_STREAM_CASE(CPLv)                                                        \
    if (CPL == CPLv) {                                                              \
        if (kpg <= 1) return launch_stream<CPLv, 1>(pl, c, a, epi, st);             \
        if (kpg <= 2) return launch_stream<CPLv, 2>(pl, c, a, epi, st);             \
        if (kpg <= 4) return launch_stream<CPLv, 4>(pl, c, a, epi, st);             \
        return launch_stream<CPLv, 8>(pl, c, a, epi, st);                           \
    }
    TG_STREAM_CASE(1)
    TG_STREAM_CASE(2)
#undef TG_STREAM_CASE
    set_error("unsupported chunks-per-lane %d", CPL);
    return TG_ERR_UNSUPPORTED;
}

int stream_spmm_store(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st) {
    return run_stream(pl, c, epi, false, st);
}
int stream_spmm_loss(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st) {
    return run_stream(pl, c, epi, true, st);
}

// ---- plan build ------------------------------------------------------------------------------------------------------------
// one thread per row: stable partition of the row's entries into (non-hub columns | hub columns); hub columns are
// rewritten to their hub slot.  Within each part the ascending column order of the CSR is kept, so a document row of a
// document-topic graph (self loop, then topics) is summed in exactly the reference's storage order.
__global__ void reorder_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    const float* __restrict__ vals, int64_t n, const int32_t* __restrict__ slot_of,
                                    int32_t hub_threshold, int2* __restrict__ dent, int32_t* __restrict__ rsplit) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = rowptr[r], e = rowptr[r + 1];
    if (e - s > hub_threshold) {  // hub rows are not read through this copy
        rsplit[r] = e;
        return;
    }
    int w = s;
    for (int p = s; p < e; ++p) {
        const int c = colidx[p];
        if (slot_of[c] < 0) dent[w++] = make_int2(c, __float_as_int(vals[p]));
    }
    rsplit[r] = w;
    for (int p = s; p < e; ++p) {
        const int c = colidx[p];
        const int sl = slot_of[c];
        if (sl >= 0) dent[w++] = make_int2(sl, __float_as_int(vals[p]));
    }
}

// one block per hub row: key = chunk * Kh + slot for each of its entries, in storage (column) order
__global__ void hub_keys_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const int32_t* __restrict__ hub_rows, const int64_t* __restrict__ hub_ofs, int Kh, int T,
                                uint32_t* __restrict__ keys, int32_t* __restrict__ src) {
    const int k = blockIdx.x;
    const int r = hub_rows[k];
    const int s = rowptr[r], e = rowptr[r + 1];
    const int64_t o = hub_ofs[k];
    for (int p = s + threadIdx.x; p < e; p += blockDim.x) {
        keys[o + (p - s)] = (uint32_t)(colidx[p] / T) * (uint32_t)Kh + (uint32_t)k;
        src[o + (p - s)] = p;
    }
}

__global__ void hub_gather_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ src,
                                  const int32_t* __restrict__ colidx, const float* __restrict__ vals, int64_t hub_nnz, int Kh,
                                  int T, int2* __restrict__ hent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hub_nnz) return;
    const int p = src[i];
    const int c = (int)(keys[i] / (uint32_t)Kh);
    hent[i] = make_int2(colidx[p] - c * T, __float_as_int(vals[p]));
}

// htab[c][k] = first sorted position with key >= c*Kh + k   (k = Kh gives the start of chunk c+1)
__global__ void hub_table_kernel(const uint32_t* __restrict__ keys, int64_t hub_nnz, int n_chunks, int Kh,
                                 int32_t* __restrict__ htab) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_chunks * (Kh + 1)) return;
    const int c = (int)(i / (Kh + 1)), k = (int)(i % (Kh + 1));
    const uint64_t want = (uint64_t)c * Kh + k;
    int64_t lo = 0, hi = hub_nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((uint64_t)keys[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    htab[i] = (int32_t)lo;
}

__global__ void chunk_desc_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ htab, int64_t n, int T,
                                  int n_chunks, int Kh, int4* __restrict__ cdesc) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const int64_t r0 = (int64_t)c * T;
    const int64_t r1 = (r0 + T < n) ? r0 + T : n;
    cdesc[c] = make_int4(rowptr[r0], rowptr[r1], htab[(int64_t)c * (Kh + 1)], htab[(int64_t)c * (Kh + 1) + Kh]);
}

void stream_plan_free(tg_plan* pl) {
    if (!pl) return;
    cudaFree(pl->colidx2); cudaFree(pl->hcol); cudaFree(pl->htab); cudaFree(pl->cdesc); cudaFree(pl->rsplit);
    pl->colidx2 = nullptr; pl->hcol = nullptr; pl->hval = nullptr; pl->htab = nullptr; pl->cdesc = nullptr;
    pl->rsplit = nullptr;
    pl->stream_ok = false;
}

int stream_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* h_rowptr, cudaStream_t st) {
    pl->stream_ok = false;
    if (env_int("TG_STREAM", 1) == 0) return TG_OK;
    if (!colidx || !vals || pl->n_rows != pl->n_cols || pl->n_hub < 1 || pl->n_hub > kMaxHubStream || pl->nnz == 0)
        return TG_OK;
    const int64_t n = pl->n_rows;
    const int Kh = pl->n_hub;
    int T = env_int("TG_STREAM_CHUNK", 128);
    if (T < 32 || T > 1024 || (T % 32) != 0) T = 128;
    const int n_chunks = (int)ceil_div64(n, T);
    if ((uint64_t)n_chunks * (uint64_t)Kh >= 0xFFFFFFFFull) return TG_OK;
    // only worth it when the hub rows carry a real share of the entries
    if (pl->hub_nnz * 8 < pl->nnz) return TG_OK;

    std::vector<int32_t> slot_of((size_t)n, -1), hub_rows;
    std::vector<int64_t> hub_ofs;
    int64_t run = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int32_t len = h_rowptr[(size_t)r + 1] - h_rowptr[(size_t)r];
        if (len > pl->hub_threshold) {
            slot_of[(size_t)r] = (int32_t)hub_rows.size();
            hub_rows.push_back((int32_t)r);
            hub_ofs.push_back(run);
            run += len;
        }
    }
    const int64_t hub_nnz = run;

    int32_t* d_slot = nullptr;
    int64_t* d_ofs = nullptr;
    uint32_t *keys_a = nullptr, *keys_b = nullptr;
    int32_t *src_a = nullptr, *src_b = nullptr;
    void* tmp = nullptr;
    cudaError_t e = cudaSuccess;
    auto fail = [&](cudaError_t err, const char* what) {
        cudaFree(d_slot); cudaFree(d_ofs); cudaFree(keys_a); cudaFree(keys_b); cudaFree(src_a); cudaFree(src_b); cudaFree(tmp);
        stream_plan_free(pl);
        return cuda_fail(err, what, __FILE__, __LINE__);
    };
#define TG_TRY(call) do { e = (call); if (e != cudaSuccess) return fail(e, #call); } while (0)
    TG_TRY(cudaMalloc((void**)&d_slot, (size_t)n * sizeof(int32_t)));
    TG_TRY(cudaMemcpyAsync(d_slot, slot_of.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TG_TRY(cudaMalloc((void**)&d_ofs, (size_t)Kh * sizeof(int64_t)));
    TG_TRY(cudaMemcpyAsync(d_ofs, hub_ofs.data(), (size_t)Kh * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    // +2 entries of padding so that the kernel's aligned 16-byte pair copies never leave the allocation
    TG_TRY(cudaMalloc((void**)&pl->colidx2, ((size_t)pl->nnz + 2) * sizeof(int2)));
    TG_TRY(cudaMemsetAsync(pl->colidx2, 0, ((size_t)pl->nnz + 2) * sizeof(int2), st));
    TG_TRY(cudaMalloc((void**)&pl->rsplit, (size_t)n * sizeof(int32_t)));
    reorder_rows_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(rowptr, colidx, vals, n, d_slot, pl->hub_threshold,
                                                                      reinterpret_cast<int2*>(pl->colidx2), pl->rsplit);
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaMalloc((void**)&keys_a, (size_t)hub_nnz * sizeof(uint32_t)));
    TG_TRY(cudaMalloc((void**)&keys_b, (size_t)hub_nnz * sizeof(uint32_t)));
    TG_TRY(cudaMalloc((void**)&src_a, (size_t)hub_nnz * sizeof(int32_t)));
    TG_TRY(cudaMalloc((void**)&src_b, (size_t)hub_nnz * sizeof(int32_t)));
    hub_keys_kernel<<<Kh, 256, 0, st>>>(rowptr, colidx, pl->hub_rows, d_ofs, Kh, T, keys_a, src_a);
    TG_TRY(cudaGetLastError());
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (uint64_t)n_chunks * Kh) ++end_bit;
    TG_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, src_a, src_b, (int)hub_nnz, 0, end_bit, st));
    TG_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    // stable: inside a (chunk, hub) segment the entries keep their column order
    TG_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, src_a, src_b, (int)hub_nnz, 0, end_bit, st));
    TG_TRY(cudaMalloc((void**)&pl->hcol, ((size_t)hub_nnz + 2) * sizeof(int2)));
    TG_TRY(cudaMemsetAsync(pl->hcol, 0, ((size_t)hub_nnz + 2) * sizeof(int2), st));
    hub_gather_kernel<<<(unsigned)ceil_div64(hub_nnz, 256), 256, 0, st>>>(keys_b, src_b, colidx, vals, hub_nnz, Kh, T,
                                                                         reinterpret_cast<int2*>(pl->hcol));
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaMalloc((void**)&pl->htab, (size_t)n_chunks * (Kh + 1) * sizeof(int32_t)));
    hub_table_kernel<<<(unsigned)ceil_div64((int64_t)n_chunks * (Kh + 1), 256), 256, 0, st>>>(keys_b, hub_nnz, n_chunks, Kh,
                                                                                              pl->htab);
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaMalloc((void**)&pl->cdesc, (size_t)n_chunks * sizeof(int4)));
    chunk_desc_kernel<<<(unsigned)ceil_div64(n_chunks, 256), 256, 0, st>>>(rowptr, pl->htab, n, T, n_chunks, Kh, pl->cdesc);
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaStreamSynchronize(st));
#undef TG_TRY
    cudaFree(d_slot); cudaFree(d_ofs); cudaFree(keys_a); cudaFree(keys_b); cudaFree(src_a); cudaFree(src_b); cudaFree(tmp);
    pl->chunk_rows = T;
    pl->n_chunks = n_chunks;
    // staged entry windows: average occupancy of a chunk plus slack (rows beyond the window take the L2 path)
    const int64_t avg_doc = ceil_div64(pl->nnz - hub_nnz, n_chunks), avg_hub = ceil_div64(hub_nnz, n_chunks);
    pl->cap_doc = (int32_t)(((avg_doc + avg_doc / 4 + 64) + 1) & ~(int64_t)1);
    pl->cap_hub = (int32_t)(((avg_hub + avg_hub / 4 + 64) + 1) & ~(int64_t)1);
    pl->stream_ok = true;
    // the kernel must fit at least the narrow configuration
    if (stream_smem_bytes(pl, 1) > kSmemBudget) stream_plan_free(pl);
    return TG_OK;
}

}  // namespace tg
