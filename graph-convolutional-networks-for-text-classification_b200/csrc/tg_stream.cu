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
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "tg_stream.cuh"
#include "tg_async.cuh"
#include "tg_finish.cuh"

namespace tg {

constexpr int kSThreads = 512;
constexpr int kGW = 8;                    // lanes per row group: a quarter warp reads one 128-byte line of a row slice
constexpr int kNG = kSThreads / kGW;      // row groups per CTA
constexpr int kMaxHubStream = 512;
constexpr int kHtabPad = 4;               // htab rows hold Kv + 1 offsets, padded to Kv + 4 ints (16-byte aligned rows)
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
    int32_t Kv;                          // virtual hub slots (table width): heavy hub rows are split into several
    const int32_t* __restrict__ vmap;    // [Kh][8] virtual slots of every hub row
    const int32_t* __restrict__ vcnt;    // [Kh]    how many
    int32_t cap_doc, cap_hub;
    int32_t n_slices, n_groups;
    float* partials;
    int64_t ldp;
};

struct StageView {
    float* Bs;       // [T][FT]
    int32_t* rp;     // [T+1]
    int32_t* rs;     // [T]
    int2* dent;      // [cap_doc]
    int2* hent;      // [cap_hub]
    int32_t* htab;   // [Kh+1]
};

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
    const size_t st_bytes = stage_bytes(a.T, FT, a.cap_doc, a.cap_hub, a.Kv);
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
        for (int idx = tid; idx <= a.Kv; idx += kSThreads)
            cp_async4(sv.htab + idx, a.htab + (int64_t)c * (a.Kv + kHtabPad) + idx, true);
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
                if (k < a.Kv) {
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
        if (k < a.Kv) {
#pragma unroll
            for (int i = 0; i < CPL; ++i)
                if (q0 + 8 * i < a.n_chunks4)
                    chunk_st<4>(a.partials + ((int64_t)cg * a.Kv + k) * a.ldp + (int64_t)(q0 + 8 * i) * 4, hacc[kk][i]);
        }
    }
}

// ---- role-specialised streaming kernel (wide rows) ----------------------------------------------------------------------
// The combined kernel above keeps BOTH the hub rows of B and the staged chunk in shared memory, which caps its column
// slice at 64 and its chunk at 128 nodes.  Splitting the two uses of B over two kinds of CTAs in ONE launch removes that:
//   hub CTAs  (64-column slices): stream 256-node chunks through shared memory (cp.async double buffer) and accumulate
//             acc[k] += sum_{j in chunk} A[k,j] * B[j]  in registers — no copy of the hub rows needed;
//   doc CTAs  (32*CPLD-column slices, CPLD = 4 -> 128 columns): keep the K hub rows of B resident in shared memory and
//             produce the short rows: hub columns from shared memory, the self loop straight from L2; no staging, no
//             block-level synchronisation after the prologue.
// Both kinds sweep the node range front to back in round-robin (chunk c -> hub CTA c mod n; row block j -> doc CTA
// j mod m), so the two fronts advance together and the second reader of a row of B hits L2: DRAM sees B once.
struct RolesArgs {
    StreamArgs s;
    int32_t hub_slices, hub_lanes;  // hub CTAs = hub_slices * hub_lanes (slice fastest)
    int32_t doc_slices, doc_lanes;  // doc CTAs = doc_slices * doc_lanes
    int32_t doc_job_rows;           // rows per doc-role job
    int32_t n_doc_jobs;
    int32_t only_role;              // debugging/profiling knob: 1 = hub CTAs only, 2 = document CTAs only, 0 = both
    int32_t use_tma;                // hub role: stage chunks with TMA bulk copies + mbarrier instead of cp.async
};

template <int THREADS, int CPLD, int KPG, class Epi>
__global__ void __launch_bounds__(THREADS, 1) stream_roles_kernel(const RolesArgs ra, const Epi epi,
                                                                  const __grid_constant__ CUtensorMap tmapB) {
    const StreamArgs& a = ra.s;
    constexpr int NGr = THREADS / kGW;  // row groups per CTA
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t mbar[2];
    // TMA destinations must be 128-byte aligned: align the dynamic base by hand (the launch requests 128 spare bytes)
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int gl = lane & (kGW - 1);
    const int grp = tid / kGW;
    const unsigned gmask = group_mask<kGW>(lane);
    const int n_hub_ctas = ra.hub_slices * ra.hub_lanes;

    if (ra.only_role == 2 && (int)blockIdx.x < n_hub_ctas) return;
    if (ra.only_role == 1 && (int)blockIdx.x >= n_hub_ctas) return;
    if ((int)blockIdx.x < n_hub_ctas) {
        // ================================ hub role ================================
        constexpr int CPL = 2, FT = 64, QS = 16;
        const int slice = blockIdx.x % ra.hub_slices;
        const int hl = blockIdx.x / ra.hub_slices;
        const int q0 = slice * QS + gl;
        const size_t bs_bytes = align128((size_t)a.T * FT * 4), he_bytes = align128((size_t)a.cap_hub * 8);
        const size_t st_bytes = bs_bytes + he_bytes + align128((size_t)(a.Kv + kHtabPad) * 4);
        auto Bs_at = [&](int buf) { return reinterpret_cast<float*>(smem_raw + (size_t)buf * st_bytes); };
        auto he_at = [&](int buf) { return reinterpret_cast<int2*>(smem_raw + (size_t)buf * st_bytes + bs_bytes); };
        auto ht_at = [&](int buf) { return reinterpret_cast<int32_t*>(smem_raw + (size_t)buf * st_bytes + bs_bytes + he_bytes); };
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        // One thread moves a whole stage: a 2-D TMA tile of B (T rows x 64 columns, zero filled past the matrix), the
        // chunk's contiguous run of hub entries and its offset-table row; the stage's mbarrier counts the bytes.
        auto issue = [&](int buf, int c, const int4 d) {
            if (tid != 0) return;
            const int h0 = d.z & ~1;
            int n_e = d.w - h0;
            n_e = n_e < 0 ? 0 : (n_e > a.cap_hub ? a.cap_hub : n_e);
            n_e = (n_e + 1) & ~1;
            const unsigned ht_b = (unsigned)((a.Kv + kHtabPad) * 4);
            fence_proxy_async();  // the buffer was last read through the generic proxy (ordered by the block barrier)
            mbar_expect_tx(&mbar[buf], (unsigned)(a.T * FT * 4) + (unsigned)n_e * 8u + ht_b);
            tma_load_2d(Bs_at(buf), &tmapB, slice * FT, c * a.T, &mbar[buf]);
            if (n_e) bulk_load_1d(he_at(buf), a.hent + h0, (unsigned)n_e * 8u, &mbar[buf]);
            bulk_load_1d(ht_at(buf), a.htab + (int64_t)c * (a.Kv + kHtabPad), ht_b, &mbar[buf]);
        };
        Chunk<4> hacc[KPG][CPL];
#pragma unroll
        for (int kk = 0; kk < KPG; ++kk)
#pragma unroll
            for (int i = 0; i < CPL; ++i) hacc[kk][i] = chunk_zero<4>();
        int c = hl;
        if (c < a.n_chunks) {
            int4 d_cur = __ldg(a.cdesc + c);
            issue(0, c, d_cur);
            int4 d_next = (c + ra.hub_lanes < a.n_chunks) ? __ldg(a.cdesc + c + ra.hub_lanes) : make_int4(0, 0, 0, 0);
            for (int it = 0; c < a.n_chunks; c += ra.hub_lanes, ++it) {
                const int buf = it & 1;
                const int cn = c + ra.hub_lanes;
                if (cn < a.n_chunks) issue(buf ^ 1, cn, d_next);
                mbar_wait(&mbar[buf], (unsigned)(it >> 1) & 1u);  // every thread observes the stage's bytes itself
                const int4 d_new = (cn + ra.hub_lanes < a.n_chunks) ? __ldg(a.cdesc + cn + ra.hub_lanes) : make_int4(0, 0, 0, 0);
                const float* Bs = Bs_at(buf);
                const int2* he = he_at(buf);
                const int32_t* ht = ht_at(buf);
                const int hwin0 = d_cur.z & ~1;
                const int hwin_end = hwin0 + a.cap_hub;
#pragma unroll
                for (int kk = 0; kk < KPG; ++kk) {
                    const int k = grp + NGr * kk;
                    if (k < a.Kv) {
                        const int h0 = ht[k], h1 = ht[k + 1];
                        if (h1 <= hwin_end) {
                            const int2* ent = he - hwin0;
#pragma unroll 4
                            for (int q = h0; q < h1; ++q) {
                                const int2 en = ent[q];
                                fma_row_smem<CPL>(hacc[kk], __int_as_float(en.y), Bs + en.x * FT, gl);
                            }
                        } else {
                            for (int q = h0; q < h1; ++q) {
                                const int2 en = (q < hwin_end) ? he[q - hwin0] : __ldg(a.hent + q);
                                fma_row_smem<CPL>(hacc[kk], __int_as_float(en.y), Bs + en.x * FT, gl);
                            }
                        }
                    }
                }
                __syncthreads();
                d_cur = d_next;
                d_next = d_new;
            }
        }
#pragma unroll
        for (int kk = 0; kk < KPG; ++kk) {
            const int k = grp + NGr * kk;
            if (k < a.Kv) {
#pragma unroll
                for (int i = 0; i < CPL; ++i)
                    if (q0 + 8 * i < a.n_chunks4)
                        chunk_st<4>(a.partials + ((int64_t)hl * a.Kv + k) * a.ldp + (int64_t)(q0 + 8 * i) * 4, hacc[kk][i]);
            }
        }
        return;
    }

    // ================================ document role ================================
    constexpr int FT = 32 * CPLD, QS = 8 * CPLD;
    const int bid = blockIdx.x - n_hub_ctas;
    const int slice = bid % ra.doc_slices;
    const int dl = bid / ra.doc_slices;
    const int q0 = slice * QS + gl;
    float* BH = reinterpret_cast<float*>(smem_raw);
    for (int idx = tid; idx < a.Kh * QS; idx += THREADS) {
        const int k = idx / QS, q = idx % QS;
        const bool ok = (slice * QS + q) < a.n_chunks4;
        cp_async16(BH + k * FT + q * 4, ok ? a.B + (int64_t)__ldg(a.hub_rows + k) * a.ldb + (int64_t)(slice * QS + q) * 4 : a.B, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    constexpr int RPG = (THREADS >= 1024) ? 2 : 4;  // rows per group per job
    for (int job = dl; job < ra.n_doc_jobs; job += ra.doc_lanes) {
        const int64_t r_base = (int64_t)job * ra.doc_job_rows;
        // row pointers of this group's rows first: they head the dependent chain rowptr -> entries -> FMA
        int rs_[RPG], re_[RPG], rm_[RPG];
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
            const int64_t row = r_base + grp + (int64_t)NGr * i;
            const bool ok = (grp + NGr * i) < ra.doc_job_rows && row < a.n;
            rs_[i] = ok ? __ldg(a.rowptr + row) : 0;
            re_[i] = ok ? __ldg(a.rowptr + row + 1) : 0;
            rm_[i] = ok ? __ldg(a.rsplit + row) : 0;
            if (re_[i] - rs_[i] > a.hub_threshold) re_[i] = rs_[i];  // hub rows belong to the hub role
        }
        int2 first[RPG];
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
            const int p = rm_[i] + gl;
            first[i] = (p < re_[i]) ? __ldg(a.dent + p) : make_int2(0, 0);
        }
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
            const int s = rs_[i], e = re_[i], m = rm_[i];
            if (e <= s) continue;
            const int64_t row = r_base + grp + (int64_t)NGr * i;
            Chunk<4> acc[CPLD];
#pragma unroll
            for (int u = 0; u < CPLD; ++u) acc[u] = chunk_zero<4>();
            for (int p = s; p < m; ++p) {  // non-hub columns (the self loop): L2
                const int2 en = __ldg(a.dent + p);
                const float v = __int_as_float(en.y);
                const float* src = a.B + (int64_t)en.x * a.ldb;
#pragma unroll
                for (int u = 0; u < CPLD; ++u)
                    if (q0 + 8 * u < a.n_chunks4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(q0 + 8 * u) * 4));
                        acc[u].v[0] = fmaf(v, b.x, acc[u].v[0]); acc[u].v[1] = fmaf(v, b.y, acc[u].v[1]);
                        acc[u].v[2] = fmaf(v, b.z, acc[u].v[2]); acc[u].v[3] = fmaf(v, b.w, acc[u].v[3]);
                    }
            }
            int2 my = first[i];
            for (int base = m; base < e; base += kGW) {
                if (base != m) {
                    const int p = base + gl;
                    my = (p < e) ? __ldg(a.dent + p) : make_int2(0, 0);
                }
                const int cnt = min(kGW, e - base);
#pragma unroll 4
                for (int j = 0; j < cnt; ++j) {
                    const int slot = __shfl_sync(gmask, my.x, j, kGW);
                    const float v = __int_as_float(__shfl_sync(gmask, my.y, j, kGW));
                    fma_row_smem<CPLD>(acc, v, BH + slot * FT, gl);
                }
            }
            epi.template apply<4, kGW, CPLD>(row, q0, gmask, a.n_chunks4, acc);
        }
    }
}

// ---- narrow rows (F <= 32): one LANE per row ------------------------------------------------------------------------
// With 8-20 classes a row is only 2-5 float4 chunks: the lane-group layout above would spend most of its instructions on
// index handling.  Here every lane owns a whole output row (document side) and a whole hub row (hub side): NV float4
// accumulators in registers, entries read at lane-private shared-memory addresses, row epilogue entirely inside the lane
// (the group epilogues are instantiated with a group of ONE lane).  Same staging, same partial/finish scheme, same
// fixed summation order as the wide kernel.
constexpr int kNThreads = 256;

template <int NV>
__device__ __forceinline__ void fma_row_lane(Chunk<4> (&acc)[NV], float v, const float* row) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 b = *reinterpret_cast<const float4*>(row + i * 4);
        acc[i].v[0] = fmaf(v, b.x, acc[i].v[0]);
        acc[i].v[1] = fmaf(v, b.y, acc[i].v[1]);
        acc[i].v[2] = fmaf(v, b.z, acc[i].v[2]);
        acc[i].v[3] = fmaf(v, b.w, acc[i].v[3]);
    }
}

template <int NV, int KPT, class Epi>
__global__ void __launch_bounds__(kNThreads) stream_narrow_kernel(const StreamArgs a, const Epi epi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int FT = a.n_chunks4 * 4;  // == n_feat (one slice)
    const int cg = blockIdx.x;
    const int c_begin = (int)((int64_t)cg * a.n_chunks / a.n_groups);
    const int c_end = (int)((int64_t)(cg + 1) * a.n_chunks / a.n_groups);
    const unsigned lmask = 1u << (tid & 31);

    float* BH = reinterpret_cast<float*>(smem_raw);
    const size_t bh_bytes = align16((size_t)a.Kh * FT * 4);
    const size_t st_bytes = stage_bytes(a.T, FT, a.cap_doc, a.cap_hub, a.Kv);
    auto stage_at = [&](int buf) {
        return stage_view(smem_raw + bh_bytes + (size_t)buf * st_bytes, a.T, FT, a.cap_doc, a.cap_hub);
    };
    auto issue_stage = [&](const StageView& sv, int c, const int4 d) {
        const int64_t c0 = (int64_t)c * a.T;
        for (int idx = tid; idx < a.T * NV; idx += kNThreads) {
            const int r = idx / NV, q = idx % NV;
            const int64_t row = c0 + r;
            const bool ok = row < a.n && q < a.n_chunks4;
            cp_async16(sv.Bs + r * FT + q * 4, ok ? a.B + row * a.ldb + q * 4 : a.B, ok);
        }
        for (int idx = tid; idx <= a.T; idx += kNThreads) {
            const bool ok = c0 + idx <= a.n;
            cp_async4(sv.rp + idx, ok ? a.rowptr + c0 + idx : a.rowptr, ok);
        }
        for (int idx = tid; idx < a.T; idx += kNThreads) {
            const bool ok = c0 + idx < a.n;
            cp_async4(sv.rs + idx, ok ? a.rsplit + c0 + idx : a.rsplit, ok);
        }
        const int64_t d0 = d.x & ~1, h0 = d.z & ~1;
        for (int idx = tid; idx < a.cap_doc / 2; idx += kNThreads) {
            const int64_t p = d0 + 2 * idx;
            const bool ok = p < d.y;
            cp_async16(sv.dent + 2 * idx, ok ? a.dent + p : a.dent, ok);
        }
        for (int idx = tid; idx < a.cap_hub / 2; idx += kNThreads) {
            const int64_t p = h0 + 2 * idx;
            const bool ok = p < d.w;
            cp_async16(sv.hent + 2 * idx, ok ? a.hent + p : a.hent, ok);
        }
        for (int idx = tid; idx <= a.Kv; idx += kNThreads)
            cp_async4(sv.htab + idx, a.htab + (int64_t)c * (a.Kv + kHtabPad) + idx, true);
    };

    Chunk<4> hacc[KPT][NV];
#pragma unroll
    for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
        for (int i = 0; i < NV; ++i) hacc[kk][i] = chunk_zero<4>();

    if (c_begin < c_end) {
        for (int idx = tid; idx < a.Kh * NV; idx += kNThreads) {
            const int k = idx / NV, q = idx % NV;
            const bool ok = q < a.n_chunks4;
            cp_async16(BH + k * FT + q * 4, ok ? a.B + (int64_t)__ldg(a.hub_rows + k) * a.ldb + q * 4 : a.B, ok);
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

            // ---- (i) short rows: one lane per row -------------------------------------------------------------------
            for (int lr = tid; lr < a.T; lr += kNThreads) {
                const int64_t row = (int64_t)c0 + lr;
                if (row >= a.n) break;
                const int s = sv.rp[lr], e = sv.rp[lr + 1];
                if (e - s > a.hub_threshold) continue;
                const int m = sv.rs[lr];
                Chunk<4> acc[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] = chunk_zero<4>();
                for (int p = s; p < e; ++p) {
                    const int2 en = (p < dwin_end) ? sv.dent[p - dwin0] : __ldg(a.dent + p);
                    const float v = __int_as_float(en.y);
                    if (p >= m) {
                        fma_row_lane<NV>(acc, v, BH + en.x * FT);
                    } else if ((unsigned)(en.x - c0) < (unsigned)a.T) {
                        fma_row_lane<NV>(acc, v, sv.Bs + (en.x - c0) * FT);
                    } else {
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(a.B + (int64_t)en.x * a.ldb + i * 4));
                            acc[i].v[0] = fmaf(v, b.x, acc[i].v[0]); acc[i].v[1] = fmaf(v, b.y, acc[i].v[1]);
                            acc[i].v[2] = fmaf(v, b.z, acc[i].v[2]); acc[i].v[3] = fmaf(v, b.w, acc[i].v[3]);
                        }
                    }
                }
                epi.template apply<4, 1, NV>(row, 0, lmask, a.n_chunks4, acc);
            }

            // ---- (ii) hub rows: one lane per hub row ----------------------------------------------------------------
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
                const int k = tid + kNThreads * kk;
                if (k < a.Kv) {
                    const int h0 = sv.htab[k], h1 = sv.htab[k + 1];
                    for (int q = h0; q < h1; ++q) {
                        const int2 en = (q < hwin_end) ? sv.hent[q - hwin0] : __ldg(a.hent + q);
                        fma_row_lane<NV>(hacc[kk], __int_as_float(en.y), sv.Bs + en.x * FT);
                    }
                }
            }
            __syncthreads();
            d_cur = d_next;
            d_next = d_new;
        }
    }
#pragma unroll
    for (int kk = 0; kk < KPT; ++kk) {
        const int k = tid + kNThreads * kk;
        if (k < a.Kv) {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i < a.n_chunks4) chunk_st<4>(a.partials + ((int64_t)cg * a.Kv + k) * a.ldp + i * 4, hacc[kk][i]);
        }
    }
}

template <class Epi>
static int finish_dispatch(const StreamArgs& a, const float* partials, int n_groups, const Epi& epi, cudaStream_t st) {
    FinishArgs f{partials, a.ldp, n_groups, a.Kh, a.Kv, a.vmap, a.vcnt, a.hub_rows, a.n_chunks4};
    return finish_run(f, epi, st);
}

// ---- launch ------------------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static size_t stream_smem_bytes(const tg_plan* pl, int CPL) {
    const int FT = 32 * CPL;
    return align16((size_t)pl->n_hub * FT * 4) + 2 * stage_bytes(pl->chunk_rows, FT, pl->cap_doc, pl->cap_hub, pl->n_vslot);
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

// ---- role-specialised launch -------------------------------------------------------------------------------------------
static inline bool epi_runs_rng(const EpiStore& e) { return e.drop_mode == 1; }
static inline bool epi_runs_rng(const EpiLoss&) { return false; }

static size_t roles_hub_smem(const tg_plan* pl) {
    return 2 * (align128((size_t)pl->chunk_rows * 64 * 4) + align128((size_t)pl->cap_hub * 8) +
                align128((size_t)(pl->n_vslot + kHtabPad) * 4));
}

static size_t roles_doc_smem(const tg_plan* pl, int CPLD) { return align16((size_t)pl->n_hub * 32 * CPLD * 4); }

static int roles_cpld(const tg_plan* pl, int n_feat) {
    for (int CPLD = 4; CPLD >= 2; CPLD -= 2) {
        if (n_feat < 32 * CPLD && CPLD > 2) continue;
        if (roles_doc_smem(pl, CPLD) + 128 <= kSmemBudget) return CPLD;
    }
    return 0;
}

static bool roles_applicable(const tg_plan* pl, int n_feat) {
    if (env_int("TG_STREAM_ROLES", 1) == 0) return false;
    if (n_feat < 64 || n_feat % 4 != 0) return false;
    if (!encode_tiled_fn() || pl->chunk_rows > 256) return false;  // the hub role stages B with TMA tiles
    if (roles_hub_smem(pl) + 128 > kSmemBudget || pl->n_vslot > kNG * 8) return false;
    return roles_cpld(pl, n_feat) != 0;
}

template <int THREADS, int CPLD, int KPG, class Epi>
static int launch_roles(const tg_plan* pl, const StreamCall& c, RolesArgs ra, const Epi& epi, cudaStream_t st) {
    size_t smem = roles_hub_smem(pl);
    if (roles_doc_smem(pl, CPLD) > smem) smem = roles_doc_smem(pl, CPLD);
    smem += 128;  // slack for the in-kernel 128-byte alignment of the dynamic base
    TG_CUDA(cudaFuncSetAttribute(stream_roles_kernel<THREADS, CPLD, KPG, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StreamArgs& a = ra.s;
    ra.hub_slices = (int)ceil_div64(a.n_chunks4, 16);
    ra.doc_slices = (int)ceil_div64(a.n_chunks4, 8 * CPLD);
    // share of the SMs given to the hub role: measured balance point on the 1M x 256 graph is 18 hub lanes x 4 slices
    // (72 SMs) without dropout and 16 (64 SMs) when the document role also runs the Philox dropout epilogue
    const int hub_pct = env_int("TG_ROLES_HUB_PCT", epi_runs_rng(epi) ? 45 : 50);
    int hub_lanes = (kNumSM * hub_pct / 100) / ra.hub_slices;
    if (hub_lanes < 1) hub_lanes = 1;
    int doc_lanes = (kNumSM - hub_lanes * ra.hub_slices) / ra.doc_slices;
    if (doc_lanes < 1) doc_lanes = 1;
    if (hub_lanes > a.n_chunks) hub_lanes = a.n_chunks;
    ra.hub_lanes = hub_lanes;
    ra.doc_lanes = doc_lanes;
    ra.doc_job_rows = 256;
    ra.only_role = env_int("TG_ROLES_ONLY", 0);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    ra.use_tma = 1;
    TG_REQUIRE(make_tensor_map(&tmap, a.B, a.n, c.n_feat, a.ldb, a.T, 64), TG_ERR_UNSUPPORTED,
               "cuTensorMapEncodeTiled failed (TMA tile of the dense operand)");
    ra.n_doc_jobs = (int)ceil_div64(a.n, ra.doc_job_rows);
    a.n_groups = hub_lanes;
    const size_t need = (size_t)hub_lanes * pl->n_vslot * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    const unsigned grid = (unsigned)(ra.hub_slices * hub_lanes + ra.doc_slices * doc_lanes);
    stream_roles_kernel<THREADS, CPLD, KPG, Epi><<<grid, THREADS, smem, st>>>(ra, epi, tmap);
    TG_LAUNCH_CHECK();
    return finish_dispatch(a, a.partials, hub_lanes, epi, st);
}

template <class Epi>
static int run_roles(const tg_plan* pl, const StreamCall& c, const Epi& epi, cudaStream_t st) {
    RolesArgs ra;
    StreamArgs& a = ra.s;
    a.rowptr = c.rowptr; a.rsplit = pl->rsplit; a.dent = reinterpret_cast<const int2*>(pl->colidx2);
    a.hent = reinterpret_cast<const int2*>(pl->hcol); a.htab = pl->htab; a.cdesc = pl->cdesc;
    a.hub_rows = pl->hub_rows; a.B = c.B; a.ldb = c.ldb; a.n = pl->n_rows; a.nnz = pl->nnz; a.hub_nnz = pl->hub_nnz;
    a.n_chunks4 = c.n_feat / 4; a.T = pl->chunk_rows; a.n_chunks = pl->n_chunks; a.Kh = pl->n_hub;
    a.Kv = pl->n_vslot; a.vmap = pl->vmap; a.vcnt = pl->vcnt;
    a.hub_threshold = pl->hub_threshold; a.cap_doc = pl->cap_doc; a.cap_hub = pl->cap_hub;
    a.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(c.workspace) + 15u) & ~(uintptr_t)15u);
    a.ldp = (int64_t)((c.n_feat + 3) / 4) * 4;
    a.n_slices = a.n_groups = 0;
    const int CPLD = roles_cpld(pl, c.n_feat);
    // 1024-thread CTAs (64 registers per thread, 32 warps per SM) hide the shared-memory latency better; they need the
    // virtual slots to fit 128 groups x 4
    const bool big = env_int("TG_ROLES_THREADS", 512) >= 1024 && pl->n_vslot <= 128 * 4;
    const int kpg = (int)ceil_div64(pl->n_vslot, big ? 128 : kNG);
#define TG_ROLES_CASE(Tv, Cv)                                                          \
    if (CPLD == Cv && big == (Tv == 1024)) {                                           \
        if (kpg <= 1) return launch_roles<Tv, Cv, 1>(pl, c, ra, epi, st);              \
        if (kpg <= 2) return launch_roles<Tv, Cv, 2>(pl, c, ra, epi, st);              \
        if (kpg <= 4) return launch_roles<Tv, Cv, 4>(pl, c, ra, epi, st);              \
        if (Tv == 512) return launch_roles<512, Cv, 8>(pl, c, ra, epi, st);            \
    }
    TG_ROLES_CASE(512, 2)
    TG_ROLES_CASE(512, 4)
    TG_ROLES_CASE(1024, 2)
    TG_ROLES_CASE(1024, 4)
#undef TG_ROLES_CASE
    set_error("no role-specialised configuration fits");
    return TG_ERR_UNSUPPORTED;
}

static size_t narrow_smem_bytes(const tg_plan* pl, int n_feat) {
    return align16((size_t)pl->n_hub * n_feat * 4) + 2 * stage_bytes(pl->chunk_rows, n_feat, pl->cap_doc, pl->cap_hub, pl->n_vslot);
}

static bool narrow_applicable(const tg_plan* pl, int n_feat) {
    return env_int("TG_STREAM_NARROW", 0) != 0 && n_feat <= 32 && n_feat % 4 == 0 && pl->n_vslot <= 2 * kNThreads &&
           narrow_smem_bytes(pl, n_feat) <= kSmemBudget;
}

bool stream_applicable(const tg_plan* pl, const StreamCall& c, bool out_vec4_ok, bool whole_row) {
    if (!pl || !pl->stream_ok) return false;
    if (!(c.n_feat % 4 == 0 && c.ldb % 4 == 0 && aligned16(c.B) && out_vec4_ok)) return false;
    if (c.n_feat > 1024) return false;
    if (narrow_applicable(pl, c.n_feat)) return true;
    if (roles2_narrow_applicable(pl, c)) return true;
    // rows of <= 32 columns: at the single-GPU sizes B (N x F x 4 B, 80 MB at C3) stays in L2 and the gather kernel
    // (tg_spmm.cu, 8 lanes per row) is faster than streaming; TG_STREAM_NARROW=1 selects the lane-per-row streaming
    // kernel, which keeps DRAM traffic at the algorithmic minimum when B outgrows L2.
    if (c.n_feat <= 32 && env_int("TG_STREAM_NARROW", 0) == 0) return false;
    return pick_cpl(pl, c.n_feat, whole_row) != 0;
}

size_t stream_workspace_bytes(const tg_plan* pl, int32_t n_feat) {
    if (!pl || !pl->stream_ok) return 0;
    const size_t ld = (size_t)((n_feat + 3) / 4) * 4;
    const size_t ws = (size_t)kNumSM * (n_feat <= 32 ? 4 : 1) * pl->n_vslot * ld * sizeof(float) + 16;
    return std::max(ws, roles2_workspace_bytes(pl, n_feat));
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
    const size_t need = (size_t)groups * pl->n_vslot * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    stream_spmm_kernel<CPL, KPG, Epi><<<(unsigned)(groups * a.n_slices), kSThreads, smem, st>>>(a, epi);
    TG_LAUNCH_CHECK();
    return finish_dispatch(a, a.partials, groups, epi, st);
}

template <int NV, int KPT, class Epi>
static int launch_narrow(const tg_plan* pl, const StreamCall& c, StreamArgs a, const Epi& epi, cudaStream_t st) {
    const size_t smem = narrow_smem_bytes(pl, c.n_feat);
    TG_CUDA(cudaFuncSetAttribute(stream_narrow_kernel<NV, KPT, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a.n_slices = 1;
    int per_sm = (int)(kSmemBudget / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int groups = kNumSM * per_sm;
    if (groups > a.n_chunks) groups = a.n_chunks;
    a.n_groups = groups;
    const size_t need = (size_t)groups * pl->n_vslot * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    stream_narrow_kernel<NV, KPT, Epi><<<(unsigned)groups, kNThreads, smem, st>>>(a, epi);
    TG_LAUNCH_CHECK();
    return finish_dispatch(a, a.partials, groups, epi, st);
}

template <class Epi>
static int run_stream(const tg_plan* pl, const StreamCall& c, const Epi& epi, bool whole_row, cudaStream_t st) {
    if (!narrow_applicable(pl, c.n_feat) && roles2_narrow_applicable(pl, c)) return roles2_narrow_run(pl, c, epi, st);
    if (narrow_applicable(pl, c.n_feat)) {
        StreamArgs a;
        a.rowptr = c.rowptr; a.rsplit = pl->rsplit; a.dent = reinterpret_cast<const int2*>(pl->colidx2);
        a.hent = reinterpret_cast<const int2*>(pl->hcol); a.htab = pl->htab; a.cdesc = pl->cdesc;
        a.hub_rows = pl->hub_rows; a.B = c.B; a.ldb = c.ldb; a.n = pl->n_rows; a.nnz = pl->nnz; a.hub_nnz = pl->hub_nnz;
        a.n_chunks4 = c.n_feat / 4; a.T = pl->chunk_rows; a.n_chunks = pl->n_chunks; a.Kh = pl->n_hub;
    a.Kv = pl->n_vslot; a.vmap = pl->vmap; a.vcnt = pl->vcnt;
        a.hub_threshold = pl->hub_threshold; a.cap_doc = pl->cap_doc; a.cap_hub = pl->cap_hub;
        a.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(c.workspace) + 15u) & ~(uintptr_t)15u);
        a.ldp = (int64_t)((c.n_feat + 3) / 4) * 4;
        a.n_slices = a.n_groups = 0;
        const bool two = pl->n_vslot > kNThreads;
#define TG_NARROW_CASE(NVv)                                                                   \
        if (a.n_chunks4 == NVv) return two ? launch_narrow<NVv, 2>(pl, c, a, epi, st) : launch_narrow<NVv, 1>(pl, c, a, epi, st);
        TG_NARROW_CASE(1) TG_NARROW_CASE(2) TG_NARROW_CASE(3) TG_NARROW_CASE(4)
        TG_NARROW_CASE(5) TG_NARROW_CASE(6) TG_NARROW_CASE(7) TG_NARROW_CASE(8)
#undef TG_NARROW_CASE
    }
    if constexpr (std::is_same<Epi, EpiStore>::value) {
        if (!whole_row && roles2_applicable(pl, c)) return roles2_run(pl, c, epi, st);
    }
    if (!whole_row && roles_applicable(pl, c.n_feat)) return run_roles(pl, c, epi, st);
    const int CPL = pick_cpl(pl, c.n_feat, whole_row);
    TG_REQUIRE(CPL != 0, TG_ERR_UNSUPPORTED, "streaming kernel not applicable");
    StreamArgs a;
    a.rowptr = c.rowptr; a.rsplit = pl->rsplit; a.dent = reinterpret_cast<const int2*>(pl->colidx2);
    a.hent = reinterpret_cast<const int2*>(pl->hcol); a.htab = pl->htab; a.cdesc = pl->cdesc;
    a.hub_rows = pl->hub_rows; a.B = c.B; a.ldb = c.ldb; a.n = pl->n_rows; a.nnz = pl->nnz; a.hub_nnz = pl->hub_nnz;
    a.n_chunks4 = c.n_feat / 4; a.T = pl->chunk_rows; a.n_chunks = pl->n_chunks; a.Kh = pl->n_hub;
    a.Kv = pl->n_vslot; a.vmap = pl->vmap; a.vcnt = pl->vcnt;
    a.hub_threshold = pl->hub_threshold; a.cap_doc = pl->cap_doc; a.cap_hub = pl->cap_hub;
    a.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(c.workspace) + 15u) & ~(uintptr_t)15u);
    a.ldp = (int64_t)((c.n_feat + 3) / 4) * 4;
    a.n_slices = a.n_groups = 0;
    const int kpg = (int)ceil_div64(pl->n_vslot, kNG);
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
                                const int32_t* __restrict__ hub_rows, const int64_t* __restrict__ hub_ofs, int Kv, int T,
                                const int32_t* __restrict__ vmap, const int32_t* __restrict__ vcnt,
                                uint32_t* __restrict__ keys, int32_t* __restrict__ src) {
    const int k = blockIdx.x;
    const int r = hub_rows[k];
    const int s = rowptr[r], e = rowptr[r + 1];
    const int64_t o = hub_ofs[k];
    const int nv = vcnt[k];
    for (int p = s + threadIdx.x; p < e; p += blockDim.x) {
        const int c = colidx[p];
        // a heavy hub row is dealt over nv virtual slots by column residue: every slot sees every chunk
        keys[o + (p - s)] = (uint32_t)(c / T) * (uint32_t)Kv + (uint32_t)vmap[k * 8 + (c % nv)];
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
    if (i >= (int64_t)n_chunks * (Kh + kHtabPad)) return;
    const int c = (int)(i / (Kh + kHtabPad));
    int k = (int)(i % (Kh + kHtabPad));
    if (k > Kh) k = Kh;  // padding entries repeat the end offset
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
    cdesc[c] = make_int4(rowptr[r0], rowptr[r1], htab[(int64_t)c * (Kh + kHtabPad)], htab[(int64_t)c * (Kh + kHtabPad) + Kh]);
}

void stream_plan_free(tg_plan* pl) {
    if (!pl) return;
    roles2_plan_free(pl);
    cudaFree(pl->colidx2); cudaFree(pl->hcol); cudaFree(pl->htab); cudaFree(pl->cdesc); cudaFree(pl->rsplit);
    cudaFree(pl->vmap); cudaFree(pl->vcnt);
    pl->colidx2 = nullptr; pl->hcol = nullptr; pl->hval = nullptr; pl->htab = nullptr; pl->cdesc = nullptr;
    pl->rsplit = nullptr; pl->vmap = nullptr; pl->vcnt = nullptr;
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
    int T = env_int("TG_STREAM_CHUNK", 256);
    if (T < 32 || T > 1024 || (T % 32) != 0) T = 256;
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

    // ---- virtual hub slots: split heavy hub rows, then balance the 64 row groups of a CTA (longest first) ----------
    // Group g of the wide kernels owns the slots {g + 64*kk}; the groups of a CTA meet at a barrier after every chunk,
    // so the slowest group sets the pace.  Topic popularity is heavily skewed (a single topic can carry more than the
    // average group's share), hence the heaviest rows are split by column residue into the spare accumulator slots, and
    // the pieces are then dealt to the groups with the LPT rule.
    std::vector<int32_t> vcnt((size_t)Kh, 1), vmap((size_t)Kh * 8, 0);
    int Kv = 0;
    {
        struct Piece { double w; int k, j; };
        std::vector<Piece> pieces;
        std::vector<double> len((size_t)Kh);
        for (int k = 0; k < Kh; ++k)
            len[(size_t)k] = (double)(h_rowptr[(size_t)hub_rows[(size_t)k] + 1] - h_rowptr[(size_t)hub_rows[(size_t)k]]);
        // slots are free: 64 groups x (smallest power of two >= Kh/64) accumulators; every spare slot is spent on
        // halving (then thirding, ...) whichever hub row currently has the heaviest pieces
        int kpg0 = 1;
        while (kpg0 * kNG < Kh) kpg0 *= 2;
        int spare = kpg0 * kNG - Kh;
        while (spare > 0) {
            int best = -1;
            for (int k = 0; k < Kh; ++k)
                if (vcnt[(size_t)k] < 8 && (best < 0 || len[(size_t)k] / vcnt[(size_t)k] > len[(size_t)best] / vcnt[(size_t)best])) best = k;
            if (best < 0) break;
            vcnt[(size_t)best] += 1;
            --spare;
        }
        for (int k = 0; k < Kh; ++k)
            for (int j = 0; j < vcnt[(size_t)k]; ++j) pieces.push_back(Piece{len[(size_t)k] / vcnt[(size_t)k], k, j});
        int kpg = 1;
        while (kpg * kNG < (int)pieces.size()) kpg *= 2;
        if (kpg > 8) return TG_OK;  // too many hub rows for the streaming layout
        Kv = kpg * kNG;
        std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return x.w > y.w; });
        std::vector<double> load((size_t)kNG, 0.0);
        std::vector<int> used((size_t)kNG, 0);
        for (const Piece& pc : pieces) {
            int best = -1;
            for (int g = 0; g < kNG; ++g)
                if (used[(size_t)g] < kpg && (best < 0 || load[(size_t)g] < load[(size_t)best])) best = g;
            vmap[(size_t)pc.k * 8 + pc.j] = best + kNG * used[(size_t)best];
            used[(size_t)best] += 1;
            load[(size_t)best] += pc.w;
        }
    }
    if ((uint64_t)n_chunks * (uint64_t)Kv >= 0xFFFFFFFFull) return TG_OK;

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
    TG_TRY(cudaMalloc((void**)&pl->vmap, (size_t)Kh * 8 * sizeof(int32_t)));
    TG_TRY(cudaMemcpyAsync(pl->vmap, vmap.data(), (size_t)Kh * 8 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TG_TRY(cudaMalloc((void**)&pl->vcnt, (size_t)Kh * sizeof(int32_t)));
    TG_TRY(cudaMemcpyAsync(pl->vcnt, vcnt.data(), (size_t)Kh * sizeof(int32_t), cudaMemcpyHostToDevice, st));
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
    hub_keys_kernel<<<Kh, 256, 0, st>>>(rowptr, colidx, pl->hub_rows, d_ofs, Kv, T, pl->vmap, pl->vcnt, keys_a, src_a);
    TG_TRY(cudaGetLastError());
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (uint64_t)n_chunks * Kv) ++end_bit;
    TG_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, src_a, src_b, (int)hub_nnz, 0, end_bit, st));
    TG_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    // stable: inside a (chunk, hub) segment the entries keep their column order
    TG_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, src_a, src_b, (int)hub_nnz, 0, end_bit, st));
    TG_TRY(cudaMalloc((void**)&pl->hcol, ((size_t)hub_nnz + 2) * sizeof(int2)));
    TG_TRY(cudaMemsetAsync(pl->hcol, 0, ((size_t)hub_nnz + 2) * sizeof(int2), st));
    hub_gather_kernel<<<(unsigned)ceil_div64(hub_nnz, 256), 256, 0, st>>>(keys_b, src_b, colidx, vals, hub_nnz, Kv, T,
                                                                         reinterpret_cast<int2*>(pl->hcol));
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaMalloc((void**)&pl->htab, (size_t)n_chunks * (Kv + kHtabPad) * sizeof(int32_t)));
    hub_table_kernel<<<(unsigned)ceil_div64((int64_t)n_chunks * (Kv + kHtabPad), 256), 256, 0, st>>>(keys_b, hub_nnz, n_chunks, Kv,
                                                                                              pl->htab);
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaMalloc((void**)&pl->cdesc, (size_t)n_chunks * sizeof(int4)));
    chunk_desc_kernel<<<(unsigned)ceil_div64(n_chunks, 256), 256, 0, st>>>(rowptr, pl->htab, n, T, n_chunks, Kv, pl->cdesc);
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaStreamSynchronize(st));
#undef TG_TRY
    cudaFree(d_slot); cudaFree(d_ofs); cudaFree(keys_a); cudaFree(keys_b); cudaFree(src_a); cudaFree(src_b); cudaFree(tmp);
    pl->chunk_rows = T;
    pl->n_chunks = n_chunks;
    pl->n_vslot = Kv;
    // staged entry windows: average occupancy of a chunk plus slack (rows beyond the window take the L2 path)
    const int64_t avg_doc = ceil_div64(pl->nnz - hub_nnz, n_chunks), avg_hub = ceil_div64(hub_nnz, n_chunks);
    pl->cap_doc = (int32_t)(((avg_doc + avg_doc / 4 + 64) + 1) & ~(int64_t)1);
    pl->cap_hub = (int32_t)(((avg_hub + avg_hub / 4 + 64) + 1) & ~(int64_t)1);
    pl->stream_ok = true;
    // the kernel must fit at least the narrow configuration
    if (stream_smem_bytes(pl, 1) > kSmemBudget) stream_plan_free(pl);
    if (pl->stream_ok) return roles2_plan_build(pl, rowptr, colidx, vals, h_rowptr, nullptr, hub_rows.data(), st);
    return TG_OK;
}

}  // namespace tg
