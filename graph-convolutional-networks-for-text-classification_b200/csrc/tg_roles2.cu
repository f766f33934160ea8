// tg_roles2.cu — role-specialised column-chunk streaming SpMM ("warp per hub slot").
//
// Product: Y = A * B for the normalised adjacency of a document-topic-topic graph (reference layer.py:106 and its autograd
// transpose product).  B is read from HBM once; its second use is served by L2.  One launch, two kinds of CTAs:
//
//   hub CTAs  (128-column slices, 16 warps): a WARP owns 16 hub slots; its 32 lanes cover the slice with one float4 each,
//             so every lane of the warp works on the same entry (no divergence inside a warp) and the sixteen
//             accumulators are sixteen float4 registers.  A CTA therefore owns a GROUP of 256 slots; graphs with more hub
//             rows (K = 1024 topics) use several groups — CTA (group g, slice s, lane l) streams the node range like every
//             other hub CTA but only walks the entries of its group.  Heavy hub rows are split over spare slots and the
//             pieces are dealt to the warps of all groups longest-first, so every warp carries the same load.  Chunks of
//             T nodes are staged with one 2-D TMA tile + two bulk copies per stage (double buffered on mbarriers).
//             Entries carry the byte offset of their row inside the staged tile: address = base + offset, one LDS.128,
//             two FFMA2.
//   doc CTAs  (slices of 32*NQ columns, 64 groups of 8 lanes): the K hub rows of B stay resident in shared memory (bulk
//             copies) — NQ = 4 / 2 / 1 float4 per lane for up to ~368 / ~736 / 1280 hub rows, so that the resident rows
//             always fit; a job is 64 consecutive rows — one row per group.  The job's entries and row descriptors arrive
//             through a bulk-copy ring (mbarrier full/empty pairs, issued two jobs ahead by one thread), and each group
//             prefetches the slice of B its NEXT row needs for the self loop into registers one job ahead, so no
//             consumer instruction waits on global memory.
//
// All floating-point additions happen in an order fixed by the plan: bitwise reproducible, no float atomics.
// Packed fp32 FMAs (FFMA2, fma.rn.f32x2) halve the issue slots of the arithmetic; the results are the IEEE fma results.
#include <cub/cub.cuh>
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "tg_async.cuh"
#include "tg_finish.cuh"
#include "tg_roles.cuh"

namespace tg {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kKPW = 16;              // hub slots per warp
constexpr int kKv = kWarps * kKPW;    // 256 slots per hub CTA = one slot group
constexpr int kMaxGroups = 5;         // slot groups of one plan (one hub CTA per group, slice and chunk lane)
constexpr int kMaxHubRows = 4096;     // hub rows the hub role handles (the document role's resident rows set a tighter limit: ~1 400)
constexpr int kHtW = kKv + 4;         // offset-table row of the 256-slot layout: 257 offsets padded to a multiple of 16 bytes
constexpr int kFT = 128;              // columns per 128-column slice (mask layout, narrow kernels' hub tile encoding)
constexpr int kRowBytes = kFT * 4;
constexpr int kHubEnc = 128;          // document-role entries address hub row h as h * kHubEnc (scaled by NQ in the kernel)
constexpr int kJobRows = 64;          // document role: rows per job = row groups per CTA
constexpr int kStages = 4;            // document role: entry ring depth (maximum)
constexpr int kPF = 2;                // ... jobs issued ahead
constexpr int kL2PF = 4;              // document role: jobs whose self-loop rows of B are prefetched into L2 ahead
constexpr int kHubL2PF = 4;           // hub role: chunks (per chunk lane) whose tiles of B are prefetched into L2 ahead
constexpr int kNStages = 4;            // narrow kernels: hub stages
constexpr int kNRing = 8;              // narrow kernels: document ring depth
constexpr int kNPF = 6;                // ... jobs issued ahead
constexpr size_t kSmemMax = 227 * 1024;
constexpr double kDropDocWeight = 1.10;  // document-role weight of the dropout epilogue in the SM split (roles2_run_t)
constexpr int32_t kNotShort = -1;      // rdesc.y of rows the document role does not produce (hub rows, padding)

struct R2Args {
    // hub role
    const int2* __restrict__ hent;
    const int32_t* __restrict__ htab;
    const int4* __restrict__ cdesc;     // {aligned base into hent, staged entries (even), first node of the chunk, -}
    int32_t T, n_chunks, cap_hub;
    int32_t groups, Kv;                 // slot groups (hub CTAs per chunk lane and slice), total slots = groups * 256
    // document role
    const int2* __restrict__ dent;
    const int2* __restrict__ rdesc;
    const int2* __restrict__ jdesc;
    int32_t n_jobs, cap_doc;
    const int32_t* __restrict__ hub_rows;
    int32_t Kh;
    // operands
    const float* __restrict__ B;
    int64_t ldb, n;
    int32_t n_chunks4;
    float* partials;
    int64_t ldp;
    int32_t hub_slices, hub_lanes, doc_slices, doc_lanes, only_role;
    int32_t n_stages;    // document-role ring depth actually used (<= kStages)
    int32_t hub_pf;      // hub role: chunks (per chunk lane) whose tiles are prefetched into L2 ahead; 0 = no prefetch
    const uint32_t* __restrict__ keep_bits;  // bit-packed dropout keep mask [n][n_feat/32] (bit b of word w = column 32w+b) or null
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// acc += v * b on packed pairs (two FFMA2)
__device__ __forceinline__ void fma4p(float4& acc, float v, const float4& b) {
    const float2 vv = make_float2(v, v);
    const float2 lo = __ffma2_rn(vv, make_float2(b.x, b.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(vv, make_float2(b.z, b.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// pull a 2-D tile of B into L2 ahead of the loads that need it (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int col, int row) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(col), "r"(row) : "memory");
}

__device__ __forceinline__ float4 lds128(const unsigned char* p) { return *reinterpret_cast<const float4*>(p); }

// One lock-step trip of the GS < 32 hub role: up to two entries (positions q and q + 1, each below h1) of this sub-group's run.
// No branch (the sub-groups of a warp disagree and a branch serialises them).  The loads carry their predicates — a skipped
// entry moves no data — and write into the caller's scratch registers `t`, which keep their previous contents when the load
// is skipped; the FMAs are unconditional, with the value of a skipped entry replaced by 0.  The scratch rows are cleared at
// the start of every slot, so what a skipped entry multiplies by 0 is either 0 or a row that already went into this very
// slot: no 0 * x product can turn a non-finite value of an unrelated row of B into a NaN of this slot (a slot that received an
// infinite term may end up NaN instead of infinite — non-finite either way, and confined to the rows the reference taints).  Shared memory is
// addressed with 32-bit shared-window addresses and the accumulators are the packed pairs the FFMA2s work on.  The C++ form
// of this trip compiled to 31 instructions, among them a generic-to-shared address conversion (S2R SR_CgaCtaId) per entry
// and six register clears; a variant with unpredicated loads from an all-zero row saved as many instructions but moved the
// skipped entries' rows through the shared-memory pipe (+50 % wavefronts: hub role 58 % -> 90 % of the pipe, no gain).
//   he_q: shared address of entry q; bl: shared address of this lane's float4 of tile row 0
struct HubScratch {
    unsigned x0, y0, x1, y1;
    unsigned long long a01, a23, b01, b23;
};
__device__ __forceinline__ void hub_trip2(unsigned long long& acc01, unsigned long long& acc23, HubScratch& t, unsigned he_q,
                                          unsigned bl, int q, int h1) {
    asm volatile(
        "{\n"
        ".reg .pred p0, p1;\n"
        ".reg .b32 q1, z0, z1;\n"
        ".reg .b64 v0, v1;\n"
        "setp.lt.s32 p0, %12, %13;\n"
        "add.s32 q1, %12, 1;\n"
        "setp.lt.s32 p1, q1, %13;\n"
        "@p0 ld.shared.v2.b32 {%2, %3}, [%10];\n"
        "@p1 ld.shared.v2.b32 {%4, %5}, [%10+8];\n"
        "@p0 add.u32 %2, %2, %11;\n"
        "@p1 add.u32 %4, %4, %11;\n"
        "@p0 ld.shared.v2.b64 {%6, %7}, [%2];\n"
        "@p1 ld.shared.v2.b64 {%8, %9}, [%4];\n"
        "selp.b32 z0, %3, 0, p0;\n"
        "selp.b32 z1, %5, 0, p1;\n"
        "mov.b64 v0, {z0, z0};\n"
        "mov.b64 v1, {z1, z1};\n"
        "fma.rn.f32x2 %0, v0, %6, %0;\n"
        "fma.rn.f32x2 %1, v0, %7, %1;\n"
        "fma.rn.f32x2 %0, v1, %8, %0;\n"
        "fma.rn.f32x2 %1, v1, %9, %1;\n"
        "}\n"
        : "+l"(acc01), "+l"(acc23), "+r"(t.x0), "+r"(t.y0), "+r"(t.x1), "+r"(t.y1), "+l"(t.a01), "+l"(t.a23), "+l"(t.b01), "+l"(t.b23)
        : "r"(he_q), "r"(bl), "r"(q), "r"(h1)
        : "memory");
}

// =============================================== hub role ===============================================================
// GS lanes per slot sub-group: the 32 lanes of a warp form NSUB = 32 / GS sub-groups; a sub-group owns 16 hub slots and its
// GS lanes cover a slice of 4 GS columns with one float4 each (sixteen float4 accumulators per lane, whatever GS is).
//   GS = 32: a warp per slot, 128-column slices, 256 slots per CTA    (K <= 256 topics)
//   GS = 16: two slots per warp step, 64-column slices, 512 slots     (K <= 512)
//   GS =  8: four slots per warp step, 32-column slices, 1 024 slots  (K <= 1 024; more hub rows: several slot groups)
// With narrower slices a tile of B holds more nodes for the same shared memory (T up to 512), so that a (chunk, slot) run
// stays a few entries long when the entries of a chunk spread over 1 024 slots, every tile is read by ONE CTA per slice
// instead of one per 256 slots, and the sub-groups of a warp walk their runs in lock step (the plan deals slots of similar
// weight to the same position, so the runs of a step have similar lengths).
template <int GS>
__device__ __forceinline__ void hub_role(const R2Args& a, const CUtensorMap* tmap, unsigned char* smem, uint64_t* bars,
                                         int bid) {
    constexpr int NSUB = 32 / GS;
    constexpr int kSlots = kKv * NSUB;          // slots per CTA (one slot group)
    constexpr int kHtWg = kSlots + 4;           // offset-table row, padded to a multiple of 16 bytes
    constexpr int kCols = 4 * GS;               // columns per CTA slice
    constexpr int kRowB = kCols * 4;            // bytes of a row of the staged tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gl = lane % GS, sub = lane / GS;
    // CTA = (chunk lane hl, slot group grp, column slice): the CTAs that share a chunk sequence are neighbours in the
    // grid, so the groups x slices readers of a node range run at the same time and share its rows of B in L2
    const int slice = bid % a.hub_slices, grp = (bid / a.hub_slices) % a.groups, hl = bid / (a.hub_slices * a.groups);
    const int4* __restrict__ cdesc = a.cdesc + (int64_t)grp * a.n_chunks;
    const int32_t* __restrict__ htab = a.htab + (int64_t)grp * a.n_chunks * kHtWg;
    const size_t bs_bytes = (size_t)a.T * kRowB;
    const size_t he_bytes = align128((size_t)a.cap_hub * 8);
    const size_t st_bytes = bs_bytes + he_bytes + align128((size_t)kHtWg * 4);
    const int box_rows = a.T < 256 ? a.T : 256;  // a TMA box holds at most 256 rows: taller tiles are several boxes
    uint64_t* full = bars;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // one thread moves a whole stage: the T x 4 GS tile of B (zero filled past the matrix), the chunk's run of hub entries
    // and its offset-table row
    auto issue = [&](int buf, int c, const int4 d) {
        unsigned char* base = smem + (size_t)buf * st_bytes;
        fence_proxy_async();  // the buffer was last read through the generic proxy (ordered by the block barrier)
        mbar_expect_tx(&full[buf], (unsigned)bs_bytes + (unsigned)d.y * 8u + (unsigned)(kHtWg * 4));
        for (int r0 = 0; r0 < a.T; r0 += box_rows) tma_load_2d(base + (size_t)r0 * kRowB, tmap, slice * kCols, d.z + r0, &full[buf]);
        if (d.y) bulk_load_1d(base + bs_bytes, a.hent + d.x, (unsigned)d.y * 8u, &full[buf]);
        bulk_load_1d(base + bs_bytes + he_bytes, htab + (int64_t)c * kHtWg, (unsigned)(kHtWg * 4), &full[buf]);
    };
    auto prefetch = [&](int row0) {
        for (int r0 = 0; r0 < a.T; r0 += box_rows) tma_prefetch_l2_2d(tmap, slice * kCols, row0 + r0);
    };
    float4 acc[kKPW];
    unsigned long long acc2[kKPW][2];   // the same accumulators as packed pairs (GS < 32: lock-step trips in PTX)
#pragma unroll
    for (int kk = 0; kk < kKPW; ++kk) {
        acc[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
        acc2[kk][0] = acc2[kk][1] = 0ull;
    }
    int c = hl;
    if (c < a.n_chunks) {
        int4 d_next = make_int4(0, 0, 0, 0);
        int d_pf = 0;  // first node of the chunk kHubL2PF steps ahead (loaded one step early: no dependent stall)
        if (tid == 0) {
            issue(0, c, __ldg(cdesc + c));
            if (c + a.hub_lanes < a.n_chunks) d_next = __ldg(cdesc + c + a.hub_lanes);
            if (grp == 0 && a.hub_pf >= 2) {
                // the tiles of the next steps: into L2 now, so that the TMA loads one step ahead see L2 latency, not HBM
                // latency (the double buffer alone does not cover an HBM round trip when a stage is ~1 us of work); one
                // slot group per chunk lane prefetches for all of them
                for (int i = 2; i < a.hub_pf; ++i)
                    if (c + i * a.hub_lanes < a.n_chunks) prefetch(__ldg(cdesc + c + i * a.hub_lanes).z);
                if (c + a.hub_pf * a.hub_lanes < a.n_chunks) d_pf = __ldg(cdesc + c + a.hub_pf * a.hub_lanes).z;
            }
        }
        for (int it = 0; c < a.n_chunks; c += a.hub_lanes, ++it) {
            const int buf = it & 1;
            const int cn = c + a.hub_lanes;
            if (tid == 0 && cn < a.n_chunks) {
                issue(buf ^ 1, cn, d_next);
                if (cn + a.hub_lanes < a.n_chunks) d_next = __ldg(cdesc + cn + a.hub_lanes);
            }
            if (tid == 0 && grp == 0 && a.hub_pf >= 2) {
                const int cp = c + a.hub_pf * a.hub_lanes;
                if (cp < a.n_chunks) prefetch(d_pf);
                if (cp + a.hub_lanes < a.n_chunks) d_pf = __ldg(cdesc + cp + a.hub_lanes).z;
            }
            mbar_wait(&full[buf], (unsigned)(it >> 1) & 1u);
            const unsigned char* base = smem + (size_t)buf * st_bytes;
            const unsigned char* Bl = base + gl * 16;
            const int2* he = reinterpret_cast<const int2*>(base + bs_bytes);
            const int32_t* ht = reinterpret_cast<const int32_t*>(base + bs_bytes + he_bytes) + (warp * NSUB + sub) * kKPW;
            // the sub-group's 17 slot offsets: four LDS.128 + one LDS.32 (the table row and 16-int steps are 16-byte aligned)
            int hofs[kKPW + 1];
            {
                const int4* ht4 = reinterpret_cast<const int4*>(ht);
#pragma unroll
                for (int i = 0; i < kKPW / 4; ++i) {
                    const int4 t = ht4[i];
                    hofs[4 * i] = t.x; hofs[4 * i + 1] = t.y; hofs[4 * i + 2] = t.z; hofs[4 * i + 3] = t.w;
                }
                hofs[kKPW] = ht[kKPW];
            }
            if (NSUB == 1) {
#pragma unroll
                for (int kk = 0; kk < kKPW; ++kk) {
                    int q = hofs[kk];
                    const int h1 = hofs[kk + 1];
#pragma unroll 1
                    for (; q + 4 <= h1; q += 4) {
                        const int2 e0 = he[q], e1 = he[q + 1], e2 = he[q + 2], e3 = he[q + 3];
                        const float4 b0 = lds128(Bl + e0.x), b1 = lds128(Bl + e1.x), b2 = lds128(Bl + e2.x), b3 = lds128(Bl + e3.x);
                        fma4p(acc[kk], __int_as_float(e0.y), b0);
                        fma4p(acc[kk], __int_as_float(e1.y), b1);
                        fma4p(acc[kk], __int_as_float(e2.y), b2);
                        fma4p(acc[kk], __int_as_float(e3.y), b3);
                    }
                    if (q + 2 <= h1) {
                        const int2 e0 = he[q], e1 = he[q + 1];
                        const float4 b0 = lds128(Bl + e0.x), b1 = lds128(Bl + e1.x);
                        fma4p(acc[kk], __int_as_float(e0.y), b0);
                        fma4p(acc[kk], __int_as_float(e1.y), b1);
                        q += 2;
                    }
                    if (q < h1) {
                        const int2 e0 = he[q];
                        fma4p(acc[kk], __int_as_float(e0.y), lds128(Bl + e0.x));
                    }
                }
            } else {
                const unsigned bl_s = smem_u32(Bl), he_s = smem_u32(he);
                HubScratch scr;
                scr.x0 = scr.y0 = scr.x1 = scr.y1 = 0u;
                // NSUB sub-groups in lock step: the warp makes as many trips as the longest of its NSUB runs needs, two entries
                // per trip; a sub-group past the end of its run sits the trip out.  Predicated, not branched: the sub-groups
                // disagree on the predicates and a branch would serialise them (measured: 22 ms against 16 ms at the
                // 6.25 M x 1 024 shard).  Walking two positions per trip for more chains in flight did not help (8.9 vs 8.5 ms).
#pragma unroll
                for (int kk = 0; kk < kKPW; ++kk) {
                    int q = hofs[kk];
                    const int h1 = hofs[kk + 1];
                    const int n_trip = (__reduce_max_sync(0xffffffffu, h1 - q) + 1) >> 1;
                    unsigned hq = he_s + 8u * (unsigned)q;
                    scr.a01 = scr.a23 = scr.b01 = scr.b23 = 0ull;   // (what a skipped entry multiplies by 0: see hub_trip2)
#pragma unroll 1   // (unrolled by two the body grows to 42 instructions for two trips — register shuffles around the asm blocks)
                    for (int t = 0; t < n_trip; ++t, q += 2, hq += 16u) hub_trip2(acc2[kk][0], acc2[kk][1], scr, hq, bl_s, q, h1);
                }
            }
            __syncthreads();
        }
    }
    float* dst = a.partials + ((int64_t)hl * a.Kv + grp * kSlots + (warp * NSUB + sub) * kKPW) * a.ldp + (int64_t)(slice * GS + gl) * 4;
    if (slice * GS + gl < a.n_chunks4) {  // (the last slice of a width that is no multiple of the slice width is narrower; TMA zero-fills it)
#pragma unroll
        for (int kk = 0; kk < kKPW; ++kk) {
            if (NSUB == 1) {
                *reinterpret_cast<float4*>(dst + (int64_t)kk * a.ldp) = acc[kk];
            } else {
                const uint2 lo = *reinterpret_cast<const uint2*>(&acc2[kk][0]), hi = *reinterpret_cast<const uint2*>(&acc2[kk][1]);
                *reinterpret_cast<float4*>(dst + (int64_t)kk * a.ldp) =
                    make_float4(__uint_as_float(lo.x), __uint_as_float(lo.y), __uint_as_float(hi.x), __uint_as_float(hi.y));
            }
        }
    }
}

// One predicated step of the document role's lock-step tail (R > 1): row r takes its entry p if it has one (p < n).  Loads
// are predicated into the caller's scratch (they keep their previous contents when skipped), the FMAs are unconditional with
// the value of a missing entry replaced by 0: what a missing entry multiplies by 0 is either 0 (the scratch is cleared at the
// start of the tail) or a hub row that already went into this very output row.  ea: shared address of the entry; bh: shared
// address of this lane's float4 of resident hub row 0; entries address hub row h as h * 128, a resident row is 128 NQ bytes.
template <int NQ>
struct DocTailScratch {
    unsigned x, y;
    unsigned long long b[NQ][2];
};
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack_f2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
template <int NQ>
__device__ __forceinline__ void doc_tail_step(unsigned long long (&acc)[NQ][2], DocTailScratch<NQ>& t, unsigned ea, unsigned bh, int p,
                                              int n) {
    static_assert(NQ == 1 || NQ == 2, "the lock-step tail is written for one or two float4 chunks per lane");
    if (NQ == 1) {
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            ".reg .b32 z;\n"
            ".reg .b64 v;\n"
            "setp.lt.s32 q, %8, %9;\n"
            "@q ld.shared.v2.b32 {%2, %3}, [%6];\n"
            "@q add.u32 %2, %2, %7;\n"
            "@q ld.shared.v2.b64 {%4, %5}, [%2];\n"
            "selp.b32 z, %3, 0, q;\n"
            "mov.b64 v, {z, z};\n"
            "fma.rn.f32x2 %0, v, %4, %0;\n"
            "fma.rn.f32x2 %1, v, %5, %1;\n"
            "}\n"
            : "+l"(acc[0][0]), "+l"(acc[0][1]), "+r"(t.x), "+r"(t.y), "+l"(t.b[0][0]), "+l"(t.b[0][1])
            : "r"(ea), "r"(bh), "r"(p), "r"(n)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            ".reg .b32 z;\n"
            ".reg .b64 v;\n"
            "setp.lt.s32 q, %12, %13;\n"
            "@q ld.shared.v2.b32 {%4, %5}, [%10];\n"
            "@q mad.lo.u32 %4, %4, 2, %11;\n"
            "@q ld.shared.v2.b64 {%6, %7}, [%4];\n"
            "@q ld.shared.v2.b64 {%8, %9}, [%4+128];\n"
            "selp.b32 z, %5, 0, q;\n"
            "mov.b64 v, {z, z};\n"
            "fma.rn.f32x2 %0, v, %6, %0;\n"
            "fma.rn.f32x2 %1, v, %7, %1;\n"
            "fma.rn.f32x2 %2, v, %8, %2;\n"
            "fma.rn.f32x2 %3, v, %9, %3;\n"
            "}\n"
            : "+l"(acc[0][0]), "+l"(acc[0][1]), "+l"(acc[NQ - 1][0]), "+l"(acc[NQ - 1][1]), "+r"(t.x), "+r"(t.y), "+l"(t.b[0][0]),
              "+l"(t.b[0][1]), "+l"(t.b[NQ - 1][0]), "+l"(t.b[NQ - 1][1])
            : "r"(ea), "r"(bh), "r"(p), "r"(n)
            : "memory");
    }
}

// =============================================== document role ===========================================================
// NQ float4 chunks per lane: a group of 8 lanes covers a slice of 32 * NQ columns of a row; the resident hub rows take
// Kh * 128 * NQ bytes of shared memory.  A group works on R = 4 / NQ rows AT ONCE (rows grp, grp + 64, ... of a super job of
// 64 * R consecutive rows), walking their entry lists in lock step: every (super) job is the same amount of work —
// 64 R rows x 32 NQ columns — and a warp always has four independent load -> FMA chains in flight per lane, whatever the
// slice width (with one row per group the narrow slices were latency bound: 1.7 us per job for a quarter of the work).
// TABLE: B is only the resident table (X * W with a sparse feature matrix X): no self loops to prefetch, every entry
// addresses the table.
template <int NQ, bool TABLE, class Epi>
__device__ __forceinline__ void doc_role(const R2Args& a, const Epi& epi, const CUtensorMap* tmap_job, unsigned char* smem,
                                         uint64_t* bars, int bid) {
    // Epi = EpiStore: elementwise epilogue with the per-thread constants in registers (any width).  Epi = EpiLoss (row-wise
    // log-softmax / cross-entropy): the row must lie inside ONE slice (n_feat <= 32 NQ), so that the 8 lanes of a group hold
    // the whole row — the class-sized products of graphs whose plan uses narrow slices (K = 1 024 topics).
    constexpr bool kStore = std::is_same<Epi, EpiStore>::value;
    constexpr int R = 4 / NQ;              // rows per group and job; plan jobs (64 rows) per super job
    constexpr int kSliceCols = 32 * NQ;
    constexpr int kRowB = kSliceCols * 4;  // bytes of a resident row
    constexpr int kSJRows = kJobRows * R;
    const int tid = threadIdx.x, lane = tid & 31, gl = lane & 7, grp = tid >> 3;
    const int slice = bid % a.doc_slices, dl = bid / a.doc_slices;
    const int q0 = slice * (8 * NQ) + gl;  // this lane owns the float4 chunks q0 + 8u, u < NQ
    // the last slice of a width that is no multiple of the slice width is narrower: chunks at or past n_chunks4 do not exist
    const int slice_bytes = min(kRowB, (a.n_chunks4 - slice * 8 * NQ) * 16);
    bool valid[NQ];
#pragma unroll
    for (int u = 0; u < NQ; ++u) valid[u] = q0 + 8 * u < a.n_chunks4;
    const int n_sj = (a.n_jobs + R - 1) / R;  // super jobs
    const size_t bh_bytes = align128((size_t)a.Kh * kRowB);
    const size_t en_bytes = align128((size_t)a.cap_doc * 8 * R);
    const size_t rd_bytes = (size_t)kSJRows * 8;
    const size_t st_bytes = en_bytes + rd_bytes + 128;  // entries | row descriptors | the R job descriptors
    unsigned char* ring = smem + bh_bytes;
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* bhbar = bars + 2 * kStages;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);
        }
        mbar_init(bhbar, 1);
        mbar_fence_init();
        mbar_expect_tx(bhbar, (unsigned)a.Kh * (unsigned)slice_bytes);
    }
    __syncthreads();
    // the hub rows of B (this slice) stay resident for the whole kernel
    for (int k = tid; k < a.Kh; k += kThreads)
        bulk_load_1d(smem + (size_t)k * kRowB, a.B + (int64_t)__ldg(a.hub_rows + k) * a.ldb + (int64_t)slice * kSliceCols, (unsigned)slice_bytes, bhbar);
    // ring depth and look-ahead: what fits next to the resident rows (plan time)
    const int ns = a.n_stages;
    const int pf = ns > kPF ? kPF : ns - 1;
    auto issue = [&](int s, int sjn) {  // s: ring stage
        const int j0 = sjn * R, j1 = min(j0 + R, a.n_jobs) - 1;  // plan jobs of this super job: their entry runs are contiguous
        const int2 jd0 = __ldg(a.jdesc + j0), jd1 = __ldg(a.jdesc + j1);
        const unsigned n_ent = (unsigned)(jd1.x + jd1.y - jd0.x);
        const unsigned n_rd = (unsigned)(j1 - j0 + 1) * (unsigned)(kJobRows * 8);
        unsigned char* base = ring + (size_t)s * st_bytes;
        fence_proxy_async();
        mbar_expect_tx(&full[s], n_ent * 8u + n_rd + (R > 1 ? 16u * ((R + 1) / 2) : 0u));
        if (n_ent) bulk_load_1d(base, a.dent + jd0.x, n_ent * 8u, &full[s]);
        bulk_load_1d(base + en_bytes, a.rdesc + (int64_t)j0 * kJobRows, n_rd, &full[s]);
        // (the job-descriptor array is padded with zeros past n_jobs: the copy of R of them never leaves it)
        if (R > 1) bulk_load_1d(base + en_bytes + rd_bytes, a.jdesc + j0, 16u * ((R + 1) / 2), &full[s]);
    };
    if (tid == 0) {
        for (int i = 0; i < pf; ++i)
            if (dl + i * a.doc_lanes < n_sj) issue(i, dl + i * a.doc_lanes);
        if (!TABLE) {
#pragma unroll
            for (int i = 1; i < kL2PF; ++i)
                if (dl + i * a.doc_lanes < n_sj) {
#pragma unroll
                    for (int r = 0; r < R; ++r) tma_prefetch_l2_2d(tmap_job, slice * kSliceCols, ((dl + i * a.doc_lanes) * R + r) * kJobRows);
                }
        }
    }
    // the self-loop operand of (super job sjx, row r) is at selfp + r * jrow_stride once selfp has been advanced to sjx:
    // running pointers instead of a 64-bit multiply per row and job
    const int64_t jrow_stride = (int64_t)kJobRows * a.ldb;
    const int64_t sj_stride = (int64_t)a.doc_lanes * kSJRows * a.ldb;
    const int n32 = (int)a.n;
    auto load_self = [&](float4(&dst)[R][NQ], int sjx, const float* selfp) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = sjx * kSJRows + r * kJobRows + grp;
            if (!TABLE && sjx < n_sj && row < n32) {
                const float* p = selfp + r * jrow_stride;
#pragma unroll
                for (int u = 0; u < NQ; ++u) dst[r][u] = valid[u] ? ldg_f4_stream(p + u * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int u = 0; u < NQ; ++u) dst[r][u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    // per-thread constants of the epilogue: this lane's bias values, the optional upstream scale
    float4 bias4[NQ];
    float gscale = 1.f;
    if constexpr (kStore) {
#pragma unroll
        for (int u = 0; u < NQ; ++u)
            bias4[u] = (epi.bias && valid[u]) ? __ldg(reinterpret_cast<const float4*>(epi.bias + (int64_t)(q0 + 8 * u) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        gscale = epi.out_scale ? __ldg(epi.out_scale) : 1.f;
    }
    // bit-packed mask: four words per 128 columns; chunk q lives in word q / 8 at bits 4 (q % 8) .. — this lane's NQ words
    // of a row are the words slice * NQ + u
    const int mask_words = ((a.n_chunks4 + 31) / 32) * 4;
    struct Bits { uint32_t w[R][NQ]; };
    auto load_bits = [&](int sjx) {
        Bits b;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < NQ; ++u) b.w[r][u] = 0u;
        if (!a.keep_bits) return b;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = sjx * kSJRows + r * kJobRows + grp;
            if (sjx < n_sj && row < n32) {
                const uint32_t* p = a.keep_bits + (int64_t)row * mask_words + slice * NQ;
                if (NQ == 4) {
                    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
                    b.w[r][0] = t.x; b.w[r][1 % NQ] = t.y; b.w[r][2 % NQ] = t.z; b.w[r][3 % NQ] = t.w;
                } else if (NQ == 2) {
                    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
                    b.w[r][0] = t.x; b.w[r][1 % NQ] = t.y;
                } else {
                    b.w[r][0] = __ldg(p);
                }
            }
        }
        return b;
    };
    const float* selfp = a.B + ((int64_t)dl * kSJRows + grp) * a.ldb + (int64_t)q0 * 4;
    float* yp = nullptr;
    int64_t yrow_stride = 0, ysj_stride = 0;
    bool plain = false;
    int fast_mode = 0;        // 1: bias + ReLU + bit-packed dropout mask, 2: bias + ReLU (see below)
    float drop_scale = 1.f;
    bool scale_pow2 = false;
    float4 bias_s[NQ];        // bias * 1 / (1 - p)
#pragma unroll
    for (int u = 0; u < NQ; ++u) bias_s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (kStore) {
        yp = epi.Y + ((int64_t)dl * kSJRows + grp) * epi.ldy + (int64_t)q0 * 4;
        yrow_stride = (int64_t)kJobRows * epi.ldy;
        ysj_stride = (int64_t)a.doc_lanes * kSJRows * epi.ldy;
        // nothing to do after the sums: plain product without bias / activation / scale / dropout
        plain = !epi.bias && !epi.relu && !epi.out_scale && !a.keep_bits && epi.drop_mode != 2;
        // The two epilogues of the layer-1 forward get paths of their own, chosen once per thread: the general path tests its
        // five flags and reloads their constants for every float4 — 40 instructions per chunk, as many per row as the whole
        // entry loop (bias alone cost 0.085 ms on top of the 0.78 ms plain product at C3).
        bool all_valid = true;
#pragma unroll
        for (int u = 0; u < NQ; ++u) all_valid = all_valid && valid[u];
        if (all_valid && epi.bias && epi.relu && !epi.out_scale && epi.drop_mode != 2) fast_mode = a.keep_bits ? 1 : 2;
        drop_scale = epi.scale;
        if (!(drop_scale > 0.f)) fast_mode = fast_mode == 1 ? 0 : fast_mode;   // (p = 1: scale 0 — the general path)
        scale_pow2 = (__float_as_uint(drop_scale) & 0x007fffffu) == 0u && drop_scale > 0.f;
#pragma unroll
        for (int u = 0; u < NQ; ++u)
            bias_s[u] = make_float4(bias4[u].x * drop_scale, bias4[u].y * drop_scale, bias4[u].z * drop_scale, bias4[u].w * drop_scale);
    }
    const unsigned gmask = group_mask<8>(lane);
    float4 cur[R][NQ];
    load_self(cur, dl, selfp);
    Bits kbits = load_bits(dl);
    mbar_wait(bhbar, 0);
    const unsigned char* BHl = smem + gl * 16;
    // ring positions as running counters (stage, phase): consumer side s / ph, producer side (pf jobs ahead) sp / php
    int s = 0, sp = pf;
    unsigned ph = 0, php = 0;
    if (sp >= ns) { sp -= ns; php ^= 1u; }
    bool wrapped = false;  // the producer has gone round the ring once: stages must be released before they are refilled
    for (int sj = dl; sj < n_sj; sj += a.doc_lanes) {
        if (tid == 0) {
            const int sjn = sj + pf * a.doc_lanes;
            if (sjn < n_sj) {
                // the stage was last read ns iterations before the one being issued
                if (wrapped || php) mbar_wait(&empty[sp], php ^ 1u);
                issue(sp, sjn);
            }
            // the rows of B the self loops of a later job need: into L2 now, so that the register prefetch one job ahead
            // sees L2 latency instead of HBM latency (bytes in flight per SM, not bandwidth, were the limit)
            const int sjp = sj + kL2PF * a.doc_lanes;
            if (!TABLE && sjp < n_sj) {
#pragma unroll
                for (int r = 0; r < R; ++r) tma_prefetch_l2_2d(tmap_job, slice * kSliceCols, (sjp * R + r) * kJobRows);
            }
        }
        mbar_wait(&full[s], ph);
        const unsigned char* base = ring + (size_t)s * st_bytes;
        const int2* jds = reinterpret_cast<const int2*>(base + en_bytes + rd_bytes);
        const int jbase0 = (R > 1) ? jds[0].x : 0;
        const int2* ent[R];
        int n_nh[R], n_hub[R];
        bool produce[R];
        int max_nh = 0, max_hub = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool live = sj * R + r < a.n_jobs;   // (a trailing super job may hold fewer than R plan jobs)
            const int2 rd = live ? reinterpret_cast<const int2*>(base + en_bytes)[r * kJobRows + grp] : make_int2(0, kNotShort);
            produce[r] = rd.y >= 0;  // hub rows and padding rows carry kNotShort; a row without entries is produced (epilogue of zero)
            const int n_tot = produce[r] ? (rd.y >> 16) : 0;
            n_nh[r] = produce[r] ? (rd.y & 0xffff) : 0;
            n_hub[r] = n_tot - n_nh[r];
            const int job_ofs = (R > 1 && live) ? jds[r].x - jbase0 : 0;
            ent[r] = reinterpret_cast<const int2*>(base) + job_ofs + rd.x;
            max_nh = max(max_nh, n_nh[r]);
            max_hub = max(max_hub, n_hub[r]);
        }
        float4 acc[R][NQ];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < NQ; ++u) acc[r][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        // columns outside the hub set: the self loop (prefetched) or a global gather
        for (int p = 0; p < max_nh; ++p) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (p < n_nh[r]) {
                    const int2 en = ent[r][p];
                    const float v = __int_as_float(en.y);
                    if (en.x == sj * kSJRows + r * kJobRows + grp) {
#pragma unroll
                        for (int u = 0; u < NQ; ++u) fma4p(acc[r][u], v, cur[r][u]);
                    } else {
                        const float* src = a.B + (int64_t)en.x * a.ldb + (int64_t)q0 * 4;
#pragma unroll
                        for (int u = 0; u < NQ; ++u)
                            if (valid[u]) fma4p(acc[r][u], v, ldg_f4(src + u * 32));
                    }
                }
            }
        }
        selfp += sj_stride;
        load_self(cur, sj + a.doc_lanes, selfp);  // next job's self-loop operand: in flight during the hub-column loop
#pragma unroll
        for (int r = 0; r < R; ++r) ent[r] += n_nh[r];
        if constexpr (R == 1) {
            // one row per group: two entries per trip
            int p = 0;
#pragma unroll 1
            for (; p + 2 <= max_hub; p += 2) {
                const int2 e0 = ent[0][p], e1 = ent[0][p + 1];
                const unsigned char* r0 = BHl + e0.x * NQ;  // entries address hub row h as h * 128; a resident row is 128 * NQ bytes
                const unsigned char* r1 = BHl + e1.x * NQ;
                float4 b0[NQ], b1[NQ];
#pragma unroll
                for (int u = 0; u < NQ; ++u) { b0[u] = lds128(r0 + 128 * u); b1[u] = lds128(r1 + 128 * u); }
                const float v0 = __int_as_float(e0.y), v1 = __int_as_float(e1.y);
#pragma unroll
                for (int u = 0; u < NQ; ++u) fma4p(acc[0][u], v0, b0[u]);
#pragma unroll
                for (int u = 0; u < NQ; ++u) fma4p(acc[0][u], v1, b1[u]);
            }
            if (p < max_hub) {
                const int2 e0 = ent[0][p];
                const unsigned char* r0 = BHl + e0.x * NQ;
                const float v0 = __int_as_float(e0.y);
#pragma unroll
                for (int u = 0; u < NQ; ++u) fma4p(acc[0][u], v0, lds128(r0 + 128 * u));
            }
        } else {
            // R rows in lock step, one entry of each per trip: R independent chains.  Rows of one graph have near-equal
            // lengths: up to the shortest row of the WARP the trips carry no predicates at all; the few entries beyond
            // it are walked row by row.
            int min_hub = n_hub[0];
#pragma unroll
            for (int r = 1; r < R; ++r) min_hub = min(min_hub, n_hub[r]);
            min_hub = __reduce_min_sync(0xffffffffu, min_hub);
#pragma unroll 1
            for (int p = 0; p < min_hub; ++p) {
                int2 e[R];
                float4 b[R][NQ];
#pragma unroll
                for (int r = 0; r < R; ++r) e[r] = ent[r][p];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned char* rp = BHl + e[r].x * NQ;
#pragma unroll
                    for (int u = 0; u < NQ; ++u) b[r][u] = lds128(rp + 128 * u);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float v = __int_as_float(e[r].y);
#pragma unroll
                    for (int u = 0; u < NQ; ++u) fma4p(acc[r][u], v, b[r][u]);
                }
            }
            // The entries beyond the warp's shortest row: all R rows of all groups stay in lock step up to the warp's LONGEST
            // row, a row that has run out loads nothing and adds 0 (doc_tail_step).  Walking them row by row — divergent
            // loops of dependent shared-memory loads, one row after the other — took as many instructions as the whole
            // unpredicated loop above and 16 % of the role's stall samples at the C4 shape.
            int max_hub_w = n_hub[0];
#pragma unroll
            for (int r = 1; r < R; ++r) max_hub_w = max(max_hub_w, n_hub[r]);
            max_hub_w = __reduce_max_sync(0xffffffffu, max_hub_w);
            if (max_hub_w > min_hub) {
                const unsigned bh_s = smem_u32(BHl);
                unsigned ea[R];
                DocTailScratch<NQ> scr[R];
                unsigned long long pa[R][NQ][2];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    ea[r] = smem_u32(ent[r] + min_hub);
                    scr[r].x = scr[r].y = 0u;
#pragma unroll
                    for (int u = 0; u < NQ; ++u) {
                        scr[r].b[u][0] = scr[r].b[u][1] = 0ull;
                        pa[r][u][0] = pack_f2(acc[r][u].x, acc[r][u].y);
                        pa[r][u][1] = pack_f2(acc[r][u].z, acc[r][u].w);
                    }
                }
#pragma unroll 1
                for (int p = min_hub; p < max_hub_w; ++p) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        doc_tail_step<NQ>(pa[r], scr[r], ea[r], bh_s, p, n_hub[r]);
                        ea[r] += 8u;
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int u = 0; u < NQ; ++u) {
                        const float2 lo = unpack_f2(pa[r][u][0]), hi = unpack_f2(pa[r][u][1]);
                        acc[r][u] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp no longer reads the stage
        if (++s == ns) { s = 0; ph ^= 1u; }
        if (++sp == ns) { sp = 0; php ^= 1u; wrapped = true; }
        const Bits kb = kbits;
        kbits = load_bits(sj + a.doc_lanes);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!produce[r]) continue;
            const int64_t row = (int64_t)sj * kSJRows + r * kJobRows + grp;
            if constexpr (!kStore) {
                // row-wise epilogue: the group holds the whole row (chunk gl + 8 u of lane gl)
                Chunk<4> out[NQ];
#pragma unroll
                for (int u = 0; u < NQ; ++u) { out[u].v[0] = acc[r][u].x; out[u].v[1] = acc[r][u].y; out[u].v[2] = acc[r][u].z; out[u].v[3] = acc[r][u].w; }
                epi.template apply<4, 8, NQ>(row, gl, gmask, a.n_chunks4, out);
            } else {
            // fused epilogue (EpiStore semantics, tg_epilogue.cuh) with the per-thread constants held in registers
            float* yrow = yp + r * yrow_stride;
            if (plain || row >= epi.raw_row_begin) {
#pragma unroll
                for (int u = 0; u < NQ; ++u)
                    if (valid[u]) *reinterpret_cast<float4*>(yrow + u * 32) = acc[r][u];
            } else if (fast_mode == 1) {
                // layer-1 forward in training: max(acc + bias, 0), kept elements times 1 / (1 - p) — nothing else to test
#pragma unroll
                for (int u = 0; u < NQ; ++u) {
                    const uint32_t w = kb.w[r][u] >> (4 * gl);
                    float4 v;
                    if (scale_pow2) {
                        // max(acc + b, 0) * s = max(acc * s + b * s, 0): two packed FMAs instead of four adds and four
                        // multiplies — only when s = 1 / (1 - p) is a power of two (p = 0.5, the reference's setting), where
                        // the two forms are bit-identical; otherwise the rounding of b * s can flip the sign of an element
                        // that cancels to almost zero, and with it the ReLU gate the backward reads from H1
                        v = bias_s[u];
                        fma4p(v, drop_scale, acc[r][u]);
                        v.x = fmaxf(v.x, 0.f);
                        v.y = fmaxf(v.y, 0.f);
                        v.z = fmaxf(v.z, 0.f);
                        v.w = fmaxf(v.w, 0.f);
                    } else {
                        v = acc[r][u];
                        v.x = fmaxf(v.x + bias4[u].x, 0.f) * drop_scale;
                        v.y = fmaxf(v.y + bias4[u].y, 0.f) * drop_scale;
                        v.z = fmaxf(v.z + bias4[u].z, 0.f) * drop_scale;
                        v.w = fmaxf(v.w + bias4[u].w, 0.f) * drop_scale;
                    }
                    v.x = (w & 1u) ? v.x : 0.f;
                    v.y = (w & 2u) ? v.y : 0.f;
                    v.z = (w & 4u) ? v.z : 0.f;
                    v.w = (w & 8u) ? v.w : 0.f;
                    *reinterpret_cast<float4*>(yrow + u * 32) = v;
                }
            } else if (fast_mode == 2) {
                // layer-1 forward in evaluation: max(acc + bias, 0)
#pragma unroll
                for (int u = 0; u < NQ; ++u) {
                    float4 v = acc[r][u];
                    v.x = fmaxf(v.x + bias4[u].x, 0.f);
                    v.y = fmaxf(v.y + bias4[u].y, 0.f);
                    v.z = fmaxf(v.z + bias4[u].z, 0.f);
                    v.w = fmaxf(v.w + bias4[u].w, 0.f);
                    *reinterpret_cast<float4*>(yrow + u * 32) = v;
                }
            } else {
#pragma unroll
                for (int u = 0; u < NQ; ++u) {
                    if (!valid[u]) continue;
                    float y[4] = {acc[r][u].x, acc[r][u].y, acc[r][u].z, acc[r][u].w};
                    const float bb[4] = {bias4[u].x, bias4[u].y, bias4[u].z, bias4[u].w};
                    if (epi.bias) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] += bb[k];
                    }
                    if (epi.relu) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] = fmaxf(y[k], 0.f);
                    }
                    if (epi.out_scale) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] *= gscale;
                    }
                    const int q = q0 + 8 * u;
                    if (a.keep_bits) {  // Philox dropout: the mask was drawn into the bit-packed side buffer (roles2_run)
                        const uint32_t w = kb.w[r][u] >> (4 * gl);
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] = ((w >> k) & 1u) ? y[k] * epi.scale : 0.f;
                    } else if (epi.drop_mode == 2) {
                        const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(epi.keep_mask + row * (int64_t)epi.n_feat + (int64_t)q * 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] = ((m >> (8 * k)) & 0xffu) ? y[k] * epi.scale : 0.f;
                    }
                    *reinterpret_cast<float4*>(yrow + u * 32) = make_float4(y[0], y[1], y[2], y[3]);
                }
            }
            }
        }
        yp += ysj_stride;
    }
}

// Bit-packed keep mask of the Philox dropout (definition: tg_common.cuh).  One thread per Philox call = (row, 64-column
// block, lane8): its eight 16-bit draws are the columns 64 blk + 32 half + 4 lane8 + k, i.e. bits 4 lane8 + k of the words
// 2 blk + half; the eight lanes of a block OR their nibbles together and lane 0 writes the two words.
__global__ void __launch_bounds__(256) r2_keep_bits_kernel(uint32_t* __restrict__ out, int64_t n_rows, int n_blk, int blk_shift, uint32_t thr,
                                                            uint64_t seed, uint64_t offset,
                                                            const unsigned long long* __restrict__ offset_dev) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_rows * (int64_t)n_blk * 8;
    const bool live = idx < total;
    const int lane8 = (int)(idx & 7);
    const int64_t rb = idx >> 3;
    // n_blk = n_feat / 64 is 2, 4 or 8 for the usual widths: a shift instead of a 64-bit division
    const int64_t row = blk_shift >= 0 ? (rb >> blk_shift) : rb / n_blk;
    const int blk = (int)(rb - row * n_blk);
    uint32_t w0 = 0, w1 = 0;
    if (live) {
        const uint64_t off = offset_dev ? offset + __ldg(offset_dev) : offset;
        const Philox4 r = dropout_philox(row, (uint32_t)lane8, (uint32_t)blk, seed, off);
        uint32_t u[4];
        dropout_u16x4(r, 0, u);
#pragma unroll
        for (int k = 0; k < 4; ++k) w0 |= (u[k] < thr ? 1u : 0u) << k;
        dropout_u16x4(r, 1, u);
#pragma unroll
        for (int k = 0; k < 4; ++k) w1 |= (u[k] < thr ? 1u : 0u) << k;
        w0 <<= 4 * lane8;
        w1 <<= 4 * lane8;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        w0 |= __shfl_xor_sync(0xffffffffu, w0, o);
        w1 |= __shfl_xor_sync(0xffffffffu, w1, o);
    }
    if (live && lane8 == 0) *reinterpret_cast<uint2*>(out + (row * n_blk + blk) * 2) = make_uint2(w0, w1);
}

// Exact-half mode (p = 0.5): the 128 random bits of a call ARE the keep bits of 128 consecutive columns — one thread per
// (row, 128-column block) writes four mask words, no comparisons, no shuffles.
__global__ void __launch_bounds__(256) r2_keep_bits_half_kernel(uint32_t* __restrict__ out, int64_t n_rows, int n_words, uint64_t seed,
                                                                 uint64_t offset, const unsigned long long* __restrict__ offset_dev) {
    const int n_c = (n_words + 3) / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * (int64_t)n_c) return;
    const int64_t row = idx / n_c;
    const int cidx = (int)(idx - row * n_c);
    const uint64_t off = offset_dev ? offset + __ldg(offset_dev) : offset;
    const Philox4 r = dropout_philox_half(row, (uint32_t)cidx, seed, off);
    uint32_t* dst = out + row * n_words + cidx * 4;
    if (cidx * 4 + 3 < n_words) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(r.x, r.y, r.z, r.w);
    } else {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        for (int j = 0; cidx * 4 + j < n_words; ++j) dst[j] = w[j];
    }
}

template <int GS, int NQ, bool TABLE, class Epi>
__global__ void __launch_bounds__(kThreads, 1) roles2_kernel(const R2Args a, const Epi epi, const __grid_constant__ CUtensorMap tmapB,
                                                            const __grid_constant__ CUtensorMap tmapJob) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2 * kStages + 2];
    // TMA destinations must be 128-byte aligned: align the dynamic base by hand (the launch requests 128 spare bytes)
    unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int n_hub_ctas = a.hub_slices * a.groups * a.hub_lanes;
    const int bid = blockIdx.x;
    if (bid < n_hub_ctas) {
        if (a.only_role == 2) return;
        hub_role<GS>(a, &tmapB, smem, bars, bid);
    } else {
        if (a.only_role == 1) return;
        doc_role<NQ, TABLE, Epi>(a, epi, &tmapJob, smem, bars, bid - n_hub_ctas);
    }
}

// =============================================== narrow operands (F <= 32) ===============================================
// The class-sized products (F = number of classes: 8, 20) of layer 2 use the same plan and the same two roles with a row
// of F floats covered by the 8 lanes of a group (lane gl owns float4 chunk gl when gl < F/4):
//   hub role: a warp still owns 16 slots and walks one slot at a time, but the four groups of the warp take every fourth
//             entry of the slot's run (trip counts differ by at most one), so one warp instruction covers four entries;
//             the four group accumulators are added once, at the end, by a fixed shuffle butterfly (deterministic);
//   doc role: identical ring / prefetch structure, one float4 accumulator per lane, whole-row epilogues (EpiLoss) work
//             because a group holds the whole row.
// Entry offsets were precomputed for 512-byte rows (x = local row * 512): rescaled here to the F*4-byte rows.
__device__ __forceinline__ void hub_role_narrow(const R2Args& a, const CUtensorMap* tmap, unsigned char* smem, uint64_t* bars,
                                                int bid) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gl = lane & 7, g = lane >> 3;
    const int grp = bid % a.groups, hl = bid / a.groups;  // slot group, chunk lane
    const int4* __restrict__ cdesc = a.cdesc + (int64_t)grp * a.n_chunks;
    const int32_t* __restrict__ htab = a.htab + (int64_t)grp * a.n_chunks * kHtW;
    const int n4 = a.n_chunks4;
    const int rowb = n4 * 16;
    const bool act = gl < n4;
    const int lane_off = (act ? gl : 0) * 16;
    const size_t tile_bytes = (size_t)a.T * rowb;
    const size_t bs_bytes = align128(tile_bytes);
    const size_t he_bytes = align128((size_t)a.cap_hub * 8);
    const size_t st_bytes = bs_bytes + he_bytes + align128((size_t)kHtW * 4);
    uint64_t* full = bars;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kNStages; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const bool contiguous = a.ldb == (int64_t)n4 * 4;  // rows of B back to back: one bulk copy instead of a 2-D tile
    auto issue = [&](int buf, int c, const int4 d) {
        unsigned char* base = smem + (size_t)buf * st_bytes;
        fence_proxy_async();
        if (contiguous) {
            const int64_t rows = (a.n - d.z < a.T) ? a.n - d.z : a.T;
            mbar_expect_tx(&full[buf], (unsigned)(rows * rowb) + (unsigned)d.y * 8u + (unsigned)(kHtW * 4));
            bulk_load_1d(base, a.B + (int64_t)d.z * a.ldb, (unsigned)(rows * rowb), &full[buf]);
        } else {
            mbar_expect_tx(&full[buf], (unsigned)tile_bytes + (unsigned)d.y * 8u + (unsigned)(kHtW * 4));
            tma_load_2d(base, tmap, 0, d.z, &full[buf]);
        }
        if (d.y) bulk_load_1d(base + bs_bytes, a.hent + d.x, (unsigned)d.y * 8u, &full[buf]);
        bulk_load_1d(base + bs_bytes + he_bytes, htab + (int64_t)c * kHtW, (unsigned)(kHtW * 4), &full[buf]);
    };
    // The slots of a warp are in descending weight order (plan: longest-first dealing).  The four heaviest are walked by the
    // whole warp, every group taking each fourth entry of the run; the other twelve are walked four at a time, one slot per
    // group — a quarter of the loop set-ups, at the price of lock-stepping four short runs of slightly different lengths.
    float4 accS[4], accG[3];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) accS[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) accG[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = hl;
    if (c < a.n_chunks) {
        // a chunk of a narrow operand is little work: keep kNStages - 1 chunks in flight to cover the HBM latency
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < kNStages - 1; ++i)
                if (c + i * a.hub_lanes < a.n_chunks) issue(i, c + i * a.hub_lanes, __ldg(cdesc + c + i * a.hub_lanes));
        }
        for (int it = 0; c < a.n_chunks; c += a.hub_lanes, ++it) {
            const int buf = it % kNStages;
            const int cn = c + (kNStages - 1) * a.hub_lanes;  // its buffer was released by the barrier of iteration it - 1
            if (tid == 0 && cn < a.n_chunks) issue((it + kNStages - 1) % kNStages, cn, __ldg(cdesc + cn));
            mbar_wait(&full[buf], (unsigned)(it / kNStages) & 1u);
            const unsigned char* base = smem + (size_t)buf * st_bytes;
            const unsigned char* Bl = base + lane_off;
            const int2* he = reinterpret_cast<const int2*>(base + bs_bytes);
            const int32_t* htw = reinterpret_cast<const int32_t*>(base + bs_bytes + he_bytes) + warp * kKPW;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                int q = htw[kk] + g;
                const int h1 = htw[kk + 1];
#pragma unroll 1
                for (; q + 4 < h1; q += 8) {
                    const int2 e0 = he[q], e1 = he[q + 4];
                    const float4 b0 = lds128(Bl + ((e0.x * n4) >> 5)), b1 = lds128(Bl + ((e1.x * n4) >> 5));
                    fma4p(accS[kk], __int_as_float(e0.y), b0);
                    fma4p(accS[kk], __int_as_float(e1.y), b1);
                }
                if (q < h1) {
                    const int2 e0 = he[q];
                    fma4p(accS[kk], __int_as_float(e0.y), lds128(Bl + ((e0.x * n4) >> 5)));
                }
            }
#pragma unroll
            for (int sx = 0; sx < 3; ++sx) {
                const int slot = 4 + 4 * sx + g;
                int q = htw[slot];
                const int h1 = htw[slot + 1];
#pragma unroll 1
                for (; q + 2 <= h1; q += 2) {
                    const int2 e0 = he[q], e1 = he[q + 1];
                    const float4 b0 = lds128(Bl + ((e0.x * n4) >> 5)), b1 = lds128(Bl + ((e1.x * n4) >> 5));
                    fma4p(accG[sx], __int_as_float(e0.y), b0);
                    fma4p(accG[sx], __int_as_float(e1.y), b1);
                }
                if (q < h1) {
                    const int2 e0 = he[q];
                    fma4p(accG[sx], __int_as_float(e0.y), lds128(Bl + ((e0.x * n4) >> 5)));
                }
            }
            __syncthreads();
        }
    }
    float* dst = a.partials + ((int64_t)hl * a.Kv + grp * kKv + warp * kKPW) * a.ldp + (int64_t)gl * 4;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        float4 v = accS[kk];
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {  // groups (0+1), (2+3), then the two pairs: a fixed tree
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
            v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
            v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
        }
        if (g == 0 && act) *reinterpret_cast<float4*>(dst + (int64_t)kk * a.ldp) = v;
    }
#pragma unroll
    for (int sx = 0; sx < 3; ++sx)
        if (act) *reinterpret_cast<float4*>(dst + (int64_t)(4 + 4 * sx + g) * a.ldp) = accG[sx];
}

template <class Epi>
__device__ __forceinline__ void doc_role_narrow(const R2Args& a, const Epi& epi, unsigned char* smem, uint64_t* bars, int dl) {
    const int tid = threadIdx.x, lane = tid & 31, gl = lane & 7, grp = tid >> 3;
    const unsigned gmask = group_mask<8>(lane);
    const int n4 = a.n_chunks4;
    const int rowb = n4 * 16;
    const bool act = gl < n4;
    const int lane_off = (act ? gl : 0) * 16;
    const size_t bh_bytes = align128((size_t)a.Kh * rowb);
    const size_t en_bytes = align128((size_t)a.cap_doc * 8);
    const size_t self_bytes = align128((size_t)kJobRows * rowb);
    const size_t st_bytes = en_bytes + (size_t)kJobRows * 8 + self_bytes;
    unsigned char* ring = smem + bh_bytes;
    uint64_t* full = bars;
    uint64_t* empty = bars + kNRing;
    uint64_t* bhbar = bars + 2 * kNRing;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kNRing; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);
        }
        mbar_init(bhbar, 1);
        mbar_fence_init();
        mbar_expect_tx(bhbar, (unsigned)a.Kh * rowb);
    }
    __syncthreads();
    for (int k = tid; k < a.Kh; k += kThreads)
        bulk_load_1d(smem + (size_t)k * rowb, a.B + (int64_t)__ldg(a.hub_rows + k) * a.ldb, rowb, bhbar);
    // a job of a narrow operand is a few hundred cycles of work: everything it reads — entries, row descriptors AND the
    // rows of B for the self loops — arrives through the ring, issued kNPF jobs ahead
    const bool contiguous = a.ldb == (int64_t)n4 * 4;
    auto issue = [&](int itn, int jobn) {
        const int s = itn % kNRing;
        const int2 jd = __ldg(a.jdesc + jobn);
        unsigned char* base = ring + (size_t)s * st_bytes;
        const int64_t r0 = (int64_t)jobn * kJobRows;
        const int rows = (int)((a.n - r0 < kJobRows) ? a.n - r0 : kJobRows);
        fence_proxy_async();
        mbar_expect_tx(&full[s], (unsigned)jd.y * 8u + (unsigned)(kJobRows * 8) + (unsigned)(rows * rowb));
        if (jd.y) bulk_load_1d(base, a.dent + jd.x, (unsigned)jd.y * 8u, &full[s]);
        bulk_load_1d(base + en_bytes, a.rdesc + r0, (unsigned)(kJobRows * 8), &full[s]);
        unsigned char* selfs = base + en_bytes + kJobRows * 8;
        if (contiguous) {
            bulk_load_1d(selfs, a.B + r0 * a.ldb, (unsigned)(rows * rowb), &full[s]);
        } else {
            for (int r = 0; r < rows; ++r) bulk_load_1d(selfs + (size_t)r * rowb, a.B + (r0 + r) * a.ldb, (unsigned)rowb, &full[s]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kNPF; ++i)
            if (dl + i * a.doc_lanes < a.n_jobs) issue(i, dl + i * a.doc_lanes);
    }
    mbar_wait(bhbar, 0);
    const unsigned char* BHl = smem + lane_off;
    int it = 0;
    for (int job = dl; job < a.n_jobs; job += a.doc_lanes, ++it) {
        if (tid == 0) {
            const int itn = it + kNPF, jobn = job + kNPF * a.doc_lanes;
            if (jobn < a.n_jobs) {
                if (itn >= kNRing) mbar_wait(&empty[itn % kNRing], (unsigned)(itn / kNRing - 1) & 1u);
                issue(itn, jobn);
            }
        }
        const int s = it % kNRing;
        mbar_wait(&full[s], (unsigned)(it / kNRing) & 1u);
        const unsigned char* base = ring + (size_t)s * st_bytes;
        const int2 rd = reinterpret_cast<const int2*>(base + en_bytes)[grp];
        const bool produce = rd.y >= 0;  // kNotShort: hub rows / padding; a row without entries still gets its epilogue
        const int n_tot = produce ? (rd.y >> 16) : 0, n_nh = produce ? (rd.y & 0xffff) : 0;
        const int2* ent = reinterpret_cast<const int2*>(base) + rd.x;
        const int64_t row = (int64_t)job * kJobRows + grp;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < n_nh; ++p) {
            const int2 en = ent[p];
            const float v = __int_as_float(en.y);
            if ((int64_t)en.x == row) fma4p(acc, v, lds128(base + en_bytes + kJobRows * 8 + (size_t)grp * rowb + lane_off));
            else fma4p(acc, v, act ? ldg_f4(a.B + (int64_t)en.x * a.ldb + gl * 4) : make_float4(0.f, 0.f, 0.f, 0.f));
        }
        int p = n_nh;
#pragma unroll 1
        for (; p + 4 <= n_tot; p += 4) {
            const int2 e0 = ent[p], e1 = ent[p + 1], e2 = ent[p + 2], e3 = ent[p + 3];
            const float4 b0 = lds128(BHl + ((e0.x * n4) >> 3)), b1 = lds128(BHl + ((e1.x * n4) >> 3));
            const float4 b2 = lds128(BHl + ((e2.x * n4) >> 3)), b3 = lds128(BHl + ((e3.x * n4) >> 3));
            fma4p(acc, __int_as_float(e0.y), b0);
            fma4p(acc, __int_as_float(e1.y), b1);
            fma4p(acc, __int_as_float(e2.y), b2);
            fma4p(acc, __int_as_float(e3.y), b3);
        }
        for (; p < n_tot; ++p) {
            const int2 e0 = ent[p];
            fma4p(acc, __int_as_float(e0.y), lds128(BHl + ((e0.x * n4) >> 3)));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (produce) {
            Chunk<4> out[1];
            out[0].v[0] = acc.x; out[0].v[1] = acc.y; out[0].v[2] = acc.z; out[0].v[3] = acc.w;
            epi.template apply<4, 8, 1>(row, gl, gmask, n4, out);
        }
    }
}

template <class Epi>
__global__ void __launch_bounds__(kThreads, 1) roles2n_kernel(const R2Args a, const Epi epi, const __grid_constant__ CUtensorMap tmapB) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2 * kNRing + 2];
    unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int bid = blockIdx.x, n_hub_ctas = a.groups * a.hub_lanes;
    if (bid < n_hub_ctas) {
        if (a.only_role == 2) return;
        hub_role_narrow(a, &tmapB, smem, bars, bid);
    } else {
        if (a.only_role == 1) return;
        doc_role_narrow(a, epi, smem, bars, bid - n_hub_ctas);
    }
}

// Lane-per-row document role for narrow operands whose rows lie back to back (ldb == n_feat): a row is NV float4 chunks
// held by ONE lane, so a warp covers 32 rows per instruction instead of 4 and the per-entry bookkeeping is amortised over
// the whole row (ncu on the group-per-row variant: 106 M warp instructions for 16 M entries at F = 20).  The CTA is cut
// into eight teams of two warps; a team owns a job sequence of its own (64 rows per job, one per lane) and a two-stage
// ring that brings a job's entries, row descriptors and the 64 rows of B for the self loops as three bulk copies.
constexpr int kTeams = 8;       // lane-per-row document role: teams of two warps, one job (64 rows) per team and stage
constexpr int kTeamStages = 2;
template <int NV, class Epi>
__device__ __forceinline__ void doc_role_narrow_lane(const R2Args& a, const Epi& epi, unsigned char* smem, uint64_t* bars, int dl) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int team = tid >> 6, tl = tid & 63;
    const int n4 = a.n_chunks4;
    const int rowb = n4 * 16;
    const size_t bh_bytes = align128((size_t)a.Kh * rowb);
    const size_t en_bytes = align128((size_t)a.cap_doc * 8);
    const size_t rd_bytes = (size_t)kJobRows * 8;
    const size_t self_bytes = align128((size_t)kJobRows * rowb);
    const size_t st_bytes = en_bytes + rd_bytes + self_bytes;
    unsigned char* ring = smem + bh_bytes + (size_t)team * kTeamStages * st_bytes;
    uint64_t* full = bars + team * 2 * kTeamStages;
    uint64_t* empty = full + kTeamStages;
    uint64_t* bhbar = bars + kTeams * 2 * kTeamStages;
    if (tid == 0) {
        for (int i = 0; i < kTeams; ++i)
            for (int s = 0; s < kTeamStages; ++s) {
                mbar_init(&bars[i * 2 * kTeamStages + s], 1);
                mbar_init(&bars[i * 2 * kTeamStages + kTeamStages + s], 2);  // the two warps of the team
            }
        mbar_init(bhbar, 1);
        mbar_fence_init();
        mbar_expect_tx(bhbar, (unsigned)a.Kh * rowb);
    }
    __syncthreads();
    for (int k = tid; k < a.Kh; k += kThreads)
        bulk_load_1d(smem + (size_t)k * rowb, a.B + (int64_t)__ldg(a.hub_rows + k) * a.ldb, rowb, bhbar);
    // every team runs its own two-stage ring over its own job sequence: 8 teams x 1 job in flight per CTA keep enough
    // bytes in flight to cover the HBM latency (one 82 KB stage per CTA did not: 4.8 us per stage, ncu/ROLE=2 timing)
    const int job_stride = a.doc_lanes * kTeams;
    auto issue = [&](int itn, int jobn) {
        const int s = itn % kTeamStages;
        const int2 jd = __ldg(a.jdesc + jobn);
        const int64_t r0 = (int64_t)jobn * kJobRows;
        const unsigned rows = (unsigned)((a.n - r0 < kJobRows) ? a.n - r0 : kJobRows);
        unsigned char* base = ring + (size_t)s * st_bytes;
        fence_proxy_async();
        mbar_expect_tx(&full[s], (unsigned)jd.y * 8u + (unsigned)rd_bytes + rows * (unsigned)rowb);
        if (jd.y) bulk_load_1d(base, a.dent + jd.x, (unsigned)jd.y * 8u, &full[s]);
        bulk_load_1d(base + en_bytes, a.rdesc + r0, (unsigned)rd_bytes, &full[s]);
        bulk_load_1d(base + en_bytes + rd_bytes, a.B + r0 * a.ldb, rows * (unsigned)rowb, &full[s]);
    };
    const int job0 = dl * kTeams + team;
    if (tl == 0) {
#pragma unroll
        for (int i = 0; i < kTeamStages - 1; ++i)
            if (job0 + i * job_stride < a.n_jobs) issue(i, job0 + i * job_stride);
    }
    mbar_wait(bhbar, 0);
    int it = 0;
    for (int job = job0; job < a.n_jobs; job += job_stride, ++it) {
        if (tl == 0) {
            const int itn = it + kTeamStages - 1, jobn = job + (kTeamStages - 1) * job_stride;
            if (jobn < a.n_jobs) {
                if (itn >= kTeamStages) mbar_wait(&empty[itn % kTeamStages], (unsigned)(itn / kTeamStages - 1) & 1u);
                issue(itn, jobn);
            }
        }
        const int s = it % kTeamStages;
        mbar_wait(&full[s], (unsigned)(it / kTeamStages) & 1u);
        const unsigned char* base = ring + (size_t)s * st_bytes;
        const int2 rd = reinterpret_cast<const int2*>(base + en_bytes)[tl];
        const bool produce = rd.y >= 0;
        const int n_tot = produce ? (rd.y >> 16) : 0, n_nh = produce ? (rd.y & 0xffff) : 0;
        const int2* ent = reinterpret_cast<const int2*>(base) + rd.x;
        const int64_t row = (int64_t)job * kJobRows + tl;
        float4 acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < n_nh; ++p) {
            const int2 en = ent[p];
            const float v = __int_as_float(en.y);
            if ((int64_t)en.x == row) {
                const unsigned char* sr = base + en_bytes + rd_bytes + (size_t)tl * rowb;
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (i < n4) fma4p(acc[i], v, lds128(sr + i * 16));
            } else {
                const float* src = a.B + (int64_t)en.x * a.ldb;
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (i < n4) fma4p(acc[i], v, ldg_f4(src + i * 4));
            }
        }
        int p = n_nh;
#pragma unroll 1
        for (; p + 2 <= n_tot; p += 2) {
            const int2 e0 = ent[p], e1 = ent[p + 1];
            const unsigned char* r0 = smem + ((e0.x * n4) >> 3);
            const unsigned char* r1 = smem + ((e1.x * n4) >> 3);
            const float v0 = __int_as_float(e0.y), v1 = __int_as_float(e1.y);
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i < n4) {
                    const float4 b0 = lds128(r0 + i * 16), b1 = lds128(r1 + i * 16);
                    fma4p(acc[i], v0, b0);
                    fma4p(acc[i], v1, b1);
                }
        }
        if (p < n_tot) {
            const int2 e0 = ent[p];
            const unsigned char* r0 = smem + ((e0.x * n4) >> 3);
            const float v0 = __int_as_float(e0.y);
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i < n4) fma4p(acc[i], v0, lds128(r0 + i * 16));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (produce) {
            Chunk<4> out[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) { out[i].v[0] = acc[i].x; out[i].v[1] = acc[i].y; out[i].v[2] = acc[i].z; out[i].v[3] = acc[i].w; }
            epi.template apply<4, 1, NV>(row, 0, 1u << lane, n4, out);
        }
    }
}

template <int NV, class Epi>
__global__ void __launch_bounds__(kThreads, 1) roles2nl_kernel(const R2Args a, const Epi epi, const __grid_constant__ CUtensorMap tmapB) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[kTeams * 2 * kTeamStages + 2];
    unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int bid = blockIdx.x, n_hub_ctas = a.groups * a.hub_lanes;
    if (bid < n_hub_ctas) {
        if (a.only_role == 2) return;
        hub_role_narrow(a, &tmapB, smem, bars, bid);
    } else {
        if (a.only_role == 1) return;
        doc_role_narrow_lane<NV>(a, epi, smem, bars, bid - n_hub_ctas);
    }
}

// =============================================== plan build =============================================================
// number of hub-row entries in every column (= entries the hub role processes for that node)
__global__ void r2_hub_deg_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                  const int32_t* __restrict__ hub_rows, int32_t* __restrict__ deg) {
    const int r = hub_rows[blockIdx.x];
    const int s = rowptr[r], e = rowptr[r + 1];
    for (int p = s + threadIdx.x; p < e; p += blockDim.x) atomicAdd(deg + colidx[p], 1);
}

// one block per hub row: key = ((group * n_chunks) + chunk) * S + slot inside the group (S = slots per group, a power of
// two), for each of its entries, in storage (column) order.  A heavy hub row is dealt over nv slots by column residue:
// every slot sees every chunk.
__global__ void r2_hub_keys_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                   const int32_t* __restrict__ hub_rows, const int64_t* __restrict__ hub_ofs,
                                   const int32_t* __restrict__ node_chunk, int n_chunks, int s_shift, const int32_t* __restrict__ vmap,
                                   const int32_t* __restrict__ vcnt, uint32_t* __restrict__ keys, int32_t* __restrict__ src) {
    const int k = blockIdx.x;
    const int r = hub_rows[k];
    const int s = rowptr[r], e = rowptr[r + 1];
    const int64_t o = hub_ofs[k];
    const int nv = vcnt[k];
    const uint32_t s_mask = (1u << s_shift) - 1u;
    for (int p = s + threadIdx.x; p < e; p += blockDim.x) {
        const int c = colidx[p];
        const uint32_t slot = (uint32_t)vmap[k * 8 + (c % nv)];
        keys[o + (p - s)] = (((slot >> s_shift) * (uint32_t)n_chunks + (uint32_t)node_chunk[c]) << s_shift) | (slot & s_mask);
        src[o + (p - s)] = p;
    }
}

__global__ void r2_hub_gather_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ src,
                                     const int32_t* __restrict__ colidx, const float* __restrict__ vals, int64_t hub_nnz,
                                     int n_chunks, int s_shift, int row_bytes, const int32_t* __restrict__ cstart, int2* __restrict__ hent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hub_nnz) return;
    const int p = src[i];
    const int c = (int)((keys[i] >> s_shift) % (uint32_t)n_chunks);
    hent[i] = make_int2((colidx[p] - cstart[c]) * row_bytes, __float_as_int(vals[p]));
}

// tab[gc][s] = first sorted position with key >= gc*S + s   (gc = group * n_chunks + chunk; s = S: start of the next run)
__global__ void r2_hub_table_kernel(const uint32_t* __restrict__ keys, int64_t hub_nnz, int n_runs, int S, int32_t* __restrict__ tab) {
    const int W = S + 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_runs * W) return;
    const int c = (int)(i / W);
    int s = (int)(i % W);
    if (s > S) s = S;
    const uint64_t want = (uint64_t)c * S + s;
    int64_t lo = 0, hi = hub_nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((uint64_t)keys[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    tab[i] = (int32_t)lo;
}

// offsets relative to the run's even-aligned base (16-byte aligned bulk copies) + run descriptors
__global__ void r2_hub_rel_kernel(const int32_t* __restrict__ tab_abs, int n_runs, int n_chunks, int S, const int32_t* __restrict__ cstart,
                                  int32_t* __restrict__ tab, int4* __restrict__ cdesc) {
    const int W = S + 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_runs * W) return;
    const int c = (int)(i / W), s = (int)(i % W);
    const int base = tab_abs[(int64_t)c * W] & ~1;
    tab[i] = tab_abs[i] - base;
    if (s == 0) cdesc[c] = make_int4(base, ((tab_abs[(int64_t)c * W + S] - base) + 1) & ~1, cstart[c % n_chunks], 0);
}

__global__ void r2_row_len_kernel(const int32_t* __restrict__ rowptr, int64_t n, int64_t n_pad, int hub_threshold, int32_t* __restrict__ len) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_pad) return;
    int l = 0;
    if (r < n) {
        l = rowptr[r + 1] - rowptr[r];
        if (l > hub_threshold) l = 0;
    }
    len[r] = l;
}

__global__ void r2_slot_kernel(const int32_t* __restrict__ hub_rows, int Kh, int32_t* __restrict__ slot_of) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < Kh) slot_of[hub_rows[k]] = k;
}

// Compact copy of the short rows' entries + row / job descriptors.  Each row lists its non-hub columns first ({column, value}:
// the self loop, or a global gather) and its hub columns after ({hub index * 128, value}: the resident rows); within each
// part the ascending column order of the CSR is kept, so a document row of a document-topic graph (self loop, then topics)
// is summed in exactly the reference's storage order.
__global__ void r2_doc_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                   const int32_t* __restrict__ slot_of, const int32_t* __restrict__ start, int64_t n, int64_t n_pad,
                                   int hub_threshold, int2* __restrict__ dent2, int2* __restrict__ rdesc, int2* __restrict__ jdesc) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    const int64_t j0 = r / kJobRows * kJobRows;
    const int jbase = start[j0] & ~1;
    int2 rd = make_int2(0, kNotShort);
    if (r < n) {
        const int s = rowptr[r], e = rowptr[r + 1];
        const int len = e - s;
        if (len <= hub_threshold) {  // (a row without entries is a short row too: its output is the epilogue of zero)
            const int o = start[r];
            int w = o;
            for (int p = s; p < e; ++p) {
                const int c = colidx[p];
                if (slot_of[c] < 0) dent2[w++] = make_int2(c, __float_as_int(vals[p]));
            }
            const int n_other = w - o;
            for (int p = s; p < e; ++p) {
                const int sl = slot_of[colidx[p]];
                if (sl >= 0) dent2[w++] = make_int2(sl * kHubEnc, __float_as_int(vals[p]));
            }
            rd = make_int2(o - jbase, (len << 16) | n_other);
        }
    }
    rdesc[r] = rd;
    if (r == j0) {
        const int64_t j1 = (j0 + kJobRows < n_pad) ? j0 + kJobRows : n_pad;
        jdesc[r / kJobRows] = make_int2(jbase, ((start[j1] - jbase) + 1) & ~1);
    }
}

int env_int2(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

size_t hub_smem(int T, int cap_hub, int gs) {
    const size_t slots = (size_t)kKv * (32 / gs);
    return 2 * ((size_t)T * 16 * gs + align128((size_t)cap_hub * 8) + align128((slots + 4) * 4));
}
size_t doc_smem(int Kh, int cap_doc, int nq, int stages) {
    const int r = 4 / nq;  // plan jobs per super job (rows per group)
    return align128((size_t)Kh * 128 * nq) + (size_t)stages * (align128((size_t)cap_doc * 8 * r) + (size_t)kJobRows * 8 * r + 128);
}

// float4 chunks per lane and ring depth of the document role: the widest slice whose resident rows leave room for a ring of
// at least three stages, else two stages
bool pick_doc_cfg(int Kh, int cap_doc, int force_nq, int* nq_out, int* stages_out) {
    for (int min_stages = 3; min_stages >= 2; --min_stages)
        for (int nq = 4; nq >= 1; nq >>= 1) {
            if (force_nq > 0 && nq != force_nq) continue;
            for (int stages = kStages; stages >= min_stages; --stages)
                if (doc_smem(Kh, cap_doc, nq, stages) + 256 <= kSmemMax) {
                    *nq_out = nq;
                    *stages_out = stages;
                    return true;
                }
        }
    return false;
}

// Slots: split the heaviest hub rows into the spare slots of G groups, deal the pieces to the G * 16 warps longest-first.
struct SlotDeal {
    std::vector<int32_t> vcnt, vmap;
    double max_warp = 0.0;
};
SlotDeal deal_slots(const std::vector<double>& len, int G, int nsub) {
    const int Kh = (int)len.size();
    SlotDeal d;
    d.vcnt.assign((size_t)Kh, 1);
    d.vmap.assign((size_t)Kh * 8, 0);
    int spare = G * kKv * nsub - Kh;
    while (spare > 0) {
        int best = -1;
        for (int k = 0; k < Kh; ++k)
            if (d.vcnt[(size_t)k] < 8 && (best < 0 || len[(size_t)k] / d.vcnt[(size_t)k] > len[(size_t)best] / d.vcnt[(size_t)best])) best = k;
        if (best < 0) break;
        d.vcnt[(size_t)best] += 1;
        --spare;
    }
    struct Piece { double w; int k, j; };
    std::vector<Piece> pieces;
    for (int k = 0; k < Kh; ++k)
        for (int j = 0; j < d.vcnt[(size_t)k]; ++j) pieces.push_back(Piece{len[(size_t)k] / d.vcnt[(size_t)k], k, j});
    std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return x.w > y.w; });
    // units of 16 slots: a warp (GS = 32) or a sub-group of a warp; unit u = (group * 16 + warp) * nsub + sub.  Pieces go
    // longest-first to the least loaded unit, so position p of every unit holds pieces of similar weight: the sub-groups of
    // a warp, which walk position p in lock step, then have runs of similar length.
    const int W = G * kWarps * nsub;
    std::vector<double> load((size_t)W, 0.0);
    std::vector<int> used((size_t)W, 0);
    for (const Piece& pc : pieces) {
        int best = -1;
        for (int w = 0; w < W; ++w)
            if (used[(size_t)w] < kKPW && (best < 0 || load[(size_t)w] < load[(size_t)best])) best = w;
        d.vmap[(size_t)pc.k * 8 + pc.j] = best * kKPW + used[(size_t)best];  // = group * S + (warp * nsub + sub) * 16 + position
        used[(size_t)best] += 1;
        load[(size_t)best] += pc.w;
    }
    // the pace of a warp is set by its slowest sub-group
    for (int w = 0; w < W; ++w) d.max_warp = std::max(d.max_warp, load[(size_t)w]);
    return d;
}

void read_knobs(tg_plan* pl) {
    pl->r2_min_rows = env_int2("TG_ROLES2_MIN_ROWS", 16384);
    pl->r2_narrow_min_rows = env_int2("TG_ROLES2_NARROW_MIN_ROWS", 131072);
    pl->r2_hub_pct = env_int2("TG_ROLES2_HUB_PCT", -1);
    pl->r2_narrow_hub_pct = env_int2("TG_ROLES2_NARROW_HUB_PCT", -1);
    pl->r2_only_role = env_int2("TG_ROLES_ONLY", 0);
    pl->r2_narrow_lane = env_int2("TG_ROLES2_NARROW_LANE", 1);
    pl->r2_narrow_ok = env_int2("TG_ROLES2_NARROW", 1) != 0;
    pl->r2_hub_pf = env_int2("TG_ROLES2_HUB_PF", -1);
}

}  // namespace

void roles2_plan_free(tg_plan* pl) {
    if (!pl) return;
    cudaFree(pl->r2_hent); cudaFree(pl->r2_htab); cudaFree(pl->r2_cdesc); cudaFree(pl->r2_vmap); cudaFree(pl->r2_vcnt);
    cudaFree(pl->r2_dent); cudaFree(pl->r2_rdesc); cudaFree(pl->r2_jdesc); cudaFree(pl->r2_ident);
    pl->r2_ident = nullptr;
    pl->r2_hent = nullptr; pl->r2_htab = nullptr; pl->r2_cdesc = nullptr; pl->r2_vmap = nullptr; pl->r2_vcnt = nullptr;
    pl->r2_dent = nullptr; pl->r2_rdesc = nullptr; pl->r2_jdesc = nullptr;
    pl->r2_ok = false;
    pl->r2_rect = 0;
}

// Builds the sub-plan of the role kernels.  h_hub_rows: host [n_hub].
int roles2_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* h_rowptr,
                      const int32_t* h_hub_rows, cudaStream_t st) {
    pl->r2_ok = false;
    read_knobs(pl);
    if (env_int2("TG_ROLES2", 1) == 0) return TG_OK;
    const bool all_hub = pl->r2_rect == 2;  // rectangular [K x N] operand whose rows are all hub rows: hub side only
    if (!colidx || !vals || pl->nnz == 0) return TG_OK;
    if (pl->n_hub < 1 || pl->n_hub > kMaxHubRows) return TG_OK;
    if (pl->hub_threshold > 16383) return TG_OK;  // row descriptors keep the entry count of a short row in 15 bits
    if (!all_hub) {
        if (pl->n_rows != pl->n_cols) return TG_OK;
        if (pl->hub_nnz * 8 < pl->nnz) return TG_OK;  // only worth it when the hub rows carry a real share of the entries
    }
    const int64_t n = pl->n_rows;
    const int64_t n_nodes = pl->n_cols;     // nodes the hub role streams over (= n for the square graphs)
    const int Kh = pl->n_hub;
    const int64_t hub_nnz = pl->hub_nnz;
    if (hub_nnz >= (int64_t)0x7fffffff) return TG_OK;

    // ---- lanes per slot sub-group (hub_role<GS>) and slot groups -----------------------------------------------------------
    // Up to 256 hub rows a warp per slot and 128-column slices; more hub rows: narrower slices, 512 / 1 024 slots per CTA.
    // One more group than strictly needed buys spare slots for splitting the heavy rows (balance) at the price of one more
    // reader of every tile of B.  Cost model in shared-memory wavefronts per 128 columns, summed over the hub CTAs of a
    // chunk lane: 5 per entry on the critical warp (x 16 warps that wait for it) + 4 per node for the TMA fill, per group.
    int gs = Kh <= kKv ? 32 : (Kh <= 2 * kKv ? 16 : 8);
    {
        const int gs_force = env_int2("TG_ROLES2_GS", 0);
        if (gs_force == 32 || gs_force == 16 || gs_force == 8) gs = gs_force;
    }
    const int nsub = 32 / gs;
    const int S = kKv * nsub;  // slots per group
    int s_shift = 8;
    while ((1 << s_shift) < S) ++s_shift;
    std::vector<int64_t> hub_ofs((size_t)Kh);
    std::vector<double> len((size_t)Kh);
    {
        int64_t run = 0;
        for (int k = 0; k < Kh; ++k) {
            const int32_t r = h_hub_rows[k];
            hub_ofs[(size_t)k] = run;
            len[(size_t)k] = (double)(h_rowptr[(size_t)r + 1] - h_rowptr[(size_t)r]);
            run += h_rowptr[(size_t)r + 1] - h_rowptr[(size_t)r];
        }
    }
    const int g_min = (Kh + S - 1) / S;
    if (g_min > kMaxGroups) return TG_OK;
    const int g_force = env_int2("TG_ROLES2_GROUPS", 0);
    // The minimum number of groups unless its best deal leaves the slowest unit more than 50 % above the mean load: a second
    // reader of every tile and twice the (slot, chunk) visits cost more than a moderate imbalance (measured at K = 1 024:
    // one group of 1 024 slots 20 % faster than two groups with the heavy rows split).
    SlotDeal deal;
    int G = 0;
    {
        double total = 0.0;
        for (double l : len) total += l;
        for (int g = g_min; g <= std::min(g_min + 1, kMaxGroups); ++g) {
            if (g_force >= g_min && g_force <= kMaxGroups && g != g_force) continue;
            deal = deal_slots(len, g, nsub);
            G = g;
            const double mean_unit = total / (double)(g * kWarps * nsub);
            if (deal.max_warp <= 1.5 * mean_unit) break;
        }
    }
    const std::vector<int32_t>& vcnt = deal.vcnt;
    const std::vector<int32_t>& vmap = deal.vmap;

    int64_t* d_ofs = nullptr;
    uint32_t *keys_a = nullptr, *keys_b = nullptr;
    int32_t *src_a = nullptr, *src_b = nullptr, *tab_abs = nullptr, *d_len = nullptr, *d_start = nullptr, *d_slot = nullptr;
    void* tmp = nullptr;
    cudaError_t e = cudaSuccess;
    auto release_tmp = [&]() {
        cudaFree(d_ofs); cudaFree(keys_a); cudaFree(keys_b); cudaFree(src_a); cudaFree(src_b); cudaFree(tab_abs);
        cudaFree(d_len); cudaFree(d_start); cudaFree(d_slot); cudaFree(tmp);
        d_ofs = nullptr; keys_a = keys_b = nullptr; src_a = src_b = tab_abs = d_len = d_start = d_slot = nullptr; tmp = nullptr;
    };
    auto fail = [&](cudaError_t err, const char* what) {
        release_tmp();
        roles2_plan_free(pl);
        return cuda_fail(err, what, __FILE__, __LINE__);
    };
    auto give_up = [&]() {  // the layout does not apply: not an error
        release_tmp();
        const int32_t rect = pl->r2_rect;
        roles2_plan_free(pl);
        pl->r2_rect = rect;
        return TG_OK;
    };
#define TG_TRY(call) do { e = (call); if (e != cudaSuccess) return fail(e, #call); } while (0)
    TG_TRY(cudaMalloc((void**)&d_ofs, (size_t)Kh * sizeof(int64_t)));
    TG_TRY(cudaMemcpyAsync(d_ofs, hub_ofs.data(), (size_t)Kh * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    TG_TRY(cudaMalloc((void**)&pl->r2_vmap, (size_t)Kh * 8 * sizeof(int32_t)));
    TG_TRY(cudaMemcpyAsync(pl->r2_vmap, vmap.data(), (size_t)Kh * 8 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TG_TRY(cudaMalloc((void**)&pl->r2_vcnt, (size_t)Kh * sizeof(int32_t)));
    TG_TRY(cudaMemcpyAsync(pl->r2_vcnt, vcnt.data(), (size_t)Kh * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    TG_TRY(cudaMalloc((void**)&keys_a, (size_t)hub_nnz * sizeof(uint32_t)));
    TG_TRY(cudaMalloc((void**)&keys_b, (size_t)hub_nnz * sizeof(uint32_t)));
    TG_TRY(cudaMalloc((void**)&src_a, (size_t)hub_nnz * sizeof(int32_t)));
    TG_TRY(cudaMalloc((void**)&src_b, (size_t)hub_nnz * sizeof(int32_t)));
    TG_TRY(cudaMalloc((void**)&pl->r2_hent, ((size_t)hub_nnz + 2) * sizeof(int2)));

    // ---- hub side: variable-height chunks ------------------------------------------------------------------------------
    // A chunk is a run of consecutive nodes with at most T rows AND at most `cap` hub entries (all groups together), so that
    // its tile of B and the entries of any one group always fit one shared-memory stage: document ranges are cut by the row
    // limit, the dense topic-topic block (every hub node carries Kh entries) by the entry limit.
    const double avg_deg = (double)hub_nnz / (double)n_nodes;
    static const int kTs[] = {512, 256, 192, 176, 160, 144, 128, 96, 64, 32};  // (taller than a TMA box of 256 rows: whole boxes)
    const int t_first = env_int2("TG_ROLES2_T", 512);
    const int row_bytes = 16 * gs;  // bytes of a row of the staged tile
    int T = 0, cap = 0;
    for (int t : kTs) {
        if (t > t_first) continue;
        const int64_t room = (int64_t)(kSmemMax - 256) / 2 - (int64_t)t * row_bytes - (int64_t)align128((size_t)(S + 4) * 4);
        const int cap_fit = (int)std::min<int64_t>(room / 8, 1 << 20) & ~1;
        if (room <= 0 || cap_fit < Kh + 2 || (double)cap_fit < 1.25 * avg_deg * t + 32) continue;
        T = t;
        cap = cap_fit;
        break;
    }
    if (T == 0) return give_up();
    std::vector<int32_t> h_deg((size_t)n_nodes), node_chunk((size_t)n_nodes), cstart;
    {
        int32_t* d_deg = nullptr;
        TG_TRY(cudaMalloc((void**)&d_deg, (size_t)n_nodes * sizeof(int32_t)));
        e = cudaMemsetAsync(d_deg, 0, (size_t)n_nodes * sizeof(int32_t), st);
        if (e == cudaSuccess) {
            r2_hub_deg_kernel<<<Kh, 256, 0, st>>>(rowptr, colidx, pl->hub_rows, d_deg);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_deg.data(), d_deg, (size_t)n_nodes * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_deg);
        if (e != cudaSuccess) return fail(e, "hub degree histogram");
    }
    {
        int rows = 0, ents = 0;
        cstart.push_back(0);
        for (int64_t j = 0; j < n_nodes; ++j) {
            const int d = h_deg[(size_t)j];
            if (rows > 0 && (rows == T || ents + d > cap - 2)) {  // -2: the staged run starts at an even entry and has even length
                cstart.push_back((int32_t)j);
                rows = 0;
                ents = 0;
            }
            node_chunk[(size_t)j] = (int32_t)cstart.size() - 1;
            rows += 1;
            ents += d;
        }
    }
    const int n_chunks = (int)cstart.size();
    cstart.push_back((int32_t)n_nodes);
    const int64_t n_runs = (int64_t)G * n_chunks;  // (group, chunk) runs of the entry list
    if ((uint64_t)n_runs * (uint64_t)S >= 0xFFFFFFFFull) return give_up();
    {
        int32_t *d_node_chunk = nullptr, *d_cstart = nullptr;
        auto drop = [&]() { cudaFree(d_node_chunk); cudaFree(d_cstart); };
#define TG_TRY2(call) do { e = (call); if (e != cudaSuccess) { drop(); return fail(e, #call); } } while (0)
        TG_TRY2(cudaMalloc((void**)&d_node_chunk, (size_t)n_nodes * sizeof(int32_t)));
        TG_TRY2(cudaMalloc((void**)&d_cstart, (size_t)(n_chunks + 1) * sizeof(int32_t)));
        TG_TRY2(cudaMemcpyAsync(d_node_chunk, node_chunk.data(), (size_t)n_nodes * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        TG_TRY2(cudaMemcpyAsync(d_cstart, cstart.data(), (size_t)(n_chunks + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        r2_hub_keys_kernel<<<Kh, 256, 0, st>>>(rowptr, colidx, pl->hub_rows, d_ofs, d_node_chunk, n_chunks, s_shift, pl->r2_vmap, pl->r2_vcnt,
                                               keys_a, src_a);
        TG_TRY2(cudaGetLastError());
        size_t tmp_bytes = 0;
        int end_bit = 1;
        while (end_bit < 32 && (1ull << end_bit) < (uint64_t)n_runs * S) ++end_bit;
        TG_TRY2(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, src_a, src_b, (int)hub_nnz, 0, end_bit, st));
        TG_TRY2(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        // stable: inside a (group, chunk, slot) run the entries keep their column order
        TG_TRY2(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, src_a, src_b, (int)hub_nnz, 0, end_bit, st));
        const int64_t tab_n = n_runs * (S + 4);
        TG_TRY2(cudaMalloc((void**)&tab_abs, (size_t)tab_n * sizeof(int32_t)));
        TG_TRY2(cudaMalloc((void**)&pl->r2_htab, (size_t)tab_n * sizeof(int32_t)));
        TG_TRY2(cudaMalloc((void**)&pl->r2_cdesc, (size_t)n_runs * sizeof(int4)));
        r2_hub_table_kernel<<<(unsigned)ceil_div64(tab_n, 256), 256, 0, st>>>(keys_b, hub_nnz, (int)n_runs, S, tab_abs);
        TG_TRY2(cudaGetLastError());
        r2_hub_rel_kernel<<<(unsigned)ceil_div64(tab_n, 256), 256, 0, st>>>(tab_abs, (int)n_runs, n_chunks, S, d_cstart, pl->r2_htab, pl->r2_cdesc);
        TG_TRY2(cudaGetLastError());
        TG_TRY2(cudaMemsetAsync(pl->r2_hent, 0, ((size_t)hub_nnz + 2) * sizeof(int2), st));
        r2_hub_gather_kernel<<<(unsigned)ceil_div64(hub_nnz, 256), 256, 0, st>>>(keys_b, src_b, colidx, vals, hub_nnz, n_chunks, s_shift, row_bytes,
                                                                                 d_cstart, pl->r2_hent);
        TG_TRY2(cudaGetLastError());
        TG_TRY2(cudaStreamSynchronize(st));
#undef TG_TRY2
        drop();
        pl->r2_T = T;
        pl->r2_n_chunks = n_chunks;
        pl->r2_cap_hub = cap;
        pl->r2_groups = G;
        pl->r2_gs = gs;
        if (pl->r2_hub_pf < 0) pl->r2_hub_pf = kHubL2PF;
    }
    // the sort scratch is not needed any more: release it before the document side allocates
    cudaFree(keys_a); cudaFree(keys_b); cudaFree(src_a); cudaFree(src_b); cudaFree(tab_abs); cudaFree(tmp);
    keys_a = keys_b = nullptr; src_a = src_b = tab_abs = nullptr; tmp = nullptr;

    if (all_hub) {  // no short rows: the hub side is the whole plan
        release_tmp();
        pl->r2_ok = true;
        return TG_OK;
    }

    // ---- document side: compact entry copy, row descriptors, job descriptors --------------------------------------------------
    const int64_t n_jobs = ceil_div64(n, kJobRows);
    const int64_t n_pad = n_jobs * kJobRows;
    const int64_t doc_nnz = pl->nnz - hub_nnz;
    TG_TRY(cudaMalloc((void**)&d_slot, (size_t)n * sizeof(int32_t)));
    TG_TRY(cudaMemsetAsync(d_slot, 0xff, (size_t)n * sizeof(int32_t), st));
    r2_slot_kernel<<<(unsigned)ceil_div64(Kh, 256), 256, 0, st>>>(pl->hub_rows, Kh, d_slot);
    TG_TRY(cudaGetLastError());
    TG_TRY(cudaMalloc((void**)&d_len, (size_t)(n_pad + 1) * sizeof(int32_t)));
    TG_TRY(cudaMalloc((void**)&d_start, (size_t)(n_pad + 1) * sizeof(int32_t)));
    r2_row_len_kernel<<<(unsigned)ceil_div64(n_pad + 1, 256), 256, 0, st>>>(rowptr, n, n_pad, pl->hub_threshold, d_len);
    TG_TRY(cudaGetLastError());
    size_t scan_bytes = 0;
    TG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_len, d_start, (int)(n_pad + 1), st));
    TG_TRY(cudaMalloc(&tmp, scan_bytes ? scan_bytes : 1));
    TG_TRY(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, d_len, d_start, (int)(n_pad + 1), st));
    TG_TRY(cudaMalloc((void**)&pl->r2_dent, ((size_t)doc_nnz + 2) * sizeof(int2)));
    TG_TRY(cudaMemsetAsync(pl->r2_dent, 0, ((size_t)doc_nnz + 2) * sizeof(int2), st));
    TG_TRY(cudaMalloc((void**)&pl->r2_rdesc, (size_t)n_pad * sizeof(int2)));
    TG_TRY(cudaMalloc((void**)&pl->r2_jdesc, (size_t)(n_jobs + 8) * sizeof(int2)));  // padded: the lane-per-row kernel copies 8 at a time
    TG_TRY(cudaMemsetAsync(pl->r2_jdesc, 0, (size_t)(n_jobs + 8) * sizeof(int2), st));
    r2_doc_fill_kernel<<<(unsigned)ceil_div64(n_pad, 256), 256, 0, st>>>(rowptr, colidx, vals, d_slot, d_start, n, n_pad, pl->hub_threshold,
                                                                          pl->r2_dent, pl->r2_rdesc, pl->r2_jdesc);
    TG_TRY(cudaGetLastError());
    std::vector<int2> h_jdesc((size_t)n_jobs);
    TG_TRY(cudaMemcpyAsync(h_jdesc.data(), pl->r2_jdesc, (size_t)n_jobs * sizeof(int2), cudaMemcpyDeviceToHost, st));
    TG_TRY(cudaStreamSynchronize(st));
#undef TG_TRY
    int cap_doc = 2;
    for (const int2& d : h_jdesc) cap_doc = std::max(cap_doc, d.y);
    release_tmp();
    pl->r2_n_jobs = (int32_t)n_jobs;
    pl->r2_cap_doc = cap_doc;
    int nq = 0, stages = 0;
    if (!pick_doc_cfg(Kh, cap_doc, env_int2("TG_ROLES2_NQ", 0), &nq, &stages)) return give_up();
    pl->r2_nq = nq;
    pl->r2_stages = stages;
    pl->r2_ok = true;
    return TG_OK;
}

// whole_row: the epilogue needs a whole output row inside one lane group (EpiLoss)
bool roles2_applicable(const tg_plan* pl, const StreamCall& c, bool whole_row) {
    if (!pl || !pl->r2_ok || pl->r2_rect != 0) return false;
    if (c.n_feat < 4 || c.n_feat % 4 != 0 || c.n_feat > 1024) return false;
    if (c.n_feat < 64) {
        // class-sized operands: plans with warp-per-slot hub CTAs (gs = 32: 128-column slices) have their own narrow kernels;
        // plans with 64 / 32-column slices run them here as a one-slice product (TMA zero-fills the rest of the box)
        if (pl->r2_gs == 32 || !pl->r2_narrow_ok || c.n_feat > 32 * pl->r2_nq) return false;
        if (pl->n_rows < (int64_t)pl->r2_narrow_min_rows) return false;  // small graphs: the operand is L2 resident, the gather kernel is faster
    } else {
        // below ~16 K rows a launch is a few microseconds of work and the 148-CTA prologue (resident hub rows, barriers, TMA
        // descriptors) costs more than it saves (R8 shape: 0.150 vs 0.113 ms per captured step on the gather kernel)
        if (pl->n_rows < (int64_t)pl->r2_min_rows) return false;
    }
    if (whole_row && (c.n_feat > 32 * pl->r2_nq || pl->r2_gs == 32)) return false;
    if (c.ldb % 4 != 0 || !aligned16(c.B) || !encode_tiled_fn()) return false;
    return true;
}

// bytes of the per-CTA hub partials: (chunk lanes x groups x slices) <= 148 CTAs, one row of `ld` floats per slot and chunk lane
static size_t partial_bytes(const tg_plan* pl, int32_t n_feat) {
    const size_t ld = (size_t)((n_feat + 3) / 4) * 4;
    const int nsub = 32 / pl->r2_gs;
    const size_t slots = (size_t)pl->r2_groups * kKv * nsub;
    const int slices = n_feat <= 32 ? 1 : (n_feat + 4 * pl->r2_gs - 1) / (4 * pl->r2_gs);
    const size_t lanes_max = std::max<size_t>(1, (size_t)kNumSM / ((size_t)pl->r2_groups * slices));
    return ((lanes_max * slots * ld * sizeof(float) + 16) + 255) & ~(size_t)255;
}

size_t roles2_workspace_bytes(const tg_plan* pl, int32_t n_feat) {
    if (!pl || !pl->r2_ok) return 0;
    if (pl->r2_rect == 1) return 16;
    const size_t ld = (size_t)((n_feat + 3) / 4) * 4;
    const size_t part = partial_bytes(pl, n_feat);
    if (pl->r2_rect == 2) return part;
    // + the bit-packed dropout keep mask (four words per row and 128 columns)
    return part + (size_t)pl->n_rows * (((ld + kFT - 1) / kFT) * 16) + 256;
}

int roles2_launches(const tg_plan* pl, const StreamCall& c, bool philox, bool whole_row) {
    if (!whole_row && roles2_rect_applicable(pl, c) && !(pl->r2_rect == 1 && philox)) return pl->r2_rect;  // resident-table product: one kernel; all-hub product: hub role + finish
    if (roles2_applicable(pl, c, whole_row)) return 2 + (philox ? 1 : 0);
    if (roles2_narrow_applicable(pl, c)) return 2;
    return 0;
}

namespace {

// Split of the 148 SMs between the roles.  Both roles are bound by the shared-memory pipe, so the split follows their
// wavefront counts (128-byte shared-memory transactions per 128 columns): hub role 5 per entry (row of B + broadcast
// entry) + 4 per node and group (TMA fill), document role 4 per hub-column entry + ~14 per row and slice (entries, self loop,
// store), weighted by the pipe utilisation each role reaches (82 % / 67 %, profiles/r01_roles2_*_only_full.md).  C3: 44 %
// hub CTAs, the measured optimum.  hub_unit / doc_unit: CTAs per chunk lane / job lane.
void split_sms(double hub_w, double doc_w, int hub_unit, int doc_unit, int n_chunks, int n_jobs, int forced_pct, int* hub_lanes_out,
               int* doc_lanes_out) {
    int best_h = 1, best_d = 1;
    double best = -1.0;
    if (forced_pct >= 0) {
        best_h = std::max(1, (kNumSM * forced_pct / 100) / hub_unit);
        best_d = std::max(1, (kNumSM - best_h * hub_unit) / doc_unit);
    } else {
        for (int h = 1; h * hub_unit + doc_unit <= kNumSM; ++h) {
            const int d = (kNumSM - h * hub_unit) / doc_unit;
            const double t = std::max(hub_w / (double)(h * hub_unit), doc_w / (double)(d * doc_unit));
            if (best < 0.0 || t < best) {
                best = t;
                best_h = h;
                best_d = d;
            }
        }
    }
    *hub_lanes_out = std::min(best_h, std::max(n_chunks, 1));
    *doc_lanes_out = std::min(best_d, std::max(n_jobs, 1));
}

template <int GS, int NQ, bool TABLE, class Epi>
int launch_roles(const R2Args& a, const Epi& epi, const CUtensorMap& tmap, const CUtensorMap& tmap_job, size_t smem, unsigned grid,
                 cudaStream_t st) {
    TG_CUDA(cudaFuncSetAttribute(roles2_kernel<GS, NQ, TABLE, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roles2_kernel<GS, NQ, TABLE, Epi><<<grid, kThreads, smem, st>>>(a, epi, tmap, tmap_job);
    TG_LAUNCH_CHECK();
    return TG_OK;
}
template <int GS>
int launch_wide_gs(int nq, const R2Args& a, const EpiStore& epi, const CUtensorMap& tmap, const CUtensorMap& tmap_job, size_t smem,
                   unsigned grid, cudaStream_t st) {
    if (nq == 4) return launch_roles<GS, 4, false, EpiStore>(a, epi, tmap, tmap_job, smem, grid, st);
    if (nq == 2) return launch_roles<GS, 2, false, EpiStore>(a, epi, tmap, tmap_job, smem, grid, st);
    return launch_roles<GS, 1, false, EpiStore>(a, epi, tmap, tmap_job, smem, grid, st);
}
int launch_wide(int gs, int nq, const R2Args& a, const EpiStore& epi, const CUtensorMap& tmap, const CUtensorMap& tmap_job, size_t smem,
                unsigned grid, cudaStream_t st) {
    if (gs == 32) return launch_wide_gs<32>(nq, a, epi, tmap, tmap_job, smem, grid, st);
    if (gs == 16) return launch_wide_gs<16>(nq, a, epi, tmap, tmap_job, smem, grid, st);
    return launch_wide_gs<8>(nq, a, epi, tmap, tmap_job, smem, grid, st);
}
// row-wise loss epilogue: plans with narrow slices only (gs = 16 / 8; the warp-per-slot plans have their own narrow kernels)
int launch_wide(int gs, int nq, const R2Args& a, const EpiLoss& epi, const CUtensorMap& tmap, const CUtensorMap& tmap_job, size_t smem,
                unsigned grid, cudaStream_t st) {
    if (gs == 16 && nq == 4) return launch_roles<16, 4, false, EpiLoss>(a, epi, tmap, tmap_job, smem, grid, st);
    if (gs == 16 && nq == 2) return launch_roles<16, 2, false, EpiLoss>(a, epi, tmap, tmap_job, smem, grid, st);
    if (gs == 16 && nq == 1) return launch_roles<16, 1, false, EpiLoss>(a, epi, tmap, tmap_job, smem, grid, st);
    if (gs == 8 && nq == 2) return launch_roles<8, 2, false, EpiLoss>(a, epi, tmap, tmap_job, smem, grid, st);
    if (gs == 8 && nq == 1) return launch_roles<8, 1, false, EpiLoss>(a, epi, tmap, tmap_job, smem, grid, st);
    set_error("role kernels: no loss-epilogue instance for gs = %d, nq = %d", gs, nq);
    return TG_ERR_UNSUPPORTED;
}
int launch_table(int nq, const R2Args& a, const EpiStore& epi, const CUtensorMap& tmap, const CUtensorMap& tmap_job, size_t smem,
                 unsigned grid, cudaStream_t st) {
    if (nq == 4) return launch_roles<32, 4, true, EpiStore>(a, epi, tmap, tmap_job, smem, grid, st);
    if (nq == 2) return launch_roles<32, 2, true, EpiStore>(a, epi, tmap, tmap_job, smem, grid, st);
    return launch_roles<32, 1, true, EpiStore>(a, epi, tmap, tmap_job, smem, grid, st);
}

void fill_common(R2Args& a, const tg_plan* pl, const StreamCall& c) {
    memset(&a, 0, sizeof(a));
    a.hent = pl->r2_hent; a.htab = pl->r2_htab; a.cdesc = pl->r2_cdesc;
    a.T = pl->r2_T; a.n_chunks = pl->r2_n_chunks; a.cap_hub = pl->r2_cap_hub;
    a.groups = pl->r2_groups; a.Kv = pl->r2_groups * kKv * (32 / pl->r2_gs);
    a.dent = pl->r2_dent; a.rdesc = pl->r2_rdesc; a.jdesc = pl->r2_jdesc;
    a.n_jobs = pl->r2_n_jobs; a.cap_doc = pl->r2_cap_doc;
    a.hub_rows = pl->hub_rows; a.Kh = pl->n_hub;
    a.B = c.B; a.ldb = c.ldb; a.n = pl->n_rows; a.n_chunks4 = c.n_feat / 4;
    a.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(c.workspace) + 15u) & ~(uintptr_t)15u);
    a.ldp = (int64_t)((c.n_feat + 3) / 4) * 4;
    a.keep_bits = nullptr;
    a.n_stages = pl->r2_stages;
    a.only_role = pl->r2_only_role;
    a.hub_pf = pl->r2_hub_pf;
}

}  // namespace

template <class Epi>
static int roles2_run_t(const tg_plan* pl, const StreamCall& c, const Epi& epi, cudaStream_t st) {
    R2Args a;
    fill_common(a, pl, c);
    const int nq = pl->r2_nq, gs = pl->r2_gs, nsub = 32 / gs;
    const int slices128 = (c.n_feat + kFT - 1) / kFT;            // mask layout: four words per 128 columns
    const int hslices = (c.n_feat + 4 * gs - 1) / (4 * gs);      // hub role: slices of 4 * gs columns
    const int dslices = (c.n_feat + 32 * nq - 1) / (32 * nq);    // document role: slices of 32 * nq columns
    a.hub_slices = hslices;
    a.doc_slices = dslices;
    // Philox dropout: the keep mask is drawn by a separate ALU-bound kernel into a bit-packed side buffer (1 bit per
    // element) instead of inside the document role, whose CTAs have no issue slots to spare (ncu: the in-kernel RNG
    // cost 0.4 ms at 1M x 256).  Same mask, bit for bit, as the definition in tg_common.cuh.
    if constexpr (std::is_same<Epi, EpiStore>::value) {
      if (epi.drop_mode == 1) {
        const size_t part_bytes = partial_bytes(pl, c.n_feat);
        const size_t mask_bytes = (size_t)pl->n_rows * (size_t)slices128 * 16;
        TG_REQUIRE(c.workspace_bytes >= part_bytes + mask_bytes + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B (dropout mask)",
                   c.workspace_bytes, part_bytes + mask_bytes + 16);
        uint32_t* bits = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(c.workspace) + part_bytes + 15u) & ~(uintptr_t)15u);
        if (epi.keep_thr == kDropoutHalfThr) {
            const int n_words = slices128 * 4;
            const int64_t th = pl->n_rows * (int64_t)((n_words + 3) / 4);
            r2_keep_bits_half_kernel<<<(unsigned)ceil_div64(th, 256), 256, 0, st>>>(bits, pl->n_rows, n_words, epi.seed, epi.offset, epi.offset_dev);
        } else {
            const int n_blk = slices128 * 2;  // 64-column blocks, padded to whole 128-column slices
            const int64_t threads = pl->n_rows * (int64_t)n_blk * 8;
            int blk_shift = -1;
            for (int sh = 0; sh < 8; ++sh)
                if ((1 << sh) == n_blk) blk_shift = sh;
            r2_keep_bits_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, st>>>(bits, pl->n_rows, n_blk, blk_shift, epi.keep_thr, epi.seed,
                                                                                    epi.offset, epi.offset_dev);
        }
        TG_LAUNCH_CHECK();
        a.keep_bits = bits;
      }
    }
    // share of the SMs given to the hub role: both fronts must advance together so that the second reader of a row of B
    // hits L2 (split_sms).  Weights in shared-memory wavefronts / issue slots per 128 columns (measured per role, profiles/):
    // lock-stepped sub-groups and one-float4 document slices pay more instructions per entry and row.
    // (narrow-slice hub role: ~2.5x the cost per entry — thin (slot, chunk) runs walked in lock step.  Role-only runs at the C4
    // shape, 1.9 M documents, with the lock-step trips of both roles in PTX: hub 144 CTA-ms on 72 CTAs (160 on 64), documents
    // 149 CTA-ms — an even split again)
    const double hub_w = ((nsub > 1 ? 12.5 : 5.0) * (double)pl->hub_nnz + 4.0 * (double)a.groups * (double)pl->n_rows) / 0.82;
    // (document role, 128-column slices: 59 CTA-ms per 1 M documents in the role-only run of round 2 -> 0.715; narrower slices keep
    // the figure their sweeps were made with)
    double doc_w = (4.0 * (double)(pl->nnz - pl->hub_nnz - pl->n_rows) + 14.0 * (4.0 / nq) * (double)pl->n_rows) / (nq == 4 ? 0.715 : 0.67);
    // the row-wise loss epilogue (exp / log / shuffles on 8 lanes per row) more than doubles the document role's work on a
    // class-sized operand (C4 shard, role-only runs: 2.46 ms against 1.08 ms on the same CTAs)
    if (!std::is_same<Epi, EpiStore>::value) doc_w *= 2.4;
    // the dropout epilogue of the layer-1 forward (mask words, select, scale on 16 values per lane and row) makes the
    // document role ~10 % heavier: at C3 the fused forward runs 0.857 ms with the plain product's 70 + 78 split and 0.823 ms
    // with 64 + 84 (plain product: 0.770 / 0.808 ms) — the split follows the epilogue
    if constexpr (std::is_same<Epi, EpiStore>::value) {
        if (epi.drop_mode != 0) doc_w *= kDropDocWeight;
    }
    split_sms(hub_w, doc_w, hslices * a.groups, dslices, a.n_chunks, (a.n_jobs + 4 / nq - 1) / (4 / nq), pl->r2_hub_pct, &a.hub_lanes,
              &a.doc_lanes);
    const size_t need = (size_t)a.hub_lanes * a.Kv * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    TG_REQUIRE(make_tensor_map(&tmap, a.B, a.n, c.n_feat, a.ldb, std::min(a.T, 256), 4 * gs), TG_ERR_UNSUPPORTED,
               "cuTensorMapEncodeTiled failed (TMA tile of the dense operand)");
    CUtensorMap tmap_job;
    memset(&tmap_job, 0, sizeof(tmap_job));
    TG_REQUIRE(make_tensor_map(&tmap_job, a.B, a.n, c.n_feat, a.ldb, kJobRows, 32 * nq), TG_ERR_UNSUPPORTED,
               "cuTensorMapEncodeTiled failed (L2 prefetch tile of the dense operand)");
    const size_t smem = std::max(hub_smem(a.T, a.cap_hub, gs), doc_smem(a.Kh, a.cap_doc, nq, a.n_stages)) + 128;
    const unsigned grid = (unsigned)(a.hub_lanes * a.groups * hslices + a.doc_lanes * dslices);
    const int rc = launch_wide(gs, nq, a, epi, tmap, tmap_job, smem, grid, st);
    if (rc != TG_OK) return rc;
    FinishArgs f{a.partials, a.ldp, a.hub_lanes, a.Kh, a.Kv, pl->r2_vmap, pl->r2_vcnt, pl->hub_rows, a.n_chunks4};
    return finish_run(f, epi, st);
}

int roles2_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st) { return roles2_run_t(pl, c, epi, st); }
int roles2_run(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st) { return roles2_run_t(pl, c, epi, st); }

static void narrow_smem(const tg_plan* pl, int n_feat, size_t* hub_s, size_t* doc_s, size_t* lane_s) {
    const size_t rowb = (size_t)n_feat * 4;
    *hub_s = (size_t)kNStages * (align128((size_t)pl->r2_T * rowb) + align128((size_t)pl->r2_cap_hub * 8) + align128((size_t)kHtW * 4));
    const size_t job_s = align128((size_t)pl->r2_cap_doc * 8) + (size_t)kJobRows * 8 + align128((size_t)kJobRows * rowb);
    *doc_s = align128((size_t)pl->n_hub * rowb) + (size_t)kNRing * job_s;
    *lane_s = align128((size_t)pl->n_hub * rowb) + (size_t)kTeams * kTeamStages * job_s;
}

bool roles2_narrow_applicable(const tg_plan* pl, const StreamCall& c) {
    if (!pl || !pl->r2_ok || pl->r2_rect != 0 || !pl->r2_narrow_ok) return false;
    if (pl->r2_gs != 32) return false;  // (the narrow hub role is written for the warp-per-slot layout: <= 256 hub rows per group)
    if (c.n_feat < 4 || c.n_feat > 32 || c.n_feat % 4 != 0) return false;
    // a narrow operand of a small graph is L2 resident and the gather kernel is faster (20NG shape: 0.033 vs 0.052 ms)
    if (pl->n_rows < (int64_t)pl->r2_narrow_min_rows) return false;
    if (c.ldb % 4 != 0 || !aligned16(c.B) || !encode_tiled_fn()) return false;
    size_t hub_s, doc_s, lane_s;
    narrow_smem(pl, c.n_feat, &hub_s, &doc_s, &lane_s);
    return std::max(hub_s, doc_s) + 256 <= kSmemMax;  // (the lane-per-row variant is chosen at launch when it fits too)
}

template <class Epi>
static int roles2_narrow_run_t(const tg_plan* pl, const StreamCall& c, const Epi& epi, cudaStream_t st) {
    R2Args a;
    fill_common(a, pl, c);
    a.hub_slices = a.doc_slices = 1;
    // measured balance points at C3, F = 20: 59 % hub CTAs for the plain product, 54 % when the document role also runs the
    // log-softmax / cross-entropy epilogue (57 / 52 before the epilogues skipped unlabelled rows and per-chunk flag tests)
    const int hub_pct = pl->r2_narrow_hub_pct >= 0 ? pl->r2_narrow_hub_pct : (std::is_same<Epi, EpiLoss>::value ? 54 : 59);
    int hub_lanes = (kNumSM * hub_pct / 100) / a.groups;
    if (hub_lanes < 1) hub_lanes = 1;
    if (hub_lanes > a.n_chunks) hub_lanes = a.n_chunks;
    int doc_lanes = kNumSM - hub_lanes * a.groups;
    if (doc_lanes < 1) doc_lanes = 1;
    if (doc_lanes > a.n_jobs) doc_lanes = a.n_jobs;
    a.hub_lanes = hub_lanes;
    a.doc_lanes = doc_lanes;
    const size_t need = (size_t)hub_lanes * a.Kv * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    TG_REQUIRE(make_tensor_map(&tmap, a.B, a.n, c.n_feat, a.ldb, a.T, c.n_feat), TG_ERR_UNSUPPORTED,
               "cuTensorMapEncodeTiled failed (TMA tile of the narrow dense operand)");
    size_t hub_s, doc_s, lane_s;
    narrow_smem(pl, c.n_feat, &hub_s, &doc_s, &lane_s);
    const bool lane_rows = c.ldb == c.n_feat && std::max(hub_s, lane_s) + 256 <= kSmemMax && pl->r2_narrow_lane != 0;
    const unsigned grid = (unsigned)(hub_lanes * a.groups + doc_lanes);
    if (lane_rows) {
        const size_t smem = std::max(hub_s, lane_s) + 128;
        const int n4 = a.n_chunks4;
#define TG_R2NL(NVv)                                                                                                         \
        if (n4 <= NVv) {                                                                                                     \
            TG_CUDA(cudaFuncSetAttribute(roles2nl_kernel<NVv, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            roles2nl_kernel<NVv, Epi><<<grid, kThreads, smem, st>>>(a, epi, tmap);                                          \
        } else
        TG_R2NL(2) TG_R2NL(4) TG_R2NL(5) TG_R2NL(6) TG_R2NL(8) { set_error("narrow role kernels: n_feat > 32"); return TG_ERR_UNSUPPORTED; }
#undef TG_R2NL
    } else {
        TG_REQUIRE(std::max(hub_s, doc_s) + 256 <= kSmemMax, TG_ERR_UNSUPPORTED, "narrow role kernels: shared memory budget exceeded");
        const size_t smem = std::max(hub_s, doc_s) + 128;
        TG_CUDA(cudaFuncSetAttribute(roles2n_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roles2n_kernel<Epi><<<grid, kThreads, smem, st>>>(a, epi, tmap);
    }
    TG_LAUNCH_CHECK();
    FinishArgs f{a.partials, a.ldp, hub_lanes, a.Kh, a.Kv, pl->r2_vmap, pl->r2_vcnt, pl->hub_rows, a.n_chunks4};
    return finish_run(f, epi, st);
}

int roles2_narrow_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st) {
    return roles2_narrow_run_t(pl, c, epi, st);
}
int roles2_narrow_run(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st) {
    return roles2_narrow_run_t(pl, c, epi, st);
}

// ---- rectangular operands ---------------------------------------------------------------------------------------------------
// table mode: every entry of a row addresses the resident table (rows of B = the dense weight): {column * 128, value}
__global__ void r2_table_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                     const int32_t* __restrict__ start, int64_t n, int64_t n_pad, int2* __restrict__ dent2,
                                     int2* __restrict__ rdesc, int2* __restrict__ jdesc) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    const int64_t j0 = r / kJobRows * kJobRows;
    const int jbase = start[j0] & ~1;
    int2 rd = make_int2(0, kNotShort);
    if (r < n) {
        const int s = rowptr[r], len = rowptr[r + 1] - s;
        const int o = start[r];
        for (int p = 0; p < len; ++p) dent2[o + p] = make_int2(colidx[s + p] * kHubEnc, __float_as_int(vals[s + p]));
        rd = make_int2(o - jbase, len << 16);  // (an all-zero feature row is produced too: the epilogue of zero)
    }
    rdesc[r] = rd;
    if (r == j0) {
        const int64_t j1 = (j0 + kJobRows < n_pad) ? j0 + kJobRows : n_pad;
        jdesc[r / kJobRows] = make_int2(jbase, ((start[j1] - jbase) + 1) & ~1);
    }
}

__global__ void r2_iota_kernel(int32_t* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// Sub-plans for the two products of a sparse feature matrix X [n x nfeat], nfeat <= 1280 (reference layer.py:102 with the
// topic features of trainer.py:197-238, and its autograd transpose product):
//   X * W      (plan of X):   every column's row of W is resident in shared memory -> the document role alone;
//   X^T * dS   (plan of X^T): every row is a hub row                               -> the hub role alone + finish.
int roles2_rect_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* h_rowptr,
                           cudaStream_t st) {
    read_knobs(pl);
    if (pl->r2_ok || env_int2("TG_ROLES2", 1) == 0 || env_int2("TG_ROLES2_RECT", 1) == 0) return TG_OK;
    if (!colidx || !vals || pl->nnz == 0 || pl->nnz >= (int64_t)0x7fffffff) return TG_OK;
    const int64_t min_rows = pl->r2_min_rows;
    const int max_k = 1280;  // resident table rows / hub rows of the transposed product
    if (pl->n_rows <= max_k && pl->n_hub == pl->n_rows && pl->n_cols >= min_rows && pl->n_cols > pl->n_rows) {
        // ---- all-hub mode: the hub side of the square plan, over the columns of this matrix ----
        std::vector<int32_t> hub_rows((size_t)pl->n_rows);
        for (int64_t i = 0; i < pl->n_rows; ++i) hub_rows[(size_t)i] = (int32_t)i;
        pl->r2_rect = 2;
        const int rc = roles2_plan_build(pl, rowptr, colidx, vals, h_rowptr, hub_rows.data(), st);
        if (rc != TG_OK || !pl->r2_ok) pl->r2_rect = 0;
        return rc;
    }
    if (pl->n_cols <= max_k && pl->n_hub == 0 && pl->n_rows >= min_rows && pl->n_rows > pl->n_cols && pl->max_row_nnz <= 16383) {
        // ---- table mode: compact entries + row / job descriptors, the slice width and ring depth that fit next to the table ----
        const int64_t n = pl->n_rows;
        const int64_t n_jobs = ceil_div64(n, kJobRows), n_pad = n_jobs * kJobRows;
        int32_t *d_len = nullptr, *d_start = nullptr;
        void* tmp = nullptr;
        cudaError_t e = cudaSuccess;
        auto fail = [&](cudaError_t err, const char* what) {
            cudaFree(d_len); cudaFree(d_start); cudaFree(tmp);
            roles2_plan_free(pl);
            return cuda_fail(err, what, __FILE__, __LINE__);
        };
#define TG_TRY(call) do { e = (call); if (e != cudaSuccess) return fail(e, #call); } while (0)
        TG_TRY(cudaMalloc((void**)&d_len, (size_t)(n_pad + 1) * sizeof(int32_t)));
        TG_TRY(cudaMalloc((void**)&d_start, (size_t)(n_pad + 1) * sizeof(int32_t)));
        r2_row_len_kernel<<<(unsigned)ceil_div64(n_pad + 1, 256), 256, 0, st>>>(rowptr, n, n_pad, 0x7fffffff, d_len);
        TG_TRY(cudaGetLastError());
        size_t scan_bytes = 0;
        TG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_len, d_start, (int)(n_pad + 1), st));
        TG_TRY(cudaMalloc(&tmp, scan_bytes ? scan_bytes : 1));
        TG_TRY(cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, d_len, d_start, (int)(n_pad + 1), st));
        TG_TRY(cudaMalloc((void**)&pl->r2_dent, ((size_t)pl->nnz + 2) * sizeof(int2)));
        TG_TRY(cudaMemsetAsync(pl->r2_dent, 0, ((size_t)pl->nnz + 2) * sizeof(int2), st));
        TG_TRY(cudaMalloc((void**)&pl->r2_rdesc, (size_t)n_pad * sizeof(int2)));
        TG_TRY(cudaMalloc((void**)&pl->r2_jdesc, (size_t)(n_jobs + 8) * sizeof(int2)));
        TG_TRY(cudaMemsetAsync(pl->r2_jdesc, 0, (size_t)(n_jobs + 8) * sizeof(int2), st));
        r2_table_fill_kernel<<<(unsigned)ceil_div64(n_pad, 256), 256, 0, st>>>(rowptr, colidx, vals, d_start, n, n_pad, pl->r2_dent,
                                                                               pl->r2_rdesc, pl->r2_jdesc);
        TG_TRY(cudaGetLastError());
        TG_TRY(cudaMalloc((void**)&pl->r2_ident, (size_t)pl->n_cols * sizeof(int32_t)));
        r2_iota_kernel<<<(unsigned)ceil_div64(pl->n_cols, 256), 256, 0, st>>>(pl->r2_ident, (int)pl->n_cols);
        TG_TRY(cudaGetLastError());
        std::vector<int2> h_jdesc((size_t)n_jobs);
        TG_TRY(cudaMemcpyAsync(h_jdesc.data(), pl->r2_jdesc, (size_t)n_jobs * sizeof(int2), cudaMemcpyDeviceToHost, st));
        TG_TRY(cudaStreamSynchronize(st));
#undef TG_TRY
        cudaFree(d_len); cudaFree(d_start); cudaFree(tmp);
        int cap_doc = 2;
        for (const int2& d : h_jdesc) cap_doc = std::max(cap_doc, d.y);
        pl->r2_n_jobs = (int32_t)n_jobs;
        pl->r2_cap_doc = cap_doc;
        int nq = 0, stages = 0;
        if (!pick_doc_cfg((int)pl->n_cols, cap_doc, env_int2("TG_ROLES2_NQ", 0), &nq, &stages)) {
            roles2_plan_free(pl);
            return TG_OK;
        }
        pl->r2_nq = nq;
        pl->r2_stages = stages;
        pl->r2_rect = 1;
        pl->r2_ok = true;
    }
    return TG_OK;
}

bool roles2_rect_applicable(const tg_plan* pl, const StreamCall& c) {
    if (!pl || !pl->r2_ok || pl->r2_rect == 0) return false;
    if (c.n_feat < 64 || c.n_feat % 4 != 0 || c.n_feat > 1024) return false;
    if (c.ldb % 4 != 0 || !aligned16(c.B) || !encode_tiled_fn()) return false;
    return true;
}

int roles2_rect_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st) {
    R2Args a;
    fill_common(a, pl, c);
    a.only_role = 0;
    CUtensorMap tmap, tmap_job;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap_job, 0, sizeof(tmap_job));
    if (pl->r2_rect == 1) {
        // X * W: the rows of W (c.B, n_cols of them) are the resident table; all CTAs run the document role
        const int nq = pl->r2_nq;
        const int dslices = (c.n_feat + 32 * nq - 1) / (32 * nq);
        a.hub_slices = 1;
        a.doc_slices = dslices;
        a.hub_rows = pl->r2_ident; a.Kh = (int32_t)pl->n_cols;
        a.groups = 1; a.Kv = kKv;
        a.hub_lanes = 0;
        int doc_lanes = kNumSM / dslices;
        const int n_sj = (a.n_jobs + 4 / nq - 1) / (4 / nq);
        if (doc_lanes < 1) doc_lanes = 1;
        if (doc_lanes > n_sj) doc_lanes = n_sj;
        a.doc_lanes = doc_lanes;
        const size_t smem = doc_smem(a.Kh, a.cap_doc, nq, a.n_stages) + 128;
        return launch_table(nq, a, epi, tmap, tmap_job, smem, (unsigned)(doc_lanes * dslices), st);
    }
    // X^T * dS: every row of this matrix is a hub row; all CTAs run the hub role over the rows of c.B, then the finish
    const int gs = pl->r2_gs;
    const int hslices = (c.n_feat + 4 * gs - 1) / (4 * gs);
    a.hub_slices = a.doc_slices = hslices;
    a.doc_lanes = 0;
    int hub_lanes = kNumSM / (hslices * a.groups);
    if (hub_lanes < 1) hub_lanes = 1;
    if (hub_lanes > a.n_chunks) hub_lanes = a.n_chunks;
    a.hub_lanes = hub_lanes;
    const size_t need = (size_t)hub_lanes * a.Kv * a.ldp * sizeof(float);
    TG_REQUIRE(c.workspace && c.workspace_bytes >= need + 16, TG_ERR_WORKSPACE, "workspace %zu B < required %zu B",
               c.workspace_bytes, need + 16);
    TG_REQUIRE(make_tensor_map(&tmap, a.B, pl->n_cols, c.n_feat, a.ldb, std::min(a.T, 256), 4 * gs), TG_ERR_UNSUPPORTED,
               "cuTensorMapEncodeTiled failed (TMA tile of the dense operand)");
    const size_t smem = hub_smem(a.T, a.cap_hub, gs) + 128;
    const int rc = launch_wide(gs, 4, a, epi, tmap, tmap_job, smem, (unsigned)(hub_lanes * a.groups * hslices), st);
    if (rc != TG_OK) return rc;
    FinishArgs f{a.partials, a.ldp, hub_lanes, a.Kh, a.Kv, pl->r2_vmap, pl->r2_vcnt, pl->hub_rows, a.n_chunks4};
    return finish_run(f, epi, st);
}

}  // namespace tg
