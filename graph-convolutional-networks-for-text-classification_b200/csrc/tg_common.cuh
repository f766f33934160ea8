// tg_common.cuh — shared helpers for the topicgcn sm_100a kernels (error plumbing, Philox, vector ld/st).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/topicgcn.h"

namespace tg {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TG_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return tg::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define TG_REQUIRE(cond, code, ...)   \
    do {                              \
        if (!(cond)) {                \
            tg::set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

#define TG_LAUNCH_CHECK() TG_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

// ---- plan (opaque in the C header) ------------------------------------------------------------------
}  // namespace tg

struct tg_plan {
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    int32_t hub_threshold = 0, segment_nnz = 0;
    int32_t n_hub = 0, n_seg = 0;
    int64_t hub_nnz = 0;
    int32_t max_row_nnz = 0;
    // device tables (owned)
    int32_t* hub_rows = nullptr;     // [n_hub]   row id of each split row
    int32_t* hub_seg_ptr = nullptr;  // [n_hub+1] first segment of each split row
    int32_t* seg_hub = nullptr;      // [n_seg]   hub slot a segment belongs to
    int32_t* seg_begin = nullptr;    // [n_seg]   first stored entry of the segment
    int32_t* seg_end = nullptr;      // [n_seg]   one past the last stored entry
    uint32_t* tickets = nullptr;     // [n_hub]   arrival counters (integer; reset by the last arriver)
    int32_t* seg_order = nullptr;    // [n_seg]   execution order: segments sorted by their first column (L2 reuse of B)

    // ---- role-specialised column-chunk streaming kernels (tg_roles2.cu) ----------------------------------------------
    // Square graphs with a compact hub set (the document-topic-topic graphs, up to ~1 400 topic rows).  A hub CTA owns a
    // GROUP of 256 * 32 / r2_gs slots of one column slice of 4 * r2_gs columns; the document role keeps all hub rows of B
    // resident in shared memory and therefore works on column slices of 32 * r2_nq columns (r2_nq = 4 / 2 / 1 for up to
    // ~368 / ~736 / ~1 400 hub rows).
    bool r2_ok = false;
    // rectangular operands (sparse feature matrices X [n x nfeat] and their transposes):
    //   1 = "table": every column's row of B is resident (document role only: X * W)
    //   2 = "all hub": every row is a hub row (hub role only: X^T * dS)
    int32_t r2_rect = 0;
    int32_t r2_gs = 32;              // lanes per hub slot sub-group (hub_role<GS>): 32 / 16 / 8 -> 256 / 512 / 1024 slots per group, 128 / 64 / 32-column slices
    int32_t r2_groups = 1;           // hub slot groups (one hub CTA per group, slice and chunk lane)
    int32_t r2_nq = 4;               // float4 chunks per lane of the document role (slice = 32 * r2_nq columns)
    int32_t r2_stages = 4;           // document-role ring depth that fits next to the resident rows
    int32_t* r2_ident = nullptr;     // [n_cols] 0,1,2,... (the table rows of mode 1)
    int32_t r2_T = 0, r2_n_chunks = 0, r2_cap_hub = 0;   // hub role: nodes per chunk, chunks, staged entries per (chunk, group) (max)
    int32_t r2_n_jobs = 0, r2_cap_doc = 0;               // document role: jobs of 64 rows, staged entries per job (max)
    int2* r2_hent = nullptr;         // [hub_nnz+2] hub entries, (group, chunk, slot)-major: {byte offset of the column's row in the staged tile, value bits}
    int32_t* r2_htab = nullptr;      // [r2_groups * r2_n_chunks][260] offsets of the 256 slots relative to the (even-aligned) base of the (group, chunk) run
    int4* r2_cdesc = nullptr;        // [r2_groups * r2_n_chunks] {aligned base into r2_hent, staged entry count (even), first node, -}
    int32_t* r2_vmap = nullptr;      // [n_hub][8] slots of each hub row (slot = group * 256 + warp * 16 + position)
    int32_t* r2_vcnt = nullptr;      // [n_hub]
    int2* r2_dent = nullptr;         // compact entries of the short rows, per row: other columns {col, v}, then hub columns {hub * 128 * r2_nq, v}
    int2* r2_rdesc = nullptr;        // [r2_n_jobs*64] {first entry relative to the job's base, n_entries << 16 | n_other}; y < 0: not a short row
    int2* r2_jdesc = nullptr;        // [r2_n_jobs] {aligned base into r2_dent, staged entry count (even)}
    // knobs, read from the environment ONCE at plan creation (profiling / tests), never on the hot path
    int32_t r2_min_rows = 16384, r2_narrow_min_rows = 131072, r2_hub_pct = -1, r2_narrow_hub_pct = -1, r2_only_role = 0;
    int32_t r2_narrow_lane = 1;
    bool r2_narrow_ok = true;
    int32_t r2_hub_pf = -1;          // hub role: tiles prefetched into L2 ahead (chunks per lane); -1 = chosen at plan build
};

namespace tg {

// ---- Philox4x32-10 (Salmon et al., SC'11) — counter-based RNG for the dropout keep mask -------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                           uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// Keep-mask definition shared by every kernel (and restated in oracle/gcn_oracle.py):
//   element (row, col):  q = col / 4 (float4 chunk), blk = q / 16 (64-column block), lane8 = q % 8, half = (q / 8) % 2
//   r = philox(counter = (row_lo, row_hi, lane8 | blk << 8, offset_lo), key = (seed_lo ^ offset_hi, seed_hi))
//   u16 #(half*4 + col%4) of the 128 random bits;  keep  <=>  u16 < keep_threshold(p)
// One Philox call therefore serves the two float4 chunks (q, q + 8) a lane of the streaming kernel owns.
__host__ __device__ __forceinline__ uint32_t dropout_keep_threshold(float p) {
    float t = (1.0f - p) * 65536.0f + 0.5f;
    if (t < 0.f) t = 0.f;
    if (t > 65536.f) t = 65536.f;
    return (uint32_t)t;
}

__host__ __device__ __forceinline__ Philox4 dropout_philox(int64_t row, uint32_t lane8, uint32_t blk,
                                                            uint64_t seed, uint64_t offset) {
    return philox4x32_10((uint32_t)(uint64_t)row, (uint32_t)((uint64_t)row >> 32), lane8 | (blk << 8),
                         (uint32_t)offset, (uint32_t)seed ^ (uint32_t)(offset >> 32),
                         (uint32_t)(seed >> 32));
}

// Exact-half mode: when keep_threshold(p) == 32768 (p = 0.5, the reference's setting, trainer.py:428) one random BIT per
// element is an exact Bernoulli(1/2) draw, so a Philox call serves 128 consecutive columns instead of 8:
//   element (row, col):  r = philox(counter = (row_lo, row_hi, (col / 128) | 0x80000000, offset_lo), same key)
//   keep  <=>  bit (col % 128) of the 128-bit little-endian value (x | y << 32 | z << 64 | w << 96) is set.
// Every other p keeps the 16-bit definition above.  (Both are restated in oracle/gcn_oracle.py.)
constexpr uint32_t kDropoutHalfThr = 32768u;
__host__ __device__ __forceinline__ Philox4 dropout_philox_half(int64_t row, uint32_t cidx, uint64_t seed, uint64_t offset) {
    return philox4x32_10((uint32_t)(uint64_t)row, (uint32_t)((uint64_t)row >> 32), cidx | 0x80000000u,
                         (uint32_t)offset, (uint32_t)seed ^ (uint32_t)(offset >> 32),
                         (uint32_t)(seed >> 32));
}
// the 32-bit word of a half-mode result that holds float4 chunk q (columns 4q .. 4q+3): word (q / 8) % 4, bits 4 (q % 8) ..
__host__ __device__ __forceinline__ uint32_t dropout_half_word(const Philox4& r, int q) {
    const int w = (q >> 3) & 3;
    return w == 0 ? r.x : w == 1 ? r.y : w == 2 ? r.z : r.w;
}

// the four u16 lanes that belong to chunk-half `half` (0/1) of a Philox result
__host__ __device__ __forceinline__ void dropout_u16x4(const Philox4& r, int half, uint32_t u[4]) {
    const uint32_t a = half ? r.z : r.x, b = half ? r.w : r.y;
    u[0] = a & 0xFFFFu; u[1] = a >> 16; u[2] = b & 0xFFFFu; u[3] = b >> 16;
}

// ---- vector helpers ----------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg_f4(const float* p) {  // read-only path, 128-bit
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {  // streamed once: do not keep in L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_f4(float* p, const float4& v) {
    *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void st_f4_stream(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void fma4(float4& acc, float a, const float4& b) {
    acc.x = fmaf(a, b.x, acc.x);
    acc.y = fmaf(a, b.y, acc.y);
    acc.z = fmaf(a, b.z, acc.z);
    acc.w = fmaf(a, b.w, acc.w);
}
__device__ __forceinline__ void add4(float4& acc, const float4& b) {
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
}
#endif

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace tg
