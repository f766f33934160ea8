// tg_csr.cu — plan-time graph preparation: torch COO -> device CSR, CSR transpose, skew plan.
//
// Replaces what ATen does on EVERY torch.spmm(adj, support) call on CUDA (coalesce() sort + COO->CSR
// row pointer, reference layer.py:106 -> s_addmm_out_sparse_dense_cuda) with a once-per-adjacency
// conversion.  The adjacency produced by utils.preprocess_adj (reference utils.py:185-213) is already
// row-major sorted and duplicate free, so the common case is a sort-free pass; arbitrary COO input gets
// coalesce() semantics (stable radix sort on (row, col), duplicates summed in input order).
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

#include "tg_roles.cuh"

namespace tg {

// flag bits used on the device
constexpr int kFlagUnsorted = 1;
constexpr int kFlagOutOfRange = 2;
constexpr int kFlagMismatch = 4;

__global__ void coo_check_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                 int64_t nnz, int64_t n_rows, int64_t n_cols, int* __restrict__ flags) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int64_t r = rows[p], c = cols[p];
    int f = 0;
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) f |= kFlagOutOfRange;
    if (p > 0) {
        const int64_t rp = rows[p - 1], cp = cols[p - 1];
        if (rp > r || (rp == r && cp >= c)) f |= kFlagUnsorted;
    }
    if (f) atomicOr(flags, f);
}

// rowptr from a sorted row-id stream: thread p owns the boundary between entry p-1 and entry p.
template <typename RowT>
__global__ void rowptr_from_sorted_rows_kernel(const RowT* __restrict__ rows, int64_t nnz, int64_t n_rows,
                                               int32_t* __restrict__ rowptr) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nnz) return;
    const int64_t r_prev = (p == 0) ? -1 : (int64_t)rows[p - 1];
    const int64_t r_cur = (p == nnz) ? n_rows : (int64_t)rows[p];
    for (int64_t q = r_prev + 1; q <= r_cur; ++q) rowptr[q] = (int32_t)p;
}

__global__ void narrow_copy_kernel(const int64_t* __restrict__ cols, const float* __restrict__ vals,
                                   int64_t nnz, int32_t* __restrict__ colidx, float* __restrict__ vals_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    colidx[p] = (int32_t)cols[p];
    vals_out[p] = vals[p];
}

__global__ void make_keys_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                 int64_t nnz, int64_t n_cols, uint64_t* __restrict__ keys,
                                 int32_t* __restrict__ perm) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    keys[p] = (uint64_t)rows[p] * (uint64_t)n_cols + (uint64_t)cols[p];
    perm[p] = (int32_t)p;
}

__global__ void mark_heads_kernel(const uint64_t* __restrict__ keys, int64_t nnz, int32_t* __restrict__ head) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    head[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1 : 0;
}

// one thread per distinct (row, col): sums its duplicates sequentially in input order (stable sort).
__global__ void merge_duplicates_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ perm,
                                        const int32_t* __restrict__ head, const int32_t* __restrict__ pos,
                                        const float* __restrict__ vals, int64_t nnz, int64_t n_cols,
                                        int32_t* __restrict__ out_rows, int32_t* __restrict__ colidx,
                                        float* __restrict__ vals_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz || !head[p]) return;
    const uint64_t k = keys[p];
    float s = vals[perm[p]];
    for (int64_t q = p + 1; q < nnz && keys[q] == k; ++q) s += vals[perm[q]];
    const int32_t o = pos[p];
    out_rows[o] = (int32_t)(k / (uint64_t)n_cols);
    colidx[o] = (int32_t)(k % (uint64_t)n_cols);
    vals_out[o] = s;
}

__global__ void expand_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                   int64_t* __restrict__ rows64) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    for (int32_t p = rowptr[r]; p < rowptr[r + 1]; ++p) rows64[p] = r;
}

__global__ void widen_cols_kernel(const int32_t* __restrict__ colidx, int64_t nnz, int64_t* __restrict__ cols64) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) cols64[p] = colidx[p];
}

__global__ void compare_csr_kernel(const int32_t* __restrict__ a_ptr, const int32_t* __restrict__ b_ptr,
                                   int64_t n_ptr, const int32_t* __restrict__ a_col,
                                   const int32_t* __restrict__ b_col, const float* __restrict__ a_val,
                                   const float* __restrict__ b_val, int64_t nnz, int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n_ptr && a_ptr[i] != b_ptr[i]) bad = true;
    if (i < nnz && (a_col[i] != b_col[i] || __float_as_uint(a_val[i]) != __float_as_uint(b_val[i]))) bad = true;
    if (bad) atomicOr(flags, kFlagMismatch);
}

__global__ void seg_first_col_kernel(const int32_t* __restrict__ colidx, const int32_t* __restrict__ seg_begin, int n_seg,
                                     int32_t* __restrict__ first_col) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) first_col[s] = colidx[seg_begin[s]];
}

struct DevBuf {  // plan-time scratch with RAII
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as() { return reinterpret_cast<T*>(p); }
};

static inline unsigned grid1d(int64_t n, int block = 256) { return (unsigned)ceil_div64(n > 0 ? n : 1, block); }

static int coo_to_csr(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz,
                      int64_t n_rows, int64_t n_cols, int32_t* rowptr, int32_t* colidx, float* vals_out,
                      int64_t* nnz_out_host, uint32_t* flags_host, cudaStream_t st) {
    TG_REQUIRE(nnz >= 0 && n_rows >= 0 && n_cols >= 0, TG_ERR_INVALID_ARG, "negative size");
    TG_REQUIRE(nnz < (int64_t)INT32_MAX && n_rows < (int64_t)INT32_MAX && n_cols < (int64_t)INT32_MAX,
               TG_ERR_OVERFLOW, "nnz/n_rows/n_cols must fit int32 (nnz=%lld)", (long long)nnz);
    TG_REQUIRE(rowptr && (nnz == 0 || (rows && cols && vals && colidx && vals_out)), TG_ERR_INVALID_ARG,
               "null pointer");
    uint32_t out_flags = 0;
    if (nnz == 0) {
        TG_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_rows + 1) * sizeof(int32_t), st));
        TG_CUDA(cudaStreamSynchronize(st));
        if (nnz_out_host) *nnz_out_host = 0;
        if (flags_host) *flags_host = TG_COO_WAS_SORTED;
        return TG_OK;
    }
    DevBuf dflags;
    TG_CUDA(dflags.alloc(sizeof(int)));
    TG_CUDA(cudaMemsetAsync(dflags.p, 0, sizeof(int), st));
    coo_check_kernel<<<grid1d(nnz), 256, 0, st>>>(rows, cols, nnz, n_rows, n_cols, dflags.as<int>());
    TG_LAUNCH_CHECK();
    int hflags = 0;
    TG_CUDA(cudaMemcpyAsync(&hflags, dflags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    TG_REQUIRE(!(hflags & kFlagOutOfRange), TG_ERR_INVALID_ARG, "COO index out of range");

    if (!(hflags & kFlagUnsorted)) {
        // fast path: the layout utils.sparse_mx_to_torch_sparse_tensor emits (scipy tocoo() of a CSR)
        narrow_copy_kernel<<<grid1d(nnz), 256, 0, st>>>(cols, vals, nnz, colidx, vals_out);
        TG_LAUNCH_CHECK();
        rowptr_from_sorted_rows_kernel<int64_t><<<grid1d(nnz + 1), 256, 0, st>>>(rows, nnz, n_rows, rowptr);
        TG_LAUNCH_CHECK();
        TG_CUDA(cudaStreamSynchronize(st));
        out_flags |= TG_COO_WAS_SORTED;
        if (nnz_out_host) *nnz_out_host = nnz;
        if (flags_host) *flags_host = out_flags;
        return TG_OK;
    }

    // general path: coalesce() semantics
    DevBuf keys_a, keys_b, perm_a, perm_b, head, pos, out_rows, tmp;
    TG_CUDA(keys_a.alloc(nnz * sizeof(uint64_t)));
    TG_CUDA(keys_b.alloc(nnz * sizeof(uint64_t)));
    TG_CUDA(perm_a.alloc(nnz * sizeof(int32_t)));
    TG_CUDA(perm_b.alloc(nnz * sizeof(int32_t)));
    TG_CUDA(head.alloc(nnz * sizeof(int32_t)));
    TG_CUDA(pos.alloc(nnz * sizeof(int32_t)));
    TG_CUDA(out_rows.alloc(nnz * sizeof(int32_t)));
    make_keys_kernel<<<grid1d(nnz), 256, 0, st>>>(rows, cols, nnz, n_cols, keys_a.as<uint64_t>(),
                                                 perm_a.as<int32_t>());
    TG_LAUNCH_CHECK();
    int end_bit = 1;
    {
        const unsigned __int128 max_key = (unsigned __int128)n_rows * (unsigned __int128)n_cols;
        TG_REQUIRE(max_key < ((unsigned __int128)1 << 63), TG_ERR_OVERFLOW, "n_rows*n_cols overflows 63 bits");
        while (end_bit < 64 && ((unsigned __int128)1 << end_bit) < max_key) ++end_bit;
    }
    size_t tmp_bytes = 0;
    TG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(),
                                            perm_a.as<int32_t>(), perm_b.as<int32_t>(), (int)nnz, 0, end_bit, st));
    size_t scan_bytes = 0;
    TG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, head.as<int32_t>(), pos.as<int32_t>(), (int)nnz, st));
    TG_CUDA(tmp.alloc(tmp_bytes > scan_bytes ? tmp_bytes : scan_bytes));
    TG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(),
                                            perm_a.as<int32_t>(), perm_b.as<int32_t>(), (int)nnz, 0, end_bit, st));
    mark_heads_kernel<<<grid1d(nnz), 256, 0, st>>>(keys_b.as<uint64_t>(), nnz, head.as<int32_t>());
    TG_LAUNCH_CHECK();
    TG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, head.as<int32_t>(), pos.as<int32_t>(), (int)nnz, st));
    int32_t last_pos = 0, last_head = 0;
    TG_CUDA(cudaMemcpyAsync(&last_pos, pos.as<int32_t>() + (nnz - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaMemcpyAsync(&last_head, head.as<int32_t>() + (nnz - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    const int64_t n_out = (int64_t)last_pos + last_head;
    merge_duplicates_kernel<<<grid1d(nnz), 256, 0, st>>>(keys_b.as<uint64_t>(), perm_b.as<int32_t>(),
                                                        head.as<int32_t>(), pos.as<int32_t>(), vals, nnz, n_cols,
                                                        out_rows.as<int32_t>(), colidx, vals_out);
    TG_LAUNCH_CHECK();
    rowptr_from_sorted_rows_kernel<int32_t><<<grid1d(n_out + 1), 256, 0, st>>>(out_rows.as<int32_t>(), n_out,
                                                                              n_rows, rowptr);
    TG_LAUNCH_CHECK();
    TG_CUDA(cudaStreamSynchronize(st));
    if (n_out != nnz) out_flags |= TG_COO_HAD_DUPLICATES;
    if (nnz_out_host) *nnz_out_host = n_out;
    if (flags_host) *flags_host = out_flags;
    return TG_OK;
}

}  // namespace tg

extern "C" {

int tg_csr_from_coo(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz, int64_t n_rows,
                    int64_t n_cols, int32_t* rowptr, int32_t* colidx, float* vals_out, int64_t* nnz_out_host,
                    uint32_t* flags_host, void* stream) {
    return tg::coo_to_csr(rows, cols, vals, nnz, n_rows, n_cols, rowptr, colidx, vals_out, nnz_out_host,
                          flags_host, tg::as_stream(stream));
}

int tg_csr_transpose(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                     int64_t n_cols, int64_t nnz, int32_t* t_rowptr, int32_t* t_colidx, float* t_vals,
                     int32_t* is_symmetric_host, void* stream) {
    using namespace tg;
    cudaStream_t st = as_stream(stream);
    TG_REQUIRE(rowptr && t_rowptr && (nnz == 0 || (colidx && vals && t_colidx && t_vals)), TG_ERR_INVALID_ARG,
               "null pointer");
    DevBuf rows64, cols64;
    TG_CUDA(rows64.alloc(nnz * sizeof(int64_t)));
    TG_CUDA(cols64.alloc(nnz * sizeof(int64_t)));
    if (nnz > 0) {
        expand_rows_kernel<<<grid1d(n_rows), 256, 0, st>>>(rowptr, n_rows, rows64.as<int64_t>());
        TG_LAUNCH_CHECK();
        widen_cols_kernel<<<grid1d(nnz), 256, 0, st>>>(colidx, nnz, cols64.as<int64_t>());
        TG_LAUNCH_CHECK();
    }
    int64_t n_out = 0;
    uint32_t flags = 0;
    // transposed COO = (col, row); a valid CSR has no duplicates so n_out == nnz
    int rc = coo_to_csr(cols64.as<int64_t>(), rows64.as<int64_t>(), vals, nnz, n_cols, n_rows, t_rowptr, t_colidx,
                        t_vals, &n_out, &flags, st);
    if (rc != TG_OK) return rc;
    TG_REQUIRE(n_out == nnz, TG_ERR_INVALID_ARG, "input CSR holds duplicate entries");
    if (is_symmetric_host) {
        *is_symmetric_host = 0;
        if (n_rows == n_cols) {
            DevBuf dflags;
            TG_CUDA(dflags.alloc(sizeof(int)));
            TG_CUDA(cudaMemsetAsync(dflags.p, 0, sizeof(int), st));
            const int64_t n = (n_rows + 1 > nnz) ? n_rows + 1 : nnz;
            compare_csr_kernel<<<grid1d(n), 256, 0, st>>>(rowptr, t_rowptr, n_rows + 1, colidx, t_colidx, vals,
                                                         t_vals, nnz, dflags.as<int>());
            TG_LAUNCH_CHECK();
            int h = 0;
            TG_CUDA(cudaMemcpyAsync(&h, dflags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            TG_CUDA(cudaStreamSynchronize(st));
            *is_symmetric_host = (h == 0) ? 1 : 0;
        }
    }
    return TG_OK;
}

// ---- skew plan -----------------------------------------------------------------------------------------
int tg_plan_create(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows, int64_t n_cols,
                   int64_t nnz, int32_t hub_threshold, int32_t segment_nnz, tg_plan** plan_out, void* stream) {
    using namespace tg;
    cudaStream_t st = as_stream(stream);
    TG_REQUIRE(rowptr && plan_out, TG_ERR_INVALID_ARG, "null pointer");
    TG_REQUIRE(n_rows >= 0 && n_rows < (int64_t)INT32_MAX && nnz < (int64_t)INT32_MAX, TG_ERR_OVERFLOW,
               "n_rows/nnz must fit int32");
    if (segment_nnz <= 0 && hub_threshold <= 0 && n_rows > 0 &&
        ((nnz <= (int64_t)32 << 20 && nnz / n_rows >= 24) || nnz <= (int64_t)4 << 20)) {
        // Small graphs (R8 / 20NG shapes: a step is ~0.1 ms of kernels) take the same budget: a 256-entry segment is 64
        // dependent batches of L2 gathers for one warp — the critical path of the whole launch — where 64 entries are 16.
        // Graphs whose TYPICAL row is long (TextGCN-style word rows: median ~200 entries, reference build_graph.py PMI edges):
        // with the 512-entry threshold almost every row would be a "short" row walked serially by one lane group (a 2-lane
        // group for an 8-column operand).  A budget of 64 entries per segment turns them into warp-cooperative segments —
        // the merge-path treatment of mid-degree rows; the partial rows stay small because such graphs are small.
        segment_nnz = 64;
        hub_threshold = 128;
    }
    if (segment_nnz <= 0) segment_nnz = 256;
    if (hub_threshold <= 0) hub_threshold = 2 * segment_nnz;
    TG_REQUIRE(hub_threshold >= segment_nnz, TG_ERR_INVALID_ARG, "hub_threshold must be >= segment_nnz");

    std::vector<int32_t> h_ptr((size_t)n_rows + 1);
    TG_CUDA(cudaMemcpyAsync(h_ptr.data(), rowptr, h_ptr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    TG_REQUIRE(h_ptr[0] == 0 && h_ptr[(size_t)n_rows] == (int32_t)nnz, TG_ERR_INVALID_ARG,
               "rowptr does not span [0, nnz]");

    std::vector<int32_t> hub_rows, hub_seg_ptr(1, 0), seg_hub, seg_begin, seg_end;
    int64_t hub_nnz = 0;
    int32_t max_row = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int32_t s = h_ptr[(size_t)r], e = h_ptr[(size_t)r + 1];
        TG_REQUIRE(e >= s, TG_ERR_INVALID_ARG, "rowptr not monotone at row %lld", (long long)r);
        const int32_t len = e - s;
        if (len > max_row) max_row = len;
        if (len > hub_threshold) {
            const int32_t slot = (int32_t)hub_rows.size();
            hub_rows.push_back((int32_t)r);
            hub_nnz += len;
            // equal-sized segments (the last one may be shorter); order inside a row = storage order
            for (int32_t b = s; b < e; b += segment_nnz) {
                seg_hub.push_back(slot);
                seg_begin.push_back(b);
                seg_end.push_back(b + segment_nnz < e ? b + segment_nnz : e);
            }
            hub_seg_ptr.push_back((int32_t)seg_hub.size());
        }
    }

    tg_plan* pl = new tg_plan();
    pl->n_rows = n_rows; pl->n_cols = n_cols; pl->nnz = nnz;
    pl->hub_threshold = hub_threshold; pl->segment_nnz = segment_nnz;
    pl->n_hub = (int32_t)hub_rows.size(); pl->n_seg = (int32_t)seg_hub.size();
    pl->hub_nnz = hub_nnz; pl->max_row_nnz = max_row;
    auto upload = [&](int32_t** dst, const std::vector<int32_t>& v) -> cudaError_t {
        cudaError_t e = cudaMalloc((void**)dst, (v.size() ? v.size() : 1) * sizeof(int32_t));
        if (e != cudaSuccess) return e;
        if (!v.empty()) e = cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        return e;
    };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = upload(&pl->hub_rows, hub_rows);
    if (e == cudaSuccess) e = upload(&pl->hub_seg_ptr, hub_seg_ptr);
    if (e == cudaSuccess) e = upload(&pl->seg_hub, seg_hub);
    if (e == cudaSuccess) e = upload(&pl->seg_begin, seg_begin);
    if (e == cudaSuccess) e = upload(&pl->seg_end, seg_end);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->tickets, (hub_rows.size() ? hub_rows.size() : 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(pl->tickets, 0, (hub_rows.size() ? hub_rows.size() : 1) * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        tg_plan_destroy(pl);
        return cuda_fail(e, "plan upload", __FILE__, __LINE__);
    }
    // Execution order of the split-row segments: by first column.  Segments of DIFFERENT hub rows that cover the same
    // stretch of columns then run at the same time, so a row of B fetched for one of them is an L2 hit for the others
    // (in storage order every hub row would re-read its ~nnz rows of B from HBM).  Only the order of execution changes:
    // partial rows stay indexed by segment id and are still added in segment order.
    {
        std::vector<int32_t> order((size_t)pl->n_seg);
        for (int32_t i = 0; i < pl->n_seg; ++i) order[(size_t)i] = i;
        if (colidx && pl->n_seg > 1) {
            DevBuf d_first;
            std::vector<int32_t> first((size_t)pl->n_seg);
            e = d_first.alloc((size_t)pl->n_seg * sizeof(int32_t));
            if (e == cudaSuccess) {
                seg_first_col_kernel<<<grid1d(pl->n_seg), 256, 0, st>>>(colidx, pl->seg_begin, pl->n_seg, d_first.as<int32_t>());
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaMemcpyAsync(first.data(), d_first.p, first.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e == cudaSuccess)
                std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return first[(size_t)x] < first[(size_t)y]; });
        }
        if (e == cudaSuccess) e = upload(&pl->seg_order, order);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            tg_plan_destroy(pl);
            return cuda_fail(e, "segment order", __FILE__, __LINE__);
        }
    }
    // optional sub-plan of the role kernels (square matrices with a compact hub set: the document-topic-topic graphs)
    if (pl->n_rows == pl->n_cols && pl->n_hub >= 1) {
        const int rc = tg::roles2_plan_build(pl, rowptr, colidx, vals, h_ptr.data(), hub_rows.data(), st);
        if (rc != TG_OK) {
            tg_plan_destroy(pl);
            return rc;
        }
    }
    // optional sub-plans for rectangular operands (sparse feature matrix / its transpose, <= 1280 features)
    if (!pl->r2_ok) {
        const int rc2 = tg::roles2_rect_plan_build(pl, rowptr, colidx, vals, h_ptr.data(), st);
        if (rc2 != TG_OK) {
            tg_plan_destroy(pl);
            return rc2;
        }
    }
    *plan_out = pl;
    return TG_OK;
}

void tg_plan_destroy(tg_plan* pl) {
    if (!pl) return;
    tg::roles2_plan_free(pl);
    cudaFree(pl->hub_rows); cudaFree(pl->hub_seg_ptr); cudaFree(pl->seg_hub);
    cudaFree(pl->seg_begin); cudaFree(pl->seg_end); cudaFree(pl->tickets); cudaFree(pl->seg_order);
    delete pl;
}

int tg_plan_info(const tg_plan* pl, int64_t info[8]) {
    TG_REQUIRE(pl && info, TG_ERR_INVALID_ARG, "null pointer");
    info[0] = pl->n_hub; info[1] = pl->n_seg; info[2] = pl->hub_nnz;
    info[3] = pl->max_row_nnz; info[4] = pl->hub_threshold; info[5] = pl->segment_nnz;
    // bit 0 / 1: role kernels on the square graph, bits 2-3: rectangular mode; info[7]: nodes per hub chunk | groups << 16 | nq << 24
    info[6] = ((pl->r2_ok && pl->r2_rect == 0) ? 3 : 0) | ((pl->r2_ok && pl->r2_rect != 0) ? 4 * pl->r2_rect : 0);
    info[7] = pl->r2_ok ? ((int64_t)pl->r2_T | ((int64_t)pl->r2_groups << 16) | ((int64_t)pl->r2_nq << 24) | ((int64_t)pl->r2_gs << 32)) : 0;
    return TG_OK;
}

size_t tg_plan_workspace_bytes(const tg_plan* pl, int32_t n_feat) {
    if (!pl || n_feat <= 0) return 0;
    const size_t ld = (size_t)((n_feat + 3) / 4) * 4;
    const size_t v1 = (size_t)pl->n_seg * ld * sizeof(float) + 16;
    const size_t v2 = tg::roles2_workspace_bytes(pl, n_feat);
    return v1 > v2 ? v1 : v2;
}

int tg_plan_spmm_launches(const tg_plan* pl, const float* B, int64_t ldb, int32_t n_feat, int32_t philox, int32_t out_vec4_ok) {
    if (!pl || n_feat <= 0) return 0;
    if (out_vec4_ok) {
        tg::StreamCall sc{nullptr, nullptr, B, ldb, n_feat, nullptr, 0};
        // philox = 2: the call is the fused loss forward (row-wise epilogue)
        const int k = tg::roles2_launches(pl, sc, philox == 1, philox == 2);
        if (k > 0) return k;
    }
    return 1;  // gather kernel: split rows are finished by the last arriver inside the same launch
}

}  // extern "C"
