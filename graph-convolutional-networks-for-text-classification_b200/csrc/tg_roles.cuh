// tg_roles.cuh — interface of the role-specialised column-chunk streaming SpMM (tg_roles2.cu) used by the entry points in
// tg_spmm.cu and by the plan builder in tg_csr.cu.
#pragma once
#include "tg_epilogue.cuh"

namespace tg {

struct StreamCall {
    const int32_t* rowptr;
    const float* vals;
    const float* B;
    int64_t ldb;
    int32_t n_feat;
    void* workspace;
    size_t workspace_bytes;
};

// Sub-plan of the role kernels for a square graph whose hub set is compact (document-topic-topic graphs, <= 1280 hub
// rows); h_hub_rows: host [pl->n_hub].  Leaves pl->r2_ok false (and returns TG_OK) when the layout does not apply.
int roles2_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* h_rowptr,
                      const int32_t* h_hub_rows, cudaStream_t st);
// rectangular operands: sparse feature matrix times dense weight (document role only) and its transpose product (hub role only)
int roles2_rect_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* h_rowptr,
                           cudaStream_t st);
void roles2_plan_free(tg_plan* pl);
size_t roles2_workspace_bytes(const tg_plan* pl, int32_t n_feat);

// kernels of this file one product launches (0: the role kernels do not apply and the gather kernel runs)
int roles2_launches(const tg_plan* pl, const StreamCall& c, bool philox, bool whole_row);

// 64 <= n_feat <= 1024 (n_feat % 4 == 0); plans with 64 / 32-column slices (more than 256 hub rows) also take class-sized
// operands (n_feat <= 32 nq) and then the row-wise loss epilogue (whole_row)
bool roles2_applicable(const tg_plan* pl, const StreamCall& c, bool whole_row);
int roles2_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
int roles2_run(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st);
bool roles2_rect_applicable(const tg_plan* pl, const StreamCall& c);
int roles2_rect_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
bool roles2_narrow_applicable(const tg_plan* pl, const StreamCall& c);   // n_feat <= 32 (class-sized operands)
int roles2_narrow_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
int roles2_narrow_run(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st);

}  // namespace tg
