// tg_stream.cuh — interface of the column-chunk streaming SpMM (tg_stream.cu) used by the entry points in tg_spmm.cu.
#pragma once
#include "tg_epilogue.cuh"

namespace tg {

constexpr int32_t kHubBit = (int32_t)0x80000000;

struct StreamCall {
    const int32_t* rowptr;
    const float* vals;
    const float* B;
    int64_t ldb;
    int32_t n_feat;
    void* workspace;
    size_t workspace_bytes;
};

// builds / frees the streaming part of a plan (called from tg_plan_create / tg_plan_destroy)
int stream_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* h_rowptr, cudaStream_t st);
void stream_plan_free(tg_plan* pl);
size_t stream_workspace_bytes(const tg_plan* pl, int32_t n_feat);

// true when the streaming kernel can run this call (plan has the sub-plan, vec4-aligned operands, width supported)
bool stream_applicable(const tg_plan* pl, const StreamCall& c, bool out_vec4_ok, bool whole_row_epilogue);

// warp-per-slot role kernels (tg_roles2.cu)
int roles2_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* h_rowptr,
                      const int32_t* d_slot_of, const int32_t* h_hub_rows, cudaStream_t st);
void roles2_plan_free(tg_plan* pl);
bool roles2_applicable(const tg_plan* pl, const StreamCall& c);
size_t roles2_workspace_bytes(const tg_plan* pl, int32_t n_feat);
int roles2_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
// rectangular operands: sparse feature matrix times dense weight (document role only) and its transpose product (hub role only)
int roles2_rect_plan_build(tg_plan* pl, const int32_t* rowptr, const int32_t* colidx, const float* vals, const int32_t* h_rowptr,
                           cudaStream_t st);
bool roles2_rect_applicable(const tg_plan* pl, const StreamCall& c);
int roles2_rect_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
bool roles2_narrow_applicable(const tg_plan* pl, const StreamCall& c);   // n_feat <= 32 (class-sized operands)
int roles2_narrow_run(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
int roles2_narrow_run(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st);

int stream_spmm_store(const tg_plan* pl, const StreamCall& c, const EpiStore& epi, cudaStream_t st);
int stream_spmm_loss(const tg_plan* pl, const StreamCall& c, const EpiLoss& epi, cudaStream_t st);

}  // namespace tg
