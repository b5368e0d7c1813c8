// Interface between the C-ABI layer (dbn_lib.cu) and the tcgen05 tensor-core engine (dbn_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dbn {

struct Blob;
struct TcEngine;

int fail(int code, const char* fmt, ...);

// Returns nullptr when the engine cannot be set up (the caller then stays on the fp32 CUDA engine).
TcEngine* tc_create(const Blob& blob);
void tc_destroy(TcEngine* e);
// d_x (float32) or d_xd (float64): [n][1024] normalised windows -> d_probs [n][n_classes]
int tc_predict(TcEngine* e, const float* d_x, const double* d_xd, int64_t n, float* d_probs,
               cudaStream_t st);
// fused call_batch front end: window w = step*n_reads + read -> d_step_probs [steps][n_reads][nc]
int tc_call_windows(TcEngine* e, const int16_t* d_samples, const int64_t* d_offsets, int n_reads,
                    int side, int steps, float* d_step_probs, cudaStream_t st);

int tc_num_jobs(const TcEngine* e);
// Host-only dump of the job table (32 ints per job, struct TcJob order); -1 if the model is unsupported.
int tc_job_table(const Blob& blob, int which, int32_t* out, int max_jobs);
int tc_packed(const Blob& blob, int which, unsigned char* w_out, int64_t w_cap, float* prm_out, int64_t prm_cap,
              int64_t* w_bytes, int64_t* prm_floats);
int tc_trace(TcEngine* e, const float* d_x, int n, float* d_probs, long long* d_trace, cudaStream_t st);
int tc_trace_call(TcEngine* e, const int16_t* d_samples, const int64_t* d_offsets, int n_reads, float* d_probs,
                  long long* d_trace, cudaStream_t st);
// Debug: run windows d_x[0..1] through jobs 0..job and dump both activation regions.
int tc_debug_dump(TcEngine* e, const float* d_x, int job, unsigned char* d_out, cudaStream_t st);

}  // namespace dbn
