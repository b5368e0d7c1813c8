/*
 * deepbinner_b200 - C ABI of the B200-native barcode-from-squiggle classifier.
 *
 * This is the drop-in boundary for the one hot path of rrwick/Deepbinner that this library
 * accelerates (raw signal window -> 1-D CNN -> softmax -> per-read barcode call).  The reference has
 * no plugin registry; its seams are Python call sites, and its only FFI
 * convention is `deepbinner/dtw/dtw.h:16-19` + `deepbinner/dtw_semi_global.py:26-58`:
 * `extern "C"` free functions, C-contiguous caller-allocated arrays, scalars by value.  The entry
 * points below follow that convention; each one cites the reference interface it replaces.
 *
 * All functions return 0 on success or a negative DBN_E* code; db_last_error() returns a
 * thread-local human-readable message for the last failure.  A handle may be used from one host
 * thread at a time.  Host pointers may be pageable or pinned; "_device" variants take device
 * pointers on the handle's device and run asynchronously on the given stream.
 *
 * There is no CPU fallback: db_create() fails (DBN_ENODEVICE) when no sm_100 GPU is usable.
 */
#ifndef DEEPBINNER_B200_H
#define DEEPBINNER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBN_ABI_VERSION 1

#if defined(__GNUC__)
#define DBN_API __attribute__((visibility("default")))
#else
#define DBN_API
#endif

#define DBN_OK 0
#define DBN_EINVAL (-1)    /* bad argument */
#define DBN_EFORMAT (-2)   /* weight blob is not a valid DBNW v1 blob / unsupported topology */
#define DBN_ENODEVICE (-3) /* no usable CUDA device (compute capability 10.x required) */
#define DBN_ECUDA (-4)     /* CUDA runtime error; see db_last_error() */
#define DBN_ENOMEM (-5)

#define DBN_SIDE_START 0 /* call_batch side='start' (classify.py:343-344) */
#define DBN_SIDE_END 1   /* call_batch side='end'   (classify.py:345-349) */

/* Engines (db_set_engine).  Both are hand-written sm_100a CUDA; results agree to ~1e-4. */
#define DBN_ENGINE_FP32 0    /* CUDA-core fp32 fused per-window kernel (parity anchor) */
#define DBN_ENGINE_TCGEN05 1 /* tcgen05/TMEM split-bf16 tensor-core kernel (default) */

typedef struct db_model db_model;

/* ABI version of the loaded library (== DBN_ABI_VERSION at build time). */
DBN_API int db_abi_version(void);

/* Thread-local message describing the last error on this thread ("" if none). */
DBN_API const char *db_last_error(void);

/*
 * Replaces keras.models.load_model at classify.py:86-103 (load_trained_model).
 * weights_blob: DBNW v1 blob (deepbinner_b200/weights.py) with the parameters of the
 * network_architecture.py:18-95 graph; copied - the caller may free it after the call.
 * device: CUDA device ordinal.  *out receives the handle.
 */
DBN_API int db_create(const void *weights_blob, size_t blob_bytes, int device, db_model **out);

/* Releases all device and host resources of the handle (NULL is a no-op). */
DBN_API void db_destroy(db_model *model);

/*
 * model.inputs[0].shape[1] and model.outputs[0].shape[1] as read at classify.py:92-99
 * (1024 and 13 for the shipped models).
 */
DBN_API int db_info(const db_model *model, int *input_size, int *n_classes);

/* Select the compute engine (DBN_ENGINE_*).  Default: the fastest engine that passed self-test. */
DBN_API int db_set_engine(db_model *model, int engine);
DBN_API int db_get_engine(const db_model *model);

/*
 * Seam b1 - model.predict(x, batch_size) at classify.py:361.
 * x: host [n, input_size] already-normalised windows (float32, C-contiguous; the _f64 variant
 * takes the float64 array the reference builds at classify.py:340 and casts to float32, as Keras
 * does).  probs: host [n, n_classes] float32 softmax rows, written in full.
 * n may be any size >= 0; the library tiles it over the device.  A pageable x (a plain numpy array) is copied
 * into pinned staging by the library's host threads (float64 cast to float32 there) and transferred
 * asynchronously from it; a page-locked x goes to the copy engine directly (float64 then cast on the device).
 */
DBN_API int db_predict_windows(db_model *model, const float *x, int64_t n, float *probs);
DBN_API int db_predict_windows_f64(db_model *model, const double *x, int64_t n, float *probs);

/* Same with device-resident input/output, asynchronous on `stream` (a cudaStream_t, may be 0). */
DBN_API int db_predict_windows_device(db_model *model, const float *d_x, int64_t n, float *d_probs,
                              void *stream);

/*
 * Seam b2 - call_batch(input_size, output_size, read_ids, signals, model, args, side) at
 * classify.py:325-384, fused: windowing (:337-349), normalise (trim_signal.py:61-69), zero padding
 * (:352-357), predict (:361), the min/max merge over scan steps (:363-377), make_sum_to_one
 * (:387-393) and get_barcode_call_from_probabilities (:285-295) all run on the device.
 *
 * samples:  host int16, the raw signals of the batch concatenated
 * offsets:  host int64 [n_reads + 1]; read i is samples[offsets[i] .. offsets[i+1])
 * side:     DBN_SIDE_START / DBN_SIDE_END
 * scan_size: args.scan_size; must be a positive multiple of input_size/2 (check_input_size,
 *            classify.py:396-407) else DBN_EINVAL
 * score_diff: args.score_diff
 * probs:    host float32 [n_reads, n_classes] - per-read probabilities after make_sum_to_one
 * calls:    host int8 [n_reads] - 0 = 'none', k = barcode 'k'
 * Only the scan region of each read - its first (side start) / last (side end)
 * min(len, scan_size + input_size/2) samples, which is all call_batch ever slices
 * (classify.py:337-349: the last step reads [scan_size - step, scan_size + step)) - is transferred
 * to the device.
 */
DBN_API int db_call_batch(db_model *model, const int16_t *samples, const int64_t *offsets, int n_reads,
                  int side, int scan_size, double score_diff, float *probs, int8_t *calls);

/*
 * The same call split in two, so that the host can overlap its own work with the GPU's (prepare the
 * next batch, submit the other model's side of this batch, format results): submit() gathers the scan
 * regions chunk by chunk into pinned staging and enqueues copy (on a dedicated copy stream) -> kernels ->
 * copy-back for every chunk (on three rotating compute streams that wait for the chunk's copy event); it
 * returns as soon as everything is enqueued and no longer references the caller's buffers.  Chunk size: 2048 network
 * windows for a lone job, half the job (3072 .. 16384 windows) when other jobs of the handle are already queued - larger
 * launches of the persistent kernel are more efficient, a lone job needs its own copies under its own kernels
 * (profiles/r02_chunk_sweep.txt; DEEPBINNER_B200_CALL_CHUNK fixes the size).  wait() blocks until the job is complete and fills probs / calls as db_call_batch
 * does.  *job receives a small index; up to 4 jobs per handle may be in flight; every submitted job must
 * be waited for exactly once.  db_call_batch == submit_packed + wait.
 *   db_call_batch_submit:        read i = signals[i][0 .. lengths[i])   (ragged host arrays)
 *   db_call_batch_submit_packed: read i = samples[offsets[i] .. offsets[i+1])
 * Exception to "no longer references the caller's buffers": if `samples` of the packed variant is
 * page-locked memory (cudaHostAlloc / cudaHostRegister), chunks whose reads all fit the scan region whole,
 * and chunks of equally long reads (rows that hold the first and the last `keep` samples of every read,
 * as the fast5 batch reader packs them: one strided 2-D copy takes this side's region out of every row),
 * are copied to the device straight from it (no staging pass) - such a buffer must stay unchanged until wait().
 */
DBN_API int db_call_batch_submit(db_model *model, const int16_t *const *signals, const int64_t *lengths,
                                 int n_reads, int side, int scan_size, double score_diff, int *job);
DBN_API int db_call_batch_submit_packed(db_model *model, const int16_t *samples, const int64_t *offsets,
                                        int n_reads, int side, int scan_size, double score_diff, int *job);
DBN_API int db_call_batch_wait(db_model *model, int job, float *probs, int8_t *calls);

/*
 * Device-resident variant: d_samples holds, for read i, its scan region (first/last
 * min(len_i, scan_size + input_size/2) samples) at d_samples[d_offsets[i] .. d_offsets[i+1]).
 * Asynchronous on `stream`.  d_step_probs receives the per-step softmax rows
 * [steps, n_reads, n_classes]; it may be NULL, in which case a scratch buffer owned by the handle
 * is used (then do not overlap calls on different streams).
 */
DBN_API int db_call_batch_device(db_model *model, const int16_t *d_samples, const int64_t *d_offsets,
                         int n_reads, int side, int scan_size, double score_diff, float *d_probs,
                         int8_t *d_calls, float *d_step_probs, void *stream);

/*
 * Profiling helpers.  db_last_gpu_ms: device time (CUDA events on the handle's stream) of the
 * network kernel(s) launched by the most recent host-buffer call.  db_kernel_launches: number of
 * this library's kernels launched on this handle since creation.
 */
DBN_API float db_last_gpu_ms(const db_model *model);
DBN_API int64_t db_kernel_launches(const db_model *model);

/*
 * Native fast5 reader (host only, no GPU needed) - replaces h5py in load_fast5s.py:25-49
 * get_read_id_and_signal and :93-98 get_root_level_keys.  Return value / status: 0 = ok,
 * 1 = unreadable or not a single-read fast5 (the reference returns (None, None)), 2 = multi-read
 * file (the reference exits with "does not (yet) support multi-read fast5 files").
 *
 * db_fast5_read: read_id (64-byte NUL-padded buffer) and the int16 Signal of one file; *length
 *   receives the signal length; samples are copied only if capacity >= *length (call with
 *   signal = NULL / capacity = 0 to query the length).
 * db_fast5_list_root: NUL-separated names of the root group's members (hdf5_file.keys()).
 * db_fast5_batch_read: parse n files on `threads` host threads; if keep > 0 only the first and last
 *   `keep` samples of longer signals are retained (all call_batch ever slices when
 *   keep >= scan_size + input_size/2).  db_fast5_batch_get exposes the packed result: samples,
 *   offsets[n+1], full (untruncated) lengths[n], read ids [n][64], status[n]; the pointers stay
 *   valid until db_fast5_batch_free.
 */
/* The decompressor the fast5 reader inflates signal chunks with (csrc/dbn_inflate.h; use_zlib = 1: the
 * system zlib it is checked against).  src: a zlib stream; dst receives at most dst_capacity bytes (a
 * longer stream is cut there, as for HDF5 edge chunks).  Returns 0, or 1 for a malformed stream.
 * Exposed for tests/test_fast5_readers.py. */
DBN_API int db_zlib_inflate(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_capacity,
                            int64_t *produced, int use_zlib);
typedef struct db_fast5_batch db_fast5_batch;
DBN_API int db_fast5_read(const char *path, char *read_id, int16_t *signal, int64_t capacity,
                          int64_t *length);
DBN_API int db_fast5_list_root(const char *path, char *names, int64_t capacity, int *count);
DBN_API int db_fast5_batch_read(const char *const *paths, int n, int threads, int64_t keep,
                                db_fast5_batch **out);
/* One row per READ instead of one per file: a multi-read fast5 (several /read_<uuid> groups in the
 * root, load_fast5s.py:67-98) contributes every read, read straight from the file - where the
 * reference unpacks it with ONT's multi_to_single_fast5 first (realtime.py:183-196).  An unreadable
 * file is one row with status 1.  db_fast5_batch_rows gives the row count and, per row, the index of
 * the file it came from; db_fast5_batch_get's arrays are then per row. */
DBN_API int db_fast5_batch_read_reads(const char *const *paths, int n, int threads, int64_t keep,
                                      db_fast5_batch **out);
/* db_fast5_batch_read_reads for a run that only looks at one side of the reads (sides: 1 = start, 2 = end,
 * 3 = both).  Start only (e.g. the SQK-RBK004 preset): every read is cut after `keep` samples and the
 * inflate of its signal chunk stops there instead of decoding the whole read. */
DBN_API int db_fast5_batch_read_sides(const char *const *paths, int n, int threads, int64_t keep, int sides,
                                      db_fast5_batch **out);
DBN_API int db_fast5_batch_rows(const db_fast5_batch *batch, int64_t *rows, const int32_t **row_file);
DBN_API int db_fast5_batch_get(const db_fast5_batch *batch, const int16_t **samples,
                               const int64_t **offsets, const int64_t **full_length,
                               const char **read_ids, const int32_t **status);
DBN_API void db_fast5_batch_free(db_fast5_batch *batch);

/*
 * Diagnostics of the tcgen05 engine (used by tests/test_gpu_tc_layers.py): number of MMA jobs, and a
 * dump of the two shared-memory activation regions (2 x 98688 bytes, split-bf16 layout documented in
 * csrc/dbn_tc.cu) after running host windows x[0..1] through jobs 0..job.
 */
DBN_API int db_tc_num_jobs(const db_model *model);
/* Host only, no GPU needed: the MMA job table built for a weight blob (which = 0), 32 int32 per job in the field order of struct TcJob
 * (csrc/dbn_tc.cu).  Returns the number of jobs or a negative DBN_E* code.  Used by the CPU tests that
 * check the schedule (accumulator-slot reuse, hand-off counts, buffer sizes). */
DBN_API int db_tc_job_table(const void *weights_blob, size_t blob_bytes, int which, int32_t *out,
                            int max_jobs);
/* Host only: the packed split-bf16 weights and the parameter block (bias / folded BatchNorm) that the
 * job table above indexes (w_goff, w_part / bias_off, bn_off).  *w_bytes and *prm_floats receive the
 * sizes; the data is copied when the capacities suffice (pass NULL / 0 to query). */
DBN_API int db_tc_packed(const void *weights_blob, size_t blob_bytes, int which, unsigned char *w_out,
                         int64_t w_cap, float *prm_out, int64_t prm_cap, int64_t *w_bytes,
                         int64_t *prm_floats);
DBN_API int db_tc_debug_dump(db_model *model, const float *x, int job, unsigned char *out);
/* Timeline of CTA 0 for n device-resident windows: trace[job][window][8] SM-clock stamps (MMA issue
 * start/end, epilogue start/end, then epilogue internals); host buffer of 32*2*8 int64. */
DBN_API int db_tc_trace(db_model *model, const float *d_x, int n, float *d_probs, int64_t *trace);
/* The same for the fused call_batch form of the kernel: n_reads device-resident int16 scan regions
 * (offsets[n_reads + 1]), side 'start', one window per read. */
DBN_API int db_tc_trace_call(db_model *model, const int16_t *d_samples, const int64_t *d_offsets, int n_reads,
                             float *d_probs, int64_t *trace);

#ifdef __cplusplus
}
#endif

#endif /* DEEPBINNER_B200_H */
