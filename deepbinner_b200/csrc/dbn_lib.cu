// libdeepbinner_b200: C-ABI (include/deepbinner_b200.h) over hand-written sm_100a CUDA kernels that
// run Deepbinner's barcode classifier (reference classify.py:325-384 call_batch /
// classify.py:361 model.predict / network_architecture.py:18-95).  No CPU fallback.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/deepbinner_b200.h"
#include "dbn_engine.h"
#include "dbn_weights.h"

namespace dbn {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define DBN_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return ::dbn::fail(DBN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                               __FILE__, __LINE__);                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------
// fp32 engine kernels (one CTA = one window; see dbn_fp32_net.cuh)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
    k_fp32_predict(Fp32Net net, const T* __restrict__ x, float* __restrict__ probs) {
    extern __shared__ __align__(16) float smem[];
    const size_t w = blockIdx.x;
    float* B = smem + kBufFloats;
    DBN_PHASE(stage_window_from_values(tid, x + w * kInputSize, B));
    fp32_forward_window(net, smem, probs + w * net.n_classes);
}

// Fused call_batch front end: window w = step * n_reads + read (the reference's step-major order,
// classify.py:337-361); slices, z-scores and pads the window straight from the int16 scan region.
__global__ void __launch_bounds__(kThreads, 1)
    k_fp32_call_windows(Fp32Net net, const int16_t* __restrict__ samples,
                        const int64_t* __restrict__ offsets, int n_reads, int side,
                        float* __restrict__ step_probs) {
    extern __shared__ __align__(16) float smem[];
    const int w = blockIdx.x;
    const int step = w / n_reads, read = w % n_reads;
    const int64_t off = offsets[read];
    const int region_len = static_cast<int>(offsets[read + 1] - off);
    const int16_t* region = samples + off;
    const WindowGeom g = window_geometry(region_len, step, side);
    long long* red = reinterpret_cast<long long*>(smem);  // buffer A is free during staging
    float* B = smem + kBufFloats;
    DBN_PHASE(window_partial_sums(tid, region, g, red));
    DBN_PHASE(window_reduce(tid, red));
    DBN_PHASE(window_normalise(tid, region, g, red, B));
    fp32_forward_window(net, smem, step_probs + static_cast<size_t>(w) * net.n_classes);
}

// ---------------------------------------------------------------------------------------------
// per-read merge + renormalise + call (classify.py:363-384, :387-393, :285-295)
// ---------------------------------------------------------------------------------------------
__global__ void k_merge_call(const float* __restrict__ step_probs, int n_reads, int steps, int nc,
                             double score_diff, float* __restrict__ probs,
                             int8_t* __restrict__ calls) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    float m[kMaxClasses];
    for (int j = 0; j < nc; ++j) m[j] = step_probs[static_cast<size_t>(r) * nc + j];
    for (int s = 1; s < steps; ++s) {
        const float* row = step_probs + (static_cast<size_t>(s) * n_reads + r) * nc;
        m[0] = fminf(m[0], row[0]);                               // :373 no-barcode: minimum
        for (int j = 1; j < nc; ++j) m[j] = fmaxf(m[j], row[j]);  // :374-375 barcodes: maximum
    }
    // make_sum_to_one (:387-393).  With the numpy of the reference's era these scalar expressions
    // promote the float32 softmax values to float64; same order of operations here.
    const double none = static_cast<double>(m[0]);
    double sum = 0.0;
    for (int j = 1; j < nc; ++j) sum += static_cast<double>(m[j]);
    const double factor = (1.0 - none) / sum;
    double p[kMaxClasses];
    p[0] = none;
    for (int j = 1; j < nc; ++j) p[j] = static_cast<double>(m[j]) * factor;
    // get_barcode_call_from_probabilities (:285-295): stable descending sort => first maximum wins
    int best = 0;
    for (int j = 1; j < nc; ++j)
        if (p[j] > p[best]) best = j;
    double second = -1.0;
    for (int j = 0; j < nc; ++j)
        if (j != best && p[j] > second) second = p[j];
    int call = 0;
    if (best != 0 && (p[best] - second) >= score_diff) call = best;
    for (int j = 0; j < nc; ++j) probs[static_cast<size_t>(r) * nc + j] = static_cast<float>(p[j]);
    calls[r] = static_cast<int8_t>(call);
}

// float64 windows -> float32 (Keras casts model inputs to floatx; Appendix B.7) is folded into
// k_fp32_predict<double>.

}  // namespace dbn

// Small persistent worker pool for the host-side gather of call_batch jobs (copying the scan regions of
// thousands of reads into pinned staging is a memory-bandwidth job that one core does at ~10 GB/s).
class GatherPool {
  public:
    explicit GatherPool(int workers) {
        for (int i = 0; i < workers; ++i) threads_.emplace_back([this, i] { run(i + 1); });
    }
    ~GatherPool() {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
            stop_flag_.store(true, std::memory_order_release);
        }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }
    int parts() const { return static_cast<int>(threads_.size()) + 1; }
    // fn(part) for part = 0 .. parts()-1; part 0 runs on the calling thread
    void parallel(const std::function<void(int)>& fn) {
        if (threads_.empty()) return fn(0);
        {
            std::lock_guard<std::mutex> lock(m_);
            fn_ = &fn;
            pending_ = static_cast<int>(threads_.size());
            ++generation_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lock(m_);
        done_.wait(lock, [this] { return pending_ == 0; });
    }

  private:
    void run(int part) {
        int seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            // A caller that comes back every 100 - 200 us (predict with 256 windows per call, the chunks of a job) should
            // not pay a futex wake-up (20 - 100 us) each time: poll for a short while before going to sleep.
            for (const auto t0 = std::chrono::steady_clock::now();
                 generation_.load(std::memory_order_acquire) == seen && !stop_flag_.load(std::memory_order_acquire) &&
                 std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(150);) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                fn = fn_;
            }
            (*fn)(part);
            {
                std::lock_guard<std::mutex> lock(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int pending_ = 0;
    std::atomic<int> generation_{0};
    std::atomic<bool> stop_flag_{false};
    bool stop_ = false;
};

// =================================================================================================
// handle
// =================================================================================================
struct db_model {
    int device = 0;
    int input_size = 0;
    int n_classes = 0;
    int engine = DBN_ENGINE_FP32;
    bool tc_available = false;
    int sm_count = 0;
    dbn::Blob blob;
    // fp32 engine
    float* d_fp32_w = nullptr;
    dbn::Fp32Net fp32{};
    // tcgen05 engine
    dbn::TcEngine* tc = nullptr;
    // streams / events
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_slot[2] = {nullptr, nullptr};
    float* h_out[2] = {nullptr, nullptr};   // pinned result staging of the pipelined host predict
    size_t h_out_bytes[2] = {0, 0};
    float* h_in[2] = {nullptr, nullptr};    // pinned fp32 input staging of the host predict for PAGEABLE caller arrays
    size_t h_in_bytes[2] = {0, 0};
    float last_ms = 0.f;
    int64_t launches = 0;
    // device scratch (grown on demand)
    void* d_in[2] = {nullptr, nullptr};
    size_t d_in_bytes[2] = {0, 0};
    float* d_out[2] = {nullptr, nullptr};
    size_t d_out_bytes[2] = {0, 0};
    int64_t* d_offsets = nullptr;
    size_t d_offsets_bytes = 0;
    float* d_step = nullptr;
    size_t d_step_bytes = 0;
    int8_t* d_calls = nullptr;
    size_t d_calls_bytes = 0;
    // in-flight call_batch jobs (db_call_batch_submit .. db_call_batch_wait): every slot owns its pinned
    // staging, device buffers and completion event, so that the host can prepare batch i+1 (or the other
    // model's side of the same batch) while batch i is on the GPU
    static constexpr int kJobSlots = 4;
    struct CallJob {
        bool busy = false;
        int n_reads = 0;
        int16_t* h_samples = nullptr;   // pinned: gathered scan regions of the whole job
        size_t h_samples_bytes = 0;
        int64_t* h_offsets = nullptr;   // pinned: per chunk, cnt + 1 offsets relative to the chunk's samples
        size_t h_offsets_bytes = 0;
        float* h_probs = nullptr;       // pinned results
        size_t h_probs_bytes = 0;
        int8_t* h_calls = nullptr;
        size_t h_calls_bytes = 0;
        int16_t* d_samples = nullptr;
        size_t d_samples_bytes = 0;
        int64_t* d_offsets = nullptr;
        size_t d_offsets_bytes = 0;
        float* d_step = nullptr;
        size_t d_step_bytes = 0;
        float* d_probs = nullptr;
        size_t d_probs_bytes = 0;
        int8_t* d_calls = nullptr;
        size_t d_calls_bytes = 0;
        cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
        cudaEvent_t ev_done[3] = {nullptr, nullptr, nullptr};   // one per job stream
    } jobs[kJobSlots];
    static constexpr int kJobStreams = 3;
    cudaStream_t job_streams[kJobStreams] = {nullptr, nullptr, nullptr};   // chunks of call_batch jobs rotate over these
    unsigned job_chunk_counter = 0;
    // Host -> device copies of the chunks go through their own stream and the chunk's compute stream waits for
    // the chunk's copy event: in stream order behind the previous chunk of the same compute stream, the three
    // streams fell into lockstep - all kernels end together and nobody's next input is on the device yet
    // (a ~100 us hole in front of every third kernel, profiles/r02_e2e_pipeline.txt).
    cudaStream_t copy_stream = nullptr;
    static constexpr int kCopyEvents = 16;
    cudaEvent_t ev_copy[kCopyEvents] = {};
    int call_chunk_windows = 0;         // network windows per pipelined chunk of a job; 0 = chosen per job (DEEPBINNER_B200_CALL_CHUNK fixes it)
    GatherPool* pool = nullptr;         // host threads of the gather (DEEPBINNER_B200_GATHER_THREADS, default 4)
};

namespace dbn {

template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    // grow with 25 % head-room (batches of ragged reads differ in size; re-allocating costs milliseconds)
    size_t want = std::max(need + need / 4, static_cast<size_t>(1) << 20);
    DBN_CUDA(cudaMalloc(reinterpret_cast<void**>(p), want));
    *cap = want;
    return 0;
}

template <typename T>
static int grow_host(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = std::max(need + need / 4, static_cast<size_t>(1) << 20);
    DBN_CUDA(cudaMallocHost(reinterpret_cast<void**>(p), want));
    *cap = want;
    return 0;
}

static constexpr size_t kFp32SmemBytes = sizeof(float) * kFp32SmemFloats;

static int launch_predict(db_model* m, const void* d_x, bool is_f64, int64_t n, float* d_probs,
                          cudaStream_t st) {
    if (n <= 0) return 0;
    if (n > 0x7fffffff) return fail(DBN_EINVAL, "too many windows in one launch");
    if (m->engine == DBN_ENGINE_TCGEN05) {
        int rc = tc_predict(m->tc, is_f64 ? nullptr : static_cast<const float*>(d_x),
                            is_f64 ? static_cast<const double*>(d_x) : nullptr, n, d_probs, st);
        if (rc) return rc;
        m->launches += 1;
        return 0;
    }
    const dim3 grid(static_cast<unsigned>(n));
    if (is_f64)
        k_fp32_predict<double><<<grid, kThreads, kFp32SmemBytes, st>>>(
            m->fp32, static_cast<const double*>(d_x), d_probs);
    else
        k_fp32_predict<float><<<grid, kThreads, kFp32SmemBytes, st>>>(
            m->fp32, static_cast<const float*>(d_x), d_probs);
    DBN_CUDA(cudaGetLastError());
    m->launches += 1;
    return 0;
}

static int launch_call_batch(db_model* m, const int16_t* d_samples, const int64_t* d_offsets,
                             int n_reads, int side, int steps, double score_diff, float* d_step,
                             float* d_probs, int8_t* d_calls, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    const int64_t windows = static_cast<int64_t>(n_reads) * steps;
    if (windows > 0x7fffffff) return fail(DBN_EINVAL, "too many windows in one launch");
    if (m->engine == DBN_ENGINE_TCGEN05) {
        int rc = tc_call_windows(m->tc, d_samples, d_offsets, n_reads, side, steps, d_step, st);
        if (rc) return rc;
    } else {
        k_fp32_call_windows<<<static_cast<unsigned>(windows), kThreads, kFp32SmemBytes, st>>>(
            m->fp32, d_samples, d_offsets, n_reads, side, d_step);
        DBN_CUDA(cudaGetLastError());
    }
    k_merge_call<<<(n_reads + 127) / 128, 128, 0, st>>>(d_step, n_reads, steps, m->n_classes,
                                                        score_diff, d_probs, d_calls);
    DBN_CUDA(cudaGetLastError());
    m->launches += 2;
    return 0;
}

static int check_scan(const db_model* m, int scan_size, int* steps) {
    const int step = m->input_size / 2;
    if (scan_size <= 0 || scan_size % step != 0)
        return fail(DBN_EINVAL, "--scan_size must be a multiple of half the model input size");
    *steps = scan_size / step;
    return 0;
}

}  // namespace dbn

using namespace dbn;

// ---- pipelined host entry of seam b2 -------------------------------------------------------------
// A job = one call_batch over n_reads host reads.  submit() cuts it into chunks of ~2048 network windows;
// for every chunk the host gathers the scan regions (the only samples call_batch ever looks at: the
// first / last scan_size + input_size/2 of each read, classify.py:337-349) into the job's pinned
// staging and enqueues H2D copy -> network kernel -> merge/call kernel -> D2H of the results on one of
// three streams (chunks rotate), so the gather and the copy of chunk i+1 run under the kernels of chunk
// i.  submit() returns once everything is enqueued - the caller's buffers are no longer referenced - and
// wait() blocks on the job's completion event and hands the results out.  Up to kJobSlots jobs per
// handle may be in flight (next batch, or several jobs queued by a driver loop).
// `pinned_base`: non-NULL when the reads are consecutive in one PINNED (page-locked / registered) host buffer
// starting there (packed variant): chunks whose reads all lie inside the scan region are then copied
// straight from the caller's buffer, without the staging pass.
template <typename GetRead>
static int submit_job(db_model* m, GetRead get_read, int n_reads, int side, int scan_size, double score_diff,
                      int* job_out, const int16_t* pinned_base = nullptr) {
    if (!m) return fail(DBN_EINVAL, "call_batch: model is NULL");
    if (!job_out) return fail(DBN_EINVAL, "call_batch: job is NULL");
    if (n_reads < 0) return fail(DBN_EINVAL, "call_batch: n_reads < 0");
    if (side != DBN_SIDE_START && side != DBN_SIDE_END) return fail(DBN_EINVAL, "call_batch: bad side");
    int steps = 0;
    int rc = check_scan(m, scan_size, &steps);
    if (rc) return rc;
    int slot = -1;
    for (int i = 0; i < db_model::kJobSlots; ++i)
        if (!m->jobs[i].busy) {
            slot = i;
            break;
        }
    if (slot < 0) return fail(DBN_EINVAL, "call_batch: %d jobs already in flight on this handle", db_model::kJobSlots);
    db_model::CallJob& J = m->jobs[slot];
    J.n_reads = n_reads;
    *job_out = slot;
    if (n_reads == 0) {
        J.busy = true;
        return DBN_OK;
    }
    DBN_CUDA(cudaSetDevice(m->device));
    const int64_t region_max = static_cast<int64_t>(scan_size) + m->input_size / 2;
    // Windows per chunk.  A launch of the persistent network kernel is the more efficient the more window pairs each
    // of its CTAs gets, a pipeline needs several chunks in flight: with other jobs of this handle already queued the
    // job is cut in two (3072 .. 16384 windows per chunk), a lone job - a synchronous caller - in chunks of 2048 so
    // that its own copies overlap its own kernels (measured: profiles/r02_chunk_sweep.txt).
    int chunk_windows = m->call_chunk_windows;
    if (chunk_windows <= 0) {
        int queued = 0;
        for (int i = 0; i < db_model::kJobSlots; ++i) queued += m->jobs[i].busy ? 1 : 0;
        const int64_t job_windows = static_cast<int64_t>(n_reads) * steps;
        chunk_windows = queued > 0 ? static_cast<int>(std::min<int64_t>(std::max<int64_t>(job_windows / 2, 3072), 16384)) : 2048;
    }
    const int chunk_max = std::max(1, chunk_windows / steps);   // reads per chunk, at most
    const int nchunks = (n_reads + chunk_max - 1) / chunk_max;
    const int chunk = (n_reads + nchunks - 1) / std::max(nchunks, 1);   // balanced chunks
    const size_t nc = m->n_classes;
    // sizes: regions are bounded by region_max per read
    int64_t bound = 0;
    for (int i = 0; i < n_reads; ++i) {
        const int16_t* p;
        int64_t len;
        get_read(i, &p, &len);
        if (len < 0) return fail(DBN_EINVAL, "call_batch: negative read length");
        bound += std::min(len, region_max);
    }
    const size_t sample_bytes = sizeof(int16_t) * std::max<int64_t>(bound, 1);
    const size_t offset_bytes = sizeof(int64_t) * (static_cast<size_t>(n_reads) + nchunks);
    if ((rc = grow_host(&J.h_samples, &J.h_samples_bytes, sample_bytes))) return rc;
    if ((rc = grow_host(&J.h_offsets, &J.h_offsets_bytes, offset_bytes))) return rc;
    if ((rc = grow_host(&J.h_probs, &J.h_probs_bytes, sizeof(float) * nc * n_reads))) return rc;
    if ((rc = grow_host(&J.h_calls, &J.h_calls_bytes, static_cast<size_t>(n_reads)))) return rc;
    if ((rc = grow(&J.d_samples, &J.d_samples_bytes, sample_bytes))) return rc;
    if ((rc = grow(&J.d_offsets, &J.d_offsets_bytes, offset_bytes))) return rc;
    if ((rc = grow(&J.d_step, &J.d_step_bytes, sizeof(float) * nc * n_reads * steps))) return rc;
    if ((rc = grow(&J.d_probs, &J.d_probs_bytes, sizeof(float) * nc * n_reads))) return rc;
    if ((rc = grow(&J.d_calls, &J.d_calls_bytes, static_cast<size_t>(n_reads)))) return rc;
    // (no cross-stream dependency: the chunks of consecutive jobs simply queue behind each other on the job
    // streams, so the first copy of this job runs under the last kernels of the previous one)
    DBN_CUDA(cudaEventRecord(J.ev_start, m->job_streams[m->job_chunk_counter % db_model::kJobStreams]));
    int64_t base = 0;   // samples gathered so far
    cudaStream_t last_stream = m->job_streams[0];
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * chunk, cnt = std::min(chunk, n_reads - r0);
        cudaStream_t st = m->job_streams[m->job_chunk_counter++ % db_model::kJobStreams];
        last_stream = st;
        int64_t* offs = J.h_offsets + r0 + c;   // cnt + 1 entries, relative to this chunk's samples
        int64_t total = 0;
        bool whole = pinned_base != nullptr;    // every read of the chunk is used whole
        bool uniform = pinned_base != nullptr;  // ... or the reads are consecutive and equally long (rows of a packed reader batch)
        const int16_t* first = nullptr;
        int64_t len0 = 0, consumed = 0;
        for (int i = 0; i < cnt; ++i) {
            const int16_t* p;
            int64_t len;
            get_read(r0 + i, &p, &len);
            if (i == 0) {
                first = p;
                len0 = len;
            }
            whole = whole && len <= region_max && p == first + total;
            uniform = uniform && len == len0 && p == first + consumed;
            consumed += len;
            offs[i] = total;
            total += std::min(len, region_max);
        }
        offs[cnt] = total;
        const int16_t* h_src = J.h_samples + base;
        auto copy_reads = [&](int part) {   // reads [lo, hi) of the chunk
            const int parts = total >= (1 << 19) ? m->pool->parts() : 1;
            if (part >= parts) return;
            const int lo = static_cast<int>(static_cast<int64_t>(cnt) * part / parts);
            const int hi = static_cast<int>(static_cast<int64_t>(cnt) * (part + 1) / parts);
            for (int i = lo; i < hi; ++i) {
                const int16_t* p;
                int64_t len;
                get_read(r0 + i, &p, &len);
                const int64_t r = offs[i + 1] - offs[i];
                if (r > 0)
                    std::memcpy(J.h_samples + base + offs[i], p + (side == DBN_SIDE_START ? 0 : len - r), sizeof(int16_t) * r);
            }
        };
        if (whole && total > 0) {
            // zero-copy: DMA straight from the caller's pinned buffer
            DBN_CUDA(cudaMemcpyAsync(J.d_samples + base, first, sizeof(int16_t) * total, cudaMemcpyHostToDevice, m->copy_stream));
        } else if (uniform && total > 0 && len0 > region_max) {
            // zero-copy, strided: equally long rows (e.g. [first keep | last keep] samples of every read, as the
            // native fast5 reader packs both ends) - one 2-D DMA takes this side's region out of every row, no
            // staging pass through host memory (which is what bounded 8 ranks on one host)
            const int16_t* src = first + (side == DBN_SIDE_START ? 0 : len0 - region_max);
            DBN_CUDA(cudaMemcpy2DAsync(J.d_samples + base, sizeof(int16_t) * region_max, src, sizeof(int16_t) * len0,
                                       sizeof(int16_t) * region_max, static_cast<size_t>(cnt), cudaMemcpyHostToDevice,
                                       m->copy_stream));
        } else {
            if (total >= (1 << 19)) m->pool->parallel(copy_reads);
            else copy_reads(0);
            if (total > 0)
                DBN_CUDA(cudaMemcpyAsync(J.d_samples + base, h_src, sizeof(int16_t) * total, cudaMemcpyHostToDevice, m->copy_stream));
        }
        DBN_CUDA(cudaMemcpyAsync(J.d_offsets + r0 + c, offs, sizeof(int64_t) * (cnt + 1), cudaMemcpyHostToDevice, m->copy_stream));
        cudaEvent_t copied = m->ev_copy[m->job_chunk_counter % db_model::kCopyEvents];   // (a wait refers to the record it follows)
        DBN_CUDA(cudaEventRecord(copied, m->copy_stream));
        DBN_CUDA(cudaStreamWaitEvent(st, copied, 0));
        rc = launch_call_batch(m, J.d_samples + base, J.d_offsets + r0 + c, cnt, side, steps, score_diff,
                               J.d_step + static_cast<size_t>(r0) * steps * nc, J.d_probs + r0 * nc, J.d_calls + r0, st);
        if (rc) return rc;
        DBN_CUDA(cudaMemcpyAsync(J.h_probs + r0 * nc, J.d_probs + r0 * nc, sizeof(float) * nc * cnt,
                                 cudaMemcpyDeviceToHost, st));
        DBN_CUDA(cudaMemcpyAsync(J.h_calls + r0, J.d_calls + r0, static_cast<size_t>(cnt), cudaMemcpyDeviceToHost, st));
        base += total;
    }
    // completion: one event per stream; wait() synchronises on all of them
    for (int i = 0; i < db_model::kJobStreams; ++i) DBN_CUDA(cudaEventRecord(J.ev_done[i], m->job_streams[i]));
    DBN_CUDA(cudaEventRecord(J.ev_stop, last_stream));
    J.busy = true;
    return DBN_OK;
}


// =================================================================================================
// C ABI
// =================================================================================================
#pragma GCC visibility push(default)
extern "C" {

int db_abi_version(void) { return DBN_ABI_VERSION; }

const char* db_last_error(void) { return g_last_error.c_str(); }

void db_destroy(db_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->tc) tc_destroy(m->tc);
    cudaFree(m->d_fp32_w);
    for (int i = 0; i < 2; ++i) {
        cudaFree(m->d_in[i]);
        cudaFree(m->d_out[i]);
        if (m->streams[i]) cudaStreamDestroy(m->streams[i]);
    }
    cudaFree(m->d_offsets);
    cudaFree(m->d_step);
    cudaFree(m->d_calls);
    for (db_model::CallJob& j : m->jobs) {
        if (j.h_samples) cudaFreeHost(j.h_samples);
        if (j.h_offsets) cudaFreeHost(j.h_offsets);
        if (j.h_probs) cudaFreeHost(j.h_probs);
        if (j.h_calls) cudaFreeHost(j.h_calls);
        cudaFree(j.d_samples);
        cudaFree(j.d_offsets);
        cudaFree(j.d_step);
        cudaFree(j.d_probs);
        cudaFree(j.d_calls);
        if (j.ev_start) cudaEventDestroy(j.ev_start);
        if (j.ev_stop) cudaEventDestroy(j.ev_stop);
        for (cudaEvent_t e : j.ev_done)
            if (e) cudaEventDestroy(e);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    for (cudaEvent_t e : m->ev_copy)
        if (e) cudaEventDestroy(e);
    for (cudaStream_t st : m->job_streams)
        if (st) cudaStreamDestroy(st);
    for (int i = 0; i < 2; ++i) {
        if (m->ev_slot[i]) cudaEventDestroy(m->ev_slot[i]);
        if (m->h_out[i]) cudaFreeHost(m->h_out[i]);
        if (m->h_in[i]) cudaFreeHost(m->h_in[i]);
    }
    if (m->ev_start) cudaEventDestroy(m->ev_start);
    if (m->ev_stop) cudaEventDestroy(m->ev_stop);
    delete m->pool;
    delete m;
}

int db_create(const void* weights_blob, size_t blob_bytes, int device, db_model** out) {
    if (!out) return fail(DBN_EINVAL, "db_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(DBN_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(DBN_EINVAL, "device %d out of range", device);
    cudaDeviceProp prop;
    DBN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DBN_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200)",
                    device, prop.major, prop.minor);
    db_model* m = new (std::nothrow) db_model();
    if (!m) return fail(DBN_ENOMEM, "out of host memory");
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    const std::string err = parse_blob(weights_blob, blob_bytes, &m->blob);
    if (!err.empty()) {
        delete m;
        return fail(DBN_EFORMAT, "%s", err.c_str());
    }
    m->input_size = m->blob.input_size;
    m->n_classes = m->blob.n_classes;
    int rc = [&]() -> int {
        DBN_CUDA(cudaSetDevice(device));
        std::vector<float> packed;
        pack_fp32(m->blob, &packed, &m->fp32.lay);
        DBN_CUDA(cudaMalloc(&m->d_fp32_w, packed.size() * sizeof(float)));
        DBN_CUDA(cudaMemcpy(m->d_fp32_w, packed.data(), packed.size() * sizeof(float),
                            cudaMemcpyHostToDevice));
        m->fp32.w = m->d_fp32_w;
        m->fp32.n_classes = m->n_classes;
        DBN_CUDA(cudaFuncSetAttribute(k_fp32_predict<float>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kFp32SmemBytes));
        DBN_CUDA(cudaFuncSetAttribute(k_fp32_predict<double>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kFp32SmemBytes));
        DBN_CUDA(cudaFuncSetAttribute(k_fp32_call_windows,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kFp32SmemBytes));
        for (int i = 0; i < 2; ++i)
            DBN_CUDA(cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking));
        DBN_CUDA(cudaEventCreate(&m->ev_start));
        DBN_CUDA(cudaEventCreate(&m->ev_stop));
        for (int i = 0; i < 2; ++i) {
            DBN_CUDA(cudaEventCreateWithFlags(&m->ev_slot[i], cudaEventDisableTiming));
            DBN_CUDA(cudaEventRecord(m->ev_slot[i], m->streams[i]));   // so that a first wait never blocks
        }
        for (db_model::CallJob& j : m->jobs) {
            DBN_CUDA(cudaEventCreate(&j.ev_start));
            DBN_CUDA(cudaEventCreate(&j.ev_stop));
            for (cudaEvent_t& e : j.ev_done) DBN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        for (cudaStream_t& st : m->job_streams) DBN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        DBN_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t& e : m->ev_copy) DBN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (const char* v = getenv("DEEPBINNER_B200_CALL_CHUNK")) m->call_chunk_windows = std::max(1, atoi(v));
        int gather_threads = 4;
        if (const char* v = getenv("DEEPBINNER_B200_GATHER_THREADS")) gather_threads = std::max(1, std::min(64, atoi(v)));
        m->pool = new GatherPool(gather_threads - 1);
        return 0;
    }();
    if (rc) {
        db_destroy(m);
        return rc;
    }
    // tensor-core engine: optional at creation (falls back to the fp32 CUDA engine - still GPU -
    // if it cannot be set up); selected by default when available.
    m->tc = tc_create(m->blob);
    m->tc_available = (m->tc != nullptr);
    m->engine = m->tc_available ? DBN_ENGINE_TCGEN05 : DBN_ENGINE_FP32;
    // DEEPBINNER_B200_ENGINE = fp32 | tcgen05 overrides the default engine of new handles
    if (const char* want = getenv("DEEPBINNER_B200_ENGINE")) {
        if (!std::strcmp(want, "fp32")) m->engine = DBN_ENGINE_FP32;
    }
    *out = m;
    return DBN_OK;
}

int db_info(const db_model* m, int* input_size, int* n_classes) {
    if (!m) return fail(DBN_EINVAL, "db_info: model is NULL");
    if (input_size) *input_size = m->input_size;
    if (n_classes) *n_classes = m->n_classes;
    return DBN_OK;
}

int db_set_engine(db_model* m, int engine) {
    if (!m) return fail(DBN_EINVAL, "db_set_engine: model is NULL");
    if (engine == DBN_ENGINE_FP32) {
        m->engine = engine;
        return DBN_OK;
    }
    if (engine == DBN_ENGINE_TCGEN05) {
        if (!m->tc_available) return fail(DBN_EINVAL, "tcgen05 engine is not available in this build");
        m->engine = engine;
        return DBN_OK;
    }
    return fail(DBN_EINVAL, "unknown engine %d", engine);
}

int db_get_engine(const db_model* m) { return m ? m->engine : DBN_EINVAL; }

// Host-buffer predict: the windows are processed in chunks on two streams so that the H2D copy of
// chunk i+1 overlaps the kernel of chunk i.  Results go D2H into pinned staging (a copy into the
// caller's pageable array would block the host and serialise the pipeline) and are copied out when
// the slot is reused / at the end.
// A PAGEABLE caller array (what numpy hands over, and what Keras' model.predict got at classify.py:361) would make
// cudaMemcpyAsync stage it through the driver on the calling thread at ~10 GB/s; instead the gather pool's threads
// copy the chunk into pinned staging - float64 windows are cast to float32 there, exactly the cast the kernel does
// otherwise (and Keras does), which halves the bytes on the wire - and the copy engine takes it from there.
static int predict_host(db_model* m, const void* x, bool is_f64, int64_t n, float* probs) {
    if (!m) return fail(DBN_EINVAL, "predict: model is NULL");
    if (n < 0) return fail(DBN_EINVAL, "predict: n < 0");
    if (n == 0) return DBN_OK;
    if (!x || !probs) return fail(DBN_EINVAL, "predict: NULL buffer");
    DBN_CUDA(cudaSetDevice(m->device));
    const size_t esz = is_f64 ? sizeof(double) : sizeof(float);
    const int64_t kChunk = 8192;   // windows per pipelined chunk (32 MiB of fp32 input)
    // The first chunks are small and double up to kChunk: the pipeline then fills after the 4 MiB copy
    // of 1024 windows instead of a 32 MiB one (the kernel of chunk i overlaps the copy of chunk i+1).
    int64_t chunk = 1024;
    const int64_t cap = std::min(kChunk, n);   // buffers are sized once for the largest chunk
    const size_t row_in = static_cast<size_t>(m->input_size) * esz;
    const size_t row_out = static_cast<size_t>(m->n_classes) * sizeof(float);
    bool pageable = true;
    {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, x) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) pageable = false;
        cudaGetLastError();
        if (getenv("DEEPBINNER_B200_NO_STAGE")) pageable = false;   // measurement knob: let the driver stage pageable arrays
    }
    const size_t row_stage = static_cast<size_t>(m->input_size) * sizeof(float);
    DBN_CUDA(cudaEventRecord(m->ev_start, m->streams[0]));
    DBN_CUDA(cudaStreamWaitEvent(m->streams[1], m->ev_start, 0));
    int64_t done = 0;
    int64_t pending_off[2] = {-1, -1}, pending_cnt[2] = {0, 0};
    auto drain = [&](int slot) -> int {
        if (pending_off[slot] < 0) return 0;
        DBN_CUDA(cudaEventSynchronize(m->ev_slot[slot]));
        std::memcpy(reinterpret_cast<char*>(probs) + pending_off[slot] * row_out, m->h_out[slot],
                    pending_cnt[slot] * row_out);
        pending_off[slot] = -1;
        return 0;
    };
    int slot = 0;
    while (done < n) {
        const int64_t cnt = std::min(chunk, n - done);
        chunk = std::min(chunk * 2, kChunk);
        cudaStream_t st = m->streams[slot];
        int rc = drain(slot);
        if (rc) return rc;
        rc = grow(&m->d_in[slot], &m->d_in_bytes[slot], cap * row_in);
        if (rc) return rc;
        rc = grow(&m->d_out[slot], &m->d_out_bytes[slot], cap * row_out);
        if (rc) return rc;
        rc = grow_host(&m->h_out[slot], &m->h_out_bytes[slot], cap * row_out);
        if (rc) return rc;
        if (pageable) {
            // (drain(slot) above waited for the previous chunk of this slot, hence for its copy out of h_in[slot])
            rc = grow_host(&m->h_in[slot], &m->h_in_bytes[slot], cap * row_stage);
            if (rc) return rc;
            float* const stage = m->h_in[slot];
            const int parts = m->pool ? m->pool->parts() : 1;
            const int64_t elems = cnt * m->input_size;
            auto stage_part = [&](int part) {
                const int64_t e0 = elems * part / parts, e1 = elems * (part + 1) / parts;
                if (is_f64) {
                    const double* src = static_cast<const double*>(x) + done * m->input_size;
                    for (int64_t e = e0; e < e1; ++e) stage[e] = static_cast<float>(src[e]);
                } else {
                    std::memcpy(stage + e0, static_cast<const float*>(x) + done * m->input_size + e0,
                                static_cast<size_t>(e1 - e0) * sizeof(float));
                }
            };
            if (m->pool && elems >= (1 << 17)) {
                m->pool->parallel(stage_part);
            } else {
                for (int part = 0; part < parts; ++part) stage_part(part);
            }
            DBN_CUDA(cudaMemcpyAsync(m->d_in[slot], stage, cnt * row_stage, cudaMemcpyHostToDevice, st));
            rc = launch_predict(m, m->d_in[slot], false, cnt, m->d_out[slot], st);
        } else {
            DBN_CUDA(cudaMemcpyAsync(m->d_in[slot], static_cast<const char*>(x) + done * row_in,
                                     cnt * row_in, cudaMemcpyHostToDevice, st));
            rc = launch_predict(m, m->d_in[slot], is_f64, cnt, m->d_out[slot], st);
        }
        if (rc) return rc;
        DBN_CUDA(cudaMemcpyAsync(m->h_out[slot], m->d_out[slot], cnt * row_out, cudaMemcpyDeviceToHost, st));
        DBN_CUDA(cudaEventRecord(m->ev_slot[slot], st));
        pending_off[slot] = done;
        pending_cnt[slot] = cnt;
        done += cnt;
        slot ^= 1;
    }
    // join stream 1 into stream 0 before the stop event, then drain both slots
    DBN_CUDA(cudaStreamWaitEvent(m->streams[0], m->ev_slot[1], 0));
    DBN_CUDA(cudaEventRecord(m->ev_stop, m->streams[0]));
    for (int sl = 0; sl < 2; ++sl) {
        int rc = drain(sl);
        if (rc) return rc;
    }
    DBN_CUDA(cudaEventSynchronize(m->ev_stop));
    DBN_CUDA(cudaEventElapsedTime(&m->last_ms, m->ev_start, m->ev_stop));
    return DBN_OK;
}

int db_predict_windows(db_model* m, const float* x, int64_t n, float* probs) {
    return predict_host(m, x, false, n, probs);
}

int db_predict_windows_f64(db_model* m, const double* x, int64_t n, float* probs) {
    return predict_host(m, x, true, n, probs);
}

int db_predict_windows_device(db_model* m, const float* d_x, int64_t n, float* d_probs,
                              void* stream) {
    if (!m) return fail(DBN_EINVAL, "predict: model is NULL");
    if (n < 0) return fail(DBN_EINVAL, "predict: n < 0");
    if (n == 0) return DBN_OK;
    if (!d_x || !d_probs) return fail(DBN_EINVAL, "predict: NULL buffer");
    DBN_CUDA(cudaSetDevice(m->device));
    return launch_predict(m, d_x, false, n, d_probs, static_cast<cudaStream_t>(stream));
}

int db_call_batch_submit(db_model* m, const int16_t* const* signals, const int64_t* lengths, int n_reads, int side,
                         int scan_size, double score_diff, int* job) {
    if (n_reads > 0 && (!signals || !lengths)) return fail(DBN_EINVAL, "call_batch: NULL buffer");
    return submit_job(m, [&](int i, const int16_t** p, int64_t* len) { *p = signals[i]; *len = lengths[i]; },
                      n_reads, side, scan_size, score_diff, job);
}

int db_call_batch_submit_packed(db_model* m, const int16_t* samples, const int64_t* offsets, int n_reads, int side,
                                int scan_size, double score_diff, int* job) {
    if (n_reads > 0 && (!samples || !offsets)) return fail(DBN_EINVAL, "call_batch: NULL buffer");
    // a page-locked caller buffer is read by the copy engine directly (see the header for the lifetime rule)
    const int16_t* pinned = nullptr;
    if (m && n_reads > 0) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, samples) == cudaSuccess && attr.type == cudaMemoryTypeHost) pinned = samples;
        cudaGetLastError();
    }
    return submit_job(m, [&](int i, const int16_t** p, int64_t* len) { *p = samples + offsets[i]; *len = offsets[i + 1] - offsets[i]; },
                      n_reads, side, scan_size, score_diff, job, pinned);
}

int db_call_batch_wait(db_model* m, int job, float* probs, int8_t* calls) {
    if (!m) return fail(DBN_EINVAL, "call_batch: model is NULL");
    if (job < 0 || job >= db_model::kJobSlots || !m->jobs[job].busy) return fail(DBN_EINVAL, "call_batch: no such job in flight");
    db_model::CallJob& J = m->jobs[job];
    J.busy = false;
    if (J.n_reads == 0) return DBN_OK;
    if (!probs || !calls) return fail(DBN_EINVAL, "call_batch: NULL buffer");
    DBN_CUDA(cudaSetDevice(m->device));
    for (cudaEvent_t e : J.ev_done) DBN_CUDA(cudaEventSynchronize(e));
    DBN_CUDA(cudaEventSynchronize(J.ev_stop));
    std::memcpy(probs, J.h_probs, sizeof(float) * m->n_classes * J.n_reads);
    std::memcpy(calls, J.h_calls, static_cast<size_t>(J.n_reads));
    DBN_CUDA(cudaEventElapsedTime(&m->last_ms, J.ev_start, J.ev_stop));
    return DBN_OK;
}

int db_call_batch(db_model* m, const int16_t* samples, const int64_t* offsets, int n_reads,
                  int side, int scan_size, double score_diff, float* probs, int8_t* calls) {
    if (n_reads > 0 && (!samples || !offsets || !probs || !calls)) return fail(DBN_EINVAL, "call_batch: NULL buffer");
    for (int i = 0; i < n_reads; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(DBN_EINVAL, "call_batch: offsets must be non-decreasing");
    int job = -1;
    int rc = db_call_batch_submit_packed(m, samples, offsets, n_reads, side, scan_size, score_diff, &job);
    if (rc) return rc;
    return db_call_batch_wait(m, job, probs, calls);
}

int db_call_batch_device(db_model* m, const int16_t* d_samples, const int64_t* d_offsets,
                         int n_reads, int side, int scan_size, double score_diff, float* d_probs,
                         int8_t* d_calls, float* d_step_probs, void* stream) {
    if (!m) return fail(DBN_EINVAL, "call_batch: model is NULL");
    if (n_reads < 0) return fail(DBN_EINVAL, "call_batch: n_reads < 0");
    if (side != DBN_SIDE_START && side != DBN_SIDE_END) return fail(DBN_EINVAL, "call_batch: bad side");
    int steps = 0;
    int rc = check_scan(m, scan_size, &steps);
    if (rc) return rc;
    if (n_reads == 0) return DBN_OK;
    if (!d_samples || !d_offsets || !d_probs || !d_calls) return fail(DBN_EINVAL, "call_batch: NULL buffer");
    DBN_CUDA(cudaSetDevice(m->device));
    float* d_step = d_step_probs;
    if (!d_step) {
        rc = grow(&m->d_step, &m->d_step_bytes,
                  sizeof(float) * static_cast<size_t>(m->n_classes) * n_reads * steps);
        if (rc) return rc;
        d_step = m->d_step;
    }
    return launch_call_batch(m, d_samples, d_offsets, n_reads, side, steps, score_diff, d_step,
                             d_probs, d_calls, static_cast<cudaStream_t>(stream));
}

int db_tc_num_jobs(const db_model* m) { return (m && m->tc) ? tc_num_jobs(m->tc) : 0; }

int db_tc_job_table(const void* weights_blob, size_t blob_bytes, int which, int32_t* out, int max_jobs) {
    if (!weights_blob || !out) return fail(DBN_EINVAL, "db_tc_job_table: NULL buffer");
    Blob blob;
    const std::string err = parse_blob(weights_blob, blob_bytes, &blob);
    if (!err.empty()) return fail(DBN_EFORMAT, "%s", err.c_str());
    const int n = tc_job_table(blob, which, out, max_jobs);
    if (n < 0) return fail(DBN_EFORMAT, "no tcgen05 job table for this model");
    return n;
}

int db_tc_packed(const void* weights_blob, size_t blob_bytes, int which, unsigned char* w_out, int64_t w_cap,
                 float* prm_out, int64_t prm_cap, int64_t* w_bytes, int64_t* prm_floats) {
    if (!weights_blob || !w_bytes || !prm_floats) return fail(DBN_EINVAL, "db_tc_packed: NULL buffer");
    Blob blob;
    const std::string err = parse_blob(weights_blob, blob_bytes, &blob);
    if (!err.empty()) return fail(DBN_EFORMAT, "%s", err.c_str());
    if (tc_packed(blob, which, w_out, w_cap, prm_out, prm_cap, w_bytes, prm_floats))
        return fail(DBN_EFORMAT, "no tcgen05 job table for this model");
    return DBN_OK;
}

int db_tc_debug_dump(db_model* m, const float* x, int job, unsigned char* out) {
    if (!m || !m->tc) return fail(DBN_EINVAL, "tcgen05 engine not available");
    if (!x || !out) return fail(DBN_EINVAL, "NULL buffer");
    DBN_CUDA(cudaSetDevice(m->device));
    const size_t dump = 2 * 98688;
    int rc = grow(&m->d_in[0], &m->d_in_bytes[0], 2 * 1024 * sizeof(float));
    if (rc) return rc;
    rc = grow(&m->d_out[0], &m->d_out_bytes[0], dump);
    if (rc) return rc;
    cudaStream_t st = m->streams[0];
    DBN_CUDA(cudaMemcpyAsync(m->d_in[0], x, 2 * 1024 * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = tc_debug_dump(m->tc, static_cast<const float*>(m->d_in[0]), job,
                       reinterpret_cast<unsigned char*>(m->d_out[0]), st);
    if (rc) return rc;
    DBN_CUDA(cudaMemcpyAsync(out, m->d_out[0], dump, cudaMemcpyDeviceToHost, st));
    DBN_CUDA(cudaStreamSynchronize(st));
    return DBN_OK;
}

int db_tc_trace(db_model* m, const float* d_x, int n, float* d_probs, int64_t* trace) {
    if (!m || !m->tc) return fail(DBN_EINVAL, "tcgen05 engine not available");
    DBN_CUDA(cudaSetDevice(m->device));
    const size_t bytes = 32 * 2 * 16 * sizeof(long long);
    int rc = grow(&m->d_step, &m->d_step_bytes, bytes);
    if (rc) return rc;
    cudaStream_t st = m->streams[0];
    DBN_CUDA(cudaMemsetAsync(m->d_step, 0, bytes, st));
    rc = tc_trace(m->tc, d_x, n, d_probs, reinterpret_cast<long long*>(m->d_step), st);
    if (rc) return rc;
    DBN_CUDA(cudaMemcpyAsync(trace, m->d_step, bytes, cudaMemcpyDeviceToHost, st));
    DBN_CUDA(cudaStreamSynchronize(st));
    return DBN_OK;
}

int db_tc_trace_call(db_model* m, const int16_t* d_samples, const int64_t* d_offsets, int n_reads, float* d_probs,
                     int64_t* trace) {
    if (!m || !m->tc) return fail(DBN_EINVAL, "tcgen05 engine not available");
    DBN_CUDA(cudaSetDevice(m->device));
    const size_t bytes = 32 * 2 * 16 * sizeof(long long);
    int rc = grow(&m->d_step, &m->d_step_bytes, bytes);
    if (rc) return rc;
    cudaStream_t st = m->streams[0];
    DBN_CUDA(cudaMemsetAsync(m->d_step, 0, bytes, st));
    rc = tc_trace_call(m->tc, d_samples, d_offsets, n_reads, d_probs, reinterpret_cast<long long*>(m->d_step), st);
    if (rc) return rc;
    DBN_CUDA(cudaMemcpyAsync(trace, m->d_step, bytes, cudaMemcpyDeviceToHost, st));
    DBN_CUDA(cudaStreamSynchronize(st));
    return DBN_OK;
}

float db_last_gpu_ms(const db_model* m) { return m ? m->last_ms : 0.f; }

int64_t db_kernel_launches(const db_model* m) { return m ? m->launches : 0; }

}  // extern "C"
#pragma GCC visibility pop
