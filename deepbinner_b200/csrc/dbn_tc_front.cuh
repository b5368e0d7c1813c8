// Persistent front kernel of the SPLIT tensor-core engine (experimental): conv1d_1 .. conv1d_4 for a
// stream of window pairs per CTA.  Included by dbn_tc.cu (same translation unit).
//
// k_tc_forward<.., kFront> runs one pair per CTA, so every pair pays the pipeline fill (input load +
// z-score + conv1d_1 before the first MMA, ~3.3 k cycles) and drain (last epilogue + copy-out + CTA
// turn-around) - about 8 k of ~44 k cycles.  Here a CTA keeps its two window slots busy over many
// pairs: as soon as slot w has finished conv1d_4 (epilogue + copy of the BatchNorm_2 tensor to the
// staging buffer) the epilogue warps load the slot's NEXT window and run its conv1d_1 into the same
// region while the tensor pipe is still working on the other slot.  The job sequence seen by the
// weight loader and the MMA issuer is simply conv1d_2, 3, 4, 2, 3, 4, ... with both slots in
// lock-step, so all barrier phases keep alternating; the arrival that would follow conv1d_4's
// epilogue is replaced by the one after the next window's conv1d_1.
// Only useful when a launch has more pairs than SMs (grid = min(pairs, SM count)).
#pragma once

namespace dbn {

// Window `idx` -> normalised samples in the thread's registers -> conv1d_1 + BN1 into region `w`.
template <bool kCallMode>
__device__ __forceinline__ void front_prepare_window(const TcParams& P, const float* __restrict__ x,
                                                     const double* __restrict__ xd,
                                                     const int16_t* __restrict__ samples,
                                                     const int64_t* __restrict__ offsets, int n_reads, int side,
                                                     int idx, int w, unsigned char* smem, uint32_t sbase, int tid) {
    // (the conv1 parameters are re-fetched per window - L1/L2 hits - rather than kept in 48 registers
    // across the MMA-epilogue passes in between)
    Conv1Params c1;
    load_conv1_params(P, tid, c1);
    WindowInput in{};
    float xv[3];
    if (kCallMode) {
        const int ewarp = tid >> 5;
        const int step = idx / n_reads, read = idx % n_reads;
        const int64_t off = offsets[read];
        in.region = samples + off;
        in.g = window_geometry(static_cast<int>(offsets[read + 1] - off), step, side);
        long long s1 = 0, s2 = 0;   // exact integer sums over the slice, reduced over the 384 threads
        for (int i = tid; i < in.g.n; i += kEpiThreads) {
            const long long v = in.region[in.g.a + i];
            s1 += v;
            s2 += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        long long* red = reinterpret_cast<long long*>(smem + kSmemBar + 128);
        epi_bar_sync();   // previous readers are done with the scratch
        if ((tid & 31) == 0) { red[ewarp] = s1; red[12 + ewarp] = s2; }
        epi_bar_sync();
        s1 = 0; s2 = 0;
        for (int i = 0; i < kEpiWarps; ++i) { s1 += red[i]; s2 += red[12 + i]; }
        in.mean = 0.0; in.stdev = 0.0;
        if (in.g.n > 0) zscore_params(s1, s2, in.g.n, &in.mean, &in.stdev);
    } else if (x) {
        in.x = x + static_cast<size_t>(idx) * kInputSize;
    } else {
        in.xd = xd + static_cast<size_t>(idx) * kInputSize;
    }
    fetch_window_inputs(in, tid, xv);
    conv1_stage(c1, xv[0], xv[1], xv[2], sbase + (w ? kSmemAct1 : kSmemAct0), smem + (w ? kSmemAct1 : kSmemAct0), tid);
    fence_proxy_async();
}

template <bool kCallMode>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_tc_front(TcParams P, const float* __restrict__ x, const double* __restrict__ xd,
               const int16_t* __restrict__ samples, const int64_t* __restrict__ offsets, int n_reads, int side,
               int n_windows) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t wbuf = sbase + kSmemWbuf;
    const uint32_t prm = sbase + kSmemPrm;
    const uint32_t bar0 = sbase + kSmemBar;
    const uint32_t bar_wfull[2] = {bar0 + 0, bar0 + 8};
    const uint32_t bar_wfree[2] = {bar0 + 16, bar0 + 24};
    const uint32_t bar_mma[2] = {bar0 + 32, bar0 + 40};
    const uint32_t bar_epi[2] = {bar0 + 48, bar0 + 56};
    const uint32_t bar_final = bar0 + 64;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 96);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) >> 5, 0);
    const bool is_epi = warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps;
    constexpr int kFrontJobs = 3;   // conv1d_2, conv1d_3, conv1d_4 (the first entries of the job table)

    const int npairs = (n_windows + 1) / 2;
    const int iters = (npairs - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (threadIdx.x == kLoadWarp * 32) {
        mbar_init(bar_wfull[0], 1);
        mbar_init(bar_wfull[1], 1);
        mbar_init(bar_wfree[0], 1);
        mbar_init(bar_wfree[1], 1);
        mbar_init(bar_mma[0], 1);
        mbar_init(bar_mma[1], 1);
        mbar_init(bar_epi[0], kEpiArrivals);
        mbar_init(bar_epi[1], kEpiArrivals);
        mbar_init(bar_final, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(sbase + kSmemBar + 96, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (is_epi) {
        // ================= epilogue / CUDA-core warps =================
        const int tid = static_cast<int>(threadIdx.x) - kEpiWarp0 * 32;
        for (int i = tid; i < P.prm_floats / 4; i += kEpiThreads)
            reinterpret_cast<float4*>(smem + kSmemPrm)[i] = __ldg(reinterpret_cast<const float4*>(P.prm) + i);
        auto window_index = [&](int it, int w) { return 2 * (static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x)) + w; };
#pragma unroll 1
        for (int w = 0; w < 2; ++w) {
            front_prepare_window<kCallMode>(P, x, xd, samples, offsets, n_reads, side,
                                            min(window_index(0, w), n_windows - 1), w, smem, sbase, tid);
            epi_arrive(bar0 + 48 + 8 * w);
        }
        epi_bar_sync();   // parameter block staged by all epilogue threads is now visible
        uint32_t mma_phase = 0;   // both slots: one pass per slot and job, so one parity serves both
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int j = 0; j < kFrontJobs; ++j) {
                const TcJob& J = c_jobs[j];
#pragma unroll 1
                for (int w = 0; w < 2; ++w) {
                    const uint32_t act = sbase + (w ? kSmemAct1 : kSmemAct0);
                    run_epilogue(P, J, act, sbase + kSmemAct0, w, prm, tmem_base + w * kTmemWindowCols, tid,
                                 bar0 + 32 + 8 * w, mma_phase, nullptr, nullptr, nullptr);
                    if (j + 1 < kFrontJobs) {
                        fence_proxy_async();
                        tc_fence_before();
                        epi_arrive(bar0 + 48 + 8 * w);
                        continue;
                    }
                    // conv1d_4 done for this slot: hand the window over to the tail kernel ...
                    epi_bar_sync();   // every epilogue thread has written its part of the tensor
                    const int idx = window_index(it, w);
                    if (idx < n_windows) {
                        const uint4* src = reinterpret_cast<const uint4*>(smem + (w ? kSmemAct1 : kSmemAct0));
                        uint4* dst = reinterpret_cast<uint4*>(P.mid + static_cast<size_t>(idx) * 49536);
                        for (int i = tid; i < 49536 / 16; i += kEpiThreads) dst[i] = src[i];
                    }
                    // ... and start the slot's next window (its conv1d_1 overwrites the region; the
                    // barriers inside conv1_stage order every thread's copy before any of those stores)
                    if (it + 1 < iters) {
                        tc_fence_before();
                        front_prepare_window<kCallMode>(P, x, xd, samples, offsets, n_reads, side,
                                                        min(window_index(it + 1, w), n_windows - 1), w, smem, sbase, tid);
                        epi_arrive(bar0 + 48 + 8 * w);
                    }
                }
                mma_phase ^= 1;
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        if (tmem_base != 0) __trap();
        if (elect_one()) {
            constexpr uint32_t leader = 1;
            uint32_t wfull_phase = 0, epi_phase[2] = {0, 0};
            const uint32_t wp16[2] = {wbuf >> 4, (wbuf + kWPart0) >> 4};
            const uint32_t act16_0 = (sbase + kSmemAct0) >> 4, act16_1 = (sbase + kSmemAct1) >> 4;
            IssueArgs nxt = load_issue_args(P.jobs);
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int j = 0; j < kFrontJobs; ++j) {
                    const IssueArgs J = nxt;
                    nxt = load_issue_args(P.jobs + (j + 1) % kFrontJobs);
                    const uint32_t blk16 = 2u * J.n;
                    const uint32_t tap16[3] = {J.tap16[0], J.tap16[1], J.tap16[2]};
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        mbar_wait(bar_epi[w], epi_phase[w]);   // input written and previous accumulators drained
                        epi_phase[w] ^= 1;
                        tc_fence_after();
                        const uint32_t dwin = w * kTmemWindowCols;
                        if (w == 0) mbar_wait(bar_wfull[0], wfull_phase);
                        issue_job_part<0>(J.ntaps, J.ncb, dwin, J.ntiles, (w ? act16_1 : act16_0), tap16, J.cb0, J.lp,
                                          J.lo16, wp16[0], blk16, J.n, J.idesc, true, leader);
                        if (w == 1) tc_commit(bar_wfree[0], leader);
                        if (w == 0) mbar_wait(bar_wfull[1], wfull_phase);
                        issue_job_part<1>(J.ntaps, J.ncb, dwin, J.ntiles, (w ? act16_1 : act16_0), tap16, J.cb0, J.lp,
                                          J.lo16, wp16[1], blk16, J.n, J.idesc, false, leader);
                        tc_commit(bar_mma[w], leader);
                        if (w == 1) tc_commit(bar_wfree[1], leader);
                    }
                    wfull_phase ^= 1;
                }
            }
            tc_commit(bar_final, leader);
            mbar_wait(bar_final, 0);
        }
    } else if (warp == kLoadWarp && elect_one()) {
        // ================= weight loader =================
        uint32_t free_phase = 0;
        int g = 0;   // jobs loaded so far
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int j = 0; j < kFrontJobs; ++j, ++g) {
                const TcJob& J = c_jobs[j];
                const unsigned char* src = P.w + J.w_goff;
                if (g > 0) mbar_wait(bar_wfree[0], free_phase);
                mbar_expect_tx(bar_wfull[0], J.w_part[0]);
                bulk_g2s(wbuf, src, J.w_part[0], bar_wfull[0]);
                if (g > 0) {
                    mbar_wait(bar_wfree[1], free_phase);
                    free_phase ^= 1;
                }
                mbar_expect_tx(bar_wfull[1], J.w_part[1]);
                bulk_g2s(wbuf + kWPart0, src + J.w_part[0], J.w_part[1], bar_wfull[1]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace dbn
