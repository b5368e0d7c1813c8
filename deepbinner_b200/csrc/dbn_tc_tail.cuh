// Tail kernel of the SPLIT tensor-core engine (experimental, DBN_ENGINE_TCGEN05_SPLIT): conv1d_5 ..
// conv1d_20 + head for FOUR windows per CTA.  Included by dbn_tc.cu (same translation unit, after the
// PTX wrappers, the job table types and the issue helpers).
//
// Why: with two windows per CTA everything after conv1d_4 is bound by latencies (MMA start-up,
// MMA <-> epilogue hand-offs, the single issuer thread), not by work.  The activations of conv1d_5 on
// are at most 49 536 B per window (split-bf16 [6][258][8] hi + lo), so four windows fit in shared
// memory once the network is cut at conv1d_4: the front kernel (k_tc_forward<.., kFront = true>)
// writes the pooled BatchNorm_2 tensor of every window to a global staging buffer, this kernel loads
// four of them and runs four independent dependency chains through the same MMA issuer / epilogue
// warps.
//
// Shared memory (232 000 B): 4 regions x 49 536 B | weight buffer 27 648 B | parameters 5 696 B |
// barriers.  Per region during the inception block: X / T15 @0 (12 672 B, T15 reuses X's slot - job
// order c12+14, c11, c10f, c15, c13, c16), T1214 @12672 (8 448 B), and ONE of the four parity arrays
// of the stacked concat tensor ([24][72][8] = 27 648 B) @21120: Ye_hi / Ye_lo / Yo_hi / Yo_lo live
// in regions 0 / 1 / 2 / 3, so the hi->lo distance is one region for both parities.
// TMEM: conv1d_5..9 use 2 tiles x 64 columns per window (4 x 128 = 512); joint jobs rotate over three
// 128-column slots (inception: one 64-column tile per window pair; conv1d_17..20: one M=128 tile for
// the four stacked windows, row = 18 w + position).
#pragma once

namespace dbn {

#ifndef DBN_TAIL_PART_MAJOR
#define DBN_TAIL_PART_MAJOR 0
#endif
constexpr int kTW = 4;                               // windows per CTA
constexpr int kTReg = 49536;                         // bytes per window region
constexpr int kTSmemWbuf = kTW * kTReg;              // 198144
constexpr int kTPrmFloats = 1424;                    // per-job bias / folded BN of conv1d_5 .. conv1d_20
constexpr int kTSmemPrm = kTSmemWbuf + kWbufBytes;   // 225792
constexpr int kTSmemBar = kTSmemPrm + kTPrmFloats * 4;   // 231488
constexpr int kTSmemBytes = kTSmemBar + 512;         // 232000
static_assert(kTSmemBytes <= 232448, "shared memory budget of the tail kernel");
constexpr int kTYOff = 21120;                        // parity array of the stacked concat tensor, per region
constexpr int kTYRows = kTW * kStackPitch;           // 72
constexpr int kTYArray = 24 * kTYRows * 16;          // 27648
static_assert(kTYOff + kTYArray <= kTReg, "parity array must fit behind the inception tensors");
constexpr int kTWinCols = 128;                       // TMEM columns per window (conv1d_5 .. 9)
constexpr int kTHeadScratch = 16384;                 // region 0: 36 x 16 floats behind the conv1d_19 output

__constant__ TcJob c_tjobs[kMaxJobs];

struct TailParams {
    int njobs;
    const TcJob* jobs;          // global copy of the job table
    const unsigned char* w;     // packed bf16 weights of conv1d_5 .. conv1d_20
    const float* prm;
    int prm_floats;
    int n_classes;
    const unsigned char* mid;   // [n_windows][kTReg]: BatchNorm_2 output of the front kernel
    long long* trace;           // diagnostics instantiation: [job][window][8] clock64 stamps of CTA 0
};

enum TailMode { T_SINGLE = 0, T_PAIR = 1, T_STACK = 2 };

// Zero rows written by a pass: the halo rows of an ordinary output tensor, or (first parity job)
// rows 16 / 17 of every window in every parity array.
__device__ __forceinline__ void tail_zero_rows(const EpiArgs& A, uint32_t sbase, uint32_t act, int tid) {
    if (A.kind == EPI_PARITY) {
        if (A.zero_y) {   // 4 arrays x 24 channel-groups x 4 windows x 2 rows = 768 rows of 16 B
            for (int item = tid; item < 768; item += kEpiThreads) {
                const int cg = item % 24, rest = item / 24, arr = rest & 3, wr = rest >> 2;   // wr = 2 w + row
                st_shared_v4(sbase + arr * kTReg + kTYOff +
                                 (cg * kTYRows + (wr >> 1) * kStackPitch + 16 + (wr & 1)) * 16,
                             make_uint4(0, 0, 0, 0));
            }
        }
    } else if (tid < 32 && (tid & 7) < A.out_ncg) {
        const int cg = tid & 7, which = tid >> 3;
        const uint32_t a0 = act + A.out_off + (which & 1 ? A.out_lo_delta : 0) +
                            (cg * A.out_lp + (which & 2 ? A.out_L + 1 : 0)) * 16;
        st_shared_v4(a0, make_uint4(0, 0, 0, 0));
    }
}

// MODE T_SINGLE: one window (`w`), M=128 tiles, TMEM lane = position within the tile.
//      T_PAIR:   two tiles = two window pairs; M=64 accumulators, lanes 0-15 of a quadrant = 16 rows of
//                the pair's first window, lanes 16-31 = the same rows of its second window.
//      T_STACK:  one M=128 tile over the tensor of the four stacked windows (region 0).
template <bool POOL, bool BN, bool PARITY, int MODE>
__device__ __forceinline__ void tail_epilogue_tiles(const EpiArgs& A, uint32_t sbase, int w, uint32_t prm,
                                                    uint32_t tmem_win, int tid, uint32_t bar, uint32_t parity) {
    constexpr int NC = 16;
    const int lane = tid & 31;
    const int q = ((tid >> 5) + kEpiWarp0) & 3, h = tid >> 7;
    const bool active = h * NC < A.n;
    constexpr bool stack = MODE == T_STACK;
    const int row = MODE == T_PAIR ? q * 16 + (lane & 15) : q * 32 + lane;
    const int ntiles = MODE == T_PAIR ? 2 : (MODE == T_STACK ? 1 : A.ntiles), L = A.L;
    const int cg0 = A.out_cg_base + (h * NC) / 8;
    const int out_lp = A.out_lp, out_lo = A.out_lo_delta;
    const uint32_t taddr0 = tmem_win + h * NC + (static_cast<uint32_t>(q * 32) << 16);
    constexpr int PN = POOL ? NC / 2 : NC;
    const int odd = lane & 1;
    const int pc0 = POOL ? odd * (NC / 2) : 0;
    float bias[PN], sc[BN ? PN : 1], sh[BN ? PN : 1];
    {
        const uint32_t bias_a = prm + (A.bias_off + h * NC + pc0) * 4;
#pragma unroll
        for (int g = 0; g < PN / 4; ++g) {
            const float4 b = ld_shared_f4(bias_a + g * 16);
            bias[4 * g] = b.x; bias[4 * g + 1] = b.y; bias[4 * g + 2] = b.z; bias[4 * g + 3] = b.w;
        }
        if (BN) {
            const uint32_t bn_a = prm + (A.bn_off + h * NC + pc0) * 4;
#pragma unroll
            for (int g = 0; g < PN / 4; ++g) {
                const float4 a = ld_shared_f4(bn_a + g * 16), b = ld_shared_f4(bn_a + 192 + g * 16);
                sc[4 * g] = a.x; sc[4 * g + 1] = a.y; sc[4 * g + 2] = a.z; sc[4 * g + 3] = a.w;
                sh[4 * g] = b.x; sh[4 * g + 1] = b.y; sh[4 * g + 2] = b.z; sh[4 * g + 3] = b.w;
            }
        }
    }
    mbar_wait(bar, parity);
    tc_fence_after();
    auto zero_rows = [&]() {
        if (MODE == T_PAIR) {
            if (A.kind == EPI_PARITY) {
                tail_zero_rows(A, sbase, sbase, tid);   // one call covers the four windows of every array
            } else {
                for (int ww = 0; ww < kTW; ++ww) tail_zero_rows(A, sbase, sbase + ww * kTReg, tid);
            }
        } else {
            tail_zero_rows(A, sbase, sbase + (MODE == T_STACK ? 0 : w) * kTReg, tid);
        }
    };
    if (!active) {   // (kept apart: a conditionally executed tcgen05.ld would push r[] into local memory)
        zero_rows();
        return;
    }
    uint32_t r[NC];
    tmem_load_cols<NC>(taddr0, r);
    zero_rows();
    for (int tile = 0; tile < ntiles; ++tile) {
        const int p = MODE == T_PAIR ? row : tile * 128 + row;
        const int win = MODE == T_PAIR ? 2 * tile + (lane >> 4) : w;   // this lane's window
        const uint32_t out_base = sbase + (MODE == T_STACK ? 0 : win) * kTReg + A.out_off;
        tmem_wait_ld();
        float acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = __uint_as_float(r[c]);
        if (tile + 1 < ntiles) tmem_load_cols<NC>(taddr0 + (tile + 1) * kTmemTileCols, r);
        const int qpos = POOL ? p >> 1 : p;
        const bool valid = stack ? (p < L && (p % kStackPitch) < 16) : (p < L);
        const float es = (A.edge15 && (p == 0 || p == L - 1)) ? 1.5f : 1.0f;
        if (POOL) {
            if (A.edge15) {
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[c] *= es;
            }
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float keep = odd ? acc[8 + e] : acc[e];
                const float send = odd ? acc[e] : acc[8 + e];
                v[e] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e] + bias[e], 0.f);
            if (BN) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(sc[e], v[e], sh[e]);
            }
            uint4 hi, lo;
            split8(v, &hi, &lo);
            if (valid) {
                if (PARITY) {   // even / odd pooled positions -> Ye (regions 0, 1) / Yo (regions 2, 3)
                    const uint32_t o = sbase + (qpos & 1) * (2 * kTReg) + kTYOff +
                                       ((cg0 + odd) * kTYRows + win * kStackPitch + (qpos >> 1)) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + kTReg, lo);
                } else {
                    const uint32_t o = out_base + ((cg0 + odd) * out_lp + qpos + 1) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + out_lo, lo);
                }
            }
        } else {
            const bool writer = stack ? (p < L) : valid;
#pragma unroll
            for (int g = 0; g < NC / 8; ++g) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(acc[g * 8 + e], es, bias[g * 8 + e]), 0.f);
                if (BN) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(sc[g * 8 + e], v[e], sh[g * 8 + e]);
                }
                uint4 hi, lo;
                split8(v, &hi, &lo);
                if (stack && !valid) {
                    hi = make_uint4(0, 0, 0, 0);
                    lo = make_uint4(0, 0, 0, 0);
                }
                if (writer) {
                    const uint32_t o = out_base + ((cg0 + g) * out_lp + qpos + 1) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + out_lo, lo);
                }
            }
        }
    }
}

// Head for the four stacked windows: conv1d_20 accumulators (row 9 w + k = position k of window w, 36
// rows: TMEM lane quadrant 0 and the first 4 lanes of quadrant 1) -> ReLU -> global average pool ->
// softmax.  Epilogue warps 0 and 1 (quadrants 0 / 1) stage the rows, then finish two windows each.
__device__ void tail_epilogue_head(int bias_off, uint32_t prm, uint32_t tmem_win, uint32_t scratch, int tid,
                                   int n_classes, float* probs, int n_windows) {
    const int lane = tid & 31, ewarp = tid >> 5;   // ewarp 0 / 1 == quadrant 0 / 1 (kEpiWarp0 % 4 == 0)
    uint32_t r[16];
    const uint32_t taddr = tmem_win + (static_cast<uint32_t>(ewarp * 32) << 16);
    tmem_ld8(taddr, r);
    tmem_ld8(taddr + 8, r + 8);
    tmem_wait_ld();
    const int row = ewarp * 32 + lane;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 b = ld_shared_f4(prm + (bias_off + 4 * g) * 4);
        float4 v;
        v.x = fmaxf(__uint_as_float(r[4 * g + 0]) + b.x, 0.f);
        v.y = fmaxf(__uint_as_float(r[4 * g + 1]) + b.y, 0.f);
        v.z = fmaxf(__uint_as_float(r[4 * g + 2]) + b.z, 0.f);
        v.w = fmaxf(__uint_as_float(r[4 * g + 3]) + b.w, 0.f);
        if (row < 40)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(scratch + (row * 16 + 4 * g) * 4), "f"(v.x),
                         "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
    }
    asm volatile("bar.sync 2, 64;" ::: "memory");   // the two head warps
    const int w = ewarp * 2 + (lane >> 4), c = lane & 15;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float t;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(scratch + ((w * 9 + k) * 16 + c) * 4));
        s += t;
    }
    const float logit = s / 8.0f;
    float m = c < n_classes ? logit : -3.0e38f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = c < n_classes ? expf(logit - m) : 0.f;
    float den = e;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    const int idx = kTW * static_cast<int>(blockIdx.x) + w;
    if (idx < n_windows && c < n_classes) probs[static_cast<size_t>(idx) * n_classes + c] = e / den;
}

__device__ __forceinline__ void tail_run_epilogue(const TailParams& P, const TcJob& J, uint32_t sbase, int w, uint32_t prm,
                                  uint32_t tmem_win, int tid, uint32_t bar, uint32_t parity, float* probs,
                                  int n_windows) {
    const EpiArgs A = load_epi_args(J);
    if (A.joint == JOINT_PAIR) {
        if (A.kind == EPI_PARITY) tail_epilogue_tiles<true, true, true, T_PAIR>(A, sbase, w, prm, tmem_win, tid, bar, parity);
        else tail_epilogue_tiles<false, false, false, T_PAIR>(A, sbase, w, prm, tmem_win, tid, bar, parity);
    } else if (A.joint == JOINT_STACK && A.kind != EPI_HEAD) {
        if (A.kind == EPI_N48_BN) tail_epilogue_tiles<false, true, false, T_STACK>(A, sbase, w, prm, tmem_win, tid, bar, parity);
        else if (A.kind == EPI_N48_POOL_BN) tail_epilogue_tiles<true, true, false, T_STACK>(A, sbase, w, prm, tmem_win, tid, bar, parity);
        else tail_epilogue_tiles<false, false, false, T_STACK>(A, sbase, w, prm, tmem_win, tid, bar, parity);
    } else if (A.kind == EPI_N48 || A.kind == EPI_N16) {
        tail_epilogue_tiles<false, false, false, T_SINGLE>(A, sbase, w, prm, tmem_win, tid, bar, parity);
    } else if (A.kind == EPI_N48_POOL_BN) {
        tail_epilogue_tiles<true, true, false, T_SINGLE>(A, sbase, w, prm, tmem_win, tid, bar, parity);
    } else {   // EPI_HEAD
        mbar_wait(bar, parity);
        tc_fence_after();
        if ((tid >> 5) < 2) tail_epilogue_head(A.bias_off, prm, tmem_win, sbase + kTHeadScratch, tid, P.n_classes, probs, n_windows);
    }
}

// kDiag: timeline stamps of CTA 0 ([0] MMA issue start, [1] issue end, [2] epilogue pass start (before the
// wait for the accumulators), [3] epilogue pass end; slot [31][0][0..1] = kernel start / inputs landed).
template <bool kDiag>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_tc_tail(TailParams P, int n_windows, float* __restrict__ probs) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t wbuf = sbase + kTSmemWbuf;
    const uint32_t prm = sbase + kTSmemPrm;
    const uint32_t bar0 = sbase + kTSmemBar;
    const uint32_t bar_wfull[2] = {bar0 + 0, bar0 + 8};
    const uint32_t bar_wfree[2] = {bar0 + 16, bar0 + 24};
    const uint32_t bar_mma = bar0 + 32;     // + 8 w
    const uint32_t bar_epi = bar0 + 64;     // + 8 w
    const uint32_t bar_final = bar0 + 96, bar_in = bar0 + 104;
    const uint32_t bar_jmma = bar0 + 128, bar_jepi = bar0 + 160;   // + 8 * (eseq & 3)
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kTSmemBar + 112);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) >> 5, 0);
    const bool is_epi = warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps;

    if (threadIdx.x == kLoadWarp * 32) {
        mbar_init(bar_wfull[0], 1);
        mbar_init(bar_wfull[1], 1);
        mbar_init(bar_wfree[0], 1);
        mbar_init(bar_wfree[1], 1);
        for (int w = 0; w < kTW; ++w) {
            mbar_init(bar_mma + 8 * w, 1);
            mbar_init(bar_epi + 8 * w, kEpiArrivals);
        }
        mbar_init(bar_final, 1);
        mbar_init(bar_in, 1);
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_jmma + 8 * i, 1);
            mbar_init(bar_jepi + 8 * i, kEpiArrivals);
        }
        fence_barrier_init();
        // the four input tensors: one bulk copy each, completion on bar_in
        mbar_expect_tx(bar_in, kTW * kTReg);
        for (int w = 0; w < kTW; ++w) {   // windows past the end of the batch re-read the last one
            const int idx = min(kTW * static_cast<int>(blockIdx.x) + w, n_windows - 1);
            bulk_g2s(sbase + w * kTReg, P.mid + static_cast<size_t>(idx) * kTReg, kTReg, bar_in);
        }
    }
    if (warp == kMmaWarp) tmem_alloc(sbase + kTSmemBar + 112, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int njobs = P.njobs;
    long long* const trace = (kDiag && blockIdx.x == 0) ? P.trace : nullptr;

    if (is_epi) {
        // ================= epilogue warps =================
        const int tid = static_cast<int>(threadIdx.x) - kEpiWarp0 * 32;
        if (trace && tid == 0) trace[(31 * kTW) * 8 + 0] = clock64();
        for (int i = tid; i < P.prm_floats / 4; i += kEpiThreads)
            reinterpret_cast<float4*>(smem + kTSmemPrm)[i] = __ldg(reinterpret_cast<const float4*>(P.prm) + i);
        mbar_wait(bar_in, 0);   // inputs landed (async proxy -> consumed by the async proxy: no fence needed)
        if (trace && tid == 0) trace[(31 * kTW) * 8 + 1] = clock64();
#pragma unroll
        for (int w = 0; w < kTW; ++w) epi_arrive(bar_epi + 8 * w);
        epi_bar_sync();         // parameter block visible to all epilogue threads
        uint32_t mma_phase = 0;   // same phase for the four windows (one pass per window and job)
        for (int j = 0; j < njobs; ++j) {
            const TcJob& J = c_tjobs[j];
            if (!J.last) continue;
            // one call site: a joint job is a single pass on the joint rings, an ordinary job four
            // passes (one per window) on the per-window barriers
            const bool joint = J.joint != 0;
            const int e = J.eseq, passes = joint ? 1 : kTW;
#pragma unroll 1
            for (int w = 0; w < passes; ++w) {
                const uint32_t bar_in_mma = joint ? bar_jmma + 8 * (e & 3) : bar_mma + 8 * w;
                const uint32_t bar_out = joint ? bar_jepi + 8 * (e & 3) : bar_epi + 8 * w;
                if (trace && tid == 0) trace[(j * kTW + w) * 8 + 2] = clock64();
                tail_run_epilogue(P, J, sbase, w, prm, tmem_base + (joint ? J.tcol : w * kTWinCols), tid, bar_in_mma,
                                  joint ? (e >> 2) & 1 : mma_phase, probs, n_windows);
                fence_proxy_async();
                tc_fence_before();
                epi_arrive(bar_out);
                if (trace && tid == 0) trace[(j * kTW + w) * 8 + 3] = clock64();
            }
            if (joint) continue;
            mma_phase ^= 1;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        if (tmem_base != 0) __trap();
        if (elect_one()) {
            constexpr uint32_t leader = 1;
            uint32_t wfull_phase = 0, epi_phase = 0;   // epi_phase: same for the four windows
            int jepi_seen = 0;
            const uint32_t wp16[2] = {wbuf >> 4, (wbuf + kWPart0) >> 4};
            const uint32_t reg16 = kTReg >> 4, act16_0 = sbase >> 4;
            IssueArgs nxt = load_issue_args(P.jobs);
            for (int j = 0; j < njobs; ++j) {
                const IssueArgs J = nxt;
                if (j + 1 < njobs) nxt = load_issue_args(P.jobs + j + 1);
                const uint32_t blk16 = 2u * J.n;
                const uint32_t tap16[3] = {J.tap16[0], J.tap16[1], J.tap16[2]};
                const bool first = J.first != 0, last = J.last != 0;
                if (J.joint) {
                    mbar_wait(bar_wfull[0], wfull_phase);
                    if (first && J.eseq == 0) {   // first joint job: all single-window epilogues done
                        for (int w = 0; w < kTW; ++w) mbar_wait(bar_epi + 8 * w, epi_phase);
                        epi_phase ^= 1;
                    }
                    for (const int need = J.need; jepi_seen < need; ++jepi_seen)
                        mbar_wait(bar_jepi + 8 * (jepi_seen & 3), (jepi_seen >> 2) & 1);
                    tc_fence_after();
                    if (trace) trace[(j * kTW) * 8 + 0] = clock64();
                    const int nw = J.joint == JOINT_PAIR ? kTW : 1;
#pragma unroll 1
                    for (int w = 0; w < nw; ++w) {   // window w: pair tile w / 2, lane half w % 2
                        const uint32_t d = J.tcol + (w >> 1) * kTmemTileCols + (static_cast<uint32_t>(16 * (w & 1)) << 16);
                        issue_job_part<0>(J.ntaps, J.ncb, d, 1, act16_0 + w * reg16, tap16, J.cb0, J.lp, J.lo16, wp16[0],
                                          blk16, J.n, J.idesc, first, leader);
                    }
                    tc_commit(bar_wfree[0], leader);
                    mbar_wait(bar_wfull[1], wfull_phase);
#pragma unroll 1
                    for (int w = 0; w < nw; ++w) {
                        const uint32_t d = J.tcol + (w >> 1) * kTmemTileCols + (static_cast<uint32_t>(16 * (w & 1)) << 16);
                        issue_job_part<1>(J.ntaps, J.ncb, d, 1, act16_0 + w * reg16, tap16, J.cb0, J.lp, J.lo16, wp16[1],
                                          blk16, J.n, J.idesc, false, leader);
                    }
                    if (last) tc_commit(bar_jmma + 8 * (J.eseq & 3), leader);
                    tc_commit(bar_wfree[1], leader);
                    if (trace) trace[(j * kTW) * 8 + 1] = clock64();
                    wfull_phase ^= 1;
                    continue;
                }
#if DBN_TAIL_PART_MAJOR
                // Variant for A/B runs (not the default): weight part 0 for all four windows, then part 1
                // for all four - part 0 is released three bursts earlier, so the next job's first part
                // streams in behind a longer stretch of MMAs; each window's epilogue starts later.
#pragma unroll 1
                for (int w = 0; w < kTW; ++w) {
                    mbar_wait(bar_epi + 8 * w, epi_phase);
                    tc_fence_after();
                    if (trace) trace[(j * kTW + w) * 8 + 0] = clock64();
                    if (w == 0) mbar_wait(bar_wfull[0], wfull_phase);
                    issue_job_part<0>(J.ntaps, J.ncb, w * kTWinCols, J.ntiles, act16_0 + w * reg16, tap16, J.cb0, J.lp,
                                      J.lo16, wp16[0], blk16, J.n, J.idesc, first, leader);
                }
                tc_commit(bar_wfree[0], leader);
                mbar_wait(bar_wfull[1], wfull_phase);
#pragma unroll 1
                for (int w = 0; w < kTW; ++w) {
                    issue_job_part<1>(J.ntaps, J.ncb, w * kTWinCols, J.ntiles, act16_0 + w * reg16, tap16, J.cb0, J.lp,
                                      J.lo16, wp16[1], blk16, J.n, J.idesc, false, leader);
                    if (last) tc_commit(bar_mma + 8 * w, leader);
                    if (trace) trace[(j * kTW + w) * 8 + 1] = clock64();
                }
                tc_commit(bar_wfree[1], leader);
#else
#pragma unroll 1
                for (int w = 0; w < kTW; ++w) {
                    mbar_wait(bar_epi + 8 * w, epi_phase);   // input written and previous accumulators drained
                    tc_fence_after();
                    if (trace) trace[(j * kTW + w) * 8 + 0] = clock64();
                    const uint32_t dwin = w * kTWinCols;
                    if (w == 0) mbar_wait(bar_wfull[0], wfull_phase);
                    issue_job_part<0>(J.ntaps, J.ncb, dwin, J.ntiles, act16_0 + w * reg16, tap16, J.cb0, J.lp, J.lo16,
                                      wp16[0], blk16, J.n, J.idesc, first, leader);
                    if (w == kTW - 1) tc_commit(bar_wfree[0], leader);
                    if (w == 0) mbar_wait(bar_wfull[1], wfull_phase);
                    issue_job_part<1>(J.ntaps, J.ncb, dwin, J.ntiles, act16_0 + w * reg16, tap16, J.cb0, J.lp, J.lo16,
                                      wp16[1], blk16, J.n, J.idesc, false, leader);
                    if (last) tc_commit(bar_mma + 8 * w, leader);
                    if (w == kTW - 1) tc_commit(bar_wfree[1], leader);
                    if (trace) trace[(j * kTW + w) * 8 + 1] = clock64();
                }
#endif
                epi_phase ^= 1;
                wfull_phase ^= 1;
            }
            tc_commit(bar_final, leader);
            mbar_wait(bar_final, 0);
        }
    } else if (warp == kLoadWarp && elect_one()) {
        // ================= weight loader =================
        uint32_t free_phase = 0;
        for (int j = 0; j < njobs; ++j) {
            const TcJob& J = c_tjobs[j];
            const unsigned char* src = P.w + J.w_goff;
            if (j > 0) mbar_wait(bar_wfree[0], free_phase);
            mbar_expect_tx(bar_wfull[0], J.w_part[0]);
            bulk_g2s(wbuf, src, J.w_part[0], bar_wfull[0]);
            if (j > 0) {
                mbar_wait(bar_wfree[1], free_phase);
                free_phase ^= 1;
            }
            mbar_expect_tx(bar_wfull[1], J.w_part[1]);
            bulk_g2s(wbuf + kWPart0, src + J.w_part[0], J.w_part[1], bar_wfull[1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace dbn
