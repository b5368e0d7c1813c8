// tcgen05 micro-benchmarks used to size the engine's schedule (not part of the product path):
//   1. burst latency: n K-blocks x 3 MMAs (the engine's hi/lo pattern) + commit + mbarrier wait,
//      for M in {128, 64} and several N - gives the fixed start-up/commit cost and the marginal
//      cost per MMA of a shape;
//   2. hand-off latency MMA -> commit -> epilogue warp (tcgen05.ld) -> mbarrier arrive -> issuer;
//   3. TMEM lane layout of an M=64 accumulator (which lanes hold which rows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_microbench tools/tc_microbench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int COLL>
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if (COLL == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else if (COLL == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd),
                     "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc16(uint32_t addr16, uint32_t lbo16) {
    return (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(lbo16 & 0x3FFF) << 16) |
           static_cast<uint64_t>(addr16 & 0x3FFF);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

constexpr int kRows = 514;                 // activation rows per channel group (as conv1d_2..4)
constexpr int kActArr = 6 * kRows * 16;    // one array (hi or lo), 6 channel groups
constexpr int kSmemA = 0;                  // hi at 0, lo at kActArr
constexpr int kSmemB = 2 * kActArr;        // weights: up to 9 K blocks x (hi, lo) x 2 x 128 rows x 16 B
constexpr int kSmemBBytes = 9 * 2 * 2 * 128 * 16;
constexpr int kSmemBar = kSmemB + kSmemBBytes;
constexpr int kSmemBytes = kSmemBar + 64;

struct Result {
    long long t_issue, t_done, t_roundtrip;
};


__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred;
}

// Fully unrolled issue of TILES x NKB K blocks x 3 MMAs (the engine's pattern), warp-uniform operands.
template <int M, int N, int NKB, int TILES, int COLL>
__device__ __forceinline__ void issue_burst(uint32_t sbase) {
    const uint32_t idesc = make_idesc(M, N);
    const uint64_t a_word = (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(kRows) << 16);
    const uint64_t b_word = (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(N) << 16);
    const uint32_t a0 = (sbase + kSmemA) >> 4, b0 = (sbase + kSmemB) >> 4;
    constexpr int kTileStride = N > 64 ? 128 : 64;
#pragma unroll
    for (int tile = 0; tile < TILES; ++tile) {
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
            const int t = kb / 3, cb = kb % 3;
            const uint32_t a16 = a0 + tile * 128 + t + 2 * cb * kRows;
            const uint64_t ad = a_word | (a16 & 0x3FFF), al = a_word | ((a16 + (kActArr >> 4)) & 0x3FFF);
            const uint32_t b16 = b0 + (kb % 2) * 2 * N;   // two alternating weight blocks (keeps N = 256 inside the buffer)
            const uint64_t bh = b_word | (b16 & 0x3FFF), bl = b_word | ((b16 + 2 * 2 * N) & 0x3FFF);
            if (COLL) {
                tc_mma<1>(tile * kTileStride, ad, bh, idesc, kb ? 1u : 0u);
                tc_mma<2>(tile * kTileStride, ad, bl, idesc, 1u);
            } else {
                tc_mma<0>(tile * kTileStride, ad, bh, idesc, kb ? 1u : 0u);
                tc_mma<0>(tile * kTileStride, ad, bl, idesc, 1u);
            }
            tc_mma<0>(tile * kTileStride, al, bh, idesc, 1u);
        }
    }
}

// MODE 0: burst (issue, commit, wait).  MODE 1: burst + epilogue-warp round trip.
// MODE 2: two-term variant (A_hi x W and A_lo x W only) for reference.
template <int M, int N, int NKB, int TILES, int COLL, int MODE>
__global__ void __launch_bounds__(160, 1) k_burst(int reps, Result* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_mma = sbase + kSmemBar, bar_epi = sbase + kSmemBar + 8, bar_done = sbase + kSmemBar + 16;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 32);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < kSmemBar / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_epi, 32);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive(bar_done);   // phase 0 of bar_done is complete from here on
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmemBar + 32), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        if (tmem != 0) __trap();
        if (elect_one()) {
            long long best_issue = 1ll << 60, best_done = 1ll << 60, best_rt = 1ll << 60;
            uint32_t ph = 0;
            for (int rep = 0; rep < reps; ++rep) {
                const long long t0 = clock64();
                issue_burst<M, N, NKB, TILES, COLL>(sbase);
                const long long t1 = clock64();
                tc_commit(bar_mma);
                if (MODE == 2) {
                    const long long ta = clock64();
                    mbar_wait(bar_done, 0);          // completed long ago: how long does the poll take behind a commit?
                    const long long tb = clock64();
                    mbar_wait(bar_mma, ph);
                    const long long t2 = clock64();
                    if (t1 - t0 < best_issue) best_issue = t1 - t0;
                    if (t2 - t0 < best_done) best_done = t2 - t0;
                    if (tb - ta < best_rt) best_rt = tb - ta;
                } else if (MODE == 0) {
                    mbar_wait(bar_mma, ph);
                    const long long t2 = clock64();
                    if (t1 - t0 < best_issue) best_issue = t1 - t0;
                    if (t2 - t0 < best_done) best_done = t2 - t0;
                } else {
                    mbar_wait(bar_epi, ph);
                    const long long t2 = clock64();
                    if (t2 - t0 < best_rt) best_rt = t2 - t0;
                }
                ph ^= 1;
                tc_fence_after();
            }
            out->t_issue = best_issue;
            out->t_done = best_done;
            out->t_roundtrip = best_rt;
        }
    } else if (warp == 1 && MODE == 1) {
        uint32_t ph = 0;
        for (int rep = 0; rep < reps; ++rep) {
            mbar_wait(bar_mma, ph);
            ph ^= 1;
            tc_fence_after();
            uint32_t r[8];
            tmem_ld8(tmem + (32u << 16), r);   // warp 1 -> lane quadrant 1
            tmem_wait_ld();
            if (r[0] == 0x12345678u) out->t_issue = 0;
            tc_fence_before();
            mbar_arrive(bar_epi);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// layout probe: A row r = (r + 1) in k = 0 (other k zero), B col c = 1 in k = 0 -> D[r][c] = r + 1
__global__ void __launch_bounds__(128, 1) k_layout(int M, int N, int dlane, float* out /*[128 lanes][8]*/) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_mma = sbase + kSmemBar;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 32);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSmemBar / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem + kSmemA);
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + kSmemB);
    for (int r = threadIdx.x; r < 128; r += blockDim.x) {
        A[r * 8] = __float2bfloat16(static_cast<float>(r + 1));   // channel group 0, row r, element 0
        B[r * 8] = __float2bfloat16(1.0f);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmemBar + 32), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    // clear the accumulator lanes first with an M=128 product of zeros so stale TMEM cannot confuse us
    if (threadIdx.x == 0) {
        const uint64_t zd = make_desc16((sbase + kSmemA + 64 * 1024) >> 4, kRows);   // zero area
        tc_mma<0>(tmem, zd, zd, make_idesc(128, N), 0u);
        const uint64_t ad = make_desc16((sbase + kSmemA) >> 4, kRows);
        const uint64_t bd = make_desc16((sbase + kSmemB) >> 4, N);
        tc_mma<0>(tmem + (static_cast<uint32_t>(dlane) << 16), ad, bd, make_idesc(M, N), 0u);
        tc_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    uint32_t r[8];
    tmem_ld8(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_wait_ld();
    for (int c = 0; c < 8; ++c) out[threadIdx.x * 8 + c] = __uint_as_float(r[c]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


// MODE 3 (separate kernel): how long does a cp.async.bulk of `bytes` (global -> shared, the engine's weight
// load) take when the tensor pipe is idle / while warp 0 streams a 108-MMA burst (M=128, N=48) / behind
// `chunks` separate copies on one barrier?  Every CTA of the grid does the same (L2 contention as in the engine).
constexpr int kSmemLand = kSmemBar + 64;          // landing zone of the copy
constexpr int kLandBytes = 27648;
constexpr int kSmemCopyBytes = kSmemLand + kLandBytes;
struct CopyResult { long long t_copy, t_burst, t_poll; };
template <int BURST>
__global__ void __launch_bounds__(160, 1) k_copy(const unsigned char* src, int bytes, int chunks, int reps, CopyResult* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_mma = sbase + kSmemBar, bar_w = sbase + kSmemBar + 8, bar_go = sbase + kSmemBar + 16;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 32);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < kSmemBar / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_go, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmemBar + 32), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    long long best_copy = 1ll << 60, best_burst = 1ll << 60, best_poll = 1ll << 60;
    uint32_t ph = 0;
    for (int rep = 0; rep < reps; ++rep) {
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                const long long t0 = clock64();
                if (BURST) {
                    issue_burst<128, 48, 9, 4, 1>(sbase);
                    const long long ta = clock64();
                    if (!mbar_try_wait(bar_go, 1)) __trap();   // a poll of an unrelated (passed) barrier phase right behind the burst, before the commit; the branch makes the clock read wait
                    const long long tb = clock64();
                    if (tb - ta < best_poll) best_poll = tb - ta;
                    tc_commit(bar_mma);
                    mbar_wait(bar_mma, ph);
                    tc_fence_after();
                }
                const long long t1 = clock64();
                if (t1 - t0 < best_burst) best_burst = t1 - t0;
            }
        } else if (warp == 1) {
            if (elect_one()) {
                const long long t0 = clock64();
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(bytes) : "memory");
                const int cb = bytes / chunks;
                for (int c = 0; c < chunks; ++c)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + kSmemLand + c * cb),
                                 "l"(src + (size_t)c * cb), "r"(cb), "r"(bar_w) : "memory");
                mbar_wait(bar_w, ph);
                const long long t1 = clock64();
                if (t1 - t0 < best_copy) best_copy = t1 - t0;
            }
        }
        ph ^= 1;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { out->t_burst = 0; }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (warp == 0 && elect_one()) { out->t_burst = best_burst; out->t_poll = best_poll; }
        if (warp == 1 && elect_one()) out->t_copy = best_copy;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


// MODE 4 (separate kernel): what does the MMA issuer (ONE elected lane) pay to learn that something is
// ready while 12 "epilogue" warps hammer shared memory with 16-byte stores?
//   (a) mbarrier.try_wait on a long-completed barrier, (b) mbarrier.test_wait, (c) a named barrier
//   (barrier.sync id, 64 by the single lane; warp 1 arrived long before with barrier.arrive id, 64).
struct PollResult { long long tw_idle, tw_busy, tt_idle, tt_busy, nb_idle, nb_busy, tw_max_busy, nb_max_busy; };
__global__ void __launch_bounds__(448, 1) k_poll(int reps, PollResult* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_done = sbase + 200000;
    volatile int* flag = reinterpret_cast<volatile int*>(smem + 200064);
    volatile int* mail = reinterpret_cast<volatile int*>(smem + 200096);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) {
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive(bar_done);
        *flag = 0;
    }
    __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            long long acc[6] = {0, 0, 0, 0, 0, 0}, mx[2] = {0, 0};
            const long long kstart = clock64();
            for (int busy = 0; busy < 2; ++busy) {
                if (busy) {                      // the store warps start 100 000 cycles into the kernel
                    while (clock64() - kstart < 103000) {}
                }
                for (int rep = 0; rep < reps; ++rep) {
                    long long t0 = clock64();
                    if (!mbar_try_wait(bar_done, 0)) __trap();   // the branch makes the clock read wait for the result
                    long long t1 = clock64();
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(bar_done), "r"(0) : "memory");
                    if (!ok) __trap();
                    long long t2 = clock64();
                    // named barrier: warp 1 arrives on 2, we wait >= 400 cycles, then sync on it
                    asm volatile("barrier.arrive 3, 64;" ::: "memory");     // go: warp 1 may arrive on barrier 2
                    long long d = clock64();
                    while (clock64() - d < 400) {}
                    long long t3 = clock64();
                    asm volatile("barrier.sync 2, 64;" ::: "memory");
                    if (*mail != busy * reps + rep + 1) __trap();   // warp 1 wrote it before arriving
                    long long t4 = clock64();
                    acc[busy] += t1 - t0;
                    acc[2 + busy] += t2 - t1;
                    acc[4 + busy] += t4 - t3;
                    if (busy && t1 - t0 > mx[0]) mx[0] = t1 - t0;
                    if (busy && t4 - t3 > mx[1]) mx[1] = t4 - t3;
                    if (ok == 12345) out->tw_idle = 1;
                }
            }
            *flag = 2;
            out->tw_idle = acc[0] / reps; out->tw_busy = acc[1] / reps;
            out->tt_idle = acc[2] / reps; out->tt_busy = acc[3] / reps;
            out->nb_idle = acc[4] / reps; out->nb_busy = acc[5] / reps;
            out->tw_max_busy = mx[0]; out->nb_max_busy = mx[1];
        }
    } else if (warp == 1) {
        for (int i = 0; i < 2 * reps; ++i) {
            asm volatile("barrier.sync 3, 64;" ::: "memory");
            if ((threadIdx.x & 31) == 0) *mail = i + 1;
            __syncwarp();
            asm volatile("barrier.arrive 2, 64;" ::: "memory");
        }
    } else {
        // "epilogue" warps: once the flag is 1, stream 16-byte stores (each warp its own 16 KB) until it is 2
        const uint32_t base = sbase + (warp - 2) * 16384 + (threadIdx.x & 31) * 16;
        const long long kstart = clock64();
        while (clock64() - kstart < 100000) {}
        uint32_t k = 0;
        while (*flag != 2) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + ((k + u) & 31) * 512), "r"(k) : "memory");
            k += 8;
        }
    }
}

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                                  \
        }                                                                                  \
    } while (0)


static Result* d_res;
template <int M, int N, int NKB, int TILES, int COLL, int MODE>
static void run_burst() {
    CK(cudaFuncSetAttribute(k_burst<M, N, NKB, TILES, COLL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    k_burst<M, N, NKB, TILES, COLL, MODE><<<1, 160, kSmemBytes>>>(20, d_res);
    CK(cudaDeviceSynchronize());
    Result r;
    CK(cudaMemcpy(&r, d_res, sizeof(r), cudaMemcpyDeviceToHost));
    const int mmas = 3 * NKB * TILES;
    if (MODE == 2)
        std::printf("poll-after-commit M %3d N %3d nkb %d tiles %d | %3d MMAs | issue %6lld done %6lld | try_wait on a completed barrier right after the commit: %lld cycles\n",
                    M, N, NKB, TILES, mmas, r.t_issue, r.t_done, r.t_roundtrip);
    else if (MODE == 0)
        std::printf("burst M %3d N %3d nkb %d tiles %d coll %d | %3d MMAs | issue %6lld done %6lld | %.1f cyc/MMA\n", M, N, NKB,
                    TILES, COLL, mmas, r.t_issue, r.t_done, mmas ? static_cast<double>(r.t_done) / mmas : 0.0);
    else
        std::printf("roundtrip M %3d N %3d nkb %d tiles %d | %lld cycles\n", M, N, NKB, TILES, r.t_roundtrip);
}
template <int M, int N>
static void run_shape() {
    run_burst<M, N, 0, 1, 0, 0>();
    run_burst<M, N, 1, 1, 1, 0>();
    run_burst<M, N, 3, 1, 1, 0>();
    run_burst<M, N, 9, 1, 1, 0>();
    run_burst<M, N, 9, 1, 0, 0>();
    if (M == 128 && N <= 128) run_burst<M, N, 9, 4, 1, 0>();
    if (M == 128 && N <= 128) run_burst<M, N, 9, 4, 0, 0>();
}


template <int BURST>
static void run_copy(const unsigned char* d_src, int bytes, int chunks, int grid) {
    static CopyResult* d_cr = nullptr;
    if (!d_cr) CK(cudaMalloc(&d_cr, sizeof(CopyResult)));
    CK(cudaMemset(d_cr, 0, sizeof(CopyResult)));
    CK(cudaFuncSetAttribute(k_copy<BURST>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCopyBytes));
    k_copy<BURST><<<grid, 160, kSmemCopyBytes>>>(d_src, bytes, chunks, 20, d_cr);
    CK(cudaDeviceSynchronize());
    CopyResult r;
    CK(cudaMemcpy(&r, d_cr, sizeof(r), cudaMemcpyDeviceToHost));
    std::printf("copy %5d B in %d chunk(s), grid %3d, %s | copy done after %6lld cycles (best of 20) | burst %6lld | poll behind the burst %lld\n",
                bytes, chunks, grid, BURST ? "with a 108-MMA burst" : "tensor pipe idle     ", r.t_copy, r.t_burst, r.t_poll);
}

int main() {
    CK(cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    CK(cudaMalloc(&d_res, sizeof(Result)));
    run_shape<128, 48>();
    run_shape<64, 48>();
    run_shape<128, 16>();
    run_shape<64, 16>();
    run_shape<128, 32>();
    run_shape<128, 96>();
    run_shape<64, 96>();
    run_shape<128, 128>();
    run_shape<128, 192>();
    run_shape<128, 256>();
    run_burst<128, 48, 0, 1, 1, 2>();
    run_burst<128, 48, 3, 1, 1, 2>();
    run_burst<128, 48, 9, 1, 1, 2>();
    run_burst<128, 48, 9, 4, 1, 2>();
    run_burst<64, 48, 9, 1, 1, 2>();
    run_burst<128, 48, 0, 1, 1, 1>();
    run_burst<128, 48, 1, 1, 1, 1>();
    run_burst<128, 48, 3, 1, 1, 1>();
    run_burst<128, 48, 9, 1, 1, 1>();


    {
        PollResult* d_pr;
        CK(cudaMalloc(&d_pr, sizeof(PollResult)));
        CK(cudaMemset(d_pr, 0, sizeof(PollResult)));
        CK(cudaFuncSetAttribute(k_poll, cudaFuncAttributeMaxDynamicSharedMemorySize, 200128));
        k_poll<<<1, 448, 200128>>>(50, d_pr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { std::printf("poll test: CUDA error %s\n", cudaGetErrorString(e)); return 0; }
        PollResult r;
        CK(cudaMemcpy(&r, d_pr, sizeof(r), cudaMemcpyDeviceToHost));
        std::printf("issuer-side cost of a readiness check (cycles, mean of 50): smem idle / 12 warps storing\n");
        std::printf("  mbarrier.try_wait (completed)   %5lld / %5lld (max %lld)\n", r.tw_idle, r.tw_busy, r.tw_max_busy);
        std::printf("  mbarrier.test_wait (completed)  %5lld / %5lld\n", r.tt_idle, r.tt_busy);
        std::printf("  named barrier.sync (single lane, partner arrived 400 cycles earlier) %5lld / %5lld (max %lld)\n", r.nb_idle, r.nb_busy, r.nb_max_busy);
    }
    {
        unsigned char* d_src;
        CK(cudaMalloc(&d_src, 1 << 20));
        CK(cudaMemset(d_src, 0, 1 << 20));
        for (int grid : {1, 148}) {
            for (int bytes : {3072, 9216, 27648}) {
                run_copy<0>(d_src, bytes, 1, grid);
                run_copy<1>(d_src, bytes, 1, grid);
            }
            run_copy<0>(d_src, 27648, 9, grid);
            run_copy<1>(d_src, 27648, 9, grid);
        }
    }
    float* d_out;
    CK(cudaMalloc(&d_out, 128 * 8 * 4));
    for (int cfg = 0; cfg < 4; ++cfg) {
        const int M = cfg == 0 ? 128 : 64, dlane = cfg == 2 ? 16 : (cfg == 3 ? 64 : 0);
        CK(cudaMemset(d_out, 0, 128 * 8 * 4));
        k_layout<<<1, 128, kSmemBytes>>>(M, 16, dlane, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            std::printf("# layout M=%d dlane=%d: CUDA error %s\n", M, dlane, cudaGetErrorString(e));
            return 0;
        }
        std::vector<float> h(128 * 8);
        CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
        std::printf("# layout M=%d D lane offset %d: lane -> row+1 (column 0)\n", M, dlane);
        for (int l = 0; l < 128; ++l) std::printf("%g%s", h[l * 8], (l % 32 == 31) ? "\n" : " ");
    }
    return 0;
}
