// "Solo" tcgen05 kernel: ONE window per CTA, TWO CTAs per SM (default tensor-core engine).
// Included by dbn_tc.cu (PTX wrappers, TcJob, split-bf16 helpers and the job builder live there).
//
// Why: with two windows marching in lock-step through one CTA (k_tc_forward) the network's back half
// (conv1d_5 .. conv1d_20: 37 % of the MACs) is a chain of small jobs - MMA burst, hand-off, epilogue,
// hand-off - in which the tensor pipe idles most of the time.  Here every window is an independent
// CTA (its own MMA issuer, weight loader and epilogue warps, 256 TMEM columns, half of the SM's shared
// memory), and the SM's warp / tensor-pipe schedulers interleave the two co-resident windows: while one
// sits in a latency-bound hand-off the other one's MMAs or epilogue run.  Same arithmetic as the pair
// kernel (three split-bf16 terms per K block, implicit im2col by descriptor offsets, epilogue in place).
//
// Shared memory per CTA (114 688 B; two CTAs + 2 x 1 KB system reserve fit the SM's 228 KB):
//   [0, 98 688)        activation region, same tensors / layouts as one window of the pair kernel;
//                      inception: X @0, T15 @12 672, T1214 @25 344, parity-split Y @33 792
//                      (4 arrays [24][18][8]); conv1d_17..20 tensors @0
//   [61 440, 98 304)   weight ring B: 12 K-block slots of 3 072 B, used from conv1d_5 on (that part of
//                      the region is dead once conv1d_4's MMAs have read their input)
//   [98 688, 114 048)  weight ring A: 5 K-block slots (conv1d_2 .. conv1d_4); slots 0-1 double as the
//                      staging area of the normalised input window during conv1d_1
//   [114 048, 114 688) mbarriers, TMEM pointer, reduction scratch
// Weights stream through the rings one K block (16 input channels of one tap: [hi | lo] x N rows x 32 B)
// at a time; a job issues K-block-outer / tile-inner so that a 3 KB slot serves all tiles of the layer.
// Per-channel parameters (bias, folded BatchNorm) are read from global memory (L2) into registers
// BEFORE a pass waits for its accumulators, so their latency is hidden.
//
// All jobs use one hand-off scheme: job j may issue once `need` epilogues have completed (conv1d_1's
// CUDA-core stage counts as epilogue 0); jobs with an epilogue own ring entry eseq & 3 of bar_mma /
// bar_epi.  Warp roles as in the pair kernel: warps 0-11 epilogue + CUDA-core stages, 12 loader, 13 MMA.
#pragma once

namespace dbn {

constexpr int kSRingA = kActBytes;               // 98 688
constexpr int kSRingASlots = 5;
constexpr int kSSlotBytes = 3072;                // one K block of an N = 48 job
constexpr int kSRingB = 61440;
constexpr int kSRingBSlots = 12;
constexpr int kSBar = kSRingA + kSRingASlots * kSSlotBytes;   // 114 048
constexpr int kSoloSmemBytes = kSBar + 640;
static_assert(kSRingB + kSRingBSlots * kSSlotBytes <= kActBytes, "ring B must stay inside the region");
static_assert(2 * (kSoloSmemBytes + 1024) <= 233472, "two CTAs per SM");
constexpr int kSoloTmemCols = 256;
// parity-split concat buffer of ONE window: 4 arrays (Ye_hi, Ye_lo, Yo_hi, Yo_lo) of [24][18][8]
constexpr int kSYRows = 18;
constexpr int kSYArray = 24 * kSYRows * 16;      // 6 912
constexpr int kSYOff = 33792;
static_assert(kSYOff + 4 * kSYArray <= kSRingB, "Y buffer must end below ring B");
// barrier block (byte offsets from kSBar)
constexpr int kSBarFullA = 0, kSBarFreeA = 40, kSBarFullB = 80, kSBarFreeB = 176;
constexpr int kSBarMma = 272, kSBarEpi = 304, kSBarFinal = 336, kSTmemPtr = 344, kSRed = 352;   // red: 24 x 8 B

__constant__ TcJob c_sjobs[kMaxJobs];

struct SoloParams {
    int njobs;
    int ring_b_job;           // first job that takes its weights from ring B
    int ring_a_blocks;        // K blocks streamed through ring A
    const TcJob* jobs;        // global copy of the job table (issuer prefetch)
    const unsigned char* w;   // packed bf16 weights, K-block major
    const float* prm;         // bias / folded BN blocks + conv1 parameters (global memory)
    int conv1_w, conv1_b, bn1_s, bn1_h;
    int n_classes;
    long long* trace;         // kDiag: [job][16] clock64 stamps of CTA 0
    int dbg_job;              // kDiag: stop after this job's epilogue and dump the ACT region
    unsigned char* dbg_out;
};

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------
// conv1d_1 (1 -> 48, k=3, stride 2, TF SAME: pad right) + ReLU + BatchNorm_1 -> T1 [6][514][8] hi/lo.
// `stage`: the 1024 normalised samples (+ one zero) as fp32 in ring-A slots 0-1.  Thread t: channel
// group t / 64, half group (4 channels) t & 1, positions (t % 64) / 2 + 32 k.  The two threads of a pair
// swap halves so that each writes one whole 16-byte row (even thread: hi array, odd thread: lo array).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void solo_conv1_stage(const SoloParams& P, const float* stage, uint32_t act, int tid) {
    const int cg = tid >> 6, half = tid & 1, p0 = (tid & 63) >> 1;
    const int c0 = cg * 8 + half * 4;
    const float4 w0 = ldg_f4(P.prm + P.conv1_w + c0), w1 = ldg_f4(P.prm + P.conv1_w + 48 + c0);
    const float4 w2 = ldg_f4(P.prm + P.conv1_w + 96 + c0), b = ldg_f4(P.prm + P.conv1_b + c0);
    const float4 sc = ldg_f4(P.prm + P.bn1_s + c0), sh = ldg_f4(P.prm + P.bn1_h + c0);
    const uint32_t row0 = act + (half ? 49344 : 0) + (cg * 514 + 1) * 16;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const int p = p0 + 32 * k;
        const float2 x01 = *reinterpret_cast<const float2*>(stage + 2 * p);
        const float x2 = stage[2 * p + 2];   // stage[1024] = 0: TF SAME pads on the right
        float v[4];
        v[0] = fmaf(sc.x, fmaxf(fmaf(w2.x, x2, fmaf(w1.x, x01.y, fmaf(w0.x, x01.x, b.x))), 0.f), sh.x);
        v[1] = fmaf(sc.y, fmaxf(fmaf(w2.y, x2, fmaf(w1.y, x01.y, fmaf(w0.y, x01.x, b.y))), 0.f), sh.y);
        v[2] = fmaf(sc.z, fmaxf(fmaf(w2.z, x2, fmaf(w1.z, x01.y, fmaf(w0.z, x01.x, b.z))), 0.f), sh.z);
        v[3] = fmaf(sc.w, fmaxf(fmaf(w2.w, x2, fmaf(w1.w, x01.y, fmaf(w0.w, x01.x, b.w))), 0.f), sh.w);
        uint32_t h[2], l[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            const float r0 = v[2 * i] - __uint_as_float(h[i] << 16);
            const float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
            l[i] = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);
        }
        // even thread keeps hi and receives the partner's hi; odd thread keeps lo and receives lo
        const uint32_t s0 = __shfl_xor_sync(0xffffffffu, half ? h[0] : l[0], 1);
        const uint32_t s1 = __shfl_xor_sync(0xffffffffu, half ? h[1] : l[1], 1);
        const uint4 row = half ? make_uint4(s0, s1, l[0], l[1]) : make_uint4(h[0], h[1], s0, s1);
        st_shared_v4(row0 + p * 16, row);
    }
    if (tid < 24) {   // zero halo rows 0 and 513 of every channel-group, hi and lo
        const int g = tid % 6, which = tid / 6;
        st_shared_v4(act + (which & 1 ? 49344 : 0) + (g * 514 + (which & 2 ? 513 : 0)) * 16, make_uint4(0, 0, 0, 0));
    }
}

// ---------------------------------------------------------------------------------------------
// epilogue passes (copy of the pair kernel's epilogue_tiles, specialised for one window: parameters
// from global memory, JOINT_PAIR = M=64 accumulator of the single window in lanes 0-15 of each quadrant)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void solo_zero_padding_rows(const EpiArgs& A, uint32_t act, int tid) {
    if (A.kind == EPI_PARITY) {
        if (A.zero_y && tid < 192) {   // rows 16, 17 of every array / channel-group
            const int cg = tid % 24, rest = tid / 24, arr = rest >> 1, rrow = rest & 1;
            st_shared_v4(act + kSYOff + arr * kSYArray + (cg * kSYRows + 16 + rrow) * 16, make_uint4(0, 0, 0, 0));
        }
    } else if (tid < 32 && (tid & 7) < A.out_ncg) {   // halo rows 0 and out_L + 1 of the output tensor
        const int cg = tid & 7, which = tid >> 3;
        st_shared_v4(act + A.out_off + (which & 1 ? A.out_lo_delta : 0) +
                         (cg * A.out_lp + (which & 2 ? A.out_L + 1 : 0)) * 16,
                     make_uint4(0, 0, 0, 0));
    }
}

template <int NC, bool POOL, bool BN, bool PARITY, int JOINT>
__device__ __forceinline__ void solo_epilogue_tiles(const EpiArgs& A, uint32_t act, const float* __restrict__ gprm,
                                                    uint32_t tmem_win, int tid, uint32_t bar, uint32_t parity,
                                                    long long* tr) {
    const int lane = tid & 31;
    const int q = ((tid >> 5) + kEpiWarp0) & 3, h = tid >> 7;
    const bool active = h * NC < A.n;
    const int row = JOINT ? q * 16 + (lane & 15) : q * 32 + lane;
    const bool lane_ok = JOINT == JOINT_NONE || lane < 16;
    const int ntiles = JOINT ? 1 : A.ntiles, L = A.L;
    const int cg0 = A.out_cg_base + (h * NC) / 8;
    const uint32_t out_base = act + A.out_off;
    const int out_lp = A.out_lp, out_lo = A.out_lo_delta;
    const uint32_t taddr0 = tmem_win + h * NC + (static_cast<uint32_t>(q * 32) << 16);
    constexpr int PN = POOL ? NC / 2 : NC;
    const int odd = lane & 1;
    const int pc0 = POOL ? odd * (NC / 2) : 0;
    float bias[PN], sc[BN ? PN : 1], sh[BN ? PN : 1];
    {
        const float* bias_a = gprm + A.bias_off + h * NC + pc0;
#pragma unroll
        for (int g = 0; g < PN / 4; ++g) {
            const float4 b = ldg_f4(bias_a + 4 * g);
            bias[4 * g] = b.x; bias[4 * g + 1] = b.y; bias[4 * g + 2] = b.z; bias[4 * g + 3] = b.w;
        }
        if (BN) {
            const float* bn_a = gprm + A.bn_off + h * NC + pc0;
#pragma unroll
            for (int g = 0; g < PN / 4; ++g) {
                const float4 a = ldg_f4(bn_a + 4 * g), b = ldg_f4(bn_a + 48 + 4 * g);
                sc[4 * g] = a.x; sc[4 * g + 1] = a.y; sc[4 * g + 2] = a.z; sc[4 * g + 3] = a.w;
                sh[4 * g] = b.x; sh[4 * g + 1] = b.y; sh[4 * g + 2] = b.z; sh[4 * g + 3] = b.w;
            }
        }
    }
    wait_accumulators(bar, parity, tr);
    if (!active) {
        solo_zero_padding_rows(A, act, tid);
        return;
    }
    uint32_t r[NC];
    tmem_load_cols<NC>(taddr0, r);
    solo_zero_padding_rows(A, act, tid);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int p = tile * 128 + row;
        tmem_wait_ld();
        float acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = __uint_as_float(r[c]);
        if (tile + 1 < ntiles) tmem_load_cols<NC>(taddr0 + (tile + 1) * kTmemTileCols, r);
        const int qpos = POOL ? p >> 1 : p;
        const bool valid = lane_ok && p < L;
        const float es = (A.edge15 && (p == 0 || p == L - 1)) ? 1.5f : 1.0f;
        if (POOL) {
            static_assert(!POOL || NC == 16, "pooling epilogue is written for 16 columns per warp");
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
                if (PARITY) {
                    const uint32_t o = act + kSYOff + (qpos & 1) * (2 * kSYArray) +
                                       ((cg0 + odd) * kSYRows + (qpos >> 1)) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + kSYArray, lo);
                } else {
                    const uint32_t o = out_base + ((cg0 + odd) * out_lp + qpos + 1) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + out_lo, lo);
                }
            }
        } else {
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
                if (valid) {
                    const uint32_t o = out_base + ((cg0 + g) * out_lp + qpos + 1) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + out_lo, lo);
                }
            }
        }
    }
}

// Head: conv1d_20 accumulators (row k = position k, 16 columns, TMEM lane quadrant 0) -> ReLU -> global
// average pool over the 8 positions -> softmax (network_architecture.py:89-91).  One warp.
__device__ __forceinline__ void solo_epilogue_head(int bias_off, const float* __restrict__ gprm, uint32_t tmem_win,
                                                   uint32_t scratch, int lane, int n_classes, float* probs) {
    uint32_t r[16];
    tmem_ld8(tmem_win, r);
    tmem_ld8(tmem_win + 8, r + 8);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 b = ldg_f4(gprm + bias_off + 4 * g);
        float4 v;
        v.x = fmaxf(__uint_as_float(r[4 * g + 0]) + b.x, 0.f);
        v.y = fmaxf(__uint_as_float(r[4 * g + 1]) + b.y, 0.f);
        v.z = fmaxf(__uint_as_float(r[4 * g + 2]) + b.z, 0.f);
        v.w = fmaxf(__uint_as_float(r[4 * g + 3]) + b.w, 0.f);
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(scratch + (lane * 16 + 4 * g) * 4), "f"(v.x),
                     "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
    __syncwarp();
    const int c = lane & 15;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float t;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(scratch + (k * 16 + c) * 4));
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
    if (probs && lane < 16 && c < n_classes) probs[c] = e / den;
}

__device__ __forceinline__ void solo_run_epilogue(const SoloParams& P, const TcJob& J, uint32_t act, uint32_t tmem_win,
                                                  int tid, uint32_t bar, uint32_t parity, float* probs, long long* tr) {
    const EpiArgs A = load_epi_args(J);
    const float* gprm = P.prm;
    if (A.joint == JOINT_PAIR) {
        if (A.kind == EPI_PARITY)
            solo_epilogue_tiles<16, true, true, true, JOINT_PAIR>(A, act, gprm, tmem_win, tid, bar, parity, tr);
        else   // EPI_N48 / EPI_N16
            solo_epilogue_tiles<16, false, false, false, JOINT_PAIR>(A, act, gprm, tmem_win, tid, bar, parity, tr);
    } else if (A.joint == JOINT_STACK && A.kind != EPI_HEAD) {
        if (A.kind == EPI_N48_BN)
            solo_epilogue_tiles<16, false, true, false, JOINT_STACK>(A, act, gprm, tmem_win, tid, bar, parity, tr);
        else if (A.kind == EPI_N48_POOL_BN)
            solo_epilogue_tiles<16, true, true, false, JOINT_STACK>(A, act, gprm, tmem_win, tid, bar, parity, tr);
        else   // EPI_N48
            solo_epilogue_tiles<16, false, false, false, JOINT_STACK>(A, act, gprm, tmem_win, tid, bar, parity, tr);
    } else if (A.kind == EPI_N48 || A.kind == EPI_N16) {
        solo_epilogue_tiles<16, false, false, false, JOINT_NONE>(A, act, gprm, tmem_win, tid, bar, parity, tr);
    } else if (A.kind == EPI_N48_POOL_BN) {
        solo_epilogue_tiles<16, true, true, false, JOINT_NONE>(A, act, gprm, tmem_win, tid, bar, parity, tr);
    } else {   // EPI_HEAD (conv1d_20, M=128 idesc: rows 0..7 in TMEM lane quadrant 0)
        wait_accumulators(bar, parity, tr);
        if ((tid >> 5) < 4 && (((tid >> 5) + kEpiWarp0) & 3) == 0)
            solo_epilogue_head(A.bias_off, gprm, tmem_win, act + 8192, tid & 31, P.n_classes, probs);
    }
}

// ---------------------------------------------------------------------------------------------
// MMA issue: K-block outer, tile inner.  Every K block is one ring slot: wait for it, issue the three
// split-bf16 terms for every tile of the job, hand the slot back.
// ---------------------------------------------------------------------------------------------
struct SoloRing {
    uint32_t base16, full0, free0;   // slot 0 address (16-byte units), barrier arrays
    uint32_t nslots, slot, parity;
};
__device__ __forceinline__ void ring_advance(SoloRing& r) {
    if (++r.slot == r.nslots) {
        r.slot = 0;
        r.parity ^= 1u;
    }
}

template <int NTAPS, int NCB>
__device__ __forceinline__ void solo_issue_job(SoloRing& ring, uint32_t d0, uint32_t ntiles, uint32_t a16,
                                               const uint32_t (&tap16)[3], uint32_t cb_first, uint32_t lp,
                                               uint32_t lo16, uint32_t n, uint32_t idesc, bool zero_first) {
    const uint64_t a_hi_word = (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(lp & 0x3FFF) << 16);
    const uint64_t b_hi_word = (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(n & 0x3FFF) << 16);
    const uint32_t a_job = a16 + 2 * cb_first * lp;
#pragma unroll
    for (int kb = 0; kb < NTAPS * NCB; ++kb) {
        const int t = kb / NCB, cb = kb % NCB;
        mbar_wait(ring.full0 + 8 * ring.slot, ring.parity);
        const uint32_t b16 = ring.base16 + ring.slot * (kSSlotBytes / 16);
        const uint64_t bd_hi = b_hi_word | (b16 & 0x3FFF);
        const uint64_t bd_lo = b_hi_word | ((b16 + 2 * n) & 0x3FFF);
        const uint32_t a_kb = a_job + tap16[t] + 2 * cb * lp;
        const uint32_t acc = (kb == 0 && zero_first) ? 0u : 1u;
        for (uint32_t tile = 0; tile < ntiles; ++tile) {
            const uint32_t d = d0 + tile * kTmemTileCols;
            const uint32_t a = a_kb + tile * 128;
            const uint64_t ad = a_hi_word | (a & 0x3FFF);
            tc_mma<1>(d, ad, bd_hi, idesc, acc, 1u);
            tc_mma<2>(d, ad, bd_lo, idesc, 1u, 1u);
            tc_mma<0>(d, a_hi_word | ((a + lo16) & 0x3FFF), bd_hi, idesc, 1u, 1u);
        }
        tc_commit(ring.free0 + 8 * ring.slot);
        ring_advance(ring);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool kCallMode, bool kDiag>
__global__ void __launch_bounds__(kTcThreads, 2)
    k_tc_solo(SoloParams P, const float* __restrict__ x, const double* __restrict__ xd,
              const int16_t* __restrict__ samples, const int64_t* __restrict__ offsets, int n_reads, int side,
              int n_windows, float* __restrict__ probs) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    long long* const trace = (kDiag && blockIdx.x == 0) ? P.trace : nullptr;
    const int dbg_job = kDiag ? P.dbg_job : -1;
    const uint32_t bar0 = sbase + kSBar;
    const uint32_t bar_mma = bar0 + kSBarMma, bar_epi = bar0 + kSBarEpi, bar_final = bar0 + kSBarFinal;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kSBar + kSTmemPtr);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) >> 5, 0);
    const bool is_epi = warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps;
    const int win = blockIdx.x;   // grid == n_windows

    // input samples of the thread (predict mode), issued before the set-up barrier
    float xv[3] = {0.f, 0.f, 0.f};
    if (is_epi && !kCallMode) {
        const int etid = static_cast<int>(threadIdx.x) - kEpiWarp0 * 32;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = etid + k * kEpiThreads;
            if (i < kInputSize)
                xv[k] = x ? __ldg(x + static_cast<size_t>(win) * kInputSize + i)
                          : static_cast<float>(__ldg(xd + static_cast<size_t>(win) * kInputSize + i));
        }
    }
    if (threadIdx.x == kLoadWarp * 32) {
        for (int i = 0; i < kSRingASlots; ++i) {
            mbar_init(bar0 + kSBarFullA + 8 * i, 1);
            mbar_init(bar0 + kSBarFreeA + 8 * i, 1);
        }
        for (int i = 0; i < kSRingBSlots; ++i) {
            mbar_init(bar0 + kSBarFullB + 8 * i, 1);
            mbar_init(bar0 + kSBarFreeB + 8 * i, 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_mma + 8 * i, 1);
            mbar_init(bar_epi + 8 * i, kEpiArrivals);
        }
        mbar_init(bar_final, 1);
        // ring A is entered at slot 2 (slots 0-1 hold the staged input first): give slots 0-1 a phantom
        // first use so that every slot's phase index equals (ring position) / 5
        for (int i = 0; i < 2; ++i) {
            mbar_arrive(bar0 + kSBarFullA + 8 * i);
            mbar_arrive(bar0 + kSBarFreeA + 8 * i);
        }
        // epilogue 0 (conv1d_1's CUDA-core stage) has no MMAs: complete its phase of bar_mma[0] here, so
        // that entry e & 3 of both rings is in phase e >> 2 for every e
        mbar_arrive(bar_mma);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(sbase + kSBar + kSTmemPtr, kSoloTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int njobs = (dbg_job >= 0 && dbg_job < P.njobs) ? dbg_job + 1 : P.njobs;

    if (is_epi) {
        // ================= epilogue / CUDA-core warps =================
        const int tid = static_cast<int>(threadIdx.x) - kEpiWarp0 * 32;
        const int ewarp = tid >> 5;
        if (trace && tid == 0) trace[31 * 16 + 0] = clock64();
        float* stage = reinterpret_cast<float*>(smem + kSRingA);
        if (kCallMode) {
            const int step = win / n_reads, read = win % n_reads;
            const int64_t off = offsets[read];
            WindowInput in{};
            in.region = samples + off;
            in.g = window_geometry(static_cast<int>(offsets[read + 1] - off), step, side);
            long long s1 = 0, s2 = 0;   // exact integer sums over the slice
            for (int i = tid; i < in.g.n; i += kEpiThreads) {
                const long long v = in.region[in.g.a + i];
                s1 += v;
                s2 += v * v;
            }
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            long long* red = reinterpret_cast<long long*>(smem + kSBar + kSRed);
            if ((tid & 31) == 0) { red[ewarp] = s1; red[12 + ewarp] = s2; }
            epi_bar_sync();
            s1 = 0; s2 = 0;
            for (int i = 0; i < kEpiWarps; ++i) { s1 += red[i]; s2 += red[12 + i]; }
            in.mean = 0.0; in.stdev = 0.0;
            if (in.g.n > 0) zscore_params(s1, s2, in.g.n, &in.mean, &in.stdev);
            fetch_window_inputs(in, tid, xv);
        }
        stage[tid] = xv[0];
        stage[tid + kEpiThreads] = xv[1];
        if (tid + 2 * kEpiThreads < kInputSize) stage[tid + 2 * kEpiThreads] = xv[2];
        if (tid == 0) stage[kInputSize] = 0.f;
        epi_bar_sync();
        solo_conv1_stage(P, stage, sbase, tid);
        fence_proxy_async();
        epi_arrive(bar_epi);   // epilogue 0
        if (trace && tid == 0) trace[31 * 16 + 1] = clock64();
        float* pout = probs + static_cast<size_t>(win) * P.n_classes;
        for (int j = 0; j < njobs; ++j) {
            const TcJob& J = c_sjobs[j];
            if (!J.last) continue;
            const int e = J.eseq;
            long long* tr = (trace && tid == 0) ? trace + j * 16 : nullptr;
            solo_run_epilogue(P, J, sbase, tmem_base + J.tcol, tid, bar_mma + 8 * (e & 3), (e >> 2) & 1, pout, tr);
            fence_proxy_async();
            tc_fence_before();
            epi_arrive(bar_epi + 8 * (e & 3));
            if (tr) tr[3] = clock64();
        }
        if (dbg_job >= 0) {   // debug: dump the ACT region after the last processed job
            mbar_wait(bar_final, 0);
            epi_bar_sync();
            if (blockIdx.x == 0)
                for (int i = tid; i < kActBytes / 16; i += kEpiThreads)
                    reinterpret_cast<uint4*>(P.dbg_out)[i] = reinterpret_cast<const uint4*>(smem)[i];
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one elected lane) =================
        if (elect_one()) {
            SoloRing ringA{(sbase + kSRingA) >> 4, bar0 + kSBarFullA, bar0 + kSBarFreeA, kSRingASlots, 2u, 0u};
            SoloRing ringB{(sbase + kSRingB) >> 4, bar0 + kSBarFullB, bar0 + kSBarFreeB, kSRingBSlots, 0u, 0u};
            const uint32_t act16 = sbase >> 4;
            int seen = 0;   // epilogues known to be complete
            IssueArgs nxt = load_issue_args(P.jobs);
            for (int j = 0; j < njobs; ++j) {
                const IssueArgs J = nxt;
                if (j + 1 < njobs) nxt = load_issue_args(P.jobs + j + 1);
                if (trace) trace[j * 16 + 11] = clock64();
                for (const int need = J.need; seen < need; ++seen) mbar_wait(bar_epi + 8 * (seen & 3), (seen >> 2) & 1);
                tc_fence_after();
                if (trace) trace[j * 16 + 0] = clock64();
                const uint32_t tap16[3] = {J.tap16[0], J.tap16[1], J.tap16[2]};
                const uint32_t d0 = tmem_base + J.tcol;
                const bool first = J.first != 0;
                if (j < P.ring_b_job)   // conv1d_2 .. conv1d_4
                    solo_issue_job<3, 3>(ringA, d0, J.ntiles, act16, tap16, J.cb0, J.lp, J.lo16, J.n, J.idesc, first);
                else if (J.ntaps == 3 && J.ncb == 3)
                    solo_issue_job<3, 3>(ringB, d0, J.ntiles, act16, tap16, J.cb0, J.lp, J.lo16, J.n, J.idesc, first);
                else if (J.ntaps == 1)
                    solo_issue_job<1, 3>(ringB, d0, J.ntiles, act16, tap16, J.cb0, J.lp, J.lo16, J.n, J.idesc, first);
                else
                    solo_issue_job<3, 1>(ringB, d0, J.ntiles, act16, tap16, J.cb0, J.lp, J.lo16, J.n, J.idesc, first);
                if (J.last) tc_commit(bar_mma + 8 * (J.eseq & 3));
                if (trace) trace[j * 16 + 1] = clock64();
            }
            tc_commit(bar_final);
            mbar_wait(bar_final, 0);
        }
    } else if (warp == kLoadWarp && elect_one()) {
        // ================= weight loader (one elected lane) =================
        // ring A starts at slot 2: slots 0-1 hold the staged input window until conv1d_1 is done
        uint32_t ca = 0, cb = 0;   // K blocks loaded into ring A / ring B so far
        for (int j = 0; j < njobs; ++j) {
            const TcJob& J = c_sjobs[j];
            const unsigned char* src = P.w + J.w_goff;
            const uint32_t bytes = J.w_part[0];
            const int nkb = J.w_part[1];
            if (j < P.ring_b_job) {
                for (int kb = 0; kb < nkb; ++kb, ++ca) {
                    const uint32_t slot = (ca + 2) % kSRingASlots, use = (ca + 2) / kSRingASlots;
                    if (ca == 3) mbar_wait(bar_epi, 0);   // first use of slot 0: conv1d_1 has consumed the staged input
                    if (use > 0) mbar_wait(bar0 + kSBarFreeA + 8 * slot, (use - 1) & 1);
                    mbar_expect_tx(bar0 + kSBarFullA + 8 * slot, bytes);
                    bulk_g2s(sbase + kSRingA + slot * kSSlotBytes, src + static_cast<size_t>(kb) * bytes, bytes,
                             bar0 + kSBarFullA + 8 * slot);
                }
            } else {
                if (cb == 0 && ca > 0) {   // ring B lies in the region conv1d_4's MMAs read: wait for the last of them
                    const uint32_t last = ca - 1 + 2;
                    mbar_wait(bar0 + kSBarFreeA + 8 * (last % kSRingASlots), (last / kSRingASlots) & 1);
                }
                for (int kb = 0; kb < nkb; ++kb, ++cb) {
                    const uint32_t slot = cb % kSRingBSlots, use = cb / kSRingBSlots;
                    if (use > 0) mbar_wait(bar0 + kSBarFreeB + 8 * slot, (use - 1) & 1);
                    mbar_expect_tx(bar0 + kSBarFullB + 8 * slot, bytes);
                    bulk_g2s(sbase + kSRingB + slot * kSSlotBytes, src + static_cast<size_t>(kb) * bytes, bytes,
                             bar0 + kSBarFullB + 8 * slot);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, kSoloTmemCols);
}

}  // namespace dbn
