// tcgen05 / TMEM tensor-core engine for the Deepbinner network (reference
// network_architecture.py:18-95; executed by model.predict at classify.py:361).
//
// One CTA processes TWO windows at a time and is persistent over the window pairs of a launch (grid =
// min(pairs, SMs)).  All activations stay in shared memory as split-bf16 pairs
// (hi = bf16_rn(x), lo = upper half of the exact remainder x - hi); every Conv1D from conv1d_2 to
// conv1d_20 is a sequence of tcgen05.mma (kind::f16, M=128 positions x N=Cout x K=16) instructions
// accumulating in fp32 in tensor memory.  A k=3 'same' convolution is three accumulating MMAs per
// 16-channel block over the SAME shared-memory tile: activations are stored [C/8][L+2][8] (one
// 16-byte row per position and channel-group, zero halo rows), which is the canonical K-major
// no-swizzle UMMA layout, so a tap shift is a +16 B start-address offset in the descriptor - no
// im2col is ever materialised.  Precision: three MMA terms per K block (A_hi*W_hi, A_hi*W_lo with
// A_hi reused from the tensor core's collector buffer, A_lo*W_hi) give ~16 mantissa bits on both
// operands, which SURVEY Appendix C shows is needed for the 1e-3 probability bar (single bf16
// fails, fp16 overflows; dropping either cross term of any single layer fails too,
// profiles/r02_precision_probe_*.txt).
//
// A whole layer's output for one window (up to 4 tiles x 48 fp32 columns) lives in TMEM, so the
// epilogue (bias, ReLU, [MaxPool2], [BatchNorm affine], hi/lo split) can overwrite the layer's
// input in place once its MMAs have completed.  The network is a table of 21 MMA jobs in constant
// memory (the issuers read a compact copy, IssueRec, from global memory), in two phases:
//   * conv1d_2 .. conv1d_9 (L = 512 .. 128): one pass per window, M=128 tiles; the two windows of
//     the CTA alternate so that the tensor pipe works on one while the epilogue warps drain the other;
//   * "joint" jobs from the inception block on (L <= 64): ONE MMA burst and ONE epilogue pass serve
//     both windows.  Inception (conv1d_10 .. 16): an M=64 MMA set per window into the same
//     accumulator columns, window 0 in TMEM lanes 0-15 and window 1 in lanes 16-31 of every lane
//     quadrant, so all 32 lanes of every epilogue warp carry rows.  conv1d_12 + conv1d_14 share one
//     job (N = 32); the AveragePooling1D in front of conv1d_10 is folded into its weights (k=3, W/3,
//     end rows rescaled by 1.5 in the epilogue).  conv1d_17 .. 20: both windows stacked in one tensor
//     (row = 18 w + position), conv1d_17 as four K-slices.  Joint jobs rotate over kJointSlots accumulator
//     slots and are ordered so that independent branches sit between dependent ones; each carries
//     the number of joint epilogues that must be complete before it may issue (`need`), which lets
//     the tensor pipe run ahead of the epilogue warps.
//
// Warp roles (480 threads): warps 0-11 = epilogue (TMEM lane quadrant = warp % 4, 16 accumulator
// columns per warp) and the CUDA-core stages (z-score + conv1d_1, softmax head); warp 12 = weight
// loader (cp.async.bulk + mbarrier; each single-window job's weights come in two K-block parts so the next
// job's first part streams in while the second is still in use; joint jobs get whole-job slots); warps 13
// and 14 = the two MMA issuers (one elected lane each; warp 13 also allocates / frees tensor memory).  An
// issuer is a single thread (one dependent instruction every ~4 cycles, slower when the epilogue warps of
// its scheduler are busy) and needs ~1 000 cycles between two jobs (queue drain before the commits,
// barrier polls, descriptor set-up), so everything it does between two MMAs matters: issue records are
// fetched one job ahead, tcgen05 descriptors are built from pre-shifted fields, all diagnostics live in a
// separate kernel instantiation - and there are two of them, so that the turn-around of one overlaps the
// burst of the other (conv1d_5..9: one window each; joint jobs: alternating).  Every role loops over the
// CTA's window pairs on its own; there is no CTA-wide barrier between two pairs (see the kernel).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <map>
#include <mutex>
#include <utility>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/deepbinner_b200.h"
#include "dbn_engine.h"
#include "dbn_weights.h"

namespace dbn {

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
// Warps 0-11: epilogue / CUDA-core stages; warp 12: weight loader; warps 13 and 14: the two MMA issuers.
constexpr int kEpiWarps = 12;
constexpr int kEpiWarp0 = 0;
constexpr int kLoadWarp = 12;
constexpr int kMmaWarp = 13;                    // issuer 0 (also allocates / frees tensor memory)
constexpr int kMmaWarpB = 14;                   // issuer 1
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTcThreads = (3 + kEpiWarps) * 32;
constexpr int kActBytes = 98688;                // 2 x [6][514][8] bf16
constexpr int kWPart0 = 15360;                  // weight part 0: up to 5 K blocks x (hi + lo) x 1536 B
constexpr int kWPart1 = 12288;                  // weight part 1: up to 4 K blocks
constexpr int kWbufBytes = kWPart0 + kWPart1;   // == hi+lo bf16 of a 48->48 k=3 layer
constexpr int kPrmFloats = 1664;                // per-job bias / folded BN, resident in smem
constexpr int kSmemAct0 = 0;
constexpr int kSmemAct1 = kActBytes;
constexpr int kSmemWbuf = 2 * kActBytes;
constexpr int kSmemPrm = kSmemWbuf + kWbufBytes;
constexpr int kSmemBar = kSmemPrm + kPrmFloats * 4;   // mbarriers, tmem pointer, reduction scratch
constexpr int kJointSlots = 7;                  // accumulator slots (64 TMEM columns each) the joint jobs rotate over
constexpr int kJointRing = 16;                  // joint MMA-done / epilogue-done mbarriers: one per joint epilogue, never reused
constexpr int kTcSmemBytes = kSmemBar + 512 + 2 * 8 * kJointRing;   // [0,72) mbarriers, 96 TMEM pointer, [128,384) joint weight barriers, 384 bar_x, [512,768) joint MMA / epilogue barriers
static_assert(kTcSmemBytes <= 232448, "shared memory budget");
constexpr int kTmemCols = 512;
constexpr int kTmemWindowCols = 256;
constexpr int kTmemTileCols = 64;
// Parity-split concat buffer (input of conv1d_17) for BOTH windows of the CTA, in window 0's ACT
// region: 4 arrays (Ye_hi, Ye_lo, Yo_hi, Yo_lo) of [24 cg][36 rows][8]; window w occupies rows
// 18w .. 18w+15, rows 18w+16/17 are zero.  From conv1d_17 on the two windows are STACKED in one
// M=128 tile (row = 18 w + position): the L=16 tail then needs one MMA pass and one epilogue pass
// per layer for both windows.  Pitch 18 keeps max-pool pairs even-aligned.
constexpr int kStackPitch = 18;
constexpr int kYRows = 2 * kStackPitch;    // 36
constexpr int kYArray = 24 * kYRows * 16;  // 13824
constexpr int kYOff = kActBytes - 4 * kYArray;   // 43392
static_assert(kYOff >= 34176, "Y buffer overlaps the inception scratch tensors");
static_assert(kYOff + 2 * kWbufBytes <= kActBytes, "joint weight slots must fit behind window 1's inception tensors");
constexpr int kMaxJobs = 32;

enum EpiKind {
    EPI_N48 = 0,          // bias + ReLU
    EPI_N48_POOL_BN = 1,  // bias + ReLU + MaxPool2 + BN
    EPI_N48_BN = 2,       // bias + ReLU + BN
    EPI_N16 = 3,          // 16-channel bottleneck, bias + ReLU
    EPI_PARITY = 4,       // bias + ReLU + MaxPool2 + BN5 -> even/odd split concat buffer
    EPI_HEAD = 5          // bias + ReLU + global average pool + softmax
};

// Joint jobs (conv1d_10 onwards) serve BOTH windows of the CTA in one MMA burst and one epilogue pass:
//   JOINT_PAIR  (inception block, L = 64): one M=64 MMA set per window into the SAME accumulator
//               columns, window 0 in TMEM lanes 0-15 and window 1 in lanes 16-31 of every lane
//               quadrant (a D lane offset of 16 is honoured for M=64; measured with
//               tools/tc_microbench.cu), so each epilogue warp drains 16 rows of either window;
//   JOINT_STACK (conv1d_17 onwards): both windows stacked in one tensor (row = 18 w + position) in
//               window 0's region, one M=64 MMA set (conv1d_20: M=128, head epilogue).
// Joint jobs rotate over kJointSlots accumulator slots and carry `need` = number of joint epilogues that
// must have completed before their MMAs may be issued (input produced / slot drained), so the
// tensor pipe runs ahead of the epilogue warps where the graph allows it.
enum JointKind { JOINT_NONE = 0, JOINT_PAIR = 1, JOINT_STACK = 2 };

struct alignas(128) TcJob {
    int n;            // MMA N (Cout padded to a multiple of 16)
    int idesc;        // tcgen05 instruction descriptor (the host builder keeps M here until finalize_jobs())
    int ntiles;       // M tiles of 128 positions
    int L;            // valid positions
    int lp;           // rows per channel-group of the input tensor
    int ntaps;
    int tap16[3];     // offset (from the window's ACT base) of row 0 of each tap, hi array, in 16-byte units
    int lo16;         // distance from the hi array to the lo array of the input, in 16-byte units
                      // (the host builder keeps byte offsets in tap16 / lo16 until finalize_jobs())
    int ncb;          // 16-channel K blocks per tap handled by this job
    int cb0;          // first K block (conv1d_17 is split in 4 jobs)
    int w_goff;       // byte offset of this job's packed weights in global memory ([part 0 | part 1])
    int tcol;         // joint jobs: first accumulator column (slot * 64)
    int w_part[2];    // bytes of each part; a part is [hi blocks | lo blocks] of its K-block range
    int first, last;  // first: zero the accumulators; last: bit 0 = run the epilogue, bit 1 (joint jobs) = an issuer waits on exactly this epilogue's barrier
    // epilogue
    int kind, bias_off, bn_off;  // float offsets into the smem parameter block
    int out_off, out_lp, out_lo_delta, out_cg_base, out_ncg, out_L;
    int edge15;       // scale the accumulators of the first / last position by 1.5 (folded average pool)
    int zero_y;
    int joint;        // JointKind
    int need;         // joint jobs: joint epilogues that must be complete before the MMAs are issued
    int eseq;         // joint jobs with an epilogue: index of that epilogue in the joint sequence (else -1)
};

__constant__ TcJob c_jobs[kMaxJobs];
static_assert(sizeof(TcJob) == 128, "one job descriptor = two 64-byte constant-cache lines");

// Touch both constant-cache lines of a job descriptor so that the accesses made one job later hit
// (the volatile asm consumes the values, which pins the two LDCs at this point of the program).
#ifndef DBN_TC_PREFETCH_JOB
#define DBN_TC_PREFETCH_JOB 1
#endif
__device__ __forceinline__ void prefetch_job(int j) {
    if (DBN_TC_PREFETCH_JOB && j < kMaxJobs) asm volatile("" ::"r"(c_jobs[j].n), "r"(c_jobs[j].eseq));
}

struct TcParams {
    int njobs;
    const uint4* issue;       // global-memory table of IssueRec (what the MMA issuer needs of every job, 64 B each)
    const unsigned char* w;   // packed bf16 weights
    const float* prm;         // global copy of the smem parameter block + conv1 parameters
    int prm_floats;
    int conv1_w, conv1_b, bn1_s, bn1_h;   // float offsets into prm
    int n_classes;
    int dbg_job;              // >= 0: stop after this job's epilogue and dump ACT of both windows
    unsigned char* dbg_out;
    long long* trace;         // optional timeline of CTA 0: [job][window][4] clock64 stamps
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Epilogue -> issuer hand-off: every epilogue thread fences its shared-memory writes towards the
// async proxy and arrives itself.  (One elected arrival per warp behind a __syncwarp was measured
// equally fast, and compute-sanitizer racecheck cannot follow that chain - it then reports the
// epilogue's stores of consecutive passes as write-write hazards - so the plain form is kept.)
constexpr uint32_t kEpiArrivals = 384;
__device__ __forceinline__ void epi_arrive(uint32_t bar) { mbar_arrive(bar); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
#ifndef DBN_TC_WAIT_HINT_NS
#define DBN_TC_WAIT_HINT_NS 2000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if DBN_TC_WAIT_HINT_NS > 0
    // suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint
    // expires) instead of re-issuing the poll, leaving issue slots to the warps that have work
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(DBN_TC_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (reported as a CUDA error), never a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef DBN_TC_DEBUG_WAIT   // debugging aid (-DDBN_TC_DEBUG_WAIT): report a wait that does not end and carry on (results are then wrong)
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 12)) {
            printf("mbarrier wait timed out: block %d thread %d barrier offset %d parity %u clock %lld\n", blockIdx.x, threadIdx.x,
                   static_cast<int>(bar & 0x3FFFF) - 1024 - kSmemBar, parity, clock64());
            return;
        }
#else
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
#endif
}
// Non-blocking probe of a phase.  The MMA issuer uses it to look at the barrier of its NEXT wait
// early: a poll queues behind the epilogue warps' shared-memory traffic (hundreds of cycles when
// they are storing), so the probe is issued before the current MMAs and its result is only read
// when the wait is due - a phase that had completed by then costs nothing.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// `leader`: the MMA warp runs its control flow converged on all 32 lanes (so that ptxas keeps the
// descriptor arithmetic on the uniform datapath) and only the tcgen05 instructions themselves are
// predicated on the one elected lane.
__device__ __forceinline__ void tc_commit(uint32_t bar, uint32_t leader = 1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
        "r"(leader)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate.  Issued by the one elected lane of
// the MMA warp from a warp-uniform region, so ptxas keeps the descriptor arithmetic on the uniform
// datapath (UTCHMMA takes uniform registers; no R2UR waterfall per instruction).
// COLL: 0 = default, 1 = collector::a::fill (keep A in the collector buffer), 2 = collector::a::lastuse
// (take A from the collector buffer instead of re-reading shared memory).
template <int COLL>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate, uint32_t leader) {
    if (COLL == 1)
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.ne.b32 q, %5, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
            : "memory");
    else if (COLL == 2)
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.ne.b32 q, %5, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.ne.b32 q, %5, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
            : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                   "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() {   // the 384 epilogue threads only
    asm volatile("bar.sync 1, 384;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): core matrices
// of 8 rows x 16 bytes; LBO = byte distance between the two 8-element K chunks of one MMA,
// SBO = byte distance between consecutive 8-row groups (always 128 here); version = 1 (Blackwell).
// `addr16` and `lbo16` are in 16-byte units.
__device__ __forceinline__ uint64_t make_desc16(uint32_t addr16, uint32_t lbo16) {
    return (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(lbo16 & 0x3FFF) << 16) |
           static_cast<uint64_t>(addr16 & 0x3FFF);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 A/B, both K-major.
__host__ __device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// split-bf16 helpers: x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi) ----------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// Packed fp32 pairs (Blackwell FFMA2: one instruction, two IEEE fp32 FMAs - bit-identical to two FFMAs).  The epilogues
// and conv1d_1 are instruction-issue bound, and most of their arithmetic is independent per channel.
#ifndef DBN_TC_F32X2
#define DBN_TC_F32X2 1
#endif
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// a * b + c on 8 values, pairwise packed
__device__ __forceinline__ void fma8(const float* a, const float* b, const float* c, float* out) {
#if DBN_TC_F32X2
#pragma unroll
    for (int i = 0; i < 4; ++i)
        unpack2(fma2(pack2(a[2 * i], a[2 * i + 1]), pack2(b[2 * i], b[2 * i + 1]), pack2(c[2 * i], c[2 * i + 1])),
                out[2 * i], out[2 * i + 1]);
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = fmaf(a[i], b[i], c[i]);
#endif
}
// a * s + c with one scalar s
__device__ __forceinline__ void fma8s(const float* a, float s, const float* c, float* out) {
#if DBN_TC_F32X2
    const uint64_t s2 = pack2(s, s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        unpack2(fma2(pack2(a[2 * i], a[2 * i + 1]), s2, pack2(c[2 * i], c[2 * i + 1])), out[2 * i], out[2 * i + 1]);
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = fmaf(a[i], s, c[i]);
#endif
}
// hi is rounded to nearest (one F2FP per pair on the XU pipe); lo is the TRUNCATED upper half of
// the exact remainder x - hi (one PRMT per pair).  The remainder's sign is symmetric around zero,
// so truncating it toward zero is unbiased with respect to x; |x - hi - lo| <= 2^-16 |x|.
__device__ __forceinline__ void split8(const float (&v)[8], uint4* hi, uint4* lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#if DBN_TC_F32X2
        float r0, r1;   // both remainders with one FFMA2: v - hi = hi * (-1) + v (exact, as the two subtractions)
        unpack2(fma2(pack2(__uint_as_float(h[i] << 16), __uint_as_float(h[i] & 0xFFFF0000u)), pack2(-1.f, -1.f),
                     pack2(v[2 * i], v[2 * i + 1])), r0, r1);
#else
        const float r0 = v[2 * i] - __uint_as_float(h[i] << 16);
        const float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
#endif
        l[i] = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);
    }
    *hi = make_uint4(h[0], h[1], h[2], h[3]);
    *lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void unpack8(uint4 hi, uint4 lo, float (&v)[8]) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
    const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
        v[2 * i + 1] = __uint_as_float(h[i] & 0xFFFF0000u) + __uint_as_float(l[i] & 0xFFFF0000u);
    }
}

// ---------------------------------------------------------------------------------------------
// CUDA-core stages (epilogue warps, 384 threads)
// ---------------------------------------------------------------------------------------------

// Input of one window: either normalised values (predict seam) or an int16 scan region that is
// z-scored on the fly (fused call_batch; classify.py:342-357, trim_signal.py:61-69).
struct WindowInput {
    const float* x;          // predict mode, float32 windows (nullptr otherwise)
    const double* xd;        // predict mode, float64 windows (cast to float32 as Keras does)
    const int16_t* region;   // call mode
    WindowGeom g;
    double mean, stdev;
    __device__ __forceinline__ float at(int i) const {
        if (x) return i < kInputSize ? __ldg(x + i) : 0.f;
        if (xd) return i < kInputSize ? static_cast<float>(__ldg(xd + i)) : 0.f;
        const int k = i - g.dst;
        if (i >= kInputSize || k < 0 || k >= g.n) return 0.f;
        const double d = static_cast<double>(region[g.a + k]) - mean;
        return static_cast<float>(stdev > 0.0 ? d / stdev : d);
    }
};

// conv1d_1 (1 -> 48, k=3, stride 2, pad right) + ReLU + BatchNorm_1 -> T1 [6][514][8] hi/lo.
// The normalised window is first staged as fp32 at the END of the window's own ACT region
// (bytes [94592, 98688), overwritten later by T1 - every thread pulls its inputs into registers
// before anyone writes).  Thread t: channel-group t / 64, positions (t % 64) + 64 k, k = 0..7, so
// the 48 per-channel parameters of its group are loaded once.
constexpr int kStageOff = kActBytes - 4096;
// Per-thread conv1d_1 / BatchNorm_1 parameters of the thread's channel group (the same for both
// windows of the CTA: fetched once, early, so the global-memory latency overlaps the kernel set-up).
struct Conv1Params {
    float w0[8], w1[8], w2[8], b[8], sc[8], sh[8];
};
__device__ __forceinline__ void load_conv1_params(const TcParams& P, int tid, Conv1Params& c) {
    const int cg = tid >> 6;
    const float4* w4 = reinterpret_cast<const float4*>(P.prm + P.conv1_w);   // [3][48]
    const float4* b4 = reinterpret_cast<const float4*>(P.prm + P.conv1_b);
    const float4* s4 = reinterpret_cast<const float4*>(P.prm + P.bn1_s);
    const float4* h4 = reinterpret_cast<const float4*>(P.prm + P.bn1_h);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float4 a0 = __ldg(w4 + cg * 2 + q), a1 = __ldg(w4 + 12 + cg * 2 + q);
        const float4 a2 = __ldg(w4 + 24 + cg * 2 + q), bb = __ldg(b4 + cg * 2 + q);
        const float4 ss = __ldg(s4 + cg * 2 + q), hh = __ldg(h4 + cg * 2 + q);
        c.w0[4 * q] = a0.x; c.w0[4 * q + 1] = a0.y; c.w0[4 * q + 2] = a0.z; c.w0[4 * q + 3] = a0.w;
        c.w1[4 * q] = a1.x; c.w1[4 * q + 1] = a1.y; c.w1[4 * q + 2] = a1.z; c.w1[4 * q + 3] = a1.w;
        c.w2[4 * q] = a2.x; c.w2[4 * q + 1] = a2.y; c.w2[4 * q + 2] = a2.z; c.w2[4 * q + 3] = a2.w;
        c.b[4 * q] = bb.x; c.b[4 * q + 1] = bb.y; c.b[4 * q + 2] = bb.z; c.b[4 * q + 3] = bb.w;
        c.sc[4 * q] = ss.x; c.sc[4 * q + 1] = ss.y; c.sc[4 * q + 2] = ss.z; c.sc[4 * q + 3] = ss.w;
        c.sh[4 * q] = hh.x; c.sh[4 * q + 1] = hh.y; c.sh[4 * q + 2] = hh.z; c.sh[4 * q + 3] = hh.w;
    }
}
// The thread's three input samples of a window (i = tid + 384 k), fetched ahead of use.
__device__ __forceinline__ void fetch_window_inputs(const WindowInput& in, int tid, float (&xv)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) xv[k] = in.at(tid + k * kEpiThreads);
}
__device__ __forceinline__ void conv1_stage(const Conv1Params& c, float x0, float x1, float x2, uint32_t act,
                                            unsigned char* act_ptr, int tid) {
    float* stage = reinterpret_cast<float*>(act_ptr + kStageOff);
    stage[tid] = x0;
    stage[tid + kEpiThreads] = x1;
    if (tid + 2 * kEpiThreads < kInputSize) stage[tid + 2 * kEpiThreads] = x2;
    epi_bar_sync();
    const int cg = tid >> 6, p0 = tid & 63;
    float xs[8][3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int p = p0 + 64 * k;
        const float2 a = *reinterpret_cast<const float2*>(stage + 2 * p);
        xs[k][0] = a.x;
        xs[k][1] = a.y;
        xs[k][2] = p < 511 ? stage[2 * p + 2] : 0.f;   // x[1024] = 0: TF SAME pads on the right
    }
    epi_bar_sync();   // all inputs are in registers; the staging area may now be overwritten
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int p = p0 + 64 * k;
        float v[8];
        fma8s(c.w0, xs[k][0], c.b, v);
        fma8s(c.w1, xs[k][1], v, v);
        fma8s(c.w2, xs[k][2], v, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
        fma8(c.sc, v, c.sh, v);
        uint4 hi, lo;
        split8(v, &hi, &lo);
        const uint32_t a0 = act + (cg * 514 + p + 1) * 16;
        st_shared_v4(a0, hi);
        st_shared_v4(a0 + 49344, lo);
    }
    if (tid < 24) {   // zero halo rows 0 and 513 of every channel-group, hi and lo
        const int g = tid % 6, which = tid / 6;
        const uint32_t a0 = act + (which & 1 ? 49344 : 0) + (g * 514 + (which & 2 ? 513 : 0)) * 16;
        st_shared_v4(a0, make_uint4(0, 0, 0, 0));
    }
}

// Epilogue of one job for one window: TMEM accumulators -> bias, ReLU, [pool], [BN], split -> smem.
// Warp w handles TMEM lane quadrant (w & 3) and column group (w >> 2): NC = 16 columns per warp
// (three groups for N = 48; for N = 16 only group 0 has work).  The
// per-channel parameters of the warp's columns are hoisted into registers once per job and the
// TMEM load of the next tile is issued before the current tile is processed.
template <int NC>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, uint32_t (&r)[NC]) {
#pragma unroll
    for (int g = 0; g < NC / 8; ++g) tmem_ld8(taddr + g * 8, &r[g * 8]);
}

// Everything a pass needs from the job descriptor, fetched into registers BEFORE waiting for the
// accumulators (the wait has slack; the constant/shared loads would otherwise sit on the critical path).
struct EpiArgs {
    int kind, n, ntiles, L, joint, zero_y, edge15;
    int bias_off, bn_off;
    int out_off, out_lp, out_lo_delta, out_cg_base, out_ncg, out_L;
};
__device__ __forceinline__ EpiArgs load_epi_args(const TcJob& J) {
    EpiArgs a;
    a.kind = J.kind; a.n = J.n; a.ntiles = J.ntiles; a.L = J.L; a.joint = J.joint; a.zero_y = J.zero_y;
    a.edge15 = J.edge15;
    a.bias_off = J.bias_off; a.bn_off = J.bn_off;
    a.out_off = J.out_off; a.out_lp = J.out_lp; a.out_lo_delta = J.out_lo_delta;
    a.out_cg_base = J.out_cg_base; a.out_ncg = J.out_ncg; a.out_L = J.out_L;
    return a;
}

// Wait until the MMAs of this (job, window) have completed.  Every epilogue thread waits every pass
// (a thread may never run more than one mbarrier phase ahead).
// (Polling with one lane per warp + __syncwarp instead of all 32 was measured 2.5 % slower.)
__device__ __forceinline__ void wait_accumulators(uint32_t bar, uint32_t parity, long long* tr) {
    mbar_wait(bar, parity);
    tc_fence_after();
    if (tr) tr[2] = clock64();
}

// Zero padding rows written by each pass (after the wait: the output may alias the layer's input).
// `act` = ACT region the output tensor lives in, `w` = window (selects the window's rows of Y).
__device__ __forceinline__ void zero_padding_rows(const EpiArgs& A, uint32_t act, uint32_t act0, int w, int tid) {
    if (A.kind == EPI_PARITY) {
        if (A.zero_y && tid < 192) {   // rows 16, 17 of this window in every array / channel-group
            const int cg = tid % 24, rest = tid / 24, arr = rest >> 1, rrow = rest & 1;
            st_shared_v4(act0 + kYOff + arr * kYArray + (cg * kYRows + w * kStackPitch + 16 + rrow) * 16,
                         make_uint4(0, 0, 0, 0));
        }
    } else if (tid < 32 && (tid & 7) < A.out_ncg) {   // halo rows 0 and out_L + 1 of the output tensor
        const int cg = tid & 7, which = tid >> 3;
        const uint32_t a0 = act + A.out_off + (which & 1 ? A.out_lo_delta : 0) +
                            (cg * A.out_lp + (which & 2 ? A.out_L + 1 : 0)) * 16;
        st_shared_v4(a0, make_uint4(0, 0, 0, 0));
    }
}

// JOINT: 0 = one window per pass (M=128 tiles: TMEM lane = position within the tile),
//        1 = JOINT_PAIR  (M=64 per window: lanes 0-15 of a quadrant = 16 rows of window 0, lanes 16-31 =
//            the same rows of window 1),
//        2 = JOINT_STACK (M=64, stacked tensor in window 0's region: lanes 0-15 = rows, 16-31 idle).
template <int NC, bool POOL, bool BN, bool PARITY, int JOINT>
__device__ __forceinline__ void epilogue_tiles(const EpiArgs& A, uint32_t act, uint32_t act0, int w, uint32_t prm,
                                               uint32_t tmem_win, int tid, uint32_t bar, uint32_t parity,
                                               long long* tr) {
    // tid is epilogue-relative (0..383); hardware warp = tid / 32 + kEpiWarp0 decides the TMEM lane
    // quadrant it may access (warp % 4); each run of 4 consecutive warps covers all quadrants
    const int lane = tid & 31;
    const int q = ((tid >> 5) + kEpiWarp0) & 3, h = tid >> 7;
    const bool active = h * NC < A.n;
    constexpr bool stack = JOINT == JOINT_STACK;
    const int row = JOINT ? q * 16 + (lane & 15) : q * 32 + lane;
    const bool lane_ok = JOINT != JOINT_STACK || lane < 16;
    if (JOINT == JOINT_PAIR) {   // this lane's window
        w = lane >> 4;
        act = act0 + w * kActBytes;
    }
    const int ntiles = JOINT ? 1 : A.ntiles, L = A.L;
    const int cg0 = A.out_cg_base + (h * NC) / 8;
    const uint32_t out_base = act + A.out_off;
    const int out_lp = A.out_lp, out_lo = A.out_lo_delta;
    const uint32_t taddr0 = tmem_win + h * NC + (static_cast<uint32_t>(q * 32) << 16);
    // Pooling passes: the two lanes of a max-pool pair share the work after the maximum - the even lane
    // finishes columns 0-7 of the warp's 16, the odd lane columns 8-15 - so each lane needs only its
    // half of the per-channel parameters (PN of them, starting at column `pc0`).
    constexpr int PN = POOL ? NC / 2 : NC;
    const int odd = lane & 1;
    const int pc0 = POOL ? odd * (NC / 2) : 0;
    float bias[PN], sc[BN ? PN : 1], sh[BN ? PN : 1];
    {   // unconditional (inactive warps read neighbouring parameters they never use) so that the
        // arrays stay in registers
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
    wait_accumulators(bar, parity, tr);
    if (!active) {
        if (JOINT == JOINT_PAIR) {
            zero_padding_rows(A, act0, act0, 0, tid);
            zero_padding_rows(A, act0 + kActBytes, act0, 1, tid);
        } else {
            zero_padding_rows(A, act, act0, w, tid);
        }
        return;
    }
    uint32_t r[NC];
    tmem_load_cols<NC>(taddr0, r);
    if (JOINT == JOINT_PAIR) {
        zero_padding_rows(A, act0, act0, 0, tid);
        zero_padding_rows(A, act0 + kActBytes, act0, 1, tid);
    } else {
        zero_padding_rows(A, act, act0, w, tid);
    }
    if (tr) tr[4] = clock64();
    for (int tile = 0; tile < ntiles; ++tile) {
        const int p = tile * 128 + row;
        tmem_wait_ld();
        if (tr && tile == 0) tr[5] = clock64();
        float acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = __uint_as_float(r[c]);
        if (tile + 1 < ntiles) tmem_load_cols<NC>(taddr0 + (tile + 1) * kTmemTileCols, r);
        const int qpos = POOL ? p >> 1 : p;
        // stacked tail: rows 18 w + i, i < 16 are positions; the other rows below L are separators
        const bool valid = lane_ok && (stack ? (p < L && (p % kStackPitch) < 16) : (p < L));
        // folded average pool (conv1d_10): TF divides the two in-range taps of the end positions by 2
        const float es = (A.edge15 && (p == 0 || p == L - 1)) ? 1.5f : 1.0f;
        if (POOL) {
            static_assert(!POOL || NC == 16, "pooling epilogue is written for 16 columns per warp");
            // relu(max(a, b) + bias) == max(relu(a + bias), relu(b + bias)): take the maximum of the raw
            // accumulators of positions 2i and 2i+1 first.  Each lane keeps the 8 columns it will
            // finish and sends the other 8 to its partner (8 shuffles instead of 16).
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
            fma8s(v, 1.0f, bias, v);   // v + bias (exact)
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            if (BN) fma8(sc, v, sh, v);
            uint4 hi, lo;
            split8(v, &hi, &lo);
            if (valid) {   // both lanes of a pair write: channel group cg0 (even lane) / cg0 + 1 (odd lane)
                if (PARITY) {   // even/odd pooled positions in separate arrays (input of conv1d_17)
                    const uint32_t o = act0 + kYOff + (qpos & 1) * (2 * kYArray) +
                                       ((cg0 + odd) * kYRows + w * kStackPitch + (qpos >> 1)) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + kYArray, lo);
                } else {
                    const uint32_t o = out_base + ((cg0 + odd) * out_lp + qpos + 1) * 16;
                    st_shared_v4(o, hi);
                    st_shared_v4(o + out_lo, lo);
                }
            }
        } else {
            const bool writer = stack ? (lane_ok && p < L) : valid;
#pragma unroll
            for (int g = 0; g < NC / 8; ++g) {
                float v[8];
                fma8s(&acc[g * 8], es, &bias[g * 8], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                if (BN) fma8(&sc[g * 8], v, &sh[g * 8], v);
                uint4 hi, lo;
                split8(v, &hi, &lo);
                if (stack && !valid) {   // separator rows of a stacked tensor are zero padding
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

// Head for the two stacked windows: conv1d_20 accumulators (row 9 w + k = position k of window w, 16
// columns) -> ReLU -> global average pool -> softmax (network_architecture.py:89-91).  One warp
// (TMEM lane quadrant 0); `scratch` = 32 x 16 floats of free shared memory.  Lanes 0-15 finish
// window 0 (class = lane), lanes 16-31 window 1.
__device__ void epilogue_head(int bias_off, uint32_t prm, uint32_t tmem_win, uint32_t scratch, int lane,
                              int n_classes, float* probs0, float* probs1) {
    uint32_t r[16];
    tmem_ld8(tmem_win, r);
    tmem_ld8(tmem_win + 8, r + 8);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 b = ld_shared_f4(prm + (bias_off + 4 * g) * 4);
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
    const int w = lane >> 4, c = lane & 15;
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
    float* out = w ? probs1 : probs0;
    if (out && c < n_classes) out[c] = e / den;
}

// One epilogue pass: fetch the descriptor, dispatch on the kind, (inside) wait for the accumulators,
// drain them.
__device__ void run_epilogue(const TcParams& P, const TcJob& J, uint32_t act, uint32_t act0, int w, uint32_t prm,
                             uint32_t tmem_win, int tid, uint32_t bar, uint32_t parity, float* probs0,
                             float* probs1, long long* tr) {
    const EpiArgs A = load_epi_args(J);
    if (A.joint == JOINT_PAIR) {
        if (A.kind == EPI_PARITY)
            epilogue_tiles<16, true, true, true, JOINT_PAIR>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
        else   // EPI_N48 / EPI_N16
            epilogue_tiles<16, false, false, false, JOINT_PAIR>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
    } else if (A.joint == JOINT_STACK && A.kind != EPI_HEAD) {
        if (A.kind == EPI_N48_BN)
            epilogue_tiles<16, false, true, false, JOINT_STACK>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
        else if (A.kind == EPI_N48_POOL_BN)
            epilogue_tiles<16, true, true, false, JOINT_STACK>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
        else   // EPI_N48
            epilogue_tiles<16, false, false, false, JOINT_STACK>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
    } else if (A.kind == EPI_N48 || A.kind == EPI_N16) {
        epilogue_tiles<16, false, false, false, JOINT_NONE>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
    } else if (A.kind == EPI_N48_POOL_BN) {
        epilogue_tiles<16, true, true, false, JOINT_NONE>(A, act, act0, w, prm, tmem_win, tid, bar, parity, tr);
    } else {   // EPI_HEAD (conv1d_20: M=128 stacked tile)
        wait_accumulators(bar, parity, tr);
        // rows 0..16 live in TMEM lane quadrant 0: one of the first four epilogue warps owns it;
        // scratch: 2 KB of window 0's ACT region behind the (tiny) conv1d_19 output
        if ((tid >> 5) < 4 && (((tid >> 5) + kEpiWarp0) & 3) == 0)
            epilogue_head(A.bias_off, prm, tmem_win, act0 + 8192, tid & 31, P.n_classes, probs0, probs1);
    }
}

// One weight part of a job for one window: for every tile, the K blocks [KB0, KB1) of the job
// (block kb = tap kb / NCB, channel block kb % NCB), each as three MMAs
//   A_hi x W_hi (A kept in the collector), A_hi x W_lo (A reused from the collector), A_lo x W_hi.
// Fully unrolled; all operands warp-uniform.
template <int NCB, int KB0, int KB1>
__device__ __forceinline__ void issue_part(uint32_t dwin, uint32_t ntiles, uint32_t a16, const uint32_t (&tap16)[3],
                                           uint32_t cb_first, uint32_t lp, uint32_t lo16, uint32_t b16,
                                           uint32_t blk16, uint32_t n, uint32_t idesc, bool zero_first, uint32_t leader) {
    const uint64_t a_hi_word = (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(lp & 0x3FFF) << 16);
    const uint64_t b_hi_word = (static_cast<uint64_t>(0x4008u) << 32) | (static_cast<uint64_t>(n & 0x3FFF) << 16);
    constexpr int NKB = KB1 - KB0;
    for (uint32_t tile = 0; tile < ntiles; ++tile) {
        const uint32_t d = dwin + tile * kTmemTileCols;
        const uint32_t a_tile = a16 + tile * 128 + 2 * cb_first * lp;
        uint32_t acc = zero_first ? 0u : 1u;
#pragma unroll
        for (int kb = KB0; kb < KB1; ++kb) {
            const int t = kb / NCB, cb = kb % NCB;
            const uint32_t a = a_tile + tap16[t] + 2 * cb * lp;
            const uint64_t ad = a_hi_word | (a & 0x3FFF);
            const uint64_t bd_hi = b_hi_word | ((b16 + (kb - KB0) * blk16) & 0x3FFF);
            const uint64_t bd_lo = b_hi_word | ((b16 + (NKB + kb - KB0) * blk16) & 0x3FFF);
            tc_mma<1>(d, ad, bd_hi, idesc, acc, leader);
            tc_mma<2>(d, ad, bd_lo, idesc, 1u, leader);
            tc_mma<0>(d, a_hi_word | ((a + lo16) & 0x3FFF), bd_hi, idesc, 1u, leader);
            acc = 1u;
        }
    }
}

template <int PART>
__device__ __forceinline__ void issue_job_part(int ntaps, int ncb, uint32_t dwin, uint32_t ntiles, uint32_t a16,
                                               const uint32_t (&tap16)[3], uint32_t cb_first, uint32_t lp,
                                               uint32_t lo16, uint32_t b16, uint32_t blk16, uint32_t n,
                                               uint32_t idesc, bool zero_first, uint32_t leader) {
    if (ntaps == 3 && ncb == 3) {          // 9 K blocks: 5 + 4
        if (PART == 0) issue_part<3, 0, 5>(dwin, ntiles, a16, tap16, cb_first, lp, lo16, b16, blk16, n, idesc, zero_first, leader);
        else issue_part<3, 5, 9>(dwin, ntiles, a16, tap16, cb_first, lp, lo16, b16, blk16, n, idesc, zero_first, leader);
    } else if (ntaps == 1) {               // 1x1 conv over 48 channels: 3 K blocks: 2 + 1
        if (PART == 0) issue_part<3, 0, 2>(dwin, ntiles, a16, tap16, cb_first, lp, lo16, b16, blk16, n, idesc, zero_first, leader);
        else issue_part<3, 2, 3>(dwin, ntiles, a16, tap16, cb_first, lp, lo16, b16, blk16, n, idesc, zero_first, leader);
    } else {                               // k=3 conv over 16 channels: 3 K blocks (one per tap): 2 + 1
        if (PART == 0) issue_part<1, 0, 2>(dwin, ntiles, a16, tap16, cb_first, lp, lo16, b16, blk16, n, idesc, zero_first, leader);
        else issue_part<1, 2, 3>(dwin, ntiles, a16, tap16, cb_first, lp, lo16, b16, blk16, n, idesc, zero_first, leader);
    }
}

// What the MMA issuer needs from a job descriptor, held in registers one job ahead.  The issuer reads it
// from a separate global table of 64-byte records in which EVERY word is used: a loaded-but-unused
// component leaves a register that ptxas recycles as scratch at once, and the write-after-write
// hazard on it makes the issuer sit out the full latency of the loads it has just issued - measured:
// ~500 cycles at the top of every job with the 128-byte TcJob read by six 16-byte loads
// (profiles/r02_issuer_waw.txt).
struct alignas(16) IssueRec {
    uint32_t n, idesc, ntiles, lp;                 // word 0
    uint32_t tap16[3], lo16;                       // word 1
    uint32_t shape, cb0, tcol, flags;              // word 2: shape = ncb | ntaps << 8 | jk << 16 (jk = index among the joint jobs);
                                                   // flags = first | last << 1 | joint << 2 | owner << 4 | both << 5
    uint32_t need, eseq, blk16, part1_16;          // word 3: blk16 = one K=16 block of B, part1_16 = offset of weight part 1 behind part 0 (joint jobs), 16-byte units
};
static_assert(sizeof(IssueRec) == 64, "IssueRec is read as four 16-byte words");
struct IssueArgs {
    uint32_t n, idesc, ntiles, lp, lo16, cb0, tcol, tap16[3], blk16, part1_16;
    int ntaps, ncb, first, last, joint, need, eseq;
    int jk;       // joint jobs: index among the joint jobs (weight slot (jk + 1) % 3)
    int owner;    // joint jobs: the issuer (0 / 1) that issues the job
    int both;     // single-window jobs: issuer 0 issues both windows (conv1d_2..4); otherwise issuer w issues window w
};
// The four raw words of a record.  They are fetched with volatile loads at the top of the PREVIOUS job's
// iteration and not touched (not even unpacked) until the job starts, so the whole latency is hidden.
// (Plain loads of the __constant__ table were sunk by the compiler to their first use, which put a
// constant-cache miss of several hundred cycles between every two jobs.)
struct IssueRaw {
    uint4 w0, w1, w2, w3;
};
__device__ __forceinline__ uint4 ldg_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ IssueRaw load_issue_raw(const uint4* q) {
    IssueRaw r;
    r.w0 = ldg_volatile_v4(q); r.w1 = ldg_volatile_v4(q + 1); r.w2 = ldg_volatile_v4(q + 2); r.w3 = ldg_volatile_v4(q + 3);
    return r;
}
__device__ __forceinline__ IssueArgs unpack_issue(const IssueRaw& r) {
    IssueArgs a;
    a.n = r.w0.x; a.idesc = r.w0.y; a.ntiles = r.w0.z; a.lp = r.w0.w;
    a.tap16[0] = r.w1.x; a.tap16[1] = r.w1.y; a.tap16[2] = r.w1.z; a.lo16 = r.w1.w;
    a.ncb = static_cast<int>(r.w2.x & 0xFFu); a.ntaps = static_cast<int>((r.w2.x >> 8) & 0xFFu); a.jk = static_cast<int>(r.w2.x >> 16);
    a.cb0 = r.w2.y; a.tcol = r.w2.z;
    a.first = static_cast<int>(r.w2.w & 1u); a.last = static_cast<int>((r.w2.w >> 1) & 1u); a.joint = static_cast<int>((r.w2.w >> 2) & 3u);
    a.owner = static_cast<int>((r.w2.w >> 4) & 1u); a.both = static_cast<int>((r.w2.w >> 5) & 1u);
    a.need = static_cast<int>(r.w3.x); a.eseq = static_cast<int>(r.w3.y); a.blk16 = r.w3.z; a.part1_16 = r.w3.w;
    return a;
}
// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// kDiag: diagnostics build of the kernel (timeline stamps, early stop + ACT dump); the production
// instantiations carry none of that code.
template <bool kCallMode, bool kDiag>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_tc_forward(TcParams P, const float* __restrict__ x, const double* __restrict__ xd,
                 const int16_t* __restrict__ samples, const int64_t* __restrict__ offsets, int n_reads,
                 int side, int n_windows, float* __restrict__ probs) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    long long* const trace0 = kDiag ? P.trace : nullptr;
    const int dbg_job = kDiag ? P.dbg_job : -1;
    const uint32_t wbuf = sbase + kSmemWbuf;
    const uint32_t prm = sbase + kSmemPrm;
    const uint32_t bar0 = sbase + kSmemBar;
    const uint32_t bar_wfull[2] = {bar0 + 0, bar0 + 8};     // weight part p landed in smem
    const uint32_t bar_wfree[2] = {bar0 + 16, bar0 + 24};   // every MMA reading weight part p completed
    const uint32_t bar_mma[2] = {bar0 + 32, bar0 + 40};
    const uint32_t bar_epi[2] = {bar0 + 48, bar0 + 56};
    const uint32_t bar_final = bar0 + 64;
    // joint phase: one MMA-done and one epilogue-done barrier per joint epilogue (kJointRing >= their number)
    const uint32_t bar_jmma = bar0 + 512, bar_jepi = bar0 + 512 + 8 * kJointRing;   // + 8 * eseq
    // joint phase weights: three whole-job slots (the normal weight buffer and two halves of the upper
    // part of window 1's ACT region, which is dead from conv1d_8 on), so that the loader runs two
    // jobs ahead of the MMA issuer and independent jobs follow each other without a weight bubble
    // one weights-landed and one weights-free barrier PER JOINT JOB (never reused: the two issuers each see only
    // their own jobs, and a parity wait is only safe for a waiter that has observed every earlier phase)
    const uint32_t bar_jwfull0 = bar0 + 128, bar_jwfree0 = bar0 + 256;   // + 8 * jk
    const uint32_t bar_x = bar0 + 384;   // both windows' epilogues of the last single-window job are done
    const uint32_t jwslot1 = sbase + kSmemAct1 + kYOff, jwslot2 = jwslot1 + kWbufBytes;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 96);
    const int tid = threadIdx.x;
    // warp index via shfl: tells the compiler it is warp-uniform, so the role branches below are
    // convergent regions and the MMA issuer's address arithmetic can use the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    const bool is_epi = warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps;
    // PERSISTENT over the window pairs of the launch (grid = min(pairs, SMs)); every role loops over the pairs on
    // its own, there is no CTA-wide barrier between two pairs: the barriers are not re-initialised, a barrier that
    // completes once per pair is waited for with parity `it & 1`, the barriers of the single-window jobs (weights,
    // MMA-done, bar_epi[w]) complete eight times per pair, so their parities restart with every pair.
    const int npairs = (n_windows + 1) / 2;
    // barrier set-up, shared by two otherwise idle threads (one thread needs ~10 cycles per mbarrier.init, and
    // there is one barrier per joint job / joint epilogue)
    if (threadIdx.x == kLoadWarp * 32) {
        mbar_init(bar_wfull[0], 1);
        mbar_init(bar_wfull[1], 1);
        mbar_init(bar_wfree[0], 2);   // one commit per window pass (issuer 0 commits twice where it issues both windows)
        mbar_init(bar_wfree[1], 2);
        mbar_init(bar_mma[0], 1);
        mbar_init(bar_mma[1], 1);
        mbar_init(bar_epi[0], kEpiArrivals);
        mbar_init(bar_epi[1], kEpiArrivals);
        mbar_init(bar_final, 2);
        for (int i = 0; i < kJointRing; ++i) {
            mbar_init(bar_jwfull0 + 8 * i, 1);
            mbar_init(bar_jwfree0 + 8 * i, 1);
        }
        fence_barrier_init();
    } else if (threadIdx.x == kMmaWarpB * 32) {
        mbar_init(bar_x, 2 * kEpiArrivals);
        for (int i = 0; i < kJointRing; ++i) {
            mbar_init(bar_jmma + 8 * i, 1);
            mbar_init(bar_jepi + 8 * i, kEpiArrivals);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(sbase + kSmemBar + 96, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int njobs = (dbg_job >= 0 && dbg_job < P.njobs) ? dbg_job + 1 : P.njobs;

    if (is_epi) {
        // ================= epilogue / CUDA-core warps =================
        const int tid = static_cast<int>(threadIdx.x) - kEpiWarp0 * 32;   // epilogue-relative thread id
        const int ewarp = tid >> 5;
        for (int pair = blockIdx.x, it = 0; pair < npairs; pair += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it) & 1u;
        long long* const trace = it == 0 ? trace0 : nullptr;   // the timeline is that of the CTA's first pair
            // Epilogue threads issue every global load of the prologue BEFORE the set-up barrier (conv1
            // parameters, in predict mode the samples of BOTH windows): their latency overlaps the barrier
            // initialisation and the TMEM allocation, which the otherwise idle loader warp / MMA warp do.
            int win[2];
            bool valid[2];
        #pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int idx = 2 * pair + w;
                valid[w] = idx < n_windows;
                win[w] = valid[w] ? idx : n_windows - 1;
            }
            Conv1Params c1;
            WindowInput in[2] = {};
            float xv[2][3];
            int raw[2][3];            // call mode: the thread's raw samples of both windows (0 outside the slice)
            WindowGeom geom[2] = {};
            {
                const int etid = tid;
                load_conv1_params(P, etid, c1);
                if (!kCallMode) {
        #pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        if (x) in[w].x = x + static_cast<size_t>(win[w]) * kInputSize;
                        else in[w].xd = xd + static_cast<size_t>(win[w]) * kInputSize;
                        fetch_window_inputs(in[w], etid, xv[w]);
                    }
                } else {
                    // fused call_batch (classify.py:342-357): the thread's samples of BOTH windows are requested
                    // here, so that the two dependent global latencies (offsets, samples) overlap the set-up
                    int64_t off[2], len[2];
        #pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int read = win[w] % n_reads;
                        off[w] = __ldg(offsets + read);
                        len[w] = __ldg(offsets + read + 1) - off[w];
                    }
        #pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        geom[w] = window_geometry(static_cast<int>(len[w]), win[w] / n_reads, side);
                        const int16_t* region = samples + off[w] + geom[w].a;
        #pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int i = etid + k * kEpiThreads - geom[w].dst;   // index into the slice
                            raw[w][k] = (etid + k * kEpiThreads < kInputSize && i >= 0 && i < geom[w].n)
                                            ? static_cast<int>(__ldg(region + i)) : 0;
                        }
                    }
                }
            }

        if (trace && blockIdx.x == 0 && tid == 0) trace[31 * 32 + 0] = clock64();
        if (it == 0)
            for (int i = tid; i < P.prm_floats / 4; i += kEpiThreads)
                reinterpret_cast<float4*>(smem + kSmemPrm)[i] = __ldg(reinterpret_cast<const float4*>(P.prm) + i);
        if (kCallMode) {
            // z-score of both windows (trim_signal.py:61-69): exact integer sums over the slices, reduced
            // over the 384 threads (samples outside a slice are 0 and do not count), mean / population
            // stdev in double as numpy computes them
            long long s[4] = {0, 0, 0, 0};
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    s[2 * w] += raw[w][k];
                    s[2 * w + 1] += static_cast<long long>(raw[w][k]) * raw[w][k];
                }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
            // scratch [12 warps][4] (384 B) at the start of window 0's region: conv1_stage writes its first row there
            // only after two more barriers of the epilogue threads
            long long* red = reinterpret_cast<long long*>(smem + kSmemAct0);
            if ((tid & 31) == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) red[ewarp * 4 + q] = s[q];
            }
            epi_bar_sync();
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] = 0;
            for (int i = 0; i < kEpiWarps; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) s[q] += red[i * 4 + q];
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                double mean = 0.0, stdev = 0.0;
                if (geom[w].n > 0) zscore_params(s[2 * w], s[2 * w + 1], geom[w].n, &mean, &stdev);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = tid + k * kEpiThreads - geom[w].dst;
                    const bool inside = tid + k * kEpiThreads < kInputSize && i >= 0 && i < geom[w].n;
                    const double d = static_cast<double>(raw[w][k]) - mean;
                    xv[w][k] = inside ? static_cast<float>(stdev > 0.0 ? d / stdev : d) : 0.f;
                }
            }
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            conv1_stage(c1, xv[w][0], xv[w][1], xv[w][2], sbase + (w ? kSmemAct1 : kSmemAct0),
                        smem + (w ? kSmemAct1 : kSmemAct0), tid);
            fence_proxy_async();
            epi_arrive(bar_epi[w]);
            if (trace && blockIdx.x == 0 && tid == 0) trace[31 * 32 + 1 + w] = clock64();
        }
        epi_bar_sync();   // parameter block staged by all epilogue threads is now visible
        uint32_t mma_phase[2] = {0, 0};
        float* p0 = valid[0] ? probs + static_cast<size_t>(win[0]) * P.n_classes : nullptr;
        float* p1 = valid[1] ? probs + static_cast<size_t>(win[1]) * P.n_classes : nullptr;
        for (int j = 0; j < njobs; ++j) {
            const TcJob& J = c_jobs[j];
            if ((tid & 31) == 0) prefetch_job(j + 1);
            if (!J.last) continue;
            if (J.joint) {   // one pass serves both windows
                const int e = J.eseq;
                long long* tr = (trace && blockIdx.x == 0 && tid == 0) ? trace + (j * 2) * 16 : nullptr;
                run_epilogue(P, J, sbase + kSmemAct0, sbase + kSmemAct0, 0, prm, tmem_base + J.tcol, tid,
                             bar_jmma + 8 * e, ph, p0, p1, tr);
                if (tr) tr[6] = clock64();
                fence_proxy_async();
                if (tr) tr[7] = clock64();
                tc_fence_before();
                if (J.last & 2) epi_arrive(bar_jepi + 8 * e);   // only where an issuer waits on exactly this epilogue (TcJob::last)
                if (tr) tr[3] = clock64();
                continue;
            }
            for (int w = 0; w < 2; ++w) {
                const uint32_t act = sbase + (w ? kSmemAct1 : kSmemAct0);
                long long* tr = (trace && blockIdx.x == 0 && tid == 0) ? trace + (j * 2 + w) * 16 : nullptr;
                run_epilogue(P, J, act, sbase + kSmemAct0, w, prm, tmem_base + w * kTmemWindowCols, tid,
                             bar_mma[w], mma_phase[w], p0, p1, tr);
                mma_phase[w] ^= 1;
                if (tr) tr[6] = clock64();
                fence_proxy_async();
                if (tr) tr[7] = clock64();
                tc_fence_before();
                // (the last single-window job is followed by the joint jobs, which wait for BOTH windows on bar_x: bar_epi[w]
                // then completes eight times per pair and every one of its phases has a waiter)
                if (j + 1 < njobs && c_jobs[j + 1].joint) epi_arrive(bar_x);
                else epi_arrive(bar_epi[w]);
                if (tr) tr[3] = clock64();
            }
        }
        if (dbg_job >= 0) {   // debug: dump both ACT regions after the last processed job
            mbar_wait(bar_final, 0);
            epi_bar_sync();
            if (blockIdx.x == 0)
                for (int i = tid; i < 2 * kActBytes / 16; i += kEpiThreads)
                    reinterpret_cast<uint4*>(P.dbg_out)[i] = reinterpret_cast<const uint4*>(smem)[i];
        }
        }   // pairs
    } else if (warp == kMmaWarp || warp == kMmaWarpB) {
        // ================= the two MMA issuers (one elected lane of each warp) =================
        // A single issuing thread is the bottleneck of everything behind conv1d_4: between two jobs it needs
        // ~1 000 cycles (wait for the queue to drain before the commits may overwrite uniform registers the
        // queued MMAs reference, barrier polls, descriptor set-up) against ~850 cycles of MMAs of a small job,
        // and the tensor pipe only has the ~8 MMAs of its queue to bridge that.  So there are two issuers:
        //   * conv1d_2..4 (`both`): issuer 0 alone, window 0 then window 1 - the tensor pipe is saturated
        //     there, and strict alternation keeps one window's epilogue under the other window's MMAs;
        //   * conv1d_5..9: issuer w issues window w, so the two windows form independent chains;
        //   * joint jobs: issued by their `owner` (the jobs alternate; the K-slices of conv1d_17 stay with one
        //     issuer so that the accumulation order is fixed).
        // tcgen05.commit only tracks the MMAs of the committing thread, which is exactly what every barrier
        // here wants.  All barrier parities follow from the job index (every single-window job completes one
        // phase of bar_epi[w] / bar_wfull[p] / bar_wfree[p]), so the issuers share no state.
        // A full 512-column allocation by the only CTA on the SM starts at TMEM address 0; using
        // the literal keeps every MMA operand derived from uniform sources.
        if (tmem_base != 0) __trap();
        const int me = warp == kMmaWarp ? 0 : 1;
        if (elect_one()) {
            constexpr uint32_t leader = 1;   // (a converged-warp variant with predicated tcgen05 ops was slower: R2UR.BROADCAST per operand)
            const uint32_t wp16[2] = {wbuf >> 4, (wbuf + kWPart0) >> 4};
            const uint32_t act16_0 = (sbase + kSmemAct0) >> 4, act16_1 = (sbase + kSmemAct1) >> 4;
            // The next job's record is fetched (raw, see IssueRec) while this job's MMAs are issued.
            for (int pair = blockIdx.x, it = 0; pair < npairs; pair += gridDim.x, ++it) {
            const uint32_t ph = static_cast<uint32_t>(it) & 1u;
            long long* const trace = it == 0 ? trace0 : nullptr;
            const bool tracing = trace && blockIdx.x == 0;
            bool in_joint = false;
            IssueRaw nxt = load_issue_raw(P.issue);
            for (int j = 0; j < njobs; ++j) {
                const IssueArgs J = unpack_issue(nxt);
                if (j + 1 < njobs) nxt = load_issue_raw(P.issue + 4 * (j + 1));
                const uint32_t blk16 = J.blk16;                  // one K=16 block of B, in 16-byte units
                const uint32_t tap16[3] = {J.tap16[0], J.tap16[1], J.tap16[2]};
                const bool first = J.first != 0, last = J.last != 0;
                const uint32_t parw = static_cast<uint32_t>(j) & 1u;           // parity of the weight barriers for job j
                const uint32_t par = parw;                                     // ... and of bar_epi[w] (eight completions per pair, too)
                if (J.joint) {
                    if (J.owner != me) continue;
                    // ---- both windows in one burst: [part 0: w0, w1] [part 1: w0, w1], one commit ----
                    // whole job's weights in slot (jk + 1) % 3 (slot 0 = the two-part buffer, which the last
                    // single-window job still uses while the first two joint jobs are loaded)
                    const int jk = J.jk, slot = (jk + 1) % 3;
                    const uint32_t jw0_16 = (slot == 0 ? wbuf : slot == 1 ? jwslot1 : jwslot2) >> 4;
                    const uint32_t jw1_16 = jw0_16 + J.part1_16;   // behind part 0
                    if (tracing) trace[(j * 2) * 16 + 11] = clock64();
                    mbar_wait(bar_jwfull0 + 8 * jk, ph);
                    if (tracing) trace[(j * 2) * 16 + 12] = clock64();
                    if (!in_joint) mbar_wait(bar_x, ph);   // this issuer's first joint job: both windows' epilogues of the last
                    in_joint = true;                      // single-window job (X of both windows written, accumulators drained)
                    // epilogues complete in order and `need` never decreases: the last needed one is enough
                    if (J.need > 0) mbar_wait(bar_jepi + 8 * (J.need - 1), ph);
                    if (tracing) trace[(j * 2) * 16 + 13] = clock64();
                    tc_fence_after();
                    if (tracing) trace[(j * 2) * 16 + 0] = clock64();
                    const int nw = J.joint == JOINT_PAIR ? 2 : 1;
                    const uint32_t dcol = J.tcol;
                    if (tracing) trace[(j * 2) * 16 + 8] = clock64();
                    for (int w = 0; w < nw; ++w)
                        issue_job_part<0>(J.ntaps, J.ncb, dcol + (static_cast<uint32_t>(16 * w) << 16), 1,
                                          (w ? act16_1 : act16_0), tap16, J.cb0, J.lp, J.lo16, jw0_16, blk16, J.n,
                                          J.idesc, first, leader);
                    if (tracing) trace[(j * 2) * 16 + 9] = clock64();
                    if (tracing) trace[(j * 2) * 16 + 10] = clock64();
                    for (int w = 0; w < nw; ++w)
                        issue_job_part<1>(J.ntaps, J.ncb, dcol + (static_cast<uint32_t>(16 * w) << 16), 1,
                                          (w ? act16_1 : act16_0), tap16, J.cb0, J.lp, J.lo16, jw1_16, blk16, J.n,
                                          J.idesc, false, leader);
                    if (tracing) trace[(j * 2) * 16 + 14] = clock64();   // last MMA issued
                    if (last) tc_commit(bar_jmma + 8 * J.eseq, leader);
                    tc_commit(bar_jwfree0 + 8 * jk, leader);
                    if (tracing) trace[(j * 2) * 16 + 1] = clock64();
                    continue;
                }
                if (J.both && me != 0) {
                    // issuer 1 only follows the barriers it will wait on later: a parity wait is only safe for a
                    // waiter that has observed every earlier phase
                    mbar_wait(bar_epi[1], par);
                    mbar_wait(bar_wfull[0], parw);
                    mbar_wait(bar_wfull[1], parw);
                    continue;
                }
                const int w_first = J.both ? 0 : me, w_last = J.both ? 1 : me;
                for (int w = w_first; w <= w_last; ++w) {
                    // issuer 0 frees the weight parts for both windows when it issues both
                    const bool free_now = w == w_last;
                    const int nfree = J.both ? 2 : 1;
                    if (tracing) trace[(j * 2 + w) * 16 + 11] = clock64();
                    // weight part 0 first (normally landed long ago): its poll then overlaps the wait for the input
                    if (w == w_first) mbar_wait(bar_wfull[0], parw);
                    mbar_wait(bar_epi[w], par);   // input written and previous accumulators drained
                    if (tracing) trace[(j * 2 + w) * 16 + 12] = clock64();
                    tc_fence_after();
                    if (tracing) trace[(j * 2 + w) * 16 + 0] = clock64();
                    const uint32_t dwin = w * kTmemWindowCols;
                    // ---- weight part 0 (first K blocks); freed early so the loader can refill it ----
                    if (tracing) trace[(j * 2 + w) * 16 + 8] = clock64();
                    issue_job_part<0>(J.ntaps, J.ncb, dwin, J.ntiles, (w ? act16_1 : act16_0), tap16, J.cb0, J.lp, J.lo16, wp16[0],
                                      blk16, J.n, J.idesc, first, leader);
                    if (free_now)
                        for (int c = 0; c < nfree; ++c) tc_commit(bar_wfree[0], leader);
                    if (tracing) trace[(j * 2 + w) * 16 + 9] = clock64();
                    // ---- weight part 1 (remaining K blocks) ----
                    if (w == w_first) mbar_wait(bar_wfull[1], parw);
                    if (tracing) trace[(j * 2 + w) * 16 + 10] = clock64();
                    issue_job_part<1>(J.ntaps, J.ncb, dwin, J.ntiles, (w ? act16_1 : act16_0), tap16, J.cb0, J.lp, J.lo16, wp16[1],
                                      blk16, J.n, J.idesc, false, leader);
                    if (tracing) trace[(j * 2 + w) * 16 + 14] = clock64();   // last MMA issued
                    if (last) tc_commit(bar_mma[w], leader);
                    if (free_now)
                        for (int c = 0; c < nfree; ++c) tc_commit(bar_wfree[1], leader);
                    if (tracing) trace[(j * 2 + w) * 16 + 1] = clock64();
                }
            }
            }   // pairs
            tc_commit(bar_final, leader);
            mbar_wait(bar_final, 0);
        }
    } else if (warp == kLoadWarp && elect_one()) {
        // ================= weight loader (one elected lane) =================
        for (int pair = blockIdx.x, it = 0; pair < npairs; pair += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it) & 1u;
        long long* const trace = it == 0 ? trace0 : nullptr;
        uint32_t free_phase = 0;
        int jk = 0;
        for (int j = 0; j < njobs; ++j) {
            const TcJob& J = c_jobs[j];
            const unsigned char* src = P.w + J.w_goff;
            if (J.joint) {   // whole job ([part 0 | part 1] as packed) into slot (jk + 1) % 3
                const int slot = (jk + 1) % 3;
                if (jk == 2) {   // first use of slot 0, the normal weight buffer: the last two-part job must be done with it
                    if (j > 2) {
                        mbar_wait(bar_wfree[0], free_phase);
                        mbar_wait(bar_wfree[1], free_phase);
                    }
                } else if (jk >= 3) {
                    mbar_wait(bar_jwfree0 + 8 * (jk - 3), ph);
                }
                // (slots 1 and 2 - used by the first two joint jobs, whose weights are therefore requested while
                // conv1d_9 is still running - lie in window 1's region above everything conv1d_8.. keeps there;
                // the loader gets here only after conv1d_9's weights were requested, i.e. after conv1d_8's
                // MMAs - and with them conv1d_7's epilogue, the last user of that space - completed)
                const uint32_t bytes = J.w_part[0] + J.w_part[1];
                if (trace && blockIdx.x == 0) trace[(j * 2 + 1) * 16 + 15] = clock64();   // when the load was issued
                mbar_expect_tx(bar_jwfull0 + 8 * jk, bytes);
                bulk_g2s(slot == 0 ? wbuf : slot == 1 ? jwslot1 : jwslot2, src, bytes, bar_jwfull0 + 8 * jk);
                ++jk;
                continue;
            }
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
        // the weight slots are reused by the next pair: the MMAs of the last three joint jobs must be done with them
        for (int k = jk > 3 ? jk - 3 : 0; k < jk; ++k) mbar_wait(bar_jwfree0 + 8 * k, ph);
        }   // pairs
    }
    // (the control warps work on one elected lane; a CTA-wide barrier counts a warp as arrived as soon as any of
    // its lanes executes it, so every warp reconverges first)
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TcEngine {
    unsigned char* d_w = nullptr;
    uint4* d_issue = nullptr;
    float* d_prm = nullptr;
    TcParams params{};
    int njobs = 0;
    int sm_count = 148;        // grid of the persistent kernel = min(window pairs, SMs)
    std::vector<TcJob> jobs;   // uploaded to constant memory (identical for every model of this topology)
};

static uint16_t bf16_rn(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);   // inf / nan
    u += 0x7FFFu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
static float bf16_to_float(uint16_t h) {
    uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct JobBuilder {
    const Blob& blob;
    std::vector<TcJob> jobs;
    std::vector<unsigned char> w;
    std::vector<float> prm;        // smem-resident block: per job bias[n] (+ scale[48] shift[48])
    std::vector<float> bn_scale[8], bn_shift[8];

    explicit JobBuilder(const Blob& b) : blob(b) {
        for (int i = 1; i <= 7; ++i) fold_bn(blob, i, &bn_scale[i], &bn_shift[i]);
    }

    // pack W[tap][cin][cout] -> two parts (K blocks [0, split) and [split, nkb), block kb = tap kb / ncb,
    // channel block kb % ncb), each part [hi blocks | lo blocks], block = [2 chunks][n rows][8] bf16
    // `layer2` (optional): a second conv with the same input and kernel size whose output channels
    // are appended along N (rows cout .. cout + cout2 - 1): one MMA job computes both.
    // `fold_avg3`: a 1x1 conv that follows AveragePooling1D(3, 1, 'same') becomes a k=3 conv with every
    // tap = W / 3 (the epilogue rescales the two end positions, where TF divides by 2).
    void pack_weights(int layer, int n, int cb0, int ncb, TcJob* J, int layer2 = 0, bool fold_avg3 = false) {
        const ConvSpec& s = kConvSpecs[layer];
        const int ktaps = fold_avg3 ? 3 : s.k;
        const int cout = s.cout ? s.cout : blob.n_classes;
        const float* k = blob.find("conv1d_" + std::to_string(layer) + "/kernel")->data;
        const int cout2 = layer2 ? kConvSpecs[layer2].cout : 0;
        const float* k2 = layer2 ? blob.find("conv1d_" + std::to_string(layer2) + "/kernel")->data : nullptr;
        const int nkb = ktaps * ncb;
        const int split = nkb == 9 ? 5 : 2;
        const size_t blk = static_cast<size_t>(2) * n * 8;   // bf16 elements per K block
        while (w.size() % 128) w.push_back(0);
        J->w_goff = static_cast<int>(w.size());
        const int range[3] = {0, split, nkb};
        for (int part = 0; part < 2; ++part) {
            const int kb0 = range[part], kb1 = range[part + 1], cnt = kb1 - kb0;
            std::vector<uint16_t> buf(2 * cnt * blk, 0);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int t = kb / ncb, cb = kb % ncb;
                for (int j = 0; j < 2; ++j)
                    for (int row = 0; row < n; ++row)
                        for (int e = 0; e < 8; ++e) {
                            const int cin = (cb0 + cb) * 16 + j * 8 + e;
                            float v = 0.f;
                            if (fold_avg3) v = row < cout && cin < s.cin ? k[cin * cout + row] / 3.0f : 0.f;
                            else if (row < cout && cin < s.cin) v = k[(t * s.cin + cin) * cout + row];
                            else if (row < cout + cout2 && cin < s.cin) v = k2[(t * s.cin + cin) * cout2 + row - cout];
                            const uint16_t hi = bf16_rn(v);
                            const uint16_t lo = bf16_rn(v - bf16_to_float(hi));
                            const size_t idx = (static_cast<size_t>(kb - kb0) * 2 + j) * n * 8 + row * 8 + e;
                            buf[idx] = hi;
                            buf[cnt * blk + idx] = lo;
                        }
            }
            J->w_part[part] = static_cast<int>(buf.size() * 2);
            const unsigned char* p = reinterpret_cast<const unsigned char*>(buf.data());
            w.insert(w.end(), p, p + buf.size() * 2);
        }
    }

    // bias (padded to n) and, if bn > 0, the folded scale/shift of channels [ch0, ch0+48) of BN `bn`
    void pack_params(int layer, int n, int bn, int ch0, TcJob* J, int layer2 = 0) {
        const BlobTensor* t = blob.find("conv1d_" + std::to_string(layer) + "/bias");
        const BlobTensor* t2 = layer2 ? blob.find("conv1d_" + std::to_string(layer2) + "/bias") : nullptr;
        const int c1 = static_cast<int>(t->count);
        J->bias_off = static_cast<int>(prm.size());
        for (int c = 0; c < n; ++c)
            prm.push_back(c < c1 ? t->data[c] : (t2 && c - c1 < static_cast<int>(t2->count) ? t2->data[c - c1] : 0.f));
        J->bn_off = 0;
        if (bn > 0) {
            J->bn_off = static_cast<int>(prm.size());
            for (int c = 0; c < 48; ++c) prm.push_back(bn_scale[bn][ch0 + c]);
            for (int c = 0; c < 48; ++c) prm.push_back(bn_shift[bn][ch0 + c]);
        }
    }

    // generic conv job; in_* describe the input tensor, out_* the output tensor
    TcJob& add(int layer, int L, int in_off, int in_lp, int in_lo_delta, int kind, int bn, int out_off,
               int out_cg_base, int layer2 = 0, bool fold_avg3 = false) {
        const ConvSpec& s = kConvSpecs[layer];
        const int cout = (s.cout ? s.cout : blob.n_classes) + (layer2 ? kConvSpecs[layer2].cout : 0);
        const int ktaps = fold_avg3 ? 3 : s.k;
        const bool pool = kind == EPI_N48_POOL_BN || kind == EPI_PARITY;
        TcJob J{};
        J.n = (cout + 15) / 16 * 16;
        J.idesc = 128;
        J.ntiles = (L + 127) / 128;
        J.L = L;
        J.lp = in_lp;
        J.ntaps = ktaps;
        for (int t = 0; t < 3; ++t) J.tap16[t] = in_off + (ktaps == 3 ? t : 1) * 16;
        J.lo16 = in_lo_delta;
        J.ncb = s.cin / 16;
        J.cb0 = 0;
        J.first = J.last = 1;
        J.kind = kind;
        J.edge15 = fold_avg3 ? 1 : 0;
        J.out_L = pool ? L / 2 : L;
        J.out_off = out_off;
        // pooled outputs get two spare rows per channel group: the two lanes of a pool pair store to channel groups
        // g and g + 1, and with a group pitch of 64 (mod 128) bytes their 16-byte rows fall on disjoint banks
        // ((L/2 + 2) * 16 is 32 mod 128 for L/2 = 256, 128, 64: a 2-way bank conflict on every pooled store)
        J.out_lp = J.out_L + (pool && kind != EPI_PARITY ? 4 : 2);
        J.out_ncg = cout / 8;
        J.out_lo_delta = J.out_ncg * J.out_lp * 16;
        J.out_cg_base = out_cg_base;
        pack_weights(layer, J.n, 0, J.ncb, &J, layer2, fold_avg3);
        pack_params(layer, J.n, bn, kind == EPI_PARITY ? out_cg_base * 8 : 0, &J, layer2);
        jobs.push_back(J);
        return jobs.back();
    }

    // Convert the builder's M / byte offsets into what the MMA issuer consumes directly.
    void finalize_jobs() {
        for (TcJob& J : jobs) {
            J.idesc = static_cast<int>(make_idesc(J.idesc, J.n));
            for (int t = 0; t < 3; ++t) J.tap16[t] >>= 4;
            J.lo16 >>= 4;
        }
    }

    // Mark the jobs from index `j0` on as the joint phase: accumulator slots rotate over kJointSlots 64-column
    // slots (K-slices of one conv share a slot) and `need` is derived from the data flow: `producer[j]`
    // = job whose epilogue writes this job's input (-1: available before the joint phase).
    void finish_joint(int j0, const std::vector<int>& producer, int slot_cols = kTmemTileCols) {
        std::vector<int> eseq_of(jobs.size(), -1), slot_of(jobs.size(), 0);
        int e = 0, slot = -1;
        std::vector<int> last_user(kJointSlots, -1);   // job with the epilogue that last drained the slot
        for (size_t j = j0; j < jobs.size(); ++j) {
            TcJob& J = jobs[j];
            if (J.first) slot = (slot + 1) % kJointSlots;
            slot_of[j] = slot;
            J.tcol = slot * slot_cols;
            int need = 0;
            const int prod = producer[j - j0];
            if (prod >= 0) need = std::max(need, eseq_of[prod] + 1);
            if (J.first && last_user[slot] >= 0) need = std::max(need, eseq_of[last_user[slot]] + 1);
            if (j > static_cast<size_t>(j0)) need = std::max(need, jobs[j - 1].need);   // waits are cumulative
            J.need = need;
            J.eseq = -1;
            if (J.last) {
                J.eseq = e;
                eseq_of[j] = e++;
                last_user[slot] = static_cast<int>(j);
            }
        }
        // the issuers wait for epilogue need - 1 only (epilogues complete in order): mark those epilogues
        for (size_t j = j0; j < jobs.size(); ++j)
            if (jobs[j].last)
                for (size_t k = j0; k < jobs.size(); ++k)
                    if (jobs[k].need - 1 == jobs[j].eseq) jobs[j].last |= 2;
    }
};

static bool build_jobs(const Blob& blob, JobBuilder* B, TcParams* P) {
    if (blob.n_classes > 16) return false;
    // T1 (BN1 output) is written by conv1_stage: [6][514][8], lo at +49344
    B->add(2, 512, 0, 514, 49344, EPI_N48, 0, 0, 0);
    B->add(3, 512, 0, 514, 49344, EPI_N48, 0, 0, 0);
    B->add(4, 512, 0, 514, 49344, EPI_N48_POOL_BN, 2, 0, 0);      // -> [6][260][8], lo +24960 (pooled: pitch L/2 + 4)
    B->add(5, 256, 0, 260, 24960, EPI_N16, 0, 0, 0);              // -> [2][258][8], lo +8256
    B->add(6, 256, 0, 258, 8256, EPI_N48, 0, 0, 0);               // -> [6][258][8], lo +24768
    B->add(7, 256, 0, 258, 24768, EPI_N48_POOL_BN, 3, 0, 0);      // -> [6][132][8], lo +12672
    B->add(8, 128, 0, 132, 12672, EPI_N48, 0, 0, 0);              // -> [6][130][8], lo +12480
    B->add(9, 128, 0, 130, 12480, EPI_N48_POOL_BN, 4, 0, 0);      // X [6][68][8], lo +6528
    // ---- joint phase (both windows per job) ----
    // Inception block, JOINT_PAIR: X @0 (pitch 68), T15 @13056, T1214 @25728 (conv1d_12 | conv1d_14 outputs as one
    // 32-channel tensor: both are 1x1 convs of X, so ONE job with N = 32 computes them); Y (parity
    // split, both windows) in window 0's region @43392.  The average pool in front of conv1d_10 is
    // folded into its weights (k=3, W/3).  Concat order [conv10, conv11, conv13, conv16]
    // (network_architecture.py:68), BN5 per channel.  Job order: the dependency chain conv1d_14 -> 15 -> 16
    // with the independent branches between its links (conv1d_10 behind conv1d_12+14, conv1d_11 and 13
    // behind conv1d_15), each in its own accumulator slot, and the jobs alternate between the two MMA
    // issuers, so that the tensor pipe always has a job whose input is ready while the epilogue warps
    // drain the previous ones.
    const int j0 = static_cast<int>(B->jobs.size());
    std::vector<int> producer;
    auto joint = [&](TcJob& J, int kind, int prod) {
        J.joint = kind;
        J.idesc = 64;
        J.ntiles = 1;
        producer.push_back(prod);
    };
    joint(B->add(12, 64, 0, 68, 6528, EPI_N16, 0, 25728, 0, 14), JOINT_PAIR, -1);              // j0: -> [4][66][8], lo +4224
    joint(B->add(10, 64, 0, 68, 6528, EPI_PARITY, 5, 0, 0, 0, true), JOINT_PAIR, -1);          // j0+1: avg pool folded
    B->jobs.back().zero_y = 1;
    joint(B->add(15, 64, 25728 + 2 * 66 * 16, 66, 4224, EPI_N48, 0, 13056, 0), JOINT_PAIR, j0);   // j0+2: groups 2-3 (conv1d_14)
    joint(B->add(11, 64, 0, 68, 6528, EPI_PARITY, 5, 0, 6), JOINT_PAIR, -1);                   // j0+3
    joint(B->add(13, 64, 25728, 66, 4224, EPI_PARITY, 5, 0, 12), JOINT_PAIR, j0);              // j0+4: groups 0-1 (conv1d_12)
    joint(B->add(16, 64, 13056, 66, 6336, EPI_PARITY, 5, 0, 18), JOINT_PAIR, j0 + 2);          // j0+5
    // conv1d_17 .. conv1d_20, JOINT_STACK: both windows stacked in one tile (row = 18 w + position).
    // conv1d_17: stride 2 on the parity-split Y (tap0 = Ye[i], tap1 = Yo[i], tap2 = Ye[i+1]),
    // K = 3 x 192 split in 4 jobs of 3 channel blocks so each weight chunk fits the buffer; slice s reads
    // the 48 channels written by one inception branch (concat order conv1d_10, 11, 13, 16), so the first
    // three slices are issued while conv1d_16 is still in flight
    const int y_writer[4] = {j0 + 1, j0 + 3, j0 + 4, j0 + 5};
    for (int s = 0; s < 4; ++s) {
        TcJob J{};
        J.n = 48; J.idesc = 64; J.ntiles = 1; J.L = 34; J.lp = kYRows; J.ntaps = 3;
        J.tap16[0] = kYOff; J.tap16[1] = kYOff + 2 * kYArray; J.tap16[2] = kYOff + 16;
        J.lo16 = kYArray; J.ncb = 3; J.cb0 = 3 * s;
        J.first = (s == 0); J.last = (s == 3);
        J.joint = JOINT_STACK;
        J.kind = EPI_N48_BN;
        J.out_L = 34; J.out_off = 0; J.out_lp = 36; J.out_ncg = 6; J.out_lo_delta = 6 * 36 * 16;
        J.out_cg_base = 0;
        B->pack_weights(17, 48, 3 * s, 3, &J);
        if (J.last) B->pack_params(17, 48, 6, 0, &J);
        B->jobs.push_back(J);
        producer.push_back(y_writer[s]);
    }
    joint(B->add(18, 34, 0, 36, 3456, EPI_N48, 0, 0, 0), JOINT_STACK, j0 + 9);
    joint(B->add(19, 34, 0, 36, 3456, EPI_N48_POOL_BN, 7, 0, 0), JOINT_STACK, j0 + 10);   // -> [6][21][8], lo +2016
    joint(B->add(20, 17, 0, 21, 2016, EPI_HEAD, 0, 0, 0), JOINT_STACK, j0 + 11);
    B->jobs.back().idesc = 128;   // the head epilogue reads all 17 rows from TMEM lane quadrant 0
    B->finish_joint(j0, producer);
    B->finalize_jobs();
    if (B->prm.size() > static_cast<size_t>(kPrmFloats) || B->jobs.size() > static_cast<size_t>(kMaxJobs))
        return false;
    for (const TcJob& J : B->jobs)
        if (J.w_part[0] > kWPart0 || J.w_part[1] > kWPart1) return false;
    while (B->prm.size() % 4) B->prm.push_back(0.f);
    P->prm_floats = static_cast<int>(B->prm.size());

    // conv1 parameters live behind the smem block in the global parameter buffer
    auto push = [&](const float* p, size_t n) {
        while (B->prm.size() % 4) B->prm.push_back(0.f);
        const int off = static_cast<int>(B->prm.size());
        B->prm.insert(B->prm.end(), p, p + n);
        return off;
    };
    P->conv1_w = push(blob.find("conv1d_1/kernel")->data, 144);
    P->conv1_b = push(blob.find("conv1d_1/bias")->data, 48);
    P->bn1_s = push(B->bn_scale[1].data(), 48);
    P->bn1_h = push(B->bn_shift[1].data(), 48);
    P->n_classes = blob.n_classes;
    return true;
}

// The MMA issuer's view of a (finalized) job.
static IssueRec issue_record(const TcJob& J, int jk, int owner, bool both) {
    IssueRec r{};
    r.n = static_cast<uint32_t>(J.n); r.idesc = static_cast<uint32_t>(J.idesc); r.ntiles = static_cast<uint32_t>(J.ntiles);
    r.lp = static_cast<uint32_t>(J.lp);
    for (int t = 0; t < 3; ++t) r.tap16[t] = static_cast<uint32_t>(J.tap16[t]);
    r.lo16 = static_cast<uint32_t>(J.lo16);
    r.shape = static_cast<uint32_t>(J.ncb | (J.ntaps << 8) | (jk << 16));
    r.cb0 = static_cast<uint32_t>(J.cb0); r.tcol = static_cast<uint32_t>(J.tcol);
    r.flags = static_cast<uint32_t>((J.first ? 1 : 0) | (J.last ? 2 : 0) | (J.joint << 2) | (owner << 4) | (both ? 32 : 0));
    r.need = static_cast<uint32_t>(J.need); r.eseq = static_cast<uint32_t>(J.eseq);
    r.blk16 = 2u * static_cast<uint32_t>(J.n);
    r.part1_16 = (J.ntaps * J.ncb == 9 ? 10u : 4u) * r.blk16;
    return r;
}

TcEngine* tc_create(const Blob& blob) {
    if (getenv("DBN_DISABLE_TC")) return nullptr;
    JobBuilder B(blob);
    TcParams P{};
    if (!build_jobs(blob, &B, &P)) return nullptr;
    TcEngine* e = new TcEngine();
    e->jobs = B.jobs;
    // issuer assignment: conv1d_2..4 (L = 512) by issuer 0 for both windows, conv1d_5..9 one window per issuer;
    // the joint jobs alternate between the issuers, the K-slices of one conv staying together
    std::vector<IssueRec> recs;
    int jk = 0, owner = 1, nepi = 0;
    for (const TcJob& J : B.jobs) {
        if (J.joint && J.first) owner ^= 1;
        recs.push_back(issue_record(J, J.joint ? jk : 0, J.joint ? owner : 0, !J.joint && J.L >= 512));
        if (J.joint) ++jk;
        if (J.joint && J.last) ++nepi;
    }
    if (nepi > kJointRing) return nullptr;
    bool ok = cudaMalloc(&e->d_w, B.w.size()) == cudaSuccess &&
              cudaMalloc(&e->d_prm, B.prm.size() * sizeof(float)) == cudaSuccess &&
              cudaMemcpy(e->d_w, B.w.data(), B.w.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(e->d_prm, B.prm.data(), B.prm.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMalloc(&e->d_issue, recs.size() * sizeof(IssueRec)) == cudaSuccess &&
              cudaMemcpy(e->d_issue, recs.data(), recs.size() * sizeof(IssueRec), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaFuncSetAttribute(k_tc_forward<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(k_tc_forward<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(k_tc_forward<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(k_tc_forward<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        tc_destroy(e);
        return nullptr;
    }
    e->njobs = static_cast<int>(B.jobs.size());
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || e->sm_count <= 0) {
        cudaGetLastError();
        e->sm_count = 148;
    }
    P.njobs = e->njobs;
    P.w = e->d_w;
    P.issue = e->d_issue;
    P.prm = e->d_prm;
    P.dbg_job = -1;
    P.dbg_out = nullptr;
    P.trace = nullptr;
    e->params = P;
    return e;
}

void tc_destroy(TcEngine* e) {
    if (!e) return;
    cudaFree(e->d_w);
    cudaFree(e->d_issue);
    cudaFree(e->d_prm);
    delete e;
}

// The job tables live in constant memory, which is PER DEVICE.  They depend only on the topology and the
// class count, so the models of a process normally share one table per kernel; the cache of what has
// been uploaded is keyed by the current device and guarded by a mutex (handles on different devices or
// threads).  If a model with a different table shows up, the device is drained before the table is
// replaced (kernels of the previous model may still be reading it).
static std::mutex g_table_mutex;
static std::map<int, std::vector<TcJob>> g_uploaded;   // device -> table
static int sync_table(const std::vector<TcJob>& jobs) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(DBN_ECUDA, "cudaGetDevice failed");
    const size_t bytes = jobs.size() * sizeof(TcJob);
    std::lock_guard<std::mutex> lock(g_table_mutex);
    std::vector<TcJob>& have = g_uploaded[dev];
    if (have.size() == jobs.size() && std::memcmp(have.data(), jobs.data(), bytes) == 0) return 0;
    if (cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpyToSymbol(c_jobs, jobs.data(), bytes) != cudaSuccess)
        return fail(DBN_ECUDA, "uploading the tcgen05 job table failed");
    have = jobs;
    return 0;
}

static int launch_check(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(DBN_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(err));
    return 0;
}

int tc_predict(TcEngine* e, const float* d_x, const double* d_xd, int64_t n, float* d_probs,
               cudaStream_t st) {
    if (int rc = sync_table(e->jobs)) return rc;
    const int grid = static_cast<int>(std::min<int64_t>((n + 1) / 2, e->sm_count));
    k_tc_forward<false, false><<<grid, kTcThreads, kTcSmemBytes, st>>>(e->params, d_x, d_xd, nullptr, nullptr,
                                                               0, 0, static_cast<int>(n), d_probs);
    return launch_check("tcgen05 kernel");
}

int tc_call_windows(TcEngine* e, const int16_t* d_samples, const int64_t* d_offsets, int n_reads,
                    int side, int steps, float* d_step_probs, cudaStream_t st) {
    const int n = n_reads * steps;
    if (int rc = sync_table(e->jobs)) return rc;
    k_tc_forward<true, false><<<std::min((n + 1) / 2, e->sm_count), kTcThreads, kTcSmemBytes, st>>>(e->params, nullptr, nullptr, d_samples,
                                                                      d_offsets, n_reads, side, n,
                                                                      d_step_probs);
    return launch_check("tcgen05 kernel");
}

int tc_num_jobs(const TcEngine* e) { return e ? static_cast<int>(e->jobs.size()) : 0; }

// Host only (no CUDA call): the MMA job table the engine would use for this model, 32 ints per job in
// the order of struct TcJob (`which` must be 0: there is one kernel).
static bool build_table(const Blob& blob, int which, JobBuilder* B) {
    if (which != 0) return false;
    TcParams P{};
    return build_jobs(blob, B, &P);
}

int tc_job_table(const Blob& blob, int which, int32_t* out, int max_jobs) {
    static_assert(sizeof(TcJob) == 32 * sizeof(int32_t), "TcJob is dumped as 32 ints");
    JobBuilder B(blob);
    if (!build_table(blob, which, &B) || static_cast<int>(B.jobs.size()) > max_jobs) return -1;
    std::memcpy(out, B.jobs.data(), B.jobs.size() * sizeof(TcJob));
    return static_cast<int>(B.jobs.size());
}

// Host only: the packed bf16 weights (what TcJob::w_goff / w_part index) and the parameter block (bias /
// folded BN, what bias_off / bn_off index) of the same table.  Returns the sizes; copies if they fit.
int tc_packed(const Blob& blob, int which, unsigned char* w_out, int64_t w_cap, float* prm_out, int64_t prm_cap,
              int64_t* w_bytes, int64_t* prm_floats) {
    JobBuilder B(blob);
    if (!build_table(blob, which, &B)) return -1;
    *w_bytes = static_cast<int64_t>(B.w.size());
    *prm_floats = static_cast<int64_t>(B.prm.size());
    if (w_out && w_cap >= *w_bytes) std::memcpy(w_out, B.w.data(), B.w.size());
    if (prm_out && prm_cap >= *prm_floats) std::memcpy(prm_out, B.prm.data(), B.prm.size() * sizeof(float));
    return 0;
}

// Diagnostics: run `n` windows with CTA 0 recording clock64 stamps per (job, window):
// [0] MMA issue start, [1] MMA issue end, [2] epilogue start (accumulators ready), [3] epilogue end.
int tc_trace(TcEngine* e, const float* d_x, int n, float* d_probs, long long* d_trace, cudaStream_t st) {
    if (int rc = sync_table(e->jobs)) return rc;
    TcParams P = e->params;
    P.trace = d_trace;
    k_tc_forward<false, true><<<(n + 1) / 2, kTcThreads, kTcSmemBytes, st>>>(P, d_x, nullptr, nullptr, nullptr, 0, 0,
                                                                      n, d_probs);
    return launch_check("tcgen05 trace");
}

// The same for the fused call_batch form of the kernel (int16 scan regions, z-score in the prologue; one window per read).
int tc_trace_call(TcEngine* e, const int16_t* d_samples, const int64_t* d_offsets, int n_reads, float* d_probs,
                  long long* d_trace, cudaStream_t st) {
    if (int rc = sync_table(e->jobs)) return rc;
    TcParams P = e->params;
    P.trace = d_trace;
    k_tc_forward<true, true><<<(n_reads + 1) / 2, kTcThreads, kTcSmemBytes, st>>>(P, nullptr, nullptr, d_samples, d_offsets,
                                                                           n_reads, 0, n_reads, d_probs);
    return launch_check("tcgen05 trace (call mode)");
}

// Debug: run windows d_x[0..1] up to and including job `job`, dump both ACT regions (2*98688 B).
int tc_debug_dump(TcEngine* e, const float* d_x, int job, unsigned char* d_out, cudaStream_t st) {
    if (int rc = sync_table(e->jobs)) return rc;
    TcParams P = e->params;
    P.dbg_job = job;
    P.dbg_out = d_out;
    k_tc_forward<false, true><<<1, kTcThreads, kTcSmemBytes, st>>>(P, d_x, nullptr, nullptr, nullptr, 0, 0, 2,
                                                            nullptr);
    return launch_check("tcgen05 debug");
}

}  // namespace dbn
