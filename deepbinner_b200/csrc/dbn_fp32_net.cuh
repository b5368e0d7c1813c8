// fp32 CUDA-core engine: the whole Deepbinner network for ONE window, executed by one CTA with all
// activations resident in shared memory.  Graph: reference network_architecture.py:18-95 (SURVEY
// Appendix A), Keras/TF inference semantics per SURVEY Appendix B.
//
// The code is written as a sequence of "phases" (DBN_PHASE) separated by CTA barriers.  On the
// device a phase body runs once per thread; when this header is compiled for the host (g++, no
// __CUDA_ARCH__) the same body is looped over all thread ids, which lets tests/ emulate the exact
// indexing of the kernel on a CPU (tests/test_fp32_emulation.py) - there is no GPU in the authoring
// container.  That emulation is test infrastructure; the product only ever runs the device build.
//
// Shared-memory activation layout: channel-major rows, `row stride = L + 8` floats, data in columns
// [4, L+4), zero halo in columns 3 and L+4 (the 'same' padding of the k=3 convolutions; padding is
// applied in post-BatchNorm space, Appendix B.5).  A thread computes 4 consecutive positions x TC
// output channels; lanes of a warp take consecutive 4-position chunks, so activation loads are
// conflict-free LDS.128 and weight loads are warp-broadcast LDS.128.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DBN_HD __host__ __device__ __forceinline__
#else
#define DBN_HD inline
#endif

namespace dbn {

constexpr int kInputSize = 1024;
constexpr int kThreads = 512;        // CTA size of the fp32 engine
constexpr int kMaxClasses = 32;
constexpr int kBufFloats = 48 * 520; // one activation buffer: [48][512+8]
constexpr int kWbufFloats = 6912;    // largest per-phase packed weight set (48->48, k=3)
constexpr int kScratchFloats = 1024; // reductions / head scratch
constexpr int kLateWOffset = 48 * 264; // late-layer weight staging area inside buffer A
static_assert(kLateWOffset + 9216 <= kBufFloats, "late weight area");
constexpr int kFp32SmemFloats = 2 * kBufFloats + kWbufFloats + kScratchFloats;

struct alignas(16) F4 {
    float x, y, z, w;
};

constexpr int round_up4(int v) { return (v + 3) & ~3; }

// Packed-weight geometry of a stride-1 conv: [CIN][COUT/TC][round_up4(K*TC)], element (t*TC + o).
template <int CIN, int COUT, int K, int TC>
struct ConvPack {
    static constexpr int kGroup = round_up4(K * TC);
    static constexpr int kNct = COUT / TC;
    static constexpr int kFloats = CIN * kNct * kGroup;
    static_assert(COUT % TC == 0, "TC must divide COUT");
};

// Offsets (in floats) of every tensor inside the engine's packed device weight buffer.
struct Fp32Layout {
    // stride-1 convs staged through shared memory
    int conv[21];       // conv[i] = offset of packed conv1d_i (i = 2..16,18,19); conv[1] = conv1d_1
    int conv17;         // Keras layout [3][192][48]
    int conv20;         // Keras layout [48][n_classes]
    int bias[21];       // bias[i] for conv1d_i
    int bn_scale[8];    // folded BatchNorm: y = scale*x + shift
    int bn_shift[8];
    int total;
};

struct Fp32Net {
    const float* w;     // packed weights (device global / host memory in emulation)
    Fp32Layout lay;
    int n_classes;
};

#if defined(__CUDA_ARCH__)
#define DBN_PHASE(...)                                                                            \
    {                                                                                             \
        const int tid = threadIdx.x;                                                              \
        (void)tid;                                                                                \
        __VA_ARGS__;                                                                              \
    }                                                                                             \
    __syncthreads();
#define DBN_LDG(p) __ldg(p)
#else
#define DBN_PHASE(...)                                                                            \
    for (int tid = 0; tid < ::dbn::kThreads; ++tid) {                                             \
        __VA_ARGS__;                                                                              \
    }
#define DBN_LDG(p) (*(p))
#endif

// Cooperative copy of `n` floats (multiple of 4, 16-byte aligned) global -> shared.
DBN_HD void stage_weights(int tid, float* dst, const float* src, int n) {
    const F4* s4 = reinterpret_cast<const F4*>(src);
    F4* d4 = reinterpret_cast<F4*>(dst);
    for (int i = tid; i < n / 4; i += kThreads) {
#if defined(__CUDA_ARCH__)
        const float4 v = __ldg(reinterpret_cast<const float4*>(s4) + i);
        d4[i] = F4{v.x, v.y, v.z, v.w};
#else
        d4[i] = s4[i];
#endif
    }
}

// Stride-1 Conv1D (+bias, ReLU) [+ MaxPool2] [+ BatchNorm affine], K in {1,3}, 'same' padding.
//   in : [CIN][L+8] activations (smem), out: [COUT][LOUT+8] (smem, already offset to the first
//   destination channel), w: packed weights in smem, bias/bn_*: global.
//   Threads [tbase, tbase+tcount) take part; the others skip.
template <int CIN, int COUT, int K, int L, int TC, bool POOL, bool BN>
DBN_HD void conv_s1(int tid, int tbase, int tcount, const float* in, float* out, const float* w,
                    const float* bias, const float* bn_scale, const float* bn_shift) {
    using Pack = ConvPack<CIN, COUT, K, TC>;
    constexpr int kChunks = L / 4;
    constexpr int kTasks = kChunks * Pack::kNct;
    constexpr int kRsIn = L + 8;
    constexpr int kLout = POOL ? L / 2 : L;
    constexpr int kRsOut = kLout + 8;
    static_assert(K == 1 || K == 3, "kernel size");
    const int lt = tid - tbase;
    if (lt < 0 || lt >= tcount) return;
    for (int task = lt; task < kTasks; task += tcount) {
        const int chunk = task % kChunks;
        const int ct = task / kChunks;
        float acc[TC][4];
#pragma unroll
        for (int o = 0; o < TC; ++o) {
            const float b = DBN_LDG(bias + ct * TC + o);
            acc[o][0] = b; acc[o][1] = b; acc[o][2] = b; acc[o][3] = b;
        }
        const float* xp = in + 4 + 4 * chunk;
        const float* wp = w + ct * Pack::kGroup;
#pragma unroll 2
        for (int c = 0; c < CIN; ++c) {
            const F4 xv = *reinterpret_cast<const F4*>(xp + c * kRsIn);
            float x[6];
            x[1] = xv.x; x[2] = xv.y; x[3] = xv.z; x[4] = xv.w;
            if (K == 3) {
                x[0] = xp[c * kRsIn - 1];
                x[5] = xp[c * kRsIn + 4];
            }
            float wv[Pack::kGroup];
            const F4* w4 = reinterpret_cast<const F4*>(wp + c * Pack::kNct * Pack::kGroup);
#pragma unroll
            for (int g = 0; g < Pack::kGroup / 4; ++g) {
                const F4 t = w4[g];
                wv[4 * g + 0] = t.x; wv[4 * g + 1] = t.y; wv[4 * g + 2] = t.z; wv[4 * g + 3] = t.w;
            }
#pragma unroll
            for (int o = 0; o < TC; ++o) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (K == 3) {
                        acc[o][j] = fmaf(wv[0 * TC + o], x[j], acc[o][j]);
                        acc[o][j] = fmaf(wv[1 * TC + o], x[j + 1], acc[o][j]);
                        acc[o][j] = fmaf(wv[2 * TC + o], x[j + 2], acc[o][j]);
                    } else {
                        acc[o][j] = fmaf(wv[o], x[j + 1], acc[o][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < TC; ++o) {
            const int ch = ct * TC + o;
            float v0 = fmaxf(acc[o][0], 0.f), v1 = fmaxf(acc[o][1], 0.f);
            float v2 = fmaxf(acc[o][2], 0.f), v3 = fmaxf(acc[o][3], 0.f);
            float sc = 1.f, sh = 0.f;
            if (BN) {
                sc = DBN_LDG(bn_scale + ch);
                sh = DBN_LDG(bn_shift + ch);
            }
            float* orow = out + ch * kRsOut;
            if (POOL) {
                float p0 = fmaxf(v0, v1), p1 = fmaxf(v2, v3);
                if (BN) { p0 = fmaf(sc, p0, sh); p1 = fmaf(sc, p1, sh); }
                orow[4 + 2 * chunk] = p0;
                orow[4 + 2 * chunk + 1] = p1;
            } else {
                if (BN) {
                    v0 = fmaf(sc, v0, sh); v1 = fmaf(sc, v1, sh);
                    v2 = fmaf(sc, v2, sh); v3 = fmaf(sc, v3, sh);
                }
                *reinterpret_cast<F4*>(orow + 4 + 4 * chunk) = F4{v0, v1, v2, v3};
            }
            if (chunk == 0) orow[3] = 0.f;
            if (chunk == kChunks - 1) orow[4 + kLout] = 0.f;
        }
    }
}

// conv1d_1: Cin=1, k=3, stride 2, TF SAME on even length = pad right only (Appendix B.1):
// y[i] = b + w0*x[2i] + w1*x[2i+1] + w2*x[2i+2], x[1024] = 0.  + ReLU + BatchNorm_1.
// in: [1][1024+8] (col 4.., col 1028 = 0), out: [48][512+8], w packed as ConvPack<1,48,3,12>.
DBN_HD void conv1_s2(int tid, const float* in, float* out, const float* w, const float* bias,
                     const float* bn_scale, const float* bn_shift) {
    constexpr int TC = 12;
    using Pack = ConvPack<1, 48, 3, TC>;
    constexpr int kChunks = 128;
    for (int task = tid; task < kChunks * Pack::kNct; task += kThreads) {
        const int chunk = task % kChunks;
        const int ct = task / kChunks;
        const float* xp = in + 4 + 8 * chunk;
        const F4 a = *reinterpret_cast<const F4*>(xp);
        const F4 b = *reinterpret_cast<const F4*>(xp + 4);
        const float x[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, xp[8]};
        const float* wp = w + ct * Pack::kGroup;
#pragma unroll
        for (int o = 0; o < TC; ++o) {
            const int ch = ct * TC + o;
            const float bs = DBN_LDG(bias + ch);
            const float sc = DBN_LDG(bn_scale + ch), sh = DBN_LDG(bn_shift + ch);
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float acc = bs;
                acc = fmaf(wp[0 * TC + o], x[2 * j], acc);
                acc = fmaf(wp[1 * TC + o], x[2 * j + 1], acc);
                acc = fmaf(wp[2 * TC + o], x[2 * j + 2], acc);
                v[j] = fmaf(sc, fmaxf(acc, 0.f), sh);
            }
            float* orow = out + ch * 520;
            *reinterpret_cast<F4*>(orow + 4 + 4 * chunk) = F4{v[0], v[1], v[2], v[3]};
            if (chunk == 0) orow[3] = 0.f;
            if (chunk == kChunks - 1) orow[4 + 512] = 0.f;
        }
    }
}

// AveragePooling1D(3, stride 1, 'same'): TF divides by the number of in-range taps (Appendix B.3).
// in/out: [48][64+8]; relies on the zero halo of `in`.
DBN_HD void avgpool3(int tid, const float* in, float* out) {
    constexpr int L = 64, RS = L + 8;
    for (int i = tid; i < 48 * L; i += kThreads) {
        const int c = i / L, p = i % L;
        const float* r = in + c * RS + 4 + p;
        const float s = r[-1] + r[0] + r[1];
        out[c * RS + 4 + p] = (p == 0 || p == L - 1) ? s / 2.0f : s / 3.0f;
        if (p == 0) out[c * RS + 3] = 0.f;
        if (p == L - 1) out[c * RS + 4 + L] = 0.f;
    }
}

// conv1d_17: 192 -> 48, k=3, stride 2 (pad right only), L 32 -> 16, + ReLU + BatchNorm_6.
// in: [192][32+8]; out: [48][16+8]; w: global, Keras layout [3][192][48].
DBN_HD void conv17_s2(int tid, const float* in, float* out, const float* w, const float* bias,
                      const float* bn_scale, const float* bn_shift) {
    for (int task = tid; task < 48 * 16; task += kThreads) {
        const int o = task % 48, p = task / 48;
        float acc0 = DBN_LDG(bias + o), acc1 = 0.f, acc2 = 0.f;
        const float* xp = in + 4 + 2 * p;
#pragma unroll 8
        for (int c = 0; c < 192; ++c) {
            acc0 = fmaf(DBN_LDG(w + (0 * 192 + c) * 48 + o), xp[c * 40 + 0], acc0);
            acc1 = fmaf(DBN_LDG(w + (1 * 192 + c) * 48 + o), xp[c * 40 + 1], acc1);
            acc2 = fmaf(DBN_LDG(w + (2 * 192 + c) * 48 + o), xp[c * 40 + 2], acc2);
        }
        const float v = fmaxf(acc0 + acc1 + acc2, 0.f);
        float* orow = out + o * 24;
        orow[4 + p] = fmaf(DBN_LDG(bn_scale + o), v, DBN_LDG(bn_shift + o));
        if (p == 0) orow[3] = 0.f;
        if (p == 15) orow[4 + 16] = 0.f;
    }
}

// The whole network for one window.  smem: kFp32SmemFloats floats, 16-byte aligned.  On entry the
// normalised window must be in bufB row 0 ([1][1024+8], data at column 4, column 1028 zero).
DBN_HD void fp32_forward_window(const Fp32Net& net, float* smem, float* probs_out) {
    float* A = smem;
    float* B = smem + kBufFloats;
    float* W = smem + 2 * kBufFloats;
    float* S = smem + 2 * kBufFloats + kWbufFloats;
    const float* gw = net.w;
    const Fp32Layout& ly = net.lay;
#define BIAS(i) (gw + ly.bias[i])
#define BNS(i) (gw + ly.bn_scale[i])
#define BNH(i) (gw + ly.bn_shift[i])

    // conv1d_1 + BN1: B(input) -> A[48][512]
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[1], ConvPack<1, 48, 3, 12>::kFloats));
    DBN_PHASE(conv1_s2(tid, B, A, W, BIAS(1), BNS(1), BNH(1)));
    // conv1d_2..4 (+MaxPool +BN2): A -> B -> A -> B[48][256]
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[2], ConvPack<48, 48, 3, 12>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 512, 12, false, false>(tid, 0, kThreads, A, B, W, BIAS(2), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[3], ConvPack<48, 48, 3, 12>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 512, 12, false, false>(tid, 0, kThreads, B, A, W, BIAS(3), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[4], ConvPack<48, 48, 3, 12>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 512, 12, true, true>(tid, 0, kThreads, A, B, W, BIAS(4), BNS(2), BNH(2)));
    // From here on every activation tensor has L <= 256, so the upper part of buffer A is free:
    // the (padded) packed weights of the later layers are staged there (up to 9216 floats).
    W = A + kLateWOffset;
    // conv1d_5 (1x1 bottleneck), conv1d_6, conv1d_7 (+MaxPool +BN3): B -> A[16][256] -> B -> A[48][128]
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[5], ConvPack<48, 16, 1, 2>::kFloats));
    DBN_PHASE(conv_s1<48, 16, 1, 256, 2, false, false>(tid, 0, kThreads, B, A, W, BIAS(5), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[6], ConvPack<16, 48, 3, 6>::kFloats));
    DBN_PHASE(conv_s1<16, 48, 3, 256, 6, false, false>(tid, 0, kThreads, A, B, W, BIAS(6), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[7], ConvPack<48, 48, 3, 6>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 256, 6, true, true>(tid, 0, kThreads, B, A, W, BIAS(7), BNS(3), BNH(3)));
    // conv1d_8, conv1d_9 (+MaxPool +BN4): A -> B[48][128] -> A[48][64] =: X
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[8], ConvPack<48, 48, 3, 3>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 128, 3, false, false>(tid, 0, kThreads, A, B, W, BIAS(8), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[9], ConvPack<48, 48, 3, 3>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 128, 3, true, true>(tid, 0, kThreads, B, A, W, BIAS(9), BNS(4), BNH(4)));

    // Inception block on X = A[48][64].  Sub-buffers (row stride 72): P = avgpool(X), T12, T14
    // (16 ch), T15 (48 ch) live in A after X; the concatenated, pooled, BN5-normalised output
    // Y[192][32+8] is written to B.  Concat order [conv10, conv11, conv13, conv16]
    // (network_architecture.py:68).
    float* X = A;
    float* P = A + 48 * 72;
    float* T12 = P + 48 * 72;
    float* T14 = T12 + 16 * 72;
    float* T15 = T14 + 16 * 72;
    static_assert(2 * 48 * 72 + 2 * 16 * 72 + 48 * 72 <= kLateWOffset, "inception scratch overlaps weights");
    float* Y = B;
    constexpr int kW1 = ConvPack<48, 48, 1, 3>::kFloats;   // conv10 / conv11
    constexpr int kW2 = ConvPack<48, 16, 1, 1>::kFloats;   // conv12 / conv14
    constexpr int kW3 = ConvPack<16, 48, 3, 3>::kFloats;   // conv13 / conv15
    DBN_PHASE(avgpool3(tid, X, P); stage_weights(tid, W, gw + ly.conv[10], kW1);
              stage_weights(tid, W + kW1, gw + ly.conv[11], kW1));
    DBN_PHASE(conv_s1<48, 48, 1, 64, 3, true, true>(tid, 0, 256, P, Y + 0 * 40, W, BIAS(10), BNS(5) + 0, BNH(5) + 0);
              conv_s1<48, 48, 1, 64, 3, true, true>(tid, 256, 256, X, Y + 48 * 40, W + kW1, BIAS(11), BNS(5) + 48, BNH(5) + 48));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[12], kW2); stage_weights(tid, W + kW2, gw + ly.conv[14], kW2));
    DBN_PHASE(conv_s1<48, 16, 1, 64, 1, false, false>(tid, 0, 256, X, T12, W, BIAS(12), nullptr, nullptr);
              conv_s1<48, 16, 1, 64, 1, false, false>(tid, 256, 256, X, T14, W + kW2, BIAS(14), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[13], kW3); stage_weights(tid, W + kW3, gw + ly.conv[15], kW3));
    DBN_PHASE(conv_s1<16, 48, 3, 64, 3, true, true>(tid, 0, 256, T12, Y + 96 * 40, W, BIAS(13), BNS(5) + 96, BNH(5) + 96);
              conv_s1<16, 48, 3, 64, 3, false, false>(tid, 256, 256, T14, T15, W + kW3, BIAS(15), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[16], ConvPack<48, 48, 3, 3>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 64, 3, true, true>(tid, 0, kThreads, T15, Y + 144 * 40, W, BIAS(16), BNS(5) + 144, BNH(5) + 144));

    // conv1d_17 (+BN6): Y[192][32] -> A[48][16]; conv1d_18: A -> B[48][16]; conv1d_19 (+MaxPool +BN7): B -> A[48][8]
    DBN_PHASE(conv17_s2(tid, Y, A, gw + ly.conv17, BIAS(17), BNS(6), BNH(6));
              stage_weights(tid, W, gw + ly.conv[18], ConvPack<48, 48, 3, 1>::kFloats));
    // (Y lives in B and is dead after conv17; conv18 writes B only after the barrier above.)
    DBN_PHASE(conv_s1<48, 48, 3, 16, 1, false, false>(tid, 0, kThreads, A, B, W, BIAS(18), nullptr, nullptr));
    DBN_PHASE(stage_weights(tid, W, gw + ly.conv[19], ConvPack<48, 48, 3, 1>::kFloats));
    DBN_PHASE(conv_s1<48, 48, 3, 16, 1, true, true>(tid, 0, kThreads, B, A, W, BIAS(19), BNS(7), BNH(7)));

    // Head: conv1d_20 (48 -> n_classes, k=1) + ReLU -> S[class][8]; GlobalAveragePooling1D over the
    // 8 positions; softmax with max subtraction (network_architecture.py:89-91).
    const int nc = net.n_classes;
    DBN_PHASE(
        for (int task = tid; task < nc * 8; task += kThreads) {
            const int o = task / 8, p = task % 8;
            float acc = DBN_LDG(BIAS(20) + o);
            for (int c = 0; c < 48; ++c)
                acc = fmaf(DBN_LDG(gw + ly.conv20 + c * nc + o), A[c * 16 + 4 + p], acc);
            S[o * 8 + p] = fmaxf(acc, 0.f);
        });
    DBN_PHASE(
        if (tid < nc) {
            float s = 0.f;
            for (int p = 0; p < 8; ++p) s += S[tid * 8 + p];
            S[8 * kMaxClasses + tid] = s / 8.0f;
        });
    DBN_PHASE(
        if (tid < nc) {
            const float* lg = S + 8 * kMaxClasses;
            float m = lg[0];
            for (int j = 1; j < nc; ++j) m = fmaxf(m, lg[j]);
            float den = 0.f;
            for (int j = 0; j < nc; ++j) den += expf(lg[j] - m);
            probs_out[tid] = expf(lg[tid] - m) / den;
        });
#undef BIAS
#undef BNS
#undef BNH
}

// ---------------------------------------------------------------------------------------------
// Window staging
// ---------------------------------------------------------------------------------------------

// predict seam: copy an already-normalised float window into bufB row 0.
template <typename T>
DBN_HD void stage_window_from_values(int tid, const T* x, float* row) {
    for (int i = tid; i < kInputSize; i += kThreads) row[4 + i] = static_cast<float>(x[i]);
    if (tid < 4) { row[tid] = 0.f; row[4 + kInputSize + tid] = 0.f; }
}

// Geometry of one scan window inside a read's scan region (classify.py:337-349).  `region_len` =
// number of samples of the read available in the scan region (first/last min(len, scan+step)).
// Returns piece = region[a, b) and where it goes inside the 1024-sample window (classify.py:352-357).
struct WindowGeom {
    int a, n, dst;
};
DBN_HD WindowGeom window_geometry(int region_len, int step, int side) {
    const int sig_start = step * (kInputSize / 2);
    const int sig_end = sig_start + kInputSize;
    WindowGeom g;
    if (side == 0) {
        const int a = sig_start < region_len ? sig_start : region_len;
        const int b = sig_end < region_len ? sig_end : region_len;
        g.a = a; g.n = b - a; g.dst = 0;                 // zero pad on the right
    } else {
        const int a = region_len - sig_end > 0 ? region_len - sig_end : 0;
        const int b = region_len - sig_start > 0 ? region_len - sig_start : 0;
        g.a = a; g.n = b - a; g.dst = kInputSize - g.n;  // zero pad on the left
    }
    return g;
}

// z-score of trim_signal.py:61-69 for int16 samples given exact integer sums:
// mean = s1/n, population variance = (n*s2 - s1^2)/n^2 (exact in int64 for n <= 1024).
DBN_HD void zscore_params(long long s1, long long s2, int n, double* mean, double* stdev) {
    const double dn = static_cast<double>(n);
    *mean = static_cast<double>(s1) / dn;
    const long long num = static_cast<long long>(n) * s2 - s1 * s1;
    *stdev = sqrt(static_cast<double>(num)) / dn;
}

// Normalise + pad one scan window of an int16 scan region into bufB row 0 (classify.py:342-358 with
// trim_signal.py:61-69 inlined).  `red` = (2*kThreads + 64) long longs of scratch.  Three phases.
DBN_HD void window_partial_sums(int tid, const int16_t* region, WindowGeom g, long long* red) {
    long long s1 = 0, s2 = 0;
    for (int i = tid; i < g.n; i += kThreads) {
        const long long v = region[g.a + i];
        s1 += v;
        s2 += v * v;
    }
    red[tid] = s1;
    red[kThreads + tid] = s2;
}
DBN_HD void window_reduce(int tid, long long* red) {
    if (tid < 64) {
        const long long* src = red + (tid < 32 ? 0 : kThreads) + (tid & 31) * (kThreads / 32);
        long long s = 0;
        for (int i = 0; i < kThreads / 32; ++i) s += src[i];
        red[2 * kThreads + tid] = s;
    }
}
DBN_HD void window_normalise(int tid, const int16_t* region, WindowGeom g, const long long* red,
                             float* row) {
    long long s1 = 0, s2 = 0;
    for (int i = 0; i < 32; ++i) { s1 += red[2 * kThreads + i]; s2 += red[2 * kThreads + 32 + i]; }
    double mean = 0.0, stdev = 0.0;
    if (g.n > 0) zscore_params(s1, s2, g.n, &mean, &stdev);
    for (int i = tid; i < kInputSize; i += kThreads) {
        const int k = i - g.dst;
        float v = 0.f;
        if (k >= 0 && k < g.n) {
            const double d = static_cast<double>(region[g.a + k]) - mean;
            v = static_cast<float>(stdev > 0.0 ? d / stdev : d);
        }
        row[4 + i] = v;
    }
    if (tid < 4) { row[tid] = 0.f; row[4 + kInputSize + tid] = 0.f; }
}

}  // namespace dbn
