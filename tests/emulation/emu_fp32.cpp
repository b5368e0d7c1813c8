// TEST INFRASTRUCTURE: host build of the fp32 engine's phase code (dbn_fp32_net.cuh compiled with
// g++, every phase looped over all 512 thread ids) so the kernel's indexing, packing and layer
// semantics can be checked against the oracle on a machine without a GPU.  Never used by the product.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../deepbinner_b200/csrc/dbn_weights.h"

using namespace dbn;

extern "C" int emu_predict(const void* blob, size_t blob_bytes, const float* x, int n, float* probs) {
    Blob b;
    std::string err = parse_blob(blob, blob_bytes, &b);
    if (!err.empty()) {
        std::fprintf(stderr, "emu: %s\n", err.c_str());
        return -1;
    }
    std::vector<float> packed;
    Fp32Net net;
    pack_fp32(b, &packed, &net.lay);
    net.w = packed.data();
    net.n_classes = b.n_classes;
    float* smem = static_cast<float*>(aligned_alloc(16, sizeof(float) * kFp32SmemFloats));
    for (int i = 0; i < n; ++i) {
        // poison to catch reads of unwritten shared memory
        for (int j = 0; j < kFp32SmemFloats; ++j) smem[j] = 1e30f;
        float* B = smem + kBufFloats;
        DBN_PHASE(stage_window_from_values(tid, x + static_cast<size_t>(i) * kInputSize, B));
        fp32_forward_window(net, smem, probs + static_cast<size_t>(i) * b.n_classes);
    }
    free(smem);
    return 0;
}

extern "C" void emu_stage_window(const int16_t* region, int region_len, int step, int side,
                                 float* row_out /* 1024 */) {
    std::vector<long long> red(2 * kThreads + 64);
    std::vector<float> row(kInputSize + 8, 1e30f);
    const WindowGeom g = window_geometry(region_len, step, side);
    DBN_PHASE(window_partial_sums(tid, region, g, red.data()));
    DBN_PHASE(window_reduce(tid, red.data()));
    DBN_PHASE(window_normalise(tid, region, g, red.data(), row.data()));
    for (int i = 0; i < kInputSize; ++i) row_out[i] = row[4 + i];
}
