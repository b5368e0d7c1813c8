// Host-side weight handling: DBNW blob parsing (format: deepbinner_b200/weights.py) and packing of
// the parameters into the layouts the engines read.  Pure C++ (no CUDA) so the CPU emulation test
// can share it.  The blob carries the tensors of the graph built by reference
// network_architecture.py:18-95; conv kernels are in Keras layout [k][Cin][Cout].
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "dbn_fp32_net.cuh"

namespace dbn {

struct BlobTensor {
    const float* data = nullptr;
    int ndim = 0;
    int dims[3] = {0, 0, 0};
    size_t count = 0;
};

struct Blob {
    int input_size = 0;
    int n_classes = 0;
    std::vector<uint8_t> bytes;  // owning copy
    std::map<std::string, BlobTensor> tensors;

    const BlobTensor* find(const std::string& name) const {
        auto it = tensors.find(name);
        return it == tensors.end() ? nullptr : &it->second;
    }
};

struct ConvSpec {
    int k, stride, cin, cout;
};
// conv1d_1 .. conv1d_20 (index 1..20); cout of conv1d_20 is the class count (0 here).
static const ConvSpec kConvSpecs[21] = {
    {0, 0, 0, 0},   {3, 2, 1, 48},  {3, 1, 48, 48}, {3, 1, 48, 48}, {3, 1, 48, 48}, {1, 1, 48, 16},
    {3, 1, 16, 48}, {3, 1, 48, 48}, {3, 1, 48, 48}, {3, 1, 48, 48}, {1, 1, 48, 48}, {1, 1, 48, 48},
    {1, 1, 48, 16}, {3, 1, 16, 48}, {1, 1, 48, 16}, {3, 1, 16, 48}, {3, 1, 48, 48}, {3, 2, 192, 48},
    {3, 1, 48, 48}, {3, 1, 48, 48}, {1, 1, 48, 0}};
static const int kBnChannels[8] = {0, 48, 48, 48, 48, 192, 48, 48};
static const double kBnEpsilon = 1e-3;

// Returns "" on success, else an error message.
inline std::string parse_blob(const void* data, size_t n, Blob* out) {
    struct Header {
        char magic[8];
        uint32_t version, input_size, n_classes, n_tensors;
    };
    struct Entry {
        char name[48];
        uint32_t ndim, dims[3];
        uint64_t offset, count;
    };
    static_assert(sizeof(Header) == 24 && sizeof(Entry) == 80, "blob struct packing");
    if (data == nullptr || n < sizeof(Header)) return "weight blob too small";
    out->bytes.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + n);
    const uint8_t* base = out->bytes.data();
    Header h;
    std::memcpy(&h, base, sizeof h);
    if (std::memcmp(h.magic, "DBNWGT1\0", 8) != 0) return "not a DBNW weight blob (bad magic)";
    if (h.version != 1) return "unsupported DBNW version";
    const size_t table_end = sizeof(Header) + static_cast<size_t>(h.n_tensors) * sizeof(Entry);
    if (h.n_tensors > 4096 || table_end > n) return "corrupt DBNW tensor table";
    if ((table_end % 4) != 0) return "misaligned DBNW payload";
    out->input_size = static_cast<int>(h.input_size);
    out->n_classes = static_cast<int>(h.n_classes);
    const size_t payload_floats = (n - table_end) / 4;
    for (uint32_t i = 0; i < h.n_tensors; ++i) {
        Entry e;
        std::memcpy(&e, base + sizeof(Header) + i * sizeof(Entry), sizeof e);
        // (written so that a crafted offset / count / dims cannot wrap around the comparisons)
        if (e.ndim > 3 || e.offset > payload_floats || e.count > payload_floats - e.offset)
            return "corrupt DBNW tensor entry";
        uint64_t prod = 1;
        for (uint32_t d = 0; d < e.ndim; ++d) {
            if (e.dims[d] > (1u << 24)) return "DBNW tensor entry: implausible dimension";
            prod *= e.dims[d];                       // <= 2^72 would overflow: check after every factor
            if (prod > payload_floats) return "DBNW tensor entry: dims do not match count";
        }
        if (prod != e.count) return "DBNW tensor entry: dims do not match count";
        BlobTensor t;
        t.data = reinterpret_cast<const float*>(base + table_end) + e.offset;
        t.ndim = static_cast<int>(e.ndim);
        for (int d = 0; d < 3; ++d) t.dims[d] = static_cast<int>(e.dims[d]);
        t.count = e.count;
        char nm[49];
        std::memcpy(nm, e.name, 48);
        nm[48] = 0;
        out->tensors[nm] = t;
    }
    // topology / shape validation
    if (out->input_size != kInputSize) return "only models with input size 1024 are supported";
    if (out->n_classes < 2 || out->n_classes > kMaxClasses) return "unsupported class count";
    for (int i = 1; i <= 20; ++i) {
        const ConvSpec& s = kConvSpecs[i];
        const int cout = s.cout ? s.cout : out->n_classes;
        const std::string nm = "conv1d_" + std::to_string(i);
        const BlobTensor* k = out->find(nm + "/kernel");
        const BlobTensor* b = out->find(nm + "/bias");
        if (!k || !b) return "missing tensor " + nm;
        if (k->ndim != 3 || k->dims[0] != s.k || k->dims[1] != s.cin || k->dims[2] != cout ||
            b->ndim != 1 || b->dims[0] != cout)
            return "unexpected shape for " + nm;
    }
    for (int i = 1; i <= 7; ++i) {
        const std::string nm = "batch_normalization_" + std::to_string(i);
        for (const char* w : {"gamma", "beta", "moving_mean", "moving_variance"}) {
            const BlobTensor* t = out->find(nm + "/" + w);
            if (!t || t->ndim != 1 || t->dims[0] != kBnChannels[i]) return "bad tensor " + nm;
        }
    }
    return "";
}

// Folded BatchNorm (inference): y = scale*x + shift, scale = gamma/sqrt(var+eps), shift = beta -
// mean*scale, computed in double and rounded once (Appendix B.5).
inline void fold_bn(const Blob& b, int i, std::vector<float>* scale, std::vector<float>* shift) {
    const std::string nm = "batch_normalization_" + std::to_string(i);
    const float* g = b.find(nm + "/gamma")->data;
    const float* be = b.find(nm + "/beta")->data;
    const float* mu = b.find(nm + "/moving_mean")->data;
    const float* var = b.find(nm + "/moving_variance")->data;
    const int n = kBnChannels[i];
    scale->resize(n);
    shift->resize(n);
    for (int c = 0; c < n; ++c) {
        const double sc = static_cast<double>(g[c]) / std::sqrt(static_cast<double>(var[c]) + kBnEpsilon);
        (*scale)[c] = static_cast<float>(sc);
        (*shift)[c] = static_cast<float>(static_cast<double>(be[c]) - static_cast<double>(mu[c]) * sc);
    }
}

// Thread-tile width (output channels per thread) used by the fp32 engine for each stride-1 conv;
// must match the template arguments in fp32_forward_window().
static const int kFp32Tc[21] = {0, 12, 12, 12, 12, 2, 6, 6, 3, 3, 3, 3, 1, 3, 1, 3, 3, 0, 1, 1, 0};

inline void pack_fp32(const Blob& b, std::vector<float>* out, Fp32Layout* lay) {
    out->clear();
    auto align4 = [&]() { while (out->size() % 4) out->push_back(0.f); };
    for (int i = 1; i <= 20; ++i) {
        const ConvSpec& s = kConvSpecs[i];
        const std::string nm = "conv1d_" + std::to_string(i);
        const float* k = b.find(nm + "/kernel")->data;
        const int cout = s.cout ? s.cout : b.n_classes;
        align4();
        if (i == 17) {
            lay->conv17 = static_cast<int>(out->size());
            out->insert(out->end(), k, k + 3 * 192 * 48);
        } else if (i == 20) {
            lay->conv20 = static_cast<int>(out->size());
            out->insert(out->end(), k, k + 48 * cout);
        } else {
            const int tc = kFp32Tc[i];
            const int group = round_up4(s.k * tc);
            const int nct = cout / tc;
            lay->conv[i] = static_cast<int>(out->size());
            out->resize(out->size() + static_cast<size_t>(s.cin) * nct * group, 0.f);
            float* dst = out->data() + lay->conv[i];
            for (int c = 0; c < s.cin; ++c)
                for (int ct = 0; ct < nct; ++ct)
                    for (int t = 0; t < s.k; ++t)
                        for (int o = 0; o < tc; ++o)
                            dst[(c * nct + ct) * group + t * tc + o] =
                                k[(t * s.cin + c) * cout + ct * tc + o];
        }
        align4();
        lay->bias[i] = static_cast<int>(out->size());
        const float* bias = b.find(nm + "/bias")->data;
        out->insert(out->end(), bias, bias + cout);
    }
    for (int i = 1; i <= 7; ++i) {
        std::vector<float> sc, sh;
        fold_bn(b, i, &sc, &sh);
        align4();
        lay->bn_scale[i] = static_cast<int>(out->size());
        out->insert(out->end(), sc.begin(), sc.end());
        align4();
        lay->bn_shift[i] = static_cast<int>(out->size());
        out->insert(out->end(), sh.begin(), sh.end());
    }
    align4();
    lay->total = static_cast<int>(out->size());
}

}  // namespace dbn
