// tcgen05 / TMEM tensor-core engine (placeholder until the kernel lands: reports "unavailable" so
// the library runs on the fp32 CUDA-core engine).
#include "dbn_engine.h"

namespace dbn {

TcEngine* tc_create(const Blob&, int) { return nullptr; }
void tc_destroy(TcEngine*) {}
int tc_predict(TcEngine*, const float*, int64_t, float*, cudaStream_t) {
    return fail(-1, "tcgen05 engine not built");
}
int tc_call_windows(TcEngine*, const int16_t*, const int64_t*, int, int, int, float*, cudaStream_t) {
    return fail(-1, "tcgen05 engine not built");
}

}  // namespace dbn
