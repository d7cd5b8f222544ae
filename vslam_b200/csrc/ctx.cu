// ctx.cu — context lifetime, error string, stream selection, profiling brackets.
#include "common.cuh"

namespace vb {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace vb

void vb_ctx::prof_begin(const char *name) {
    if (!profile) return;
    vb::ProfEntry &e = prof[name];
    if (!e.a) {
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
    }
    cudaEventRecord(e.a, stream);
}
void vb_ctx::prof_end(const char *name) {
    if (!profile) return;
    vb::ProfEntry &e = prof[name];
    cudaEventRecord(e.b, stream);
    e.used = true;
}

extern "C" {

int vb_version(void) { return 100; }
const char *vb_last_error(void) { return vb::g_err; }

int vb_create(int device, vb_ctx **out) {
    VB_REQUIRE(out != nullptr, VB_ERR_INVALID, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        vb::set_error("no CUDA device (%s): this library has no CPU fallback", cudaGetErrorString(e));
        return VB_ERR_CUDA;
    }
    VB_REQUIRE(device >= 0 && device < count, VB_ERR_INVALID, "device index out of range");
    VB_CUDA(cudaSetDevice(device));
    vb_ctx *c = new vb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    VB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    VB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    *out = c;
    return VB_OK;
}

int vb_destroy(vb_ctx *c) {
    if (!c) return VB_OK;
    if (c->twin) {
        vb_destroy(c->twin);
        c->twin = nullptr;
    }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    vb::pairs_stream_release(c);
    for (auto &b : c->ws) b.release();
    for (auto &b : c->pin) b.release();
    for (auto &kv : c->prof) {
        if (kv.second.a) cudaEventDestroy(kv.second.a);
        if (kv.second.b) cudaEventDestroy(kv.second.b);
    }
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return VB_OK;
}

// Option names. The first group selects between equivalent code paths (every one returns the same bits; the GPU tests
// force each of them); the second group exists only in a -DVB_TUNING build (measurement aids: timing floors, schedules).
static const char *const kOptions[] = {
    "hamming_tc", "hamming_fp4", "tc_fix8", "hamming_qpt", "l2_tc", "ransac_lazy", "ransac_prune", "prune_first_chunks",
    "prune_first16", "prune_growth16", "prune_rounds", "prune_item_chunks", "count_packed", "score_packed", "kd_lanes_per_query", "tc_drain", "tc_svc_hi", "tc_issuers", "tc_fix_skip", "pairs_overlap",
#ifdef VB_TUNING
    "tc_dbg", "prune_ctas_per_sm", "pairs_twin", "pairs_split",
#endif
};

int vb_set_option(vb_ctx *c, const char *name, long long value) {
    VB_REQUIRE(c && name, VB_ERR_INVALID, "NULL argument");
    for (const char *k : kOptions)
        if (!strcmp(k, name)) {
            c->opts[name] = value;
            return VB_OK;
        }
    vb::set_error("vb_set_option: unknown option '%s'", name);
    return VB_ERR_INVALID;
}

int vb_reset_options(vb_ctx *c) {
    VB_REQUIRE(c != nullptr, VB_ERR_INVALID, "ctx is NULL");
    c->opts.clear();
    return VB_OK;
}

int vb_set_stream(vb_ctx *c, void *s) {
    VB_REQUIRE(c != nullptr, VB_ERR_INVALID, "ctx is NULL");
    c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->own_stream;
    return VB_OK;
}

int vb_synchronize(vb_ctx *c) {
    VB_REQUIRE(c != nullptr, VB_ERR_INVALID, "ctx is NULL");
    VB_CUDA(cudaStreamSynchronize(c->stream));
    return VB_OK;
}

uint64_t vb_launch_count(const vb_ctx *c) { return c ? c->launches + (c->twin ? c->twin->launches : 0) : 0; }

int vb_profile_enable(vb_ctx *c, int on) {
    VB_REQUIRE(c != nullptr, VB_ERR_INVALID, "ctx is NULL");
    c->profile = on != 0;
    return VB_OK;
}

float vb_profile_last_ms(vb_ctx *c, const char *name) {
    if (!c || !name) return -1.f;
    auto it = c->prof.find(name);
    if (it == c->prof.end() || !it->second.used) return -1.f;
    if (cudaEventSynchronize(it->second.b) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, it->second.a, it->second.b) != cudaSuccess) return -1.f;
    return ms;
}

}  // extern "C"
