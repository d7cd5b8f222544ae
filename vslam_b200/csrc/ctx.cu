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

uint64_t vb_launch_count(const vb_ctx *c) { return c ? c->launches : 0; }

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
