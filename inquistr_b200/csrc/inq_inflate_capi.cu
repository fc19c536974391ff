// inq_inflate_capi.cu -- extern "C" entry point of include/inqbgzf.h (part of libinqcall.so)
#include "../../include/inqbgzf.h"
#include "../../include/inqcall.h"
#include "inq_inflate.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_z_error;

int zfail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_z_error = buf;
    return code;
}

struct ZGuard {
    void *d_comp = nullptr, *d_out = nullptr, *d_blocks = nullptr, *d_status = nullptr;
    cudaStream_t s = nullptr;
    cudaEvent_t e[4] = {};
    ~ZGuard()
    {
        if (d_comp) cudaFree(d_comp);
        if (d_out) cudaFree(d_out);
        if (d_blocks) cudaFree(d_blocks);
        if (d_status) cudaFree(d_status);
        for (auto x : e)
            if (x) cudaEventDestroy(x);
        if (s) cudaStreamDestroy(s);
    }
};

#define Z_TRY(call)                                                                             \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return zfail(e_ == cudaErrorMemoryAllocation ? INQ_ERR_NOMEM : INQ_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

extern "C" {

const char *inq_bgzf_last_error(void) { return g_z_error.c_str(); }

int inq_bgzf_inflate(int device, const uint8_t *comp, uint64_t comp_bytes, const inq_zblock *blocks, uint32_t n_blocks,
                     uint8_t *out, uint64_t out_bytes, uint32_t *status, float *ms_h2d, float *ms_kernel, float *ms_d2h)
{
    using namespace inqz;
    static_assert(sizeof(inq_zblock) == sizeof(BlockDesc), "descriptor layout");
    if (ms_h2d) *ms_h2d = 0.f;
    if (ms_kernel) *ms_kernel = 0.f;
    if (ms_d2h) *ms_d2h = 0.f;
    if (n_blocks == 0) return INQ_OK;
    if (!comp || !blocks || !out || !status) return zfail(INQ_ERR_ARG, "inq_bgzf_inflate: NULL array");
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (blocks[b].in_off + blocks[b].in_len > comp_bytes || blocks[b].out_off + blocks[b].out_len > out_bytes)
            return zfail(INQ_ERR_ARG, "inq_bgzf_inflate: block %u lies outside the buffers", b);
        if (blocks[b].out_len > 65536u) return zfail(INQ_ERR_ARG, "inq_bgzf_inflate: block %u is larger than a BGZF block", b);
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return zfail(INQ_ERR_CUDA, "no CUDA device available; libinqcall has no CPU fallback");
    if (device < 0 || device >= n) return zfail(INQ_ERR_ARG, "device %d out of range [0,%d)", device, n);
    Z_TRY(cudaSetDevice(device));
    ZGuard g;
    Z_TRY(cudaStreamCreateWithFlags(&g.s, cudaStreamNonBlocking));
    for (auto &x : g.e) Z_TRY(cudaEventCreate(&x));
    Z_TRY(cudaMalloc(&g.d_comp, comp_bytes + 64));                         // the bit reader loads whole 8-byte words
    Z_TRY(cudaMalloc(&g.d_out, std::max<uint64_t>(out_bytes, 1) + 64));
    Z_TRY(cudaMalloc(&g.d_blocks, (size_t)n_blocks * sizeof(BlockDesc)));
    Z_TRY(cudaMalloc(&g.d_status, (size_t)n_blocks * sizeof(uint32_t)));
    Z_TRY(cudaEventRecord(g.e[0], g.s));
    Z_TRY(cudaMemsetAsync(static_cast<uint8_t *>(g.d_comp) + comp_bytes, 0, 64, g.s));
    Z_TRY(cudaMemcpyAsync(g.d_comp, comp, comp_bytes, cudaMemcpyHostToDevice, g.s));
    Z_TRY(cudaMemcpyAsync(g.d_blocks, blocks, (size_t)n_blocks * sizeof(BlockDesc), cudaMemcpyHostToDevice, g.s));
    Z_TRY(cudaEventRecord(g.e[1], g.s));
    int sms = 0;
    Z_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const size_t smem = sizeof(WarpSmem) * kWarpsPerCta;
    Z_TRY(cudaFuncSetAttribute(k_bgzf_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    Z_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bgzf_inflate, kWarpsPerCta * 32, smem));
    const unsigned grid = (unsigned)std::min<uint64_t>((n_blocks + kWarpsPerCta - 1) / kWarpsPerCta, (uint64_t)sms * std::max(per_sm, 1));
    k_bgzf_inflate<<<grid, kWarpsPerCta * 32, smem, g.s>>>(static_cast<const uint8_t *>(g.d_comp), static_cast<const BlockDesc *>(g.d_blocks), n_blocks,
                                                          static_cast<uint8_t *>(g.d_out), static_cast<uint32_t *>(g.d_status));
    Z_TRY(cudaGetLastError());
    Z_TRY(cudaEventRecord(g.e[2], g.s));
    Z_TRY(cudaMemcpyAsync(out, g.d_out, out_bytes, cudaMemcpyDeviceToHost, g.s));
    Z_TRY(cudaMemcpyAsync(status, g.d_status, (size_t)n_blocks * sizeof(uint32_t), cudaMemcpyDeviceToHost, g.s));
    Z_TRY(cudaEventRecord(g.e[3], g.s));
    Z_TRY(cudaStreamSynchronize(g.s));
    if (ms_h2d) cudaEventElapsedTime(ms_h2d, g.e[0], g.e[1]);
    if (ms_kernel) cudaEventElapsedTime(ms_kernel, g.e[1], g.e[2]);
    if (ms_d2h) cudaEventElapsedTime(ms_d2h, g.e[2], g.e[3]);
    return INQ_OK;
}

// ---- persistent engine ------------------------------------------------------------------------------------------
}  // extern "C"

struct inq_bgzf_engine {
    int device = 0;
    int sms = 0, per_sm = 1;
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint8_t *d_comp = nullptr, *d_out = nullptr;
    inqz::BlockDesc *d_blocks = nullptr, *h_blocks = nullptr;     // h_blocks: pinned staging of the rebased descriptors
    uint32_t *d_status = nullptr, *h_status = nullptr;
    uint64_t cap_comp = 0, cap_out = 0;
    uint32_t cap_blocks = 0;
};

extern "C" {

void inq_bgzf_engine_destroy(inq_bgzf_engine *eng)
{
    if (!eng) return;
    cudaSetDevice(eng->device);
    if (eng->s) cudaStreamSynchronize(eng->s);
    if (eng->d_comp) cudaFree(eng->d_comp);
    if (eng->d_out) cudaFree(eng->d_out);
    if (eng->d_blocks) cudaFree(eng->d_blocks);
    if (eng->d_status) cudaFree(eng->d_status);
    if (eng->h_blocks) cudaFreeHost(eng->h_blocks);
    if (eng->h_status) cudaFreeHost(eng->h_status);
    if (eng->e0) cudaEventDestroy(eng->e0);
    if (eng->e1) cudaEventDestroy(eng->e1);
    if (eng->s) cudaStreamDestroy(eng->s);
    delete eng;
}

int inq_bgzf_engine_create(int device, uint64_t max_comp_bytes, uint64_t max_out_bytes, uint32_t max_blocks, inq_bgzf_engine **out)
{
    using namespace inqz;
    if (!out) return zfail(INQ_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return zfail(INQ_ERR_CUDA, "no CUDA device available; libinqcall has no CPU fallback");
    if (device < 0 || device >= n) return zfail(INQ_ERR_ARG, "device %d out of range [0,%d)", device, n);
    Z_TRY(cudaSetDevice(device));
    inq_bgzf_engine *g = new inq_bgzf_engine();
    g->device = device;
    g->cap_comp = max_comp_bytes;
    g->cap_out = max_out_bytes;
    g->cap_blocks = max_blocks;
    auto bail = [&](const char *what, cudaError_t err) {
        const int rc = zfail(err == cudaErrorMemoryAllocation ? INQ_ERR_NOMEM : INQ_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        inq_bgzf_engine_destroy(g);
        return rc;
    };
    if ((e = cudaStreamCreateWithFlags(&g->s, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&g->e0)) != cudaSuccess || (e = cudaEventCreate(&g->e1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&g->d_comp, max_comp_bytes + 64)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&g->d_out, max_out_bytes + 64)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&g->d_blocks, (size_t)max_blocks * sizeof(BlockDesc))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&g->d_status, (size_t)max_blocks * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMallocHost(&g->h_blocks, (size_t)max_blocks * sizeof(BlockDesc))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost(&g->h_status, (size_t)max_blocks * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMemset(g->d_comp, 0, max_comp_bytes + 64)) != cudaSuccess) return bail("cudaMemset", e);
    if ((e = cudaDeviceGetAttribute(&g->sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail("cudaDeviceGetAttribute", e);
    const size_t smem = sizeof(WarpSmem) * kWarpsPerCta;
    if ((e = cudaFuncSetAttribute(k_bgzf_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g->per_sm, k_bgzf_inflate, kWarpsPerCta * 32, smem)) != cudaSuccess) return bail("occupancy", e);
    *out = g;
    return INQ_OK;
}

int inq_bgzf_engine_run(inq_bgzf_engine *g, const uint8_t *comp, uint64_t comp_bytes, const inq_zblock *blocks, uint32_t n_blocks,
                        uint8_t *out, uint32_t *status, float *ms_kernel)
{
    using namespace inqz;
    if (ms_kernel) *ms_kernel = 0.f;
    if (!g) return INQ_ERR_ARG;
    if (n_blocks == 0) return INQ_OK;
    if (!comp || !blocks || !out || !status) return zfail(INQ_ERR_ARG, "inq_bgzf_engine_run: NULL array");
    if (n_blocks > g->cap_blocks) return zfail(INQ_ERR_ARG, "inq_bgzf_engine_run: %u blocks, engine holds %u", n_blocks, g->cap_blocks);
    // the byte ranges of comp / out this run touches; descriptors are rebased onto the device buffers
    uint64_t in_lo = UINT64_MAX, in_hi = 0, out_lo = UINT64_MAX, out_hi = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (blocks[b].out_len > 65536u) return zfail(INQ_ERR_ARG, "inq_bgzf_engine_run: block %u is larger than a BGZF block", b);
        in_lo = std::min(in_lo, blocks[b].in_off);
        in_hi = std::max(in_hi, blocks[b].in_off + blocks[b].in_len);
        out_lo = std::min(out_lo, blocks[b].out_off);
        out_hi = std::max(out_hi, blocks[b].out_off + blocks[b].out_len);
    }
    if (in_hi > comp_bytes) return zfail(INQ_ERR_ARG, "inq_bgzf_engine_run: a block lies outside comp");
    in_lo &= ~7ull;                                                    // keep the 8-byte phase of the payload addresses
    if (in_hi - in_lo > g->cap_comp || out_hi - out_lo > g->cap_out) return zfail(INQ_ERR_ARG, "inq_bgzf_engine_run: run larger than the engine's buffers");
    Z_TRY(cudaSetDevice(g->device));
    for (uint32_t b = 0; b < n_blocks; ++b) {
        g->h_blocks[b].in_off = blocks[b].in_off - in_lo;
        g->h_blocks[b].out_off = blocks[b].out_off - out_lo;
        g->h_blocks[b].in_len = blocks[b].in_len;
        g->h_blocks[b].out_len = blocks[b].out_len;
    }
    Z_TRY(cudaMemcpyAsync(g->d_comp, comp + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, g->s));
    Z_TRY(cudaMemcpyAsync(g->d_blocks, g->h_blocks, (size_t)n_blocks * sizeof(BlockDesc), cudaMemcpyHostToDevice, g->s));
    Z_TRY(cudaEventRecord(g->e0, g->s));
    const unsigned grid = (unsigned)std::min<uint64_t>((n_blocks + kWarpsPerCta - 1) / kWarpsPerCta, (uint64_t)g->sms * std::max(g->per_sm, 1));
    k_bgzf_inflate<<<grid, kWarpsPerCta * 32, sizeof(WarpSmem) * kWarpsPerCta, g->s>>>(g->d_comp, g->d_blocks, n_blocks, g->d_out, g->d_status);
    Z_TRY(cudaGetLastError());
    Z_TRY(cudaEventRecord(g->e1, g->s));
    if (out_hi > out_lo) Z_TRY(cudaMemcpyAsync(out + out_lo, g->d_out, out_hi - out_lo, cudaMemcpyDeviceToHost, g->s));
    Z_TRY(cudaMemcpyAsync(g->h_status, g->d_status, (size_t)n_blocks * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->s));
    Z_TRY(cudaStreamSynchronize(g->s));
    memcpy(status, g->h_status, (size_t)n_blocks * sizeof(uint32_t));
    if (ms_kernel) cudaEventElapsedTime(ms_kernel, g->e0, g->e1);
    return INQ_OK;
}

int inq_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return INQ_ERR_ARG;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return INQ_OK; }      // (two engines met on one buffer)
    if (e != cudaSuccess) { cudaGetLastError(); return zfail(INQ_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    return INQ_OK;
}

int inq_host_unregister(void *p)
{
    if (!p) return INQ_OK;
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return INQ_ERR_CUDA; }
    return INQ_OK;
}

}  // extern "C"
