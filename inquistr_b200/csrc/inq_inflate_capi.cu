// inq_inflate_capi.cu -- extern "C" entry point of include/inqbgzf.h (part of libinqcall.so)
#include "../../include/inqbgzf.h"
#include "../../include/inqcall.h"
#include "inq_inflate.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
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

}  // extern "C"
