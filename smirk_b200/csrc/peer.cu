// Peer-mapped gather buffers: the multi-GPU all-gather of the final outputs (the reference itself has no
// multi-GPU code) done with copy-engine pushes over NVLink instead of a collective kernel.
//
// Why not only NCCL: the compute kernels of this library are persistent (148 CTAs striding over their work items, one per
// SM).  A collective kernel that occupies a few SMs while they run turns every concurrent persistent launch into two waves
// (measured on 8 x B200: 0.85 end-to-end efficiency with the NCCL all-gather in the timed region, 0.99 without it).  A push
// from the copy engines takes no SM: each rank owns one gather buffer [world][shard], maps every peer's buffer through CUDA
// IPC once, and after each batch copies its packed shard into slot `rank` of every peer's buffer on a communication stream.
#include "common.cuh"
#include <cstring>

extern "C" int smk_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    SMK_REQUIRE(ptr && handle64 && bytes > 0, "smk_peer_alloc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* p = nullptr;
    SMK_CHECK_CUDA(cudaMalloc(&p, bytes));           // set-up time only; a plain (not pooled) allocation so the handle's base is `p`
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); SMK_CHECK_CUDA(e); }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return 0;
}

extern "C" int smk_peer_free(void* ptr) {
    if (ptr) SMK_CHECK_CUDA(cudaFree(ptr));
    return 0;
}

// Maps another process's buffer (same node) into this process's current device context; peer access is enabled on demand.
extern "C" int smk_peer_open(const unsigned char* handle64, void** ptr) {
    SMK_REQUIRE(ptr && handle64, "smk_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SMK_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int smk_peer_close(void* ptr) {
    if (ptr) SMK_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

// dst may be local or peer-mapped; the copy runs on a copy engine, ordered on `stream`.
extern "C" int smk_peer_push(void* dst, const void* src, size_t bytes, void* stream) {
    SMK_REQUIRE(dst && src, "smk_peer_push: null pointer");
    if (bytes) SMK_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
