// Shared helpers for the smirk_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/smirk_b200.h"

namespace smk {

void set_error(const char* fmt, ...);

#define SMK_CHECK_CUDA(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            smk::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                             \
        }                                                                               \
    } while (0)

#define SMK_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) { smk::set_error(__VA_ARGS__); return -1; }                        \
    } while (0)

// Launch bookkeeping: every kernel launch site calls SMK_TAG(...) first.  It counts launches (the
// `gpu_launches` figure bench.py reports) and, when the built-in profiler is enabled
// (smk_profiler_enable), brackets the launch with CUDA events on the launching stream so bench.py can
// report per-kernel time, algorithmic bytes and FLOPs without an external profiler.
void prof_begin(const char* tag, double bytes, double flops, cudaStream_t st);
void prof_end();
extern bool g_prof_detail;
bool profiling();                      // true while the event profiler is on (callers then serialise side streams)
const char* prof_shape_tag(const char* base, long m, long k, long n);   // interned "base:M.._K.._N.."
#define SMK_TAG(tag, bytes, flops, st) smk::prof_begin(tag, (double)(bytes), (double)(flops), st)
#define SMK_CHECK_LAUNCH() do { smk::prof_end(); SMK_CHECK_CUDA(cudaGetLastError()); } while (0)

// Device buffers owned by a handle (constants only; forwards never allocate).
struct DeviceArena {
    std::vector<void*> ptrs;
    ~DeviceArena() { for (void* p : ptrs) cudaFree(p); }
    template <typename T>
    cudaError_t upload(const T* host, size_t n, T** out) {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, n * sizeof(T) + 16);
        if (e != cudaSuccess) return e;
        ptrs.push_back(d);
        e = cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice);
        *out = (T*)d;
        return e;
    }
    template <typename T>
    cudaError_t upload(const std::vector<T>& v, T** out) { return upload(v.data(), v.size(), out); }
};

// Bump allocator over the caller-provided workspace (256-byte aligned slices).
struct Workspace {
    char* base; size_t size; size_t off = 0;
    Workspace(void* p, size_t n) : base((char*)p), size(n) {}
    template <typename T> T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        if (off + bytes > size) return nullptr;
        T* r = (T*)(base + off); off += bytes; return r;
    }
};
static inline size_t ws_round(size_t bytes) { return (bytes + 255) & ~size_t(255); }

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts
// with pdl_sync(): `griddepcontrol.wait` blocks until the preceding kernel on the stream has completed and
// flushed (so every global read AND write of this kernel stays ordered after it), `griddepcontrol.
// launch_dependents` lets the following kernel's CTAs be scheduled as soon as all CTAs of this one have
// started.  Net effect: launch latency, parameter/tensor-map fetch, barrier init and TMEM allocation of
// kernel N+1 overlap the tail of kernel N — which is what a chain of ~100 short kernels per batch is bound
// by.  Inside a stream capture these become programmatic graph edges.  The attribute is opt-in (SMK_PDL=1, see
// common.cu); without it the device-side instructions are no-ops.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}
#define SMK_LAUNCH(kernel, grid, block, smem, st, ...) (void)smk::launch_pdl(kernel, grid, block, smem, st, __VA_ARGS__)
#endif

// TF32 rounding (round-to-nearest, ties away: PTX cvt.rna).  The tensor cores read fp32 words from shared
// memory and simply ignore the 13 low mantissa bits (truncation, a systematic bias); operands that feed
// a tcgen05 layer are therefore rounded once, where they are produced: weights on the host at pack
// time, activations in the epilogue of the kernel that writes them — what cuDNN/CUTLASS TF32 kernels
// do with cvt.rna in registers before mma.
#ifdef __CUDACC__
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// Packed fp32 FMA (Blackwell FFMA2: two IEEE fused multiply-adds per instruction on 64-bit register pairs).
// Same rounding as two scalar fmaf calls; halves the issue slots of the CUDA-core inner loops.
__device__ __forceinline__ void fma4_acc(float4& acc, const float4& x, const float4& k) {    // acc += x * k
    uint64_t* a = reinterpret_cast<uint64_t*>(&acc);
    const uint64_t* xx = reinterpret_cast<const uint64_t*>(&x);
    const uint64_t* kk = reinterpret_cast<const uint64_t*>(&k);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[0]) : "l"(xx[0]), "l"(kk[0]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[1]) : "l"(xx[1]), "l"(kk[1]));
}
__device__ __forceinline__ void fma4_s(float4& acc, float x, const float4& k) {              // acc += x * k (scalar x)
    uint64_t xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    uint64_t* a = reinterpret_cast<uint64_t*>(&acc);
    const uint64_t* kk = reinterpret_cast<const uint64_t*>(&k);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[0]) : "l"(xx), "l"(kk[0]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[1]) : "l"(xx), "l"(kk[1]));
}
__device__ __forceinline__ float4 fma4(const float4& x, const float4& s, const float4& b) {  // x * s + b
    float4 o;
    uint64_t* oo = reinterpret_cast<uint64_t*>(&o);
    const uint64_t* xx = reinterpret_cast<const uint64_t*>(&x);
    const uint64_t* ss = reinterpret_cast<const uint64_t*>(&s);
    const uint64_t* bb = reinterpret_cast<const uint64_t*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(oo[0]) : "l"(xx[0]), "l"(ss[0]), "l"(bb[0]));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(oo[1]) : "l"(xx[1]), "l"(ss[1]), "l"(bb[1]));
    return o;
}
#endif
static inline float round_tf32_host(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;          // inf / nan
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
}

}  // namespace smk
