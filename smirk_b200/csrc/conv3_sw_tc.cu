// 3x3 stride-1 convolution as a "shifted-window" TF32 tcgen05 implicit GEMM.
//
// gemm_tc.cu feeds the nine filter taps of a 3x3 convolution with nine separate im2col TMA loads per
// 32-channel chunk: every input pixel crosses L2 -> shared memory nine times, and the generator's wide,
// shallow layers (224^2 x 32..64 channels, src/smirk_generator.py:56-76) end up bound by that traffic.
// Here the input window of an output tile is loaded ONCE per channel chunk and the nine taps are nine
// *views* of the same shared-memory patch:
//
//   output tile   TH x TW pixels of one image, TW = PW - 2, TH = 128 / PW  (PW = 32: 4 x 30, PW = 16: 8 x 14)
//   patch         (TH + 2) x PW pixels x 32 channels, one 4-D tiled TMA box (out-of-image halo = zero padding),
//                 rows of 128 bytes in SWIZZLE_128B order, row index r = py * PW + px
//   tap (dy,dx)   the A operand is the 128 consecutive patch rows starting at row dy*PW + dx: the UMMA
//                 descriptor's start address simply moves by whole 128-byte rows (the tensor core takes the swizzle
//                 phase from the address itself).  GEMM row m is output pixel (m / PW, m % PW); the two columns m % PW >= TW
//                 of each row wrap into the halo and are simply not stored (94 % / 88 % of the MMA rows are useful).
//   weights       [N][9*Cin] K-major, one BN x 32 box per (tap, chunk) through a small ring of their own.
//
// L2 -> shared-memory traffic per tile and chunk drops from 9 x 16 KiB to 24 KiB (+ weights); everything
// else (TMEM accumulator, mbarrier rings, staged coalesced epilogue) is as in gemm_tc.cu.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace smk {
namespace {

using namespace ptx;

constexpr int NUM_THREADS = 192;
constexpr int NB = 6;                              // weight-ring depth

struct SwArgs {
    int H, W;                  // output size (= unpadded input size)
    int N, Cin, nchunks;       // Cout, Cin, Cin / 32
    int origin;                // -1: zero padding by TMA out-of-bounds fill; 0: input buffer is already padded by 1
    int n_tiles_n;             // blockIdx.z = img * n_tiles_n + n_tile
    const float* scale; const float* bias;
    const float* res; int ld_res; int res_pad;
    int relu;
    float* out; int ld_out;
    int store;                 // 0 plain NHWC, 2 interior of a (H+2)x(W+2) padded buffer
    int round_out;
};

template <int BN, int PW>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3_sw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const SwArgs a) {
    constexpr int TH = 128 / PW, TW = PW - 2;
    constexpr int A_BYTES = (TH + 2) * PW * 128;           // 24 KiB (PW = 32) / 20 KiB (PW = 16): multiples of 1 KiB
    constexpr int B_BYTES = BN * 128;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t IDESC = make_idesc_tf32(128, BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                    // 2 stages
    uint8_t* sB = smem + 2 * A_BYTES;                      // NB stages (also the legal overrun area of tap views)
    uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + NB * B_BYTES);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_empty + 2;
    uint64_t* b_empty = b_full + NB;
    uint64_t* tmem_full = b_empty + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.z / a.n_tiles_n, n0 = (blockIdx.z - img * a.n_tiles_n) * BN;
    const int h0 = blockIdx.y * TH, w0 = blockIdx.x * TW;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmX); prefetch_tensormap(&tmW);
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: one patch per channel chunk, nine weight boxes behind it =====
            int itb = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                const int sa = c & 1;
                mbar_wait(&a_empty[sa], ((uint32_t)(c >> 1) & 1u) ^ 1u);
                mbar_expect_tx(&a_full[sa], (uint32_t)A_BYTES);
                tma_load_4d(&tmX, sA + sa * A_BYTES, &a_full[sa], c * 32, w0 + a.origin, h0 + a.origin, img);
                for (int tap = 0; tap < 9; ++tap, ++itb) {
                    const int s = itb % NB;
                    mbar_wait(&b_empty[s], ((uint32_t)(itb / NB) & 1u) ^ 1u);
                    mbar_expect_tx(&b_full[s], (uint32_t)B_BYTES);
                    tma_load_2d(&tmW, sB + s * B_BYTES, &b_full[s], tap * a.Cin + c * 32, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int itb = 0;
            for (int c = 0; c < a.nchunks; ++c) {
                const int sa = c & 1;
                mbar_wait(&a_full[sa], (uint32_t)(c >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(sA + sa * A_BYTES);
                for (int tap = 0; tap < 9; ++tap, ++itb) {
                    const int s = itb % NB;
                    mbar_wait(&b_full[s], (uint32_t)(itb / NB) & 1u);
                    tcgen05_fence_after();
                    const uint32_t a_tap = a_base + (uint32_t)(((tap / 3) * PW + (tap % 3)) * 128);
                    const uint32_t b_tap = smem_u32(sB + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(tmem_base, make_smem_desc(a_tap + k * 32), make_smem_desc(b_tap + k * 32), IDESC,
                                  (c | tap | k) != 0 ? 1u : 0u);
                    tcgen05_commit(&b_empty[s]);
                }
                tcgen05_commit(&a_empty[sa]);
            }
            tcgen05_commit(tmem_full);
        }
    } else {
        // ===== epilogue (warps 2..5): TMEM -> swizzled slab -> coalesced stores, as in gemm_tc.cu =====
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        const int quarter = warp & 3;
        uint8_t* slab = smem + quarter * 4096;                 // patch stage 0 is idle by now
        const int sub = lane >> 3, jj = lane & 7;
        int opix[8], rpix[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = quarter * 32 + 4 * i + sub;
            const int oy = m / PW, ox = m - oy * PW;
            const int oh = h0 + oy, ow = w0 + ox;
            int o = -1, r = -1;
            if (ox < TW && oh < a.H && ow < a.W) {
                const int plain = (img * a.H + oh) * a.W + ow;
                const int padded = (img * (a.H + 2) + oh + 1) * (a.W + 2) + ow + 1;
                o = a.store == 2 ? padded : plain;
                r = a.res_pad ? padded : plain;
            }
            opix[i] = o; rpix[i] = r;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            const int n = n0 + c0;
            if (n >= a.N) break;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int nc = n + jj * 4;
            if (nc < a.N) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + nc));
                const float4 bi = __ldg(reinterpret_cast<const float4*>(a.bias + nc));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (opix[i] < 0) continue;
                    const int r = 4 * i + sub;
                    const float4 x = *reinterpret_cast<const float4*>(slab + r * 128 + ((jj ^ (r & 7)) << 4));
                    float4 o;
                    o.x = fmaf(x.x, sc.x, bi.x); o.y = fmaf(x.y, sc.y, bi.y); o.z = fmaf(x.z, sc.z, bi.z); o.w = fmaf(x.w, sc.w, bi.w);
                    if (a.res) {
                        const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)rpix[i] * a.ld_res + nc));
                        o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                    }
                    if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (a.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                    *reinterpret_cast<float4*>(a.out + (size_t)opix[i] * a.ld_out + nc) = o;
                }
            }
            __syncwarp();
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int load_encoder() {
    if (g_encode) return 0;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    SMK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    SMK_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

template <int BN, int PW>
int launch(const CUtensorMap& tmX, const CUtensorMap& tmW, const SwArgs& a, dim3 grid, cudaStream_t st) {
    constexpr int TH = 128 / PW;
    constexpr size_t smem = 2 * (size_t)(TH + 2) * PW * 128 + (size_t)NB * BN * 128 + 1024 + 512;
    static unsigned long long configured_mask = 0;
    int dev = 0;
    SMK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !(configured_mask & (1ull << dev))) {
        SMK_CHECK_CUDA(cudaFuncSetAttribute(conv3_sw_kernel<BN, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) configured_mask |= 1ull << dev;
    }
    SMK_LAUNCH((conv3_sw_kernel<BN, PW>), dim3(grid), dim3(NUM_THREADS), smem, st, tmX, tmW, a);
    SMK_CHECK_LAUNCH();
    return 0;
}

}  // namespace

// p uses TcConv semantics with mode 1 (zero padding 1) or 2 (input already padded by 1), store 0 or 2.
int conv3_sw(const TcConv& p, cudaStream_t st) {
    if (int rc = load_encoder()) return rc;
    SMK_REQUIRE((p.mode == 1 || p.mode == 2) && p.Cin % 32 == 0 && p.K == 9 * p.Cin, "conv3_sw: needs a 3x3 conv with Cin %% 32 == 0");
    SMK_REQUIRE(p.store == 0 || p.store == 2, "conv3_sw: store must be 0 or 2");
    SMK_REQUIRE(p.N % 4 == 0 && p.ld_in % 4 == 0 && p.ld_out % 4 == 0, "conv3_sw: N and strides must be multiples of 4");
    const int PW = p.W <= 14 ? 16 : 32;
    const int TH = 128 / PW, TW = PW - 2;
    const int BN = p.N <= 32 ? 32 : (p.N <= 64 ? 64 : 128);
    const int Hin = p.mode == 2 ? p.H + 2 : p.H, Win = p.mode == 2 ? p.W + 2 : p.W;
    CUtensorMap tmX, tmW;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)p.B};
        cuuint64_t strides[3] = {(cuuint64_t)p.ld_in * 4, (cuuint64_t)Win * p.ld_in * 4, (cuuint64_t)Hin * Win * p.ld_in * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)PW, (cuuint32_t)(TH + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p.in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SMK_REQUIRE(r == CUDA_SUCCESS, "conv3_sw: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
        cuuint64_t strides[1] = {(cuuint64_t)p.K * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p.wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SMK_REQUIRE(r == CUDA_SUCCESS, "conv3_sw: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
    }
    SwArgs a{};
    a.H = p.H; a.W = p.W; a.N = p.N; a.Cin = p.Cin; a.nchunks = p.Cin / 32; a.origin = p.mode == 2 ? 0 : -1;
    a.n_tiles_n = cdiv(p.N, BN);
    a.scale = p.scale; a.bias = p.bias; a.res = p.res; a.ld_res = p.ld_res; a.res_pad = p.res_pad; a.relu = p.relu;
    a.out = p.out; a.ld_out = p.ld_out; a.store = p.store; a.round_out = p.round_out;
    dim3 grid(cdiv(p.W, TW), cdiv(p.H, TH), p.B * a.n_tiles_n);
    {
        const double M = (double)p.B * p.H * p.W;
        const char* tag = "conv3x3_sw_tc";
        if (g_prof_detail) tag = prof_shape_tag(tag, (long)M, p.K, p.N);
        SMK_TAG(tag, 4.0 * (M * p.Cin + (double)p.K * p.N + M * p.N * (p.res ? 2 : 1) + 2.0 * p.N), 2.0 * M * p.N * p.K, st);
    }
    if (PW == 32) {
        if (BN == 32) return launch<32, 32>(tmX, tmW, a, grid, st);
        if (BN == 64) return launch<64, 32>(tmX, tmW, a, grid, st);
        return launch<128, 32>(tmX, tmW, a, grid, st);
    }
    if (BN == 32) return launch<32, 16>(tmX, tmW, a, grid, st);
    if (BN == 64) return launch<64, 16>(tmX, tmW, a, grid, st);
    return launch<128, 16>(tmX, tmW, a, grid, st);
}

}  // namespace smk

extern "C" int smk_debug_conv3_sw(const float* in, int ld_in, int B, int H, int W, int Cin, const float* wt, const float* scale,
                                  const float* bias, int N, int mode, int relu, const float* res, int ld_res, int res_pad,
                                  float* out, int ld_out, int store, void* stream) {
    smk::TcConv p{};
    p.in = in; p.ld_in = ld_in; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.wt = wt; p.scale = scale; p.bias = bias; p.N = N; p.K = 9 * Cin;
    p.mode = mode; p.relu = relu; p.res = res; p.ld_res = ld_res; p.res_pad = res_pad; p.out = out; p.ld_out = ld_out; p.store = store;
    return smk::conv3_sw(p, (cudaStream_t)stream);
}
