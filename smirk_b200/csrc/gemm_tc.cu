// TF32 tcgen05 implicit-GEMM convolution for sm_100a:  C[M,N] = epi( A[M,K] * W[N,K]^T ).
//
// This is the tensor-core path of the generator's 3x3 convolutions / transposed convolutions and of
// the encoder's 1x1 convolutions (precision = 1).  The reference executes these layers as cuDNN
// TF32 implicit GEMMs (src/smirk_generator.py:56-76,147-178; torch default cudnn.allow_tf32=True);
// here they are one hand-written kernel:
//
//   * A (activations, NHWC fp32) is fetched tile-by-tile by TMA: `cp.async.bulk.tensor.2d` for 1x1
//     convs / plain GEMMs, `cp.async.bulk.tensor.4d...im2col` for 3x3 convs — the hardware walks 128
//     consecutive output pixels (crossing rows and images), applies the filter-tap offset and zero
//     fills the padding halo, writing a 128 x 32-channel (128-byte rows) SWIZZLE_128B tile into
//     shared memory.  No im2col matrix is ever materialised.
//   * W ([N][K], K-major, K ordered (tap, channel)) comes in through a 2-D TMA box of BN x 32.
//   * One elected thread issues `tcgen05.mma.cta_group::1.kind::tf32` (M=128, N=BN, K=8 per
//     instruction, 4 per 128-byte k-block); the fp32 accumulator lives in TMEM (BN columns).
//   * A STAGES-deep mbarrier ring decouples the TMA producer warp from the MMA warp;
//     `tcgen05.commit` releases shared-memory stages and finally signals the epilogue warps, which
//     read the accumulator with `tcgen05.ld.32x32b`, apply folded BatchNorm scale/bias, optional
//     residual and ReLU, and store NHWC rows — plain, into a channel slice of a concat buffer, into
//     the interior of a reflection-padded buffer, or pixel-shuffled (ConvTranspose2d k2 s2).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp_id % 4 selects the TMEM lane quarter each may access).
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>

namespace smk {
namespace {

constexpr int BM = 128;
constexpr int BKB = 128;                  // bytes of K per k-block = one SWIZZLE_128B row = 32 fp32
constexpr int BK = 32;
constexpr int UMMA_K = 8;                 // tf32
constexpr int A_STAGE_BYTES = BM * BKB;   // 16 KiB
constexpr int NUM_THREADS = 192;
constexpr int SLAB_BYTES = 4 * 4096;      // epilogue staging: 32 rows x 128 B per epilogue warp

using namespace ptx;                      // PTX wrappers shared by the tcgen05 kernels (tc_ptx.cuh)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return make_idesc_tf32(M, N); }

struct TcArgs {
    int M, N, nkb;                 // nkb = number of 32-wide k-blocks
    int tiles_n, n_tiles;          // column tiles per row of tiles, total tiles (tile t -> row tile t / tiles_n)
    int mode;                      // 0 = 2-D tiled A; 1 = im2col A
    int H, W;                      // output spatial dims (im2col tile origin decode; shuffle / padded stores)
    int cpb;                       // k-blocks per filter tap (Cin / 32) for im2col
    int lc;                        // im2col lower corner (-1: zero padding 1; 0: input already reflection-padded)
    // Two problems of identical shape may share one launch (the encoder's two "large" backbones run the same layer
    // list on different weights and activations): tiles [0, n_tiles_g) belong to problem 0, [n_tiles_g, 2 n_tiles_g) to
    // problem 1, each with its own tensor maps (TcMaps) and epilogue pointers.  Twice the work per launch for kernels
    // that sit on the launch/latency floor.
    int n_tiles_g;                 // tiles per problem (= n_tiles when there is only one)
    const float* scale[2]; const float* bias[2];
    const float* res[2]; int ld_res; int res_pad;   // residual (optionally read from the interior of a padded buffer)
    int relu;
    float* out[2]; int ld_out;
    int store;                     // 0 plain, 1 pixel-shuffle (N = 4*Cout), 2 interior of a (H+2)x(W+2) padded buffer,
                                   // 3 fused head: out[b, co, h, w] = sigmoid(head_b[co] + sum_n act[m, n] * head_w[n][co]) (NCHW, N <= 32)
    int round_out;                 // 1: round the stored activations to TF32 (round-to-nearest) for the next tensor-core layer
    int x3_trunc;                  // 3xTF32, tails in TMEM: 1 = rely on the tensor core TRUNCATING fp32 words to TF32 (head = a & ~0x1fff is
                                   // implicit, nothing is written back to shared memory); 0 = rewrite the heads (round-to-nearest) in place
    const float* head_w; const float* head_b; int head_c;      // store 3: 1x1 head weights [N][head_c], bias [head_c], head_c <= 4
};

struct TcMaps { CUtensorMap a[2], b[2], blo[2]; };       // per problem: activations, weights (TF32 heads), weight tails (3xTF32)

template <int BN, int STAGES, int MINB, bool PERSIST, int X3>
__global__ void __launch_bounds__(NUM_THREADS, MINB)
gemm_tc_kernel(const __grid_constant__ TcMaps mp, const TcArgs a) {
    // Output tiles (128 rows x BN columns) are strided over the grid.
    //   PERSIST = true : grid = resident CTAs; the smem ring runs across tile boundaries and the accumulator is
    //                    double-buffered in TMEM, so the TMA loads and MMAs of tile i+1 overlap the epilogue of
    //                    tile i.  Wins for the deep-K 3x3 convolutions (measured, profiles/r01_gemm_tc_persist.txt).
    //   PERSIST = false: grid = tiles, one tile per CTA, the epilogue staging slab aliases the (by then idle)
    //                    first ring stage and TMEM holds one accumulator: smaller footprint -> 3-5 CTAs per SM,
    //                    which is what the shallow-K streaming 1x1 layers want (more epilogue warps in flight).
    //   X3 = 2         : error-compensated "3xTF32" arithmetic (fp32-equivalent results on the tensor cores).  Every fp32
    //                    operand is split into a TF32 head and a TF32 tail, a = a_hi + a_lo, and three products are
    //                    accumulated: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (the dropped a_lo*w_lo term is ~2^-22 relative).
    //                    Weights are split on the host (TcMaps::blo maps the tails).  The activation tile is split after TMA has
    //                    landed it, by the epilogue warps — otherwise idle during the main loop of a single-tile CTA: thread =
    //                    tile row = TMEM lane; the tails go to TENSOR MEMORY (32 columns per ring stage next to the
    //                    accumulator) and are multiplied with the A-from-TMEM form of tcgen05.mma; the heads are what the
    //                    tensor core reads from the fp32 words itself (it truncates to TF32; x3_trunc = 0 rewrites them in
    //                    place, round to nearest).  The shared-memory footprint is that of the plain TF32 kernel plus the
    //                    weight tails, so as many CTAs share an SM as before — which is what the concurrent pipeline's
    //                    throughput hangs on (DESIGN.md "footprint beats per-kernel speed").  A first cut with the tails in a
    //                    second shared-memory tile cost 10 % end to end (27.2k vs 30.2k faces/s) and was removed.
    static_assert(!X3 || !PERSIST, "the 3xTF32 split borrows the epilogue warps: single-tile CTAs only");
    constexpr int B_STAGE_BYTES = BN * BKB;
    constexpr int HALF_STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int STAGE_BYTES = X3 ? HALF_STAGE_BYTES + B_STAGE_BYTES : HALF_STAGE_BYTES;         // [A][B] (+ [B tails])
    constexpr int BLO_OFF = HALF_STAGE_BYTES;                                                       // weight tails within a stage
    constexpr uint32_t ALO_COL = PERSIST ? 2 * BN : BN;                                            // X3 == 2: first TMEM column of the A tails
    constexpr uint32_t TMEM_NEED = (PERSIST ? 2 * BN : BN) + (X3 == 2 ? STAGES * 32 : 0);
    constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
    constexpr uint32_t IDESC = make_idesc(BM, BN);
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B; offset arithmetic keeps the shared address space visible to the
    // compiler (LDS/STS for the staging slab instead of generic LD/ST).
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* slabs = PERSIST ? smem + STAGES * STAGE_BYTES : smem;      // 4 x 4 KiB epilogue staging, one per epilogue warp
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + (PERSIST ? SLAB_BYTES : 0));
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* split = acc_empty + 2;                                    // X3: "stage s has been split" (4 epilogue warps arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(split + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        const int g0 = (!PERSIST && (int)blockIdx.x >= a.n_tiles_g) ? 1 : 0;
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mp.a[g0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mp.b[g0]) : "memory");
        if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&mp.blo[g0]) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split[s], 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();                                // barriers, tensor maps and TMEM are set up; now wait for the producer layer

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int it = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const int g = t >= a.n_tiles_g ? 1 : 0, tt = t - g * a.n_tiles_g;
                const CUtensorMap* tmA = &mp.a[g]; const CUtensorMap* tmB = &mp.b[g]; const CUtensorMap* tmBlo = &mp.blo[g];
                const int tm = tt / a.tiles_n, tn = tt - tm * a.tiles_n;
                const int m0 = tm * BM, n0 = tn * BN;
                int q0 = 0, p0 = 0, img = 0;
                if (a.mode == 1) { int hw = a.H * a.W; img = m0 / hw; int r = m0 - img * hw; p0 = r / a.W; q0 = r - p0 * a.W; }
                for (int kb = 0; kb < a.nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((uint32_t)(it / STAGES) & 1u) ^ 1u);
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(&full[s], (uint32_t)(HALF_STAGE_BYTES + (X3 ? B_STAGE_BYTES : 0)));
                    if (X3) tma_load_2d(tmBlo, sa + BLO_OFF, &full[s], kb * BK, n0);
                    if (a.mode == 0) {
                        tma_load_2d(tmA, sa, &full[s], kb * BK, m0);
                    } else {
                        int tap = kb / a.cpb, ch = kb - tap * a.cpb;
                        int r = tap / 3, sx = tap - r * 3;
                        tma_load_im2col(tmA, sa, &full[s], ch * BK, q0 + a.lc, p0 + a.lc, img, (uint16_t)sx, (uint16_t)r);
                    }
                    tma_load_2d(tmB, sb, &full[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int it = 0, tc = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++tc) {
                const int buf = tc & 1;
                mbar_wait(&acc_empty[buf], ((uint32_t)(tc >> 1) & 1u) ^ 1u);      // epilogue has drained this accumulator
                tcgen05_fence_after();
                for (int kb = 0; kb < a.nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(X3 ? &split[s] : &full[s], (uint32_t)(it / STAGES) & 1u);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        uint64_t da = make_smem_desc(sa + k * UMMA_K * 4);
                        uint64_t db = make_smem_desc(sb + k * UMMA_K * 4);
                        umma_tf32(tmem_base + (uint32_t)(buf * BN), da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
                        if (X3) {
                            uint64_t dbl = make_smem_desc(sa + BLO_OFF + k * UMMA_K * 4);
                            umma_tf32_ts(tmem_base + (uint32_t)(buf * BN), tmem_base + ALO_COL + (uint32_t)(s * 32 + k * UMMA_K), db, IDESC, 1u);
                            umma_tf32(tmem_base + (uint32_t)(buf * BN), da, dbl, IDESC, 1u);      // a_hi * w_lo
                        }
                    }
                    tcgen05_commit(&empty[s]);          // frees this smem stage once the MMAs above have read it
                }
                tcgen05_commit(&acc_full[buf]);         // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        // Two phases per 32-column chunk, both private to the warp (it owns TMEM lanes / tile rows
        // [32*quarter, +32)), so only __syncwarp separates them:
        //   1. lane = row: tcgen05.ld 32 accumulator columns and park them in a 32 x 128-byte staging
        //      slab in shared memory (16-byte chunks XOR-swizzled by row: conflict-free both ways);
        //   2. 8 lanes per row, 4 rows per instruction: read the slab back transposed, apply
        //      scale/bias (+residual) (+ReLU) (+TF32 rounding) and store — every global access of the
        //      warp now covers whole 128-byte lines of 4 output rows instead of 16 bytes of 32 rows.
        const int quarter = warp & 3;
        if (X3 == 2) {
            // ===== 3xTF32, tails in tensor memory: thread = tile row (its TMEM lane); heads rewritten in place =====
            const int r = quarter * 32 + lane;
            for (int kb = 0; kb < a.nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full[s], (uint32_t)(kb / STAGES) & 1u);
                uint8_t* row = smem + s * STAGE_BYTES + r * 128;
                float lo[32];
                if (a.x3_trunc) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(row + ((j ^ (r & 7)) << 4));
                        lo[4 * j] = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo[4 * j + 1] = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        lo[4 * j + 2] = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo[4 * j + 3] = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4* p = reinterpret_cast<float4*>(row + ((j ^ (r & 7)) << 4));       // 16-byte chunk j of the row (SWIZZLE_128B)
                        const float4 v = *p;
                        float4 hi;
                        hi.x = smk::round_tf32(v.x); hi.y = smk::round_tf32(v.y); hi.z = smk::round_tf32(v.z); hi.w = smk::round_tf32(v.w);
                        lo[4 * j] = v.x - hi.x; lo[4 * j + 1] = v.y - hi.y; lo[4 * j + 2] = v.z - hi.z; lo[4 * j + 3] = v.w - hi.w;
                        *p = hi;
                    }
                }
                tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + ALO_COL + (uint32_t)(s * 32), lo);   // warp-collective, waits for completion
                if (!a.x3_trunc) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
            tcgen05_fence_after();
        }
        uint8_t* slab = slabs + quarter * 4096;
        const int sub = lane >> 3, jj = lane & 7;              // phase-2 role: row-in-group, 16-byte chunk
        if (a.store == 3) {                                    // head constants of channel `lane` -> this warp's slab (PERSIST layout: slabs are private)
            float* par = reinterpret_cast<float*>(slab);
            par[lane * 8 + 0] = __ldg(a.scale[0] + lane); par[lane * 8 + 1] = __ldg(a.bias[0] + lane);
#pragma unroll
            for (int co = 0; co < 4; ++co) par[lane * 8 + 2 + co] = co < a.head_c ? __ldg(a.head_w + (size_t)lane * a.head_c + co) : 0.f;
            if (lane < 4) par[256 + lane] = lane < a.head_c ? __ldg(a.head_b + lane) : 0.f;
            __syncwarp();
        }
        int tc = 0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++tc) {
            const int g = t >= a.n_tiles_g ? 1 : 0, tt = t - g * a.n_tiles_g;
            const float* __restrict__ g_scale = a.scale[g]; const float* __restrict__ g_bias = a.bias[g];
            const float* __restrict__ g_res = a.res[g]; float* __restrict__ g_out = a.out[g];
            const int tm = tt / a.tiles_n, tn = tt - tm * a.tiles_n;
            const int m0 = tm * BM, n0 = tn * BN;
            const int buf = tc & 1;
            int opix[8], rpix[8];                              // destination / residual pixel of my 8 phase-2 rows (-1: past M)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + quarter * 32 + 4 * i + sub;
                int o = -1, r = -1;
                if (m < a.M) {
                    o = r = m;
                    if (a.store != 0 || a.res_pad) {
                        const int hw = a.H * a.W, b = m / hw, rem = m - b * hw, h = rem / a.W, w = rem - h * a.W;
                        const int padded = (b * (a.H + 2) + h + 1) * (a.W + 2) + w + 1;
                        if (a.store == 2) o = padded;
                        else if (a.store == 1) o = (b * (2 * a.H) + 2 * h) * (2 * a.W) + 2 * w;
                        if (a.res_pad) r = padded;
                    }
                }
                opix[i] = o; rpix[i] = r;
            }
            mbar_wait(&acc_full[buf], (uint32_t)(tc >> 1) & 1u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int n = n0 + c0;
                if (n >= a.N) break;                           // warp-uniform
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c0), v);     // warp-collective
                if (c0 + 32 >= BN || n + 32 >= a.N) {          // last read of this accumulator: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (a.store == 3) {
                    // Fused 1x1 head + sigmoid (smirk_generator.py:77-78 -> :86), N == BN == 32: right after tcgen05.ld every lane holds
                    // ALL 32 accumulators of its own pixel, so BN + ReLU + the 32 -> head_c dot products need no exchange between
                    // lanes; per-channel constants come from this warp's slab (filled once, broadcast LDS), consecutive lanes are
                    // consecutive pixels -> coalesced NCHW stores.  The activation tensor itself is never written.
                    const float* par = reinterpret_cast<const float*>(slab);       // [32][8]: scale, bias, w0..w3, pad
                    float a0 = par[256], a1 = par[257], a2 = par[258], a3 = par[259];       // head bias
#pragma unroll
                    for (int nn = 0; nn < 32; ++nn) {
                        const float4 p0 = *reinterpret_cast<const float4*>(par + nn * 8);
                        const float2 p1 = *reinterpret_cast<const float2*>(par + nn * 8 + 4);
                        float x = fmaf(v[nn], p0.x, p0.y);
                        if (a.relu) x = fmaxf(x, 0.f);
                        a0 = fmaf(x, p0.z, a0); a1 = fmaf(x, p0.w, a1); a2 = fmaf(x, p1.x, a2); a3 = fmaf(x, p1.y, a3);
                    }
                    const int m = m0 + quarter * 32 + lane;
                    if (m < a.M) {
                        const int hw_px = a.H * a.W, b = m / hw_px, rem = m - b * hw_px;
                        float* dst = g_out + (size_t)b * a.head_c * hw_px + rem;
                        dst[0] = 1.f / (1.f + __expf(-a0));
                        if (a.head_c > 1) dst[hw_px] = 1.f / (1.f + __expf(-a1));
                        if (a.head_c > 2) dst[2 * (size_t)hw_px] = 1.f / (1.f + __expf(-a2));
                        if (a.head_c > 3) dst[3 * (size_t)hw_px] = 1.f / (1.f + __expf(-a3));
                    }
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                const int nc = n + jj * 4;                     // my 4 columns
                if (nc < a.N) {
                    const float4 sc = __ldg(reinterpret_cast<const float4*>(g_scale + nc));
                    const float4 bi = __ldg(reinterpret_cast<const float4*>(g_bias + nc));
                    int col = nc, pix_off = 0;
                    if (a.store == 1) {                        // ConvTranspose2d k2 s2: n = (dy*2+dx)*Cout + co
                        const int cout = a.N >> 2, q = nc / cout;
                        col = nc - q * cout; pix_off = (q >> 1) * (2 * a.W) + (q & 1);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (opix[i] < 0) continue;
                        const int r = 4 * i + sub;
                        const float4 x = *reinterpret_cast<const float4*>(slab + r * 128 + ((jj ^ (r & 7)) << 4));
                        float4 o;
                        o.x = fmaf(x.x, sc.x, bi.x); o.y = fmaf(x.y, sc.y, bi.y); o.z = fmaf(x.z, sc.z, bi.z); o.w = fmaf(x.w, sc.w, bi.w);
                        if (g_res) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(g_res + (size_t)rpix[i] * a.ld_res + nc));
                            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                        }
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (a.round_out) { o.x = smk::round_tf32(o.x); o.y = smk::round_tf32(o.y); o.z = smk::round_tf32(o.z); o.w = smk::round_tf32(o.w); }
                        *reinterpret_cast<float4*>(g_out + (size_t)(opix[i] + pix_off) * a.ld_out + col) = o;
                    }
                }
                __syncwarp();
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- reflection halo of a [B, H+2, W+2, C] buffer whose interior has been written -------------------
__global__ void __launch_bounds__(256)
reflect_halo_kernel(float* __restrict__ buf, int B, int H, int W, int C) {
    pdl_sync();
    const int Hp = H + 2, Wp = W + 2, C4 = C >> 2;
    const int halo = 2 * Wp + 2 * H;                       // halo pixels per image
    long total = (long)B * halo * C4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4); long t = i / C4; int k = (int)(t % halo); int b = (int)(t / halo);
        int ph, pw;
        if (k < Wp) { ph = 0; pw = k; }
        else if (k < 2 * Wp) { ph = Hp - 1; pw = k - Wp; }
        else { int r = k - 2 * Wp; ph = 1 + (r >> 1); pw = (r & 1) ? Wp - 1 : 0; }
        // ReflectionPad2d(1): padded index p -> interior index |p-1| mirrored at the far edge
        int sh = ph - 1, sw = pw - 1;
        sh = sh < 0 ? 1 : (sh >= H ? H - 2 : sh);
        sw = sw < 0 ? 1 : (sw >= W ? W - 2 : sw);
        const float4* src = reinterpret_cast<const float4*>(buf + (((size_t)b * Hp + sh + 1) * Wp + sw + 1) * C) + c4;
        float4* dst = reinterpret_cast<float4*>(buf + (((size_t)b * Hp + ph) * Wp + pw) * C) + c4;
        *dst = *src;
    }
}

// ---- host: tensor-map encoding through the driver entry points (no link-time libcuda dependency) ------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

int load_driver_fns() {
    if (g_encode_tiled && g_encode_im2col) return 0;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    SMK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    SMK_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_encode_tiled = (EncodeTiledFn)fn;
    SMK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    SMK_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col not available from the driver");
    g_encode_im2col = (EncodeIm2colFn)fn;
    return 0;
}

int encode_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SMK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
    return 0;
}

// NHWC activation tensor [B][Hin][Win][C] (pixel stride ld) read as 3x3 windows; lower/upper corner per CUTLASS
// conventions: lower = -pad, upper = pad - (3-1).
int encode_im2col(CUtensorMap* map, const float* base, int B, int Hin, int Win, int C, int ld, int pad) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)Win * ld * 4, (cuuint64_t)Hin * Win * ld * 4};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - 2, pad - 2};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, lower, upper,
                                 (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SMK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d): B=%d H=%d W=%d C=%d ld=%d pad=%d", (int)r, B, Hin, Win, C, ld, pad);
    return 0;
}

template <int BN, int STAGES, int MINB, bool PERSIST, int X3 = 0>
int launch(const TcMaps& mp, const TcArgs& a_in, cudaStream_t st, int groups) {
    constexpr size_t stage = X3 ? A_STAGE_BYTES + 2 * BN * BKB : A_STAGE_BYTES + BN * BKB;
    constexpr size_t smem = (size_t)STAGES * stage + (PERSIST ? SLAB_BYTES : 0) + 1024 + 256;
    static_assert(MINB * (smem + 1024) <= 228 * 1024, "shared memory budget of MINB resident CTAs");
    static_assert(MINB * ((PERSIST ? 2 : 1) * BN + (X3 == 2 ? STAGES * 32 : 0)) <= 512, "TMEM budget of MINB resident CTAs");
    // The attribute is per device and per function; set it once per (device, instantiation).  One bit per
    // device ordinal; a benign race (two threads setting it twice) is harmless.
    static unsigned long long configured_mask = 0;
    int dev = 0;
    SMK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !(configured_mask & (1ull << dev))) {
        SMK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, MINB, PERSIST, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) configured_mask |= 1ull << dev;
    }
    TcArgs a = a_in;
    a.tiles_n = cdiv(a.N, BN);
    a.n_tiles_g = cdiv(a.M, BM) * a.tiles_n;
    a.n_tiles = groups * a.n_tiles_g;
    dim3 grid((unsigned)(PERSIST ? std::min(a.n_tiles, MINB * 148) : a.n_tiles));
    SMK_LAUNCH((gemm_tc_kernel<BN, STAGES, MINB, PERSIST, X3>), dim3(grid), dim3(NUM_THREADS), smem, st, mp, a);
    SMK_CHECK_LAUNCH();
    return 0;
}

}  // namespace

int tc_init() { return load_driver_fns(); }

int tc_conv(const TcConv& p, cudaStream_t st, const TcConv* p2) {
    if (int rc = load_driver_fns()) return rc;
    if (!p2 && conv3_win_supported(p)) return conv3_win(p, st);       // 224^2 / 112^2, Cout 32 / 64: one patch load per chunk, resident weights
    const int M = p.B * p.H * p.W;
    const int groups = p2 ? 2 : 1;
    SMK_REQUIRE(p.N % 4 == 0 && p.K % 4 == 0 && p.ld_in % 4 == 0 && p.ld_out % 4 == 0, "tc_conv: N, K, ld must be multiples of 4");
    SMK_REQUIRE(p.mode == 0 || (p.Cin % BK == 0 && p.K == 9 * p.Cin), "tc_conv: 3x3 mode needs Cin %% 32 == 0 (got %d)", p.Cin);
    SMK_REQUIRE(p.store != 1 || ((p.N / 4) % 32 == 0), "tc_conv: pixel-shuffle store needs Cout %% 32 == 0");
    SMK_REQUIRE(!p2 || (p2->B == p.B && p2->H == p.H && p2->W == p.W && p2->Cin == p.Cin && p2->N == p.N && p2->K == p.K && p2->mode == p.mode &&
                        p2->relu == p.relu && p2->ld_in == p.ld_in && p2->ld_out == p.ld_out && p2->ld_res == p.ld_res && p2->store == p.store &&
                        p2->res_pad == p.res_pad && p2->round_out == p.round_out && !p2->wt_lo == !p.wt_lo && !p2->res == !p.res && p.store != 3),
                "tc_conv: paired problems must have identical shapes and epilogue options");
    int BN = p.N <= 32 ? 32 : (p.N <= 64 ? 64 : 128);
    // Wide layers (N >= 256: the generator's 28^2 / 14^2 convolutions, 75 % of its FLOPs): a 128 x 256 tile halves the
    // A-operand bytes the tensor core pulls from shared memory per FLOP.  TF32 operands are 4 bytes, so at BN = 128 the
    // MMA reads (A 4 KB + B 4 KB per 64-cycle instruction) plus the TMA fill already ask for ~2x the 128 B/clk an SM's
    // shared memory delivers; at BN = 256 the demand drops to ~1.5x.  One persistent CTA per SM, 4-stage ring (192 KB),
    // both accumulators (2 x 256 columns) fill the SM's tensor memory.  SMK_TC_BN256=0 restores 128-wide tiles.
    static const int bn256 = []() { const char* e = getenv("SMK_TC_BN256"); return e ? atoi(e) : 1; }();
    if (bn256 && p.mode != 0 && p.N % 256 == 0 && !p.wt_lo) BN = 256;
    // Few-tile, deep-K problems (the encoder's 7x7 / 14x14 projections: M = 1568..6272, K up to 960) are a serial
    // chain of k-blocks on a handful of SMs: narrower N tiles put more CTAs to work.  An 8-stage ring with one
    // CTA per SM (SMK_TC_DEEP_SMALL=1) makes such a kernel ~20 % faster when it runs ALONE, but its 177 KB of
    // shared memory evict every other kernel from the SM; the pipeline runs three backbones x several batches
    // concurrently, where the small-footprint configuration is worth +11 % end to end (profiles/r01_footprint_sweep.txt).
    const int nkb_all = cdiv(p.K, BK);
    bool deep_small = false;
    if (nkb_all >= 6 && p.store != 1) {
        // (narrowing is off by default: fewer, wider tiles re-read A less often, and in the concurrent pipeline idle SMs
        //  are filled by other kernels anyway; SMK_TC_NARROW=148 restores the latency-oriented choice.)
        static const int narrow_target = []() { const char* e = getenv("SMK_TC_NARROW"); return e ? atoi(e) : 0; }();
        while (BN > 32 && (long)groups * cdiv(M, BM) * cdiv(p.N, BN) < narrow_target) BN >>= 1;
        static const int deep_small_on = []() { const char* e = getenv("SMK_TC_DEEP_SMALL"); return e ? atoi(e) : 0; }();
        deep_small = deep_small_on && (long)groups * cdiv(M, BM) * cdiv(p.N, BN) <= 2 * 148;
    }
    TcMaps mp;
    TcArgs a{};
    a.M = M; a.N = p.N; a.nkb = cdiv(p.K, BK); a.mode = p.mode == 0 ? 0 : 1; a.H = p.H; a.W = p.W;
    a.cpb = p.mode == 0 ? 1 : p.Cin / BK; a.lc = p.mode == 2 ? 0 : -1;
    a.ld_res = p.ld_res; a.res_pad = p.res_pad; a.relu = p.relu;
    a.ld_out = p.ld_out; a.store = p.store; a.round_out = p.round_out;
    a.head_w = p.head_w; a.head_b = p.head_b; a.head_c = p.head_c;
    {
        static const int x3_trunc = []() { const char* e = getenv("SMK_X3_TRUNC"); return e ? atoi(e) : 1; }();
        a.x3_trunc = x3_trunc;
    }
    SMK_REQUIRE(p.store != 3 || (p.N == 32 && p.mode != 0 && p.head_w && p.head_b && p.head_c >= 1 && p.head_c <= 4 && !p.res && !p2),
                "tc_conv: the fused 1x1 head needs a 3x3 conv with N == 32 (persistent kernel, one full column tile) and 1..4 head channels");
    SMK_REQUIRE(!p.wt_lo || (p.mode == 0 && p.store == 0), "tc_conv: the 3xTF32 path covers plain 1x1 convolutions / GEMMs");
    for (int g = 0; g < groups; ++g) {
        const TcConv& q = g ? *p2 : p;
        a.scale[g] = q.scale; a.bias[g] = q.bias; a.res[g] = q.res; a.out[g] = q.out;
        if (q.mode == 0) {
            if (int rc = encode_2d(&mp.a[g], q.in, (uint64_t)M, (uint64_t)q.K, (uint64_t)q.ld_in, BM)) return rc;
        } else if (q.mode == 1) {
            if (int rc = encode_im2col(&mp.a[g], q.in, q.B, q.H, q.W, q.Cin, q.ld_in, 1)) return rc;
        } else {                                            // input buffer is [B, H+2, W+2, C], already reflection padded
            if (int rc = encode_im2col(&mp.a[g], q.in, q.B, q.H + 2, q.W + 2, q.Cin, q.ld_in, 0)) return rc;
        }
        if (int rc = encode_2d(&mp.b[g], q.wt, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.K, (uint32_t)BN)) return rc;
        if (q.wt_lo) { if (int rc = encode_2d(&mp.blo[g], q.wt_lo, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.K, (uint32_t)BN)) return rc; }
        else mp.blo[g] = mp.b[g];
    }
    if (groups == 1) { mp.a[1] = mp.a[0]; mp.b[1] = mp.b[0]; mp.blo[1] = mp.blo[0]; a.scale[1] = a.scale[0]; a.bias[1] = a.bias[0]; a.res[1] = a.res[0]; a.out[1] = a.out[0]; }
    {
        const double cin_eff = p.mode == 0 ? p.K : p.Cin;
        const char* tag = p.mode == 0 ? (p.store == 1 ? "upconv_gemm_tc" : (p.wt_lo ? "pw_gemm_tc3x" : "pw_gemm_tc"))
                                      : (p.store == 3 ? "conv3x3_head_gemm_tc" : "conv3x3_gemm_tc");
        if (g_prof_detail) tag = prof_shape_tag(tag, (long)groups * M, p.K, p.N);
        SMK_TAG(tag,
                groups * 4.0 * ((double)M * cin_eff + (double)p.K * p.N + (double)M * p.N * (p.res ? 2 : 1) + 2.0 * p.N),
                groups * 2.0 * (double)M * p.N * p.K, st);
    }
    if (p.wt_lo) {                                      // 3xTF32: fp32-equivalent arithmetic (encoder precision 3)
        // deep-K layers (the 14x14 / 7x7 projections, K = 480..960): per k-block the chain TMA -> split -> MMA -> free is
        // ~1.5 us with a 2-stage ring (measured: 47 us for K = 960); a 4-stage ring keeps three loads in flight and is
        // faster alone, but its footprint costs the concurrent pipeline 3 % (29.8k vs 30.6k faces/s): opt-in, SMK_X3_DEEP=1
        static const int x3_deep = []() { const char* e = getenv("SMK_X3_DEEP"); return e ? atoi(e) : 0; }();
        if (x3_deep && a.nkb >= 8) {
            if (BN == 32) return launch<32, 4, 2, false, 2>(mp, a, st, groups);
            if (BN == 64) return launch<64, 4, 1, false, 2>(mp, a, st, groups);
            return launch<128, 4, 1, false, 2>(mp, a, st, groups);
        }
        if (BN == 32) return launch<32, 2, 4, false, 2>(mp, a, st, groups);
        if (BN == 64) return launch<64, 2, 3, false, 2>(mp, a, st, groups);
        return launch<128, 2, 2, false, 2>(mp, a, st, groups);
    }
    if (deep_small && BN == 32) return launch<32, 8, 1, false>(mp, a, st, groups);
    if (deep_small && BN == 64) return launch<64, 8, 1, false>(mp, a, st, groups);
    static const int persist_mode = []() { const char* e = getenv("SMK_TC_PERSIST"); return e ? atoi(e) : -1; }();   // -1 auto, 0 never, 1 always
    const long n_tiles = (long)groups * cdiv(M, BM) * cdiv(p.N, BN);
    // 3x3 convolutions (deep K): persistent CTAs for the narrow-N layers and for the few-tile 14x14 layers;
    // the wide-N layers are bound by the shared-memory fill rate either way and keep two single-tile CTAs per SM.
    const bool persist = p.store == 3 || (persist_mode >= 0 ? persist_mode != 0 : (p.mode != 0 && (BN <= 64 || n_tiles <= 2 * 148)));   // (the fused head lives in the persistent layout: private slabs)
    if (BN == 256) return launch<256, 4, 1, true>(mp, a, st, groups);
    if (persist) {
        if (BN == 32) return launch<32, 4, 2, true>(mp, a, st, groups);
        if (BN == 64) return launch<64, 3, 2, true>(mp, a, st, groups);
        return a.nkb > 8 ? launch<128, 5, 1, true>(mp, a, st, groups) : launch<128, 2, 2, true>(mp, a, st, groups);
    }
    // One tile per CTA, 2-stage ring: 41-66 KB per CTA, so 3-5 CTAs of this kernel — or CTAs of the other
    // backbones' and batches' kernels — share an SM and hide each other's prologue/epilogue.  Deeper rings
    // (SMK_TC_SHALLOW_NKB=2 restores them for K > 64) win a few percent per kernel in isolation and lose
    // 5 % end to end for the same co-residency reason as above.
    static const int shallow_nkb = []() { const char* e = getenv("SMK_TC_SHALLOW_NKB"); return e ? atoi(e) : (1 << 30); }();
    const bool shallow = a.nkb <= shallow_nkb;
    if (BN == 32) return shallow ? launch<32, 2, 5, false>(mp, a, st, groups) : launch<32, 4, 2, false>(mp, a, st, groups);
    if (BN == 64) return shallow ? launch<64, 2, 4, false>(mp, a, st, groups) : launch<64, 4, 2, false>(mp, a, st, groups);
    return shallow ? launch<128, 2, 3, false>(mp, a, st, groups) : launch<128, 3, 2, false>(mp, a, st, groups);
}

int reflect_halo(float* buf, int B, int H, int W, int C, cudaStream_t st) {
    long total = (long)B * (2 * (W + 2) + 2 * H) * (C / 4);
    SMK_TAG("reflect_halo", 8.0 * (double)total * 4, 0.0, st);
    SMK_LAUNCH(reflect_halo_kernel, dim3((int)std::min<long>((total + 255) / 256, 148L * 8)), dim3(256), 0, st, buf, B, H, W, C);
    SMK_CHECK_LAUNCH();
    return 0;
}

}  // namespace smk

// ---- debug / unit-test entry points (tests/test_gpu_kernels.py) -----------------------------------------
extern "C" int smk_debug_conv_tc(const float* in, int ld_in, int B, int H, int W, int Cin, const float* wt, const float* scale,
                                 const float* bias, int N, int K, int mode, int relu, const float* res, int ld_res, int res_pad,
                                 float* out, int ld_out, int store, void* stream) {
    smk::TcConv p{};
    p.in = in; p.ld_in = ld_in; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.wt = wt; p.scale = scale; p.bias = bias; p.N = N; p.K = K;
    p.mode = mode; p.relu = relu; p.res = res; p.ld_res = ld_res; p.res_pad = res_pad; p.out = out; p.ld_out = ld_out; p.store = store;
    return smk::tc_conv(p, (cudaStream_t)stream);
}
// 1x1 conv / GEMM on the 3xTF32 path: wt_hi / wt_lo are the TF32 heads and tails of the [N][K] weights.
extern "C" int smk_debug_gemm_tc3x(const float* in, int ld_in, int M, const float* wt_hi, const float* wt_lo, const float* scale,
                                   const float* bias, int N, int K, int relu, const float* res, int ld_res, float* out, int ld_out, void* stream) {
    smk::TcConv p{};
    p.in = in; p.ld_in = ld_in; p.B = 1; p.H = 1; p.W = M; p.Cin = K; p.wt = wt_hi; p.wt_lo = wt_lo; p.scale = scale; p.bias = bias; p.N = N; p.K = K;
    p.mode = 0; p.relu = relu; p.res = res; p.ld_res = ld_res; p.res_pad = 0; p.out = out; p.ld_out = ld_out; p.store = 0; p.round_out = 0;
    return smk::tc_conv(p, (cudaStream_t)stream);
}
extern "C" int smk_debug_reflect_halo(float* buf, int B, int H, int W, int C, void* stream) {
    return smk::reflect_halo(buf, B, H, W, C, (cudaStream_t)stream);
}
