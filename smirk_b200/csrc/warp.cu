// Crop / warp front end and back end of the per-frame path.
//
// Replaces, for frames that are already on the device,
//   demo.py:97 / demo_video.py:128   cropped = skimage.transform.warp(image, tform.inverse, output_shape=(224,224),
//                                              preserve_range=True).astype(np.uint8)
//   demo.py:103-105                  cv2.cvtColor(BGR2RGB) -> torch [1,3,224,224] float / 255
//   demo_video.py:148-149            rendered -> (x * 255).astype(uint8) -> warp(rendered, tform, output_shape=(H, W),
//                                              preserve_range=True).astype(np.uint8)
// The reference does this on the CPU per frame (skimage's Cython `_warp_fast`) between a D2H and an H2D copy; here
// it is one gather-bilinear pass per direction.
//
// Arithmetic follows skimage 0.2x `_warp_fast` / `bilinear_interpolation` / `_clip_warp_output` for order = 1,
// mode = 'constant', cval = 0, clip = True, evaluated in float64 like the reference (preserve_range=True converts
// the uint8 frame to float64):
//   (c, r) = M (tfc, tfr, 1)            M = 3x3 inverse map, row-major float64; affine rows only (M[2] = 0 0 1)
//   minr = floor(r), maxr = ceil(r), dr = r - minr (same for c); pixels outside the source read cval = 0
//   top = (1 - dc) tl + dc tr ; bottom = (1 - dc) bl + dc br ; v = (1 - dr) top + dr bottom
//   clip: v = clamp(v, min(src), max(src)) unless v == cval and cval lies outside [min(src), max(src)]
//   .astype(np.uint8): truncation towards zero
// skimage is not installed in this image, so this restatement is pinned only by the properties in tests/ (identity,
// integer shifts, scipy cross-check away from the border): "parity unpinned" for the border/clip rules.
// HBM traffic: the gather touches each source texel it needs once through L2; output 150 KB (crop) or H*W*3 (back).
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace {

struct MinMax { unsigned int mn, mx; };

// per-frame min / max of the uint8 source over all channels (skimage clips per warp() call = per whole image)
__global__ void __launch_bounds__(256)
u8_minmax_kernel(const uint8_t* __restrict__ src, size_t n_per_frame, MinMax* __restrict__ mm) {
    smk::pdl_sync();
    const int b = blockIdx.y;
    const uint8_t* s = src + (size_t)b * n_per_frame;
    unsigned int mn = 255u, mx = 0u;
    // 16 bytes per load where aligned
    const size_t n16 = (((uintptr_t)s & 15) == 0) ? n_per_frame / 16 : 0;
    const uint4* s16 = reinterpret_cast<const uint4*>(s);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(s16 + i);
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned int t = (w[k] >> (8 * j)) & 255u; mn = min(mn, t); mx = max(mx, t); }
    }
    for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_frame; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned int t = s[i]; mn = min(mn, t); mx = max(mx, t);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[b].mn, mn); atomicMax(&mm[b].mx, mx); }
}

__global__ void minmax_init_kernel(MinMax* mm, int B) {
    smk::pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { mm[i].mn = 255u; mm[i].mx = 0u; }
}

// rendered float [B,3,S,S] -> uint8 [B,S,S,3]: (x * 255.0f).astype(uint8) in float32 like numpy (demo_video.py:148)
__global__ void __launch_bounds__(256)
f32chw_to_u8hwc_kernel(const float* __restrict__ in, int B, int S, uint8_t* __restrict__ out) {
    smk::pdl_sync();
    const size_t n = (size_t)B * S * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)S * S), p = i - b * (size_t)S * S;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fmul_rn(in[(b * 3 + c) * (size_t)S * S + p], 255.0f);
            out[i * 3 + c] = (uint8_t)(int)v;                      // C cast: truncation, values are in [0, 255]
        }
    }
}

// One thread per output pixel, all three channels (the four source texels are adjacent 3-byte groups).
//   OUT_F32 = false: dst uint8 [B, Hd, Wd, 3], channel order kept.
//   OUT_F32 = true : dst float [B, 3, Hd, Wd] = uint8 result / 255, channels reversed when swap_rb (BGR frame -> RGB).
template <bool OUT_F32>
__global__ void __launch_bounds__(256)
warp_bilinear_kernel(const uint8_t* __restrict__ src, int Hs, int Ws, const double* __restrict__ M, const MinMax* __restrict__ mm,
                     int Hd, int Wd, int swap_rb, void* __restrict__ dst) {
    smk::pdl_sync();
    const int b = blockIdx.z;
    const int tfc = blockIdx.x * 32 + (threadIdx.x & 31), tfr = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (tfc >= Wd || tfr >= Hd) return;
    const double* m = M + (size_t)b * 9;
    // _transform_affine: c = M00 x + M01 y + M02 (left to right, no contraction)
    const double c = __dadd_rn(__dadd_rn(__dmul_rn(m[0], (double)tfc), __dmul_rn(m[1], (double)tfr)), m[2]);
    const double r = __dadd_rn(__dadd_rn(__dmul_rn(m[3], (double)tfc), __dmul_rn(m[4], (double)tfr)), m[5]);
    const double fr = floor(r), fc = floor(c);
    const long long minr = (long long)fr, minc = (long long)fc, maxr = (long long)ceil(r), maxc = (long long)ceil(c);
    const double dr = __dsub_rn(r, (double)minr), dc = __dsub_rn(c, (double)minc);
    const uint8_t* s = src + (size_t)b * Hs * Ws * 3;
    const bool r0 = minr >= 0 && minr < Hs, r1 = maxr >= 0 && maxr < Hs, c0 = minc >= 0 && minc < Ws, c1 = maxc >= 0 && maxc < Ws;
    const double lo = (double)mm[b].mn, hi = (double)mm[b].mx;
    const bool keep_cval = !(lo <= 0.0 && 0.0 <= hi);              // cval = 0 outside the source's range: exact zeros survive the clip
    uint8_t res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double tl = (r0 && c0) ? (double)s[((size_t)minr * Ws + minc) * 3 + ch] : 0.0;
        const double tr = (r0 && c1) ? (double)s[((size_t)minr * Ws + maxc) * 3 + ch] : 0.0;
        const double bl = (r1 && c0) ? (double)s[((size_t)maxr * Ws + minc) * 3 + ch] : 0.0;
        const double br = (r1 && c1) ? (double)s[((size_t)maxr * Ws + maxc) * 3 + ch] : 0.0;
        const double omc = __dsub_rn(1.0, dc), omr = __dsub_rn(1.0, dr);
        const double top = __dadd_rn(__dmul_rn(omc, tl), __dmul_rn(dc, tr));
        const double bot = __dadd_rn(__dmul_rn(omc, bl), __dmul_rn(dc, br));
        double v = __dadd_rn(__dmul_rn(omr, top), __dmul_rn(dr, bot));
        if (!(keep_cval && v == 0.0)) v = fmin(fmax(v, lo), hi);
        res[ch] = (uint8_t)(int)v;
    }
    if (OUT_F32) {
        float* o = reinterpret_cast<float*>(dst) + (size_t)b * 3 * Hd * Wd + (size_t)tfr * Wd + tfc;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) o[(size_t)(swap_rb ? 2 - ch : ch) * Hd * Wd] = __fdiv_rn((float)res[ch], 255.0f);
    } else {
        uint8_t* o = reinterpret_cast<uint8_t*>(dst) + (((size_t)b * Hd + tfr) * Wd + tfc) * 3;
        o[0] = res[0]; o[1] = res[1]; o[2] = res[2];
    }
}

int launch_warp(const uint8_t* src, int B, int Hs, int Ws, const double* M, int Hd, int Wd, bool out_f32, int swap_rb, void* dst,
                void* ws, size_t ws_bytes, cudaStream_t st) {
    SMK_REQUIRE(ws && ws_bytes >= (size_t)B * sizeof(MinMax), "warp: workspace too small");
    MinMax* mm = reinterpret_cast<MinMax*>(ws);
    SMK_TAG("warp_minmax", (double)B * Hs * Ws * 3, 0.0, st);
    SMK_LAUNCH(minmax_init_kernel, dim3(smk::cdiv(B, 64)), dim3(64), 0, st, mm, B);
    SMK_CHECK_LAUNCH();
    const size_t n = (size_t)Hs * Ws * 3;
    SMK_TAG("warp_minmax", 0.0, 0.0, st);
    SMK_LAUNCH(u8_minmax_kernel, dim3((unsigned)std::min<size_t>((n / 16 + 255) / 256 + 1, 296), B), dim3(256), 0, st, src, n, mm);
    SMK_CHECK_LAUNCH();
    SMK_TAG("warp_bilinear", (double)B * Hd * Wd * (out_f32 ? 12.0 : 3.0) + (double)B * Hd * Wd * 12.0, 30.0 * B * Hd * Wd, st);
    dim3 grid(smk::cdiv(Wd, 32), smk::cdiv(Hd, 8), B);
    if (out_f32) SMK_LAUNCH((warp_bilinear_kernel<true>), grid, dim3(256), 0, st, src, Hs, Ws, M, (const MinMax*)mm, Hd, Wd, swap_rb, dst);
    else SMK_LAUNCH((warp_bilinear_kernel<false>), grid, dim3(256), 0, st, src, Hs, Ws, M, (const MinMax*)mm, Hd, Wd, swap_rb, dst);
    SMK_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" size_t smk_warp_workspace_bytes(int B) { return smk::ws_round((size_t)(B > 0 ? B : 1) * sizeof(MinMax)); }

extern "C" int smk_crop_warp(const uint8_t* frames, int B, int H, int W, const double* minv, int S, int swap_rb, float* out,
                             void* ws, size_t ws_bytes, void* stream) {
    if (B == 0) return 0;
    SMK_REQUIRE(frames && minv && out, "smk_crop_warp: null argument");
    SMK_REQUIRE(B > 0 && H > 0 && W > 0 && S > 0, "smk_crop_warp: bad sizes");
    return launch_warp(frames, B, H, W, minv, S, S, true, swap_rb ? 1 : 0, out, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int smk_warp_u8(const uint8_t* src, int B, int Hs, int Ws, const double* m, int Hd, int Wd, uint8_t* dst,
                           void* ws, size_t ws_bytes, void* stream) {
    if (B == 0) return 0;
    SMK_REQUIRE(src && m && dst, "smk_warp_u8: null argument");
    SMK_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, "smk_warp_u8: bad sizes");
    return launch_warp(src, B, Hs, Ws, m, Hd, Wd, false, 0, dst, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int smk_f32chw_to_u8hwc(const float* in, int B, int S, uint8_t* out, void* stream) {
    if (B == 0) return 0;
    SMK_REQUIRE(in && out && B > 0 && S > 0, "smk_f32chw_to_u8hwc: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    SMK_TAG("f32chw_to_u8hwc", 15.0 * B * S * S, 0.0, st);
    SMK_LAUNCH(f32chw_to_u8hwc_kernel, dim3((unsigned)std::min<size_t>(((size_t)B * S * S + 255) / 256, 148 * 8)), dim3(256), 0, st, in, B, S, out);
    SMK_CHECK_LAUNCH();
    return 0;
}
