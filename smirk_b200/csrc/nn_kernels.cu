// fp32 CUDA-core kernels for the encoder / generator (see nn_kernels.cuh).
#include "nn_kernels.cuh"

namespace smk {
namespace {

constexpr int BK = 16;
constexpr int NT = 256;

// ---------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution, BMxBN tile, BK = 16, 256 threads as a 16x16 grid of (BM/16)x(BN/16)
// register tiles; global->register prefetch of the next k-slab overlaps the FMAs of the current one.
template <int BM, int BN>
__global__ void __launch_bounds__(NT)
conv_gemm_kernel(ConvProblem p, int M) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int A_LD = BM + 4, B_LD = BN + 4;
    constexpr int A_PER = BM / 64;                 // float4 loads of A per thread per slab (BM*BK/4/NT)
    constexpr int B_PER = (BN * BK / 4 + NT - 1) / NT;
    __shared__ __align__(16) float As[BK][A_LD];
    __shared__ __align__(16) float Bs[BK][B_LD];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int HW = p.H * p.W;
    pdl_sync();

    // per-thread A rows: r = tid/4 + 64*j, k-quad kq = tid%4
    const int kq = tid & 3;
    int a_b[A_PER], a_h[A_PER], a_w[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
        int m = m0 + (tid >> 2) + 64 * j;
        a_ok[j] = m < M;
        int mm = a_ok[j] ? m : 0;
        a_b[j] = mm / HW; int r = mm - a_b[j] * HW; a_h[j] = r / p.W; a_w[j] = r - a_h[j] * p.W;
    }
    float4 a_reg[A_PER], b_reg[B_PER];

    auto load_slab = [&](int k0) {
        const int k = k0 + kq * 4;
#pragma unroll
        for (int j = 0; j < A_PER; ++j) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[j] && k < p.K) {
                if (p.mode == 0) {
                    v = *reinterpret_cast<const float4*>(p.in + ((size_t)(a_b[j] * HW + a_h[j] * p.W + a_w[j])) * p.ld_in + k);
                } else {
                    int tap = k / p.Cin, c = k - tap * p.Cin;
                    int ky = tap / 3, kx = tap - ky * 3;
                    int hh = a_h[j] + ky - 1, ww = a_w[j] + kx - 1;
                    bool inside = hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
                    if (p.mode == 2) {                                  // ReflectionPad2d(1)
                        hh = hh < 0 ? 1 : (hh >= p.H ? p.H - 2 : hh);
                        ww = ww < 0 ? 1 : (ww >= p.W ? p.W - 2 : ww);
                        inside = true;
                    }
                    if (inside)
                        v = *reinterpret_cast<const float4*>(p.in + ((size_t)(a_b[j] * p.H + hh) * p.W + ww) * p.ld_in + c);
                }
            }
            a_reg[j] = v;
        }
#pragma unroll
        for (int j = 0; j < B_PER; ++j) {
            int idx = tid + j * NT;                     // float4 index within the BK x BN slab
            int kk = idx / (BN / 4), nq = idx - kk * (BN / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < BK && k0 + kk < p.K && n0 + nq * 4 < p.N)
                v = *reinterpret_cast<const float4*>(p.w + (size_t)(k0 + kk) * p.N + n0 + nq * 4);
            b_reg[j] = v;
        }
    };
    auto store_slab = [&]() {
#pragma unroll
        for (int j = 0; j < A_PER; ++j) {
            int r = (tid >> 2) + 64 * j;
            As[kq * 4 + 0][r] = a_reg[j].x; As[kq * 4 + 1][r] = a_reg[j].y;
            As[kq * 4 + 2][r] = a_reg[j].z; As[kq * 4 + 3][r] = a_reg[j].w;
        }
#pragma unroll
        for (int j = 0; j < B_PER; ++j) {
            int idx = tid + j * NT;
            int kk = idx / (BN / 4), nq = idx - kk * (BN / 4);
            if (kk < BK) *reinterpret_cast<float4*>(&Bs[kk][nq * 4]) = b_reg[j];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_slab(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        __syncthreads();
        store_slab();
        __syncthreads();
        if (k0 + BK < p.K) load_slab(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
        size_t opix = m;
        int nbase = n0 + tx * TN;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = nbase + j;
            if (n >= p.N) continue;
            float v = fmaf(acc[i][j], p.scale[n], p.bias[n]);
            if (p.res) v += p.res[(size_t)m * p.ld_res + n];
            if (p.relu) v = fmaxf(v, 0.f);
            if (p.round_out) v = round_tf32(v);
            if (p.shuffle) {
                int cout = p.N >> 2, q = n / cout, co = n - q * cout;
                int b = m / HW, r = m - b * HW, h = r / p.W, w = r - h * p.W;
                size_t dp = ((size_t)b * (2 * p.H) + 2 * h + (q >> 1)) * (2 * p.W) + 2 * w + (q & 1);
                p.out[dp * p.ld_out + co] = v;
            } else {
                p.out[opix * p.ld_out + n] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const float* __restrict__ in, int B, int H, int W, int C, int stride, int pad, int Ho, int Wo,
                 const float* __restrict__ w9c, const float* __restrict__ scale, const float* __restrict__ bias,
                 float* __restrict__ out) {
    pdl_sync();
    const int C4 = C >> 2;
    const long total = (long)B * Ho * Wo * C4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4); long pix = i / C4;
        int ow = (int)(pix % Wo); long t = pix / Wo; int oh = (int)(t % Ho); int b = (int)(t / Ho);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            int ih = oh * stride + ky - pad;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                int iw = ow * stride + kx - pad;
                if (iw < 0 || iw >= W) continue;
                float4 x = *reinterpret_cast<const float4*>(in + (((size_t)b * H + ih) * W + iw) * C + c4 * 4);
                float4 k = *reinterpret_cast<const float4*>(w9c + (size_t)(ky * 3 + kx) * C + c4 * 4);
                acc.x = fmaf(x.x, k.x, acc.x); acc.y = fmaf(x.y, k.y, acc.y);
                acc.z = fmaf(x.z, k.z, acc.z); acc.w = fmaf(x.w, k.w, acc.w);
            }
        }
        float4 s = *reinterpret_cast<const float4*>(scale + c4 * 4), bb = *reinterpret_cast<const float4*>(bias + c4 * 4);
        float4 o;
        o.x = fmaxf(fmaf(acc.x, s.x, bb.x), 0.f); o.y = fmaxf(fmaf(acc.y, s.y, bb.y), 0.f);
        o.z = fmaxf(fmaf(acc.z, s.z, bb.z), 0.f); o.w = fmaxf(fmaf(acc.w, s.w, bb.w), 0.f);
        *reinterpret_cast<float4*>(out + (size_t)pix * C + c4 * 4) = o;
    }
}

// Depthwise 3x3, 4 output pixels along W x 4 channels per thread: the 3 x (3 + 3*STRIDE) input window is
// loaded once (float4 per pixel) and reused by the 4 outputs; weights stay in registers.
template <int STRIDE>
__global__ void __launch_bounds__(256)
dwconv3x3_px4_kernel(const float* __restrict__ in, int B, int H, int W, int C, int pad, int Ho, int Wo,
                     const float* __restrict__ w9c, const float* __restrict__ scale, const float* __restrict__ bias,
                     float* __restrict__ out, int round_out) {
    constexpr int PX = 4, NC = 3 + (PX - 1) * STRIDE;
    const int C4 = C >> 2, WG = (Wo + PX - 1) / PX;
    const long total = (long)B * Ho * WG * C4;
    pdl_sync();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4); long t = i / C4;
        int wg = (int)(t % WG); t /= WG; int oh = (int)(t % Ho); int b = (int)(t / Ho);
        const int ow0 = wg * PX, iw0 = ow0 * STRIDE - pad;
        float4 k[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) k[q] = __ldg(reinterpret_cast<const float4*>(w9c + (size_t)q * C) + c4);
        float4 acc[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            int ih = oh * STRIDE + ky - pad;
            if (ih < 0 || ih >= H) continue;
            const float4* row = reinterpret_cast<const float4*>(in + ((size_t)b * H + ih) * W * C) + c4;
            float4 x[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                int iw = iw0 + j;
                x[j] = (iw >= 0 && iw < W) ? __ldg(row + (size_t)iw * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int p = 0; p < PX; ++p)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 xv = x[p * STRIDE + kx], kv = k[ky * 3 + kx];
                    acc[p].x = fmaf(xv.x, kv.x, acc[p].x); acc[p].y = fmaf(xv.y, kv.y, acc[p].y);
                    acc[p].z = fmaf(xv.z, kv.z, acc[p].z); acc[p].w = fmaf(xv.w, kv.w, acc[p].w);
                }
        }
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale) + c4), bb = __ldg(reinterpret_cast<const float4*>(bias) + c4);
        float4* orow = reinterpret_cast<float4*>(out + (((size_t)b * Ho + oh) * Wo) * C) + c4;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            if (ow0 + p >= Wo) break;
            float4 o;
            o.x = fmaxf(fmaf(acc[p].x, s.x, bb.x), 0.f); o.y = fmaxf(fmaf(acc[p].y, s.y, bb.y), 0.f);
            o.z = fmaxf(fmaf(acc[p].z, s.z, bb.z), 0.f); o.w = fmaxf(fmaf(acc[p].w, s.w, bb.w), 0.f);
            if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
            orow[(size_t)(ow0 + p) * C4] = o;
        }
    }
}

// Global average pool, split over channel chunks so the grid fills the machine: block = 64 channels x 4
// pixel lanes; pooled[b][c] = mean over HW.
__global__ void __launch_bounds__(256)
gap_kernel(const float* __restrict__ feat, int HW, int C, float* __restrict__ pooled) {
    __shared__ float part[4][64];
    pdl_sync();
    const int b = blockIdx.x, c = blockIdx.y * 64 + (threadIdx.x & 63), pl = threadIdx.x >> 6;
    float s = 0.f;
    if (c < C) {
        const float* f = feat + (size_t)b * HW * C + c;
        for (int p = pl; p < HW; p += 4) s += f[(size_t)p * C];
    }
    part[pl][threadIdx.x & 63] = s;
    __syncthreads();
    if (pl == 0 && c < C)
        pooled[(size_t)b * C + c] = (part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x]) * (1.f / (float)HW);
}

// out[b][o] = clamp(bias[o] + pooled[b] . w[o]); one warp per output, 8 outputs per block.
__global__ void __launch_bounds__(256)
head_linear_kernel(const float* __restrict__ pooled, int C, const float* __restrict__ w, const float* __restrict__ bias,
                   int n_out, const uint8_t* __restrict__ codes, float* __restrict__ out) {
    pdl_sync();
    const int b = blockIdx.x, lane = threadIdx.x & 31, o = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (o >= n_out) return;
    const float* wr = w + (size_t)o * C;
    const float* pb = pooled + (size_t)b * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(pb[c], wr[c], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        float v = acc + bias[o];
        int code = codes ? codes[o] : 0;
        if (code == 1) v = fminf(fmaxf(v, 0.f), 1.f);
        else if (code == 2) v = fmaxf(v, 0.f);
        else if (code == 3) v = fminf(fmaxf(v, -0.2f), 0.2f);
        out[(size_t)b * n_out + o] = v;
    }
}

// Global average pool + linear head + clamps in ONE launch for one or two backbones (blockIdx.z): every CTA pools its
// image's feature map into shared memory (the map is L2-resident: 188 KB at 7x7x960) and computes 32 head outputs, one
// warp per output (smirk_encoder.py:34-45,66-73,95-110).  Replaces gap_kernel + head_linear_kernel: 6 launches per
// encoder pass become 2.
struct GapHead { const float* feat[2]; const float* w[2]; const float* bias[2]; const uint8_t* codes[2]; float* out[2]; int n_out[2]; };
__global__ void __launch_bounds__(256)
gap_head_kernel(const __grid_constant__ GapHead g, int HW, int C) {
    extern __shared__ float pooled[];                 // [C]
    pdl_sync();
    const int q = blockIdx.z, b = blockIdx.x;
    const int n_out = g.n_out[q];
    if ((int)blockIdx.y * 32 >= n_out) return;
    const float* f = g.feat[q] + (size_t)b * HW * C;
    const float inv = 1.f / (float)HW;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int p = 0;
        for (; p + 4 <= HW; p += 4) {
            s0 += f[(size_t)p * C + c]; s1 += f[(size_t)(p + 1) * C + c]; s2 += f[(size_t)(p + 2) * C + c]; s3 += f[(size_t)(p + 3) * C + c];
        }
        for (; p < HW; ++p) s0 += f[(size_t)p * C + c];
        pooled[c] = ((s0 + s1) + (s2 + s3)) * inv;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* w = g.w[q]; const float* bias = g.bias[q]; const uint8_t* codes = g.codes[q];
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const int o = blockIdx.y * 32 + warp * 4 + j;
        if (o >= n_out) break;
        const float* wr = w + (size_t)o * C;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) acc = fmaf(pooled[c], __ldg(wr + c), acc);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            float v = acc + bias[o];
            const int code = codes ? codes[o] : 0;
            if (code == 1) v = fminf(fmaxf(v, 0.f), 1.f);
            else if (code == 2) v = fmaxf(v, 0.f);
            else if (code == 3) v = fminf(fmaxf(v, -0.2f), 0.2f);
            g.out[q][(size_t)b * n_out + o] = v;
        }
    }
}

__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ img, int B, int H, int W, int Ho, int Wo, int pad,
                 const float* __restrict__ w /*[27][16]*/, const float* __restrict__ scale,
                 const float* __restrict__ bias, float* __restrict__ out) {
    __shared__ float sw[27 * 16];
    __shared__ float ss[16], sb[16];
    for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 16) { ss[threadIdx.x] = scale[threadIdx.x]; sb[threadIdx.x] = bias[threadIdx.x]; }
    __syncthreads();
    pdl_sync();
    long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long)B * Ho * Wo) return;
    int ow = (int)(pix % Wo); long t = pix / Wo; int oh = (int)(t % Ho); int b = (int)(t / Ho);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            int ih = oh * 2 + ky - pad;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                int iw = ow * 2 + kx - pad;
                float x = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? img[(((size_t)b * 3 + c) * H + ih) * W + iw] : 0.f;
                const float* wk = sw + ((c * 3 + ky) * 3 + kx) * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(x, wk[i], acc[i]);
            }
        }
    float4* o = reinterpret_cast<float4*>(out + (size_t)pix * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 v;
        v.x = fmaxf(fmaf(acc[q * 4 + 0], ss[q * 4 + 0], sb[q * 4 + 0]), 0.f);
        v.y = fmaxf(fmaf(acc[q * 4 + 1], ss[q * 4 + 1], sb[q * 4 + 1]), 0.f);
        v.z = fmaxf(fmaf(acc[q * 4 + 2], ss[q * 4 + 2], sb[q * 4 + 2]), 0.f);
        v.w = fmaxf(fmaf(acc[q * 4 + 3], ss[q * 4 + 3], sb[q * 4 + 3]), 0.f);
        o[q] = v;
    }
}

// The three backbones' stems read the same image: one pass over the pixels produces all 3 x 16 channels.
struct Stem3 { const float* w[3]; const float* scale[3]; const float* bias[3]; float* out[3]; };

// CTA = one output row (b, oh): the 3 channels x 3 input rows it needs are staged in shared memory with
// coalesced 16-byte loads (the direct version issued 27 strided 4-byte loads per thread and was
// latency-bound at 22 % occupancy); thread = output pixel, 48 accumulators.
constexpr int STEM_MAXW = 512;                     // staged row length (floats), >= W + 2
__global__ void __launch_bounds__(128)
stem_conv3_kernel(const float* __restrict__ img, int B, int H, int W, int Ho, int Wo, int pad, Stem3 p) {
    __shared__ float sw[27 * 48];                  // [tap][group*16 + co]
    __shared__ float ss[48], sb[48];
    __shared__ __align__(16) float sin_[9][STEM_MAXW];   // [c*3+ky][1 + iw]  (column 0 = left padding)
    const int oh = blockIdx.x % Ho, b = blockIdx.x / Ho;
    for (int i = threadIdx.x; i < 27 * 48; i += blockDim.x) { int tap = i / 48, gc = i % 48; sw[i] = p.w[gc / 16][tap * 16 + gc % 16]; }
    if (threadIdx.x < 48) { ss[threadIdx.x] = p.scale[threadIdx.x / 16][threadIdx.x % 16]; sb[threadIdx.x] = p.bias[threadIdx.x / 16][threadIdx.x % 16]; }
    pdl_sync();                                    // weights are constants; the image and the outputs are not
    {   // stage rows: element iw of row (c,ky) lands at sin_[c*3+ky][4 + iw] so 16-byte stores stay aligned; halo columns zeroed
        const int W4 = W >> 2;
        for (int i = threadIdx.x; i < 9 * W4; i += blockDim.x) {
            const int r = i / W4, q = i - r * W4, c = r / 3, ky = r - c * 3;
            const int ih = oh * 2 + ky - pad;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ih >= 0 && ih < H) v = __ldg(reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * H + ih) * W) + q);
            *reinterpret_cast<float4*>(&sin_[r][4 + 4 * q]) = v;
        }
        if (threadIdx.x < 9) { sin_[threadIdx.x][3] = 0.f; sin_[threadIdx.x][4 + W] = 0.f; sin_[threadIdx.x][5 + W] = 0.f; }
    }
    __syncthreads();
    for (int ow = threadIdx.x; ow < Wo; ow += blockDim.x) {
    const long pix = ((long)b * Ho + oh) * Wo + ow;
    float acc[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float x = sin_[c * 3 + ky][4 + ow * 2 + kx - pad];
                const float* wk = sw + ((c * 3 + ky) * 3 + kx) * 48;
#pragma unroll
                for (int i = 0; i < 48; ++i) acc[i] = fmaf(x, wk[i], acc[i]);
            }
        }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        float4* o = reinterpret_cast<float4*>(p.out[g] + (size_t)pix * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = g * 16 + q * 4;
            float4 v;
            v.x = fmaxf(fmaf(acc[i + 0], ss[i + 0], sb[i + 0]), 0.f); v.y = fmaxf(fmaf(acc[i + 1], ss[i + 1], sb[i + 1]), 0.f);
            v.z = fmaxf(fmaf(acc[i + 2], ss[i + 2], sb[i + 2]), 0.f); v.w = fmaxf(fmaf(acc[i + 3], ss[i + 3], sb[i + 3]), 0.f);
            o[q] = v;
        }
    }
    }   // ow
}

__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const float* __restrict__ in, int ld_in, int B, int H, int W, int C, float* __restrict__ out) {
    const int Ho = H >> 1, Wo = W >> 1, C4 = C >> 2;
    const long total = (long)B * Ho * Wo * C4;
    pdl_sync();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4); long pix = i / C4;
        int ow = (int)(pix % Wo); long t = pix / Wo; int oh = (int)(t % Ho); int b = (int)(t / Ho);
        const float* p00 = in + (((size_t)b * H + 2 * oh) * W + 2 * ow) * ld_in + c4 * 4;
        float4 a = *reinterpret_cast<const float4*>(p00), bq = *reinterpret_cast<const float4*>(p00 + ld_in);
        float4 c = *reinterpret_cast<const float4*>(p00 + (size_t)W * ld_in), d = *reinterpret_cast<const float4*>(p00 + (size_t)W * ld_in + ld_in);
        float4 o;
        o.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(c.x, d.x)); o.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(c.y, d.y));
        o.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(c.z, d.z)); o.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(c.w, d.w));
        *reinterpret_cast<float4*>(out + (size_t)pix * C + c4 * 4) = o;
    }
}

__global__ void __launch_bounds__(256)
nchw_to_nhwc_pad_kernel(const float* __restrict__ in, int B, int C, int HW, int Cp, int round, float* __restrict__ out) {
    // thread = (pixel, channel quad), quads of a pixel in adjacent lanes: every warp stores 512 contiguous bytes
    pdl_sync();
    const int Q = Cp >> 2;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * HW * Q) return;
    const int q = (int)(i % Q); const long pix = i / Q;
    const int b = (int)(pix / HW); const int r = (int)(pix - (long)b * HW);
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = q * 4 + k;
        v[k] = c < C ? __ldg(in + ((size_t)b * C + c) * HW + r) : 0.f;
        if (round) v[k] = round_tf32(v[k]);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(256)
conv1x1_sigmoid_kernel(const float* __restrict__ in, int B, int HW, int Cin, const float* __restrict__ w,
                       const float* __restrict__ bias, int Cout, float* __restrict__ out) {
    extern __shared__ float sw[];                 // [Cin][Cout] + [Cout]
    for (int i = threadIdx.x; i < Cin * Cout; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[Cin * Cout + i] = bias[i];
    __syncthreads();
    pdl_sync();
    long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long)B * HW) return;
    int b = (int)(pix / HW); int r = (int)(pix - (long)b * HW);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* x4 = reinterpret_cast<const float4*>(in + (size_t)pix * Cin);
    for (int c4 = 0; c4 < Cin / 4; ++c4) {
        float4 x = x4[c4];
        float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            for (int co = 0; co < Cout; ++co) acc[co] = fmaf(xs[q], sw[(c4 * 4 + q) * Cout + co], acc[co]);
    }
    for (int co = 0; co < Cout; ++co) {
        float v = acc[co] + sw[Cin * Cout + co];
        out[((size_t)b * Cout + co) * HW + r] = 1.f / (1.f + expf(-v));
    }
}

}  // namespace

int conv_gemm(const ConvProblem& p, cudaStream_t st) {
    const int M = p.B * p.H * p.W;
    SMK_REQUIRE(p.K % 4 == 0 && p.N % 4 == 0 && p.ld_in % 4 == 0, "conv_gemm: K, N, ld_in must be multiples of 4");
    SMK_REQUIRE(p.mode == 0 || p.Cin % 4 == 0, "conv_gemm: Cin must be a multiple of 4");
    {
        const double cin_eff = p.mode == 0 ? p.K : p.Cin;       // unique input bytes (not im2col-expanded)
        const char* tag = p.mode == 0 ? (p.shuffle ? "upconv_gemm_f32" : "pw_gemm_f32") : "conv3x3_gemm_f32";
        if (g_prof_detail) tag = prof_shape_tag(tag, M, p.K, p.N);
        SMK_TAG(tag,
                4.0 * ((double)M * cin_eff + (double)p.K * p.N + (double)M * p.N * (p.res ? 2 : 1) + 2.0 * p.N),
                2.0 * (double)M * p.N * p.K, st);
    }
    if (p.N <= 32) {
        dim3 grid(cdiv(M, 128), cdiv(p.N, 32));
        SMK_LAUNCH((conv_gemm_kernel<128, 32>), dim3(grid), dim3(NT), 0, st, p, M);
    } else {
        dim3 grid(cdiv(M, 64), cdiv(p.N, 64));
        SMK_LAUNCH((conv_gemm_kernel<64, 64>), dim3(grid), dim3(NT), 0, st, p, M);
    }
    SMK_CHECK_LAUNCH();
    return 0;
}

static inline int same_pad_begin(int H, int stride) {
    int out = (H + stride - 1) / stride;
    int total = (out - 1) * stride + 3 - H;
    if (total < 0) total = 0;
    return total / 2;
}

int dwconv3x3(const float* in, int B, int H, int W, int C, int stride, const float* w9c, const float* scale,
              const float* bias, float* out, cudaStream_t st, bool round_out) {
    SMK_REQUIRE(C % 4 == 0, "dwconv3x3: C must be a multiple of 4");
    int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    SMK_REQUIRE(stride == 1 || stride == 2, "dwconv3x3: stride must be 1 or 2");
    long total = (long)B * Ho * ((Wo + 3) / 4) * (C / 4);
    int blocks = (int)std::min<long>((total + 255) / 256, 148L * 32);
    SMK_TAG(g_prof_detail ? prof_shape_tag("dwconv3x3", (long)B * Ho * Wo, stride, C) : "dwconv3x3", 4.0 * ((double)B * H * W * C + (double)B * Ho * Wo * C + 11.0 * C), 18.0 * B * Ho * Wo * C, st);
    if (stride == 1)
        SMK_LAUNCH((dwconv3x3_px4_kernel<1>), dim3(blocks), dim3(256), 0, st, in, B, H, W, C, same_pad_begin(H, 1), Ho, Wo, w9c, scale, bias, out, round_out ? 1 : 0);
    else
        SMK_LAUNCH((dwconv3x3_px4_kernel<2>), dim3(blocks), dim3(256), 0, st, in, B, H, W, C, same_pad_begin(H, 2), Ho, Wo, w9c, scale, bias, out, round_out ? 1 : 0);
    SMK_CHECK_LAUNCH();
    return 0;
}

int stem_conv(const float* img, int B, int H, int W, const float* w, const float* scale, const float* bias, float* out,
              cudaStream_t st) {
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    SMK_TAG("stem_conv", 4.0 * ((double)B * 3 * H * W + (double)B * Ho * Wo * 16 + 27 * 16 + 32), 2.0 * 27 * 16 * (double)B * Ho * Wo, st);
    SMK_LAUNCH(stem_conv_kernel, dim3(cdiv((long)B * Ho * Wo, 128)), dim3(128), 0, st, img, B, H, W, Ho, Wo, same_pad_begin(H, 2), w, scale, bias, out);
    SMK_CHECK_LAUNCH();
    return 0;
}

int stem_conv3(const float* img, int B, int H, int W, const float* const w[3], const float* const scale[3],
               const float* const bias[3], float* const out[3], cudaStream_t st) {
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    Stem3 p;
    for (int g = 0; g < 3; ++g) { p.w[g] = w[g]; p.scale[g] = scale[g]; p.bias[g] = bias[g]; p.out[g] = out[g]; }
    SMK_TAG("stem_conv3", 4.0 * ((double)B * 3 * H * W + (double)B * Ho * Wo * 48 + 27 * 48 + 96), 2.0 * 27 * 48 * (double)B * Ho * Wo, st);
    SMK_REQUIRE(W % 4 == 0 && W + 8 <= STEM_MAXW && same_pad_begin(H, 2) <= 1, "stem_conv3: unsupported image width %d", W);
    SMK_LAUNCH(stem_conv3_kernel, dim3(B * Ho), dim3(128), 0, st, img, B, H, W, Ho, Wo, same_pad_begin(H, 2), p);
    SMK_CHECK_LAUNCH();
    return 0;
}

// ---- fused stem + first block --------------------------------------------------------------------------
// conv_stem 3x3 s2 (3 -> 16) + BN + ReLU  ->  block 0 of tf_mobilenetv3_*_minimal_100, a depthwise-separable
// block (depthwise 3x3 s{1,2} + BN + ReLU -> 1x1 16 -> 16 + BN, + skip when stride 1); reference
// src/smirk_encoder.py:7-12 -> timm MobileNetV3.conv_stem/bn1 + DepthwiseSeparableConv.  These are the three
// largest activations of a backbone (112 x 112 x 16); unfused they cost a 96 MB stem pass plus a depthwise and
// a 1x1 kernel per backbone.  Here a CTA owns a 16 x 16 tile of the stem output (+1 halo = 18 x 18):
//   1. stage the 37 x 37 x 3 image patch in shared memory, even/odd columns de-interleaved so the stride-2 taps
//      of neighbouring pixels hit neighbouring banks;
//   2. stem conv: thread = two stem pixels x 16 channels (the tap weights, broadcast from shared memory, are
//      used twice), BN + ReLU, zero outside the 112 x 112 map (= the depthwise conv's zero padding) -> S;
//   3. depthwise 3x3 over S: thread = (pixel, channel quad), taps in registers, BN + ReLU -> D (aliases the patch);
//   4. 1x1 conv on D: thread = (pixel, output-channel quad), 16 x 4 weights in registers, BN, + S (skip),
//      optional TF32 rounding, one coalesced 16-byte store per thread.
// All arithmetic is fp32 FMA in the reference's tap order; the image is the only HBM read, the block output the
// only write.
namespace {
constexpr int SD_T = 16, SD_ST = SD_T + 2;                 // stem-resolution tile edge, with halo
constexpr int SD_PR = 2 * SD_ST + 1;                       // image patch rows / cols (37)
constexpr int SD_PP = 20;                                  // patch row pitch per column parity (19 even + 18 odd columns)
struct StemDs {                              // one or two backbones share a launch (blockIdx.z selects; both read the same image)
    StemDsProblem q[2];
    int round_out;
};

template <int STRIDE>
__global__ void __launch_bounds__(256, 3)
stem_ds_kernel(const float* __restrict__ img, int H, int W, int Hs, int Ws, int pad, const __grid_constant__ StemDs pp) {
    struct { const float *stem_w, *stem_s, *stem_b, *dw_w, *dw_s, *dw_b, *pw_w, *pw_s, *pw_b; float* out; int round_out; } p;
    {
        const StemDsProblem& q = pp.q[blockIdx.z];
        p.stem_w = q.stem_w; p.stem_s = q.stem_s; p.stem_b = q.stem_b; p.dw_w = q.dw_w; p.dw_s = q.dw_s; p.dw_b = q.dw_b;
        p.pw_w = q.pw_w; p.pw_s = q.pw_s; p.pw_b = q.pw_b; p.out = q.out; p.round_out = pp.round_out;
    }
    constexpr int PADO = STRIDE == 1 ? 1 : 0;              // depthwise TF-SAME pad_begin on an even-sized map
    constexpr int TO = SD_T / STRIDE;                      // output tile edge
    __shared__ __align__(16) float sP[3 * SD_PR * 2 * SD_PP];          // image patch [c][row][parity][col/2]; later D [256][16]
    __shared__ __align__(16) float sS[SD_ST * SD_ST * 16];             // stem tile [pixel][16]
    __shared__ __align__(16) float sW[27 * 16];
    __shared__ __align__(16) float sK[9 * 16 + 16 * 16];               // depthwise taps, 1x1 weights
    __shared__ __align__(16) float sB[6 * 16];                         // stem / dw / pw scale, bias
    const int tid = threadIdx.x, b = blockIdx.y;
    const int tiles_x = Ws / SD_T;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int sy0 = ty * SD_T - PADO, sx0 = tx * SD_T - PADO;          // stem-tile origin
    const int iy0 = 2 * sy0 - pad, ix0 = 2 * sx0 - pad;                // image-patch origin
    for (int i = tid; i < 27 * 16; i += 256) sW[i] = p.stem_w[i];
    for (int i = tid; i < 9 * 16; i += 256) sK[i] = p.dw_w[i];
    sK[9 * 16 + tid] = p.pw_w[tid];
    if (tid < 16) {
        sB[tid] = p.stem_s[tid]; sB[16 + tid] = p.stem_b[tid]; sB[32 + tid] = p.dw_s[tid]; sB[48 + tid] = p.dw_b[tid];
        sB[64 + tid] = p.pw_s[tid]; sB[80 + tid] = p.pw_b[tid];
    }
    pdl_sync();                                            // weights are constants; the image and the outputs are not
    // ix0 and W are even: one 8-byte load fetches an (even, odd) column pair — exactly the de-interleaved layout.
    // 19 pairs per row cover the 37 columns (the odd half of the last pair is never read).
    constexpr int SD_PAIRS = (SD_PR + 1) / 2;
#pragma unroll 4
    for (int i = tid; i < 3 * SD_PR * SD_PAIRS; i += 256) {
        const int row = i / SD_PAIRS, j = i - row * SD_PAIRS, c = row / SD_PR, r = row - c * SD_PR;
        const int iy = iy0 + r, ix = ix0 + 2 * j;
        float2 v = make_float2(0.f, 0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(reinterpret_cast<const float2*>(img + (((size_t)b * 3 + c) * H + iy) * W + ix));
        sP[(row * 2) * SD_PP + j] = v.x;
        sP[(row * 2 + 1) * SD_PP + j] = v.y;
    }
    __syncthreads();
    // -- 2. stem conv: pixels (syp, sx) and (syp + 9, sx)
    if (tid < SD_ST * SD_ST / 2) {
        const int syp = tid / SD_ST, sx = tid - syp * SD_ST;
        float4 a0[4], a1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a0[q] = a1[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int off = (kx & 1) * SD_PP + sx + (kx >> 1);
                    const float x0 = sP[(c * SD_PR + 2 * syp + ky) * 2 * SD_PP + off];
                    const float x1 = sP[(c * SD_PR + 2 * (syp + SD_ST / 2) + ky) * 2 * SD_PP + off];
                    const float4* wk = reinterpret_cast<const float4*>(sW + ((c * 3 + ky) * 3 + kx) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const float4 w4 = wk[q]; fma4_s(a0[q], x0, w4); fma4_s(a1[q], x1, w4); }
                }
        const bool in_x = sx0 + sx >= 0 && sx0 + sx < Ws;
        const bool v0 = in_x && sy0 + syp >= 0 && sy0 + syp < Hs;
        const bool v1 = in_x && sy0 + syp + SD_ST / 2 >= 0 && sy0 + syp + SD_ST / 2 < Hs;
        float4* s0 = reinterpret_cast<float4*>(sS + (syp * SD_ST + sx) * 16);
        float4* s1 = reinterpret_cast<float4*>(sS + ((syp + SD_ST / 2) * SD_ST + sx) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 sc = *reinterpret_cast<const float4*>(sB + 4 * q), bi = *reinterpret_cast<const float4*>(sB + 16 + 4 * q);
            float4 o0 = fma4(a0[q], sc, bi), o1 = fma4(a1[q], sc, bi);
            o0.x = v0 ? fmaxf(o0.x, 0.f) : 0.f; o0.y = v0 ? fmaxf(o0.y, 0.f) : 0.f; o0.z = v0 ? fmaxf(o0.z, 0.f) : 0.f; o0.w = v0 ? fmaxf(o0.w, 0.f) : 0.f;
            o1.x = v1 ? fmaxf(o1.x, 0.f) : 0.f; o1.y = v1 ? fmaxf(o1.y, 0.f) : 0.f; o1.z = v1 ? fmaxf(o1.z, 0.f) : 0.f; o1.w = v1 ? fmaxf(o1.w, 0.f) : 0.f;
            s0[q] = o0; s1[q] = o1;
        }
    }
    __syncthreads();
    // -- 3. depthwise 3x3: thread = (pixel, channel quad)
    float* sD = sP;
    const int q = tid & 3;
    {
        float4 k[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) k[t] = *reinterpret_cast<const float4*>(sK + t * 16 + 4 * q);
        const float4 sc = *reinterpret_cast<const float4*>(sB + 32 + 4 * q), bi = *reinterpret_cast<const float4*>(sB + 48 + 4 * q);
#pragma unroll
        for (int j = 0; j < TO * TO / 64; ++j) {
            const int px = (tid >> 2) + 64 * j, oy = px / TO, ox = px - oy * TO;
            const float* s = sS + ((oy * STRIDE) * SD_ST + ox * STRIDE) * 16 + 4 * q;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                    fma4_acc(acc, *reinterpret_cast<const float4*>(s + (ky * SD_ST + kx) * 16), k[ky * 3 + kx]);
            float4 o = fma4(acc, sc, bi);
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            *reinterpret_cast<float4*>(sD + px * 16 + 4 * q) = o;
        }
    }
    __syncthreads();
    // -- 4. 1x1 conv: thread = (pixel, output-channel quad)
    {
        constexpr int NJ = TO * TO / 64;                   // pixels per thread
        const float4 sc = *reinterpret_cast<const float4*>(sB + 64 + 4 * q), bi = *reinterpret_cast<const float4*>(sB + 80 + 4 * q);
        const int Ho = Hs / STRIDE, Wo = Ws / STRIDE;
        float4 accs[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) accs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < 4; ++g) {                      // four input channels at a time: their weights serve all NJ pixels
            float4 w[4];
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) w[ci] = *reinterpret_cast<const float4*>(sK + 9 * 16 + (4 * g + ci) * 16 + 4 * q);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 d = *reinterpret_cast<const float4*>(sD + ((tid >> 2) + 64 * j) * 16 + 4 * g);
                fma4_s(accs[j], d.x, w[0]); fma4_s(accs[j], d.y, w[1]); fma4_s(accs[j], d.z, w[2]); fma4_s(accs[j], d.w, w[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int px = (tid >> 2) + 64 * j, oy = px / TO, ox = px - oy * TO;
            float4 o = fma4(accs[j], sc, bi);
            if (STRIDE == 1) {                             // skip connection: block input = stem output at the same pixel
                const float4 r = *reinterpret_cast<const float4*>(sS + ((oy + 1) * SD_ST + ox + 1) * 16 + 4 * q);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
            const int oh = ty * TO + oy, ow = tx * TO + ox;
            *reinterpret_cast<float4*>(p.out + (((size_t)b * Ho + oh) * Wo + ow) * 16 + 4 * q) = o;
        }
    }
}
}  // namespace

int stem_ds(const float* img, int B, int H, int W, const StemDsProblem* probs, int n, int stride, int round_out, cudaStream_t st) {
    const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
    SMK_REQUIRE(n == 1 || n == 2, "stem_ds: one or two backbones per launch");
    SMK_REQUIRE(stride == 1 || stride == 2, "stem_ds: stride must be 1 or 2");
    SMK_REQUIRE(H % 2 == 0 && W % 2 == 0 && Hs % SD_T == 0 && Ws % SD_T == 0, "stem_ds: image size %dx%d must be a multiple of 32", H, W);
    SMK_REQUIRE(((uintptr_t)img & 7) == 0, "stem_ds: image pointer must be 8-byte aligned");
    StemDs p{};
    p.q[0] = probs[0]; p.q[1] = probs[n - 1]; p.round_out = round_out;
    const double px_o = (double)B * (Hs / stride) * (Ws / stride);
    SMK_TAG("stem_ds_fused", 4.0 * ((double)B * 3 * H * W + n * (px_o * 16 + 27 * 16 + 9 * 16 + 256 + 96)),
            n * 2.0 * ((double)B * Hs * Ws * 16 * 27 + px_o * 16 * (9 + 16)), st);
    dim3 grid((Hs / SD_T) * (Ws / SD_T), B, n);
    if (stride == 1) SMK_LAUNCH((stem_ds_kernel<1>), grid, dim3(256), 0, st, img, H, W, Hs, Ws, same_pad_begin(H, 2), p);
    else SMK_LAUNCH((stem_ds_kernel<2>), grid, dim3(256), 0, st, img, H, W, Hs, Ws, same_pad_begin(H, 2), p);
    SMK_CHECK_LAUNCH();
    return 0;
}
int stem_ds(const float* img, int B, int H, int W, const float* stem_w, const float* stem_s, const float* stem_b,
            const float* dw_w, const float* dw_s, const float* dw_b, const float* pw_w, const float* pw_s, const float* pw_b,
            int stride, int round_out, float* out, cudaStream_t st) {
    StemDsProblem q{stem_w, stem_s, stem_b, dw_w, dw_s, dw_b, pw_w, pw_s, pw_b, out};
    return stem_ds(img, B, H, W, &q, 1, stride, round_out, st);
}

int maxpool2x2(const float* in, int ld_in, int B, int H, int W, int C, float* out, cudaStream_t st) {
    long total = (long)B * (H / 2) * (W / 2) * (C / 4);
    int blocks = (int)std::min<long>((total + 255) / 256, 148L * 16);
    SMK_TAG("maxpool2x2", 4.0 * 1.25 * (double)B * H * W * C, 0.0, st);
    SMK_LAUNCH(maxpool2x2_kernel, dim3(blocks), dim3(256), 0, st, in, ld_in, B, H, W, C, out);
    SMK_CHECK_LAUNCH();
    return 0;
}

int nchw_to_nhwc_pad(const float* in, int B, int C, int H, int W, int Cp, float* out, cudaStream_t st, bool round_out) {
    SMK_REQUIRE(Cp % 4 == 0 && Cp >= C, "nchw_to_nhwc_pad: padded channel count must be a multiple of 4 and >= C");
    SMK_TAG("nchw_to_nhwc", 4.0 * (double)B * H * W * (C + Cp), 0.0, st);
    SMK_LAUNCH(nchw_to_nhwc_pad_kernel, dim3(cdiv((long)B * H * W * (Cp / 4), 256)), dim3(256), 0, st, in, B, C, H * W, Cp, round_out ? 1 : 0, out);
    SMK_CHECK_LAUNCH();
    return 0;
}

int conv1x1_sigmoid_nchw(const float* in, int B, int HW, int Cin, const float* w, const float* bias, int Cout, float* out,
                         cudaStream_t st) {
    SMK_REQUIRE(Cout <= 4 && Cin % 4 == 0, "conv1x1_sigmoid_nchw: Cout <= 4 and Cin %% 4 == 0 required");
    size_t smem = (size_t)(Cin * Cout + Cout) * 4;
    SMK_TAG("conv1x1_sigmoid", 4.0 * (double)B * HW * (Cin + Cout), 2.0 * (double)B * HW * Cin * Cout, st);
    SMK_LAUNCH(conv1x1_sigmoid_kernel, dim3(cdiv((long)B * HW, 256)), dim3(256), smem, st, in, B, HW, Cin, w, bias, Cout, out);
    SMK_CHECK_LAUNCH();
    return 0;
}

int gap_linear(const float* feat, int B, int HW, int C, const float* w, const float* bias, int n_out, const uint8_t* codes,
               float* pooled_scratch, float* out, cudaStream_t st) {
    SMK_TAG("gap_pool", 4.0 * ((double)B * HW * C + (double)B * C), (double)B * C * HW, st);
    SMK_LAUNCH(gap_kernel, dim3(dim3(B, cdiv(C, 64))), dim3(256), 0, st, feat, HW, C, pooled_scratch);
    SMK_CHECK_LAUNCH();
    SMK_TAG("head_linear", 4.0 * ((double)B * C + (double)n_out * C + (double)B * n_out), 2.0 * (double)B * C * n_out, st);
    SMK_LAUNCH(head_linear_kernel, dim3(dim3(B, cdiv(n_out, 8))), dim3(256), 0, st, pooled_scratch, C, w, bias, n_out, codes, out);
    SMK_CHECK_LAUNCH();
    return 0;
}

int gap_head(const GapHeadProblem* probs, int n, int B, int HW, int C, cudaStream_t st) {
    SMK_REQUIRE(n == 1 || n == 2, "gap_head: one or two backbones per launch");
    SMK_REQUIRE((size_t)C * 4 <= 48 * 1024, "gap_head: feature width too large for the shared-memory pool");
    GapHead g{};
    int max_out = 0;
    for (int k = 0; k < 2; ++k) {
        const GapHeadProblem& q = probs[k < n ? k : n - 1];
        g.feat[k] = q.feat; g.w[k] = q.w; g.bias[k] = q.bias; g.codes[k] = q.codes; g.out[k] = q.out; g.n_out[k] = q.n_out;
        max_out = std::max(max_out, q.n_out);
    }
    double by = 0, fl = 0;
    for (int k = 0; k < n; ++k) { by += 4.0 * ((double)B * HW * C + (double)probs[k].n_out * C + (double)B * probs[k].n_out); fl += (double)B * C * HW + 2.0 * B * C * probs[k].n_out; }
    SMK_TAG("gap_head", by, fl, st);
    SMK_LAUNCH(gap_head_kernel, dim3(B, cdiv(max_out, 32), n), dim3(256), (size_t)C * 4, st, g, HW, C);
    SMK_CHECK_LAUNCH();
    return 0;
}

}  // namespace smk

extern "C" int smk_debug_conv_f32(const float* in, int ld_in, int B, int H, int W, int Cin, const float* w_kn, const float* scale,
                                  const float* bias, int N, int K, int mode, int relu, const float* res, int ld_res,
                                  float* out, int ld_out, int shuffle, void* stream) {
    smk::ConvProblem p{};
    p.in = in; p.ld_in = ld_in; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.w = w_kn; p.scale = scale; p.bias = bias; p.N = N; p.K = K;
    p.mode = mode; p.relu = relu; p.res = res; p.ld_res = ld_res; p.out = out; p.ld_out = ld_out; p.shuffle = shuffle;
    return smk::conv_gemm(p, (cudaStream_t)stream);
}

extern "C" int smk_debug_stem_ds(const float* img, int B, int H, int W, const float* stem_w, const float* stem_s, const float* stem_b,
                                 const float* dw_w, const float* dw_s, const float* dw_b, const float* pw_w, const float* pw_s,
                                 const float* pw_b, int stride, int round_out, float* out, void* stream) {
    if (B == 0) return 0;
    SMK_REQUIRE(img && stem_w && stem_s && stem_b && dw_w && dw_s && dw_b && pw_w && pw_s && pw_b && out, "smk_debug_stem_ds: null argument");
    return smk::stem_ds(img, B, H, W, stem_w, stem_s, stem_b, dw_w, dw_s, dw_b, pw_w, pw_s, pw_b, stride, round_out, out, (cudaStream_t)stream);
}
