// SmirkGenerator forward (UNet + ResNet blocks), eval-mode BatchNorm folded into conv epilogues.
//
// Replaces SmirkGenerator.forward (reference src/smirk_generator.py:51-86), `_block` (:88-119) and
// ResnetBlock (:121-178).  NHWC fp32 activations.  No tensor is materialised for torch.cat (:66-75):
// each skip tensor and each transposed-conv output is written straight into its channel slice of the
// decoder's input buffer by the producing kernel's strided epilogue; ConvTranspose2d(k2,s2) (:30-44)
// is one GEMM with N = 4*Cout and a pixel-shuffle store; reflection padding (:147-171) is resolved in
// the A-operand loader.
#include "nn_kernels.cuh"
#include <math.h>

namespace {

using smk::ConvProblem;
constexpr float kBnEps = 1e-5f;

struct Conv3 { float* w; float* scale; float* bias; int cin, cin_p, cout; };    // W[(9*cin_p)][cout]
struct UpConv { float* w; float* scale; float* bias; int cin, cout; };          // W[cin][4*cout]

struct TensorCursor {
    const float* const* t; int n; int i = 0;
    const float* next() { return i < n ? t[i++] : nullptr; }
};

bool fold_conv3(TensorCursor& cur, int cin, int cin_p, int cout, smk::DeviceArena& arena, Conv3* out, cudaError_t* err) {
    const float* w = cur.next(); const float* g = cur.next(); const float* b = cur.next();
    const float* mu = cur.next(); const float* var = cur.next();
    if (!w || !g || !b || !mu || !var) return false;
    std::vector<float> W((size_t)9 * cin_p * cout, 0.f), S(cout), Bi(cout);
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
            for (int k = 0; k < 9; ++k) W[((size_t)k * cin_p + c) * cout + o] = w[((size_t)o * cin + c) * 9 + k];
    for (int o = 0; o < cout; ++o) {
        float s = g[o] / sqrtf(var[o] + kBnEps);
        S[o] = s; Bi[o] = b[o] - mu[o] * s;
    }
    out->cin = cin; out->cin_p = cin_p; out->cout = cout;
    cudaError_t e = arena.upload(W, &out->w);
    if (e == cudaSuccess) e = arena.upload(S, &out->scale);
    if (e == cudaSuccess) e = arena.upload(Bi, &out->bias);
    *err = e;
    return e == cudaSuccess;
}

bool fold_upconv(TensorCursor& cur, int cin, int cout, smk::DeviceArena& arena, UpConv* out, cudaError_t* err) {
    const float* w = cur.next(); const float* b = cur.next();        // weight [cin, cout, 2, 2], bias [cout]
    if (!w || !b) return false;
    std::vector<float> W((size_t)cin * 4 * cout), S((size_t)4 * cout, 1.f), Bi((size_t)4 * cout);
    for (int c = 0; c < cin; ++c)
        for (int o = 0; o < cout; ++o)
            for (int q = 0; q < 4; ++q) W[(size_t)c * 4 * cout + q * cout + o] = w[((size_t)c * cout + o) * 4 + q];
    for (int q = 0; q < 4; ++q) for (int o = 0; o < cout; ++o) Bi[q * cout + o] = b[o];
    out->cin = cin; out->cout = cout;
    cudaError_t e = arena.upload(W, &out->w);
    if (e == cudaSuccess) e = arena.upload(S, &out->scale);
    if (e == cudaSuccess) e = arena.upload(Bi, &out->bias);
    *err = e;
    return e == cudaSuccess;
}

}  // namespace

struct SmkGenerator {
    int cin, cin_p, cout, f, nres, precision;
    Conv3 enc[5][2];                 // encoder1..4, bottleneck
    std::vector<Conv3> res;          // 2 per ResnetBlock
    UpConv up[4];                    // upconv4..1  (index 0 = level 4)
    Conv3 dec[4][2];                 // decoder4..1
    float *fw = nullptr, *fb = nullptr;   // final 1x1: W[f][cout], bias
    smk::DeviceArena arena;
};

extern "C" int smk_generator_create(const SmkGeneratorDesc* desc, SmkGenerator** out) {
    SMK_REQUIRE(desc && out && desc->tensors, "smk_generator_create: null argument");
    SMK_REQUIRE(desc->precision == 0, "smk_generator_create: precision %d not available yet", desc->precision);
    SMK_REQUIRE(desc->init_features % 8 == 0 && desc->out_channels <= 4 && desc->in_channels >= 1,
                "smk_generator_create: need init_features %% 8 == 0 and out_channels <= 4");
    SmkGenerator* h = new SmkGenerator();
    h->cin = desc->in_channels; h->cin_p = (desc->in_channels + 7) & ~7; h->cout = desc->out_channels;
    h->f = desc->init_features; h->nres = desc->res_blocks; h->precision = desc->precision;
    const int f = h->f;
    TensorCursor cur{desc->tensors, desc->n_tensors};
    cudaError_t e = cudaSuccess;
    bool ok = true;
    int c_in = h->cin, c_in_p = h->cin_p;
    for (int l = 0; ok && l < 5; ++l) {
        int co = f << l;
        ok = fold_conv3(cur, c_in, c_in_p, co, h->arena, &h->enc[l][0], &e) && fold_conv3(cur, co, co, co, h->arena, &h->enc[l][1], &e);
        c_in = c_in_p = co;
    }
    h->res.resize((size_t)2 * h->nres);
    for (int r = 0; ok && r < 2 * h->nres; ++r) ok = fold_conv3(cur, 16 * f, 16 * f, 16 * f, h->arena, &h->res[r], &e);
    for (int l = 0; ok && l < 4; ++l) {          // level 4 -> 1
        int ci = (16 * f) >> l, co = ci / 2;
        ok = fold_upconv(cur, ci, co, h->arena, &h->up[l], &e) && fold_conv3(cur, 2 * co, 2 * co, co, h->arena, &h->dec[l][0], &e) &&
             fold_conv3(cur, co, co, co, h->arena, &h->dec[l][1], &e);
    }
    if (ok) {
        const float* w = cur.next(); const float* b = cur.next();
        ok = w && b;
        if (ok) {
            std::vector<float> W((size_t)f * h->cout);
            for (int o = 0; o < h->cout; ++o) for (int c = 0; c < f; ++c) W[(size_t)c * h->cout + o] = w[(size_t)o * f + c];
            e = h->arena.upload(W, &h->fw);
            if (e == cudaSuccess) e = h->arena.upload(b, (size_t)h->cout, &h->fb);
            ok = e == cudaSuccess;
        }
    }
    if (!ok || cur.i != cur.n) {
        if (e != cudaSuccess) smk::set_error("smk_generator_create: upload failed: %s", cudaGetErrorString(e));
        else smk::set_error("smk_generator_create: consumed %d tensors but %d were given (state_dict order, num_batches_tracked removed)", cur.i, cur.n);
        delete h; return e != cudaSuccess ? (int)e : -1;
    }
    *out = h;
    return 0;
}

extern "C" void smk_generator_destroy(SmkGenerator* h) { delete h; }

namespace {
// floats per image for every activation buffer (224x224 input)
struct Plan {
    size_t x8, cat[4], t[4], d[4], p[4], tb, b0, b1;
    size_t total() const {
        size_t s = x8 + tb + b0 + b1;
        for (int i = 0; i < 4; ++i) s += cat[i] + t[i] + d[i] + p[i];
        return s;
    }
};
Plan make_plan(const SmkGenerator* h) {
    Plan P{};
    const size_t S = 224;
    P.x8 = S * S * h->cin_p;
    for (int l = 0; l < 4; ++l) {
        size_t s = S >> l, c = (size_t)h->f << l;
        P.cat[l] = s * s * 2 * c; P.t[l] = s * s * c; P.d[l] = s * s * c; P.p[l] = (s / 2) * (s / 2) * c;
    }
    size_t sb = S >> 4, cb = (size_t)h->f * 16;
    P.tb = P.b0 = P.b1 = sb * sb * cb;
    return P;
}
}  // namespace

extern "C" size_t smk_generator_workspace_bytes(const SmkGenerator* h, int B) {
    return (make_plan(h).total() * (size_t)B * sizeof(float)) + 32 * 256;
}

static int conv3(const Conv3& c, const float* in, int ld_in, int B, int S, int mode, bool relu, const float* res,
                 float* out, int ld_out, cudaStream_t st) {
    ConvProblem p{};
    p.in = in; p.ld_in = ld_in; p.B = B; p.H = S; p.W = S; p.Cin = c.cin_p;
    p.w = c.w; p.scale = c.scale; p.bias = c.bias; p.N = c.cout; p.K = 9 * c.cin_p; p.mode = mode; p.relu = relu ? 1 : 0;
    p.res = res; p.ld_res = c.cout; p.out = out; p.ld_out = ld_out; p.shuffle = 0;
    return smk::conv_gemm(p, st);
}

extern "C" int smk_generator_forward(const SmkGenerator* h, const float* x, int B, float* y,
                                     void* ws, size_t ws_bytes, void* stream) {
    if (B == 0) return 0;                      // empty batch: nothing to do (pointers may be null)
    SMK_REQUIRE(h && x && y, "smk_generator_forward: null argument");
    SMK_REQUIRE(B > 0, "smk_generator_forward: negative batch");
    SMK_REQUIRE(ws && ws_bytes >= smk_generator_workspace_bytes(h, B), "smk_generator_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    smk::Workspace w(ws, ws_bytes);
    const Plan P = make_plan(h);
    const int f = h->f;
    float* x8 = w.take<float>(P.x8 * B);
    float *cat[4], *t[4], *d[4], *p[4];
    for (int l = 0; l < 4; ++l) {
        cat[l] = w.take<float>(P.cat[l] * B); t[l] = w.take<float>(P.t[l] * B);
        d[l] = w.take<float>(P.d[l] * B); p[l] = w.take<float>(P.p[l] * B);
    }
    float* tb = w.take<float>(P.tb * B); float* b0 = w.take<float>(P.b0 * B); float* b1 = w.take<float>(P.b1 * B);
    SMK_REQUIRE(b1 != nullptr, "smk_generator_forward: workspace carve-up failed");
    int rc = smk::nchw_to_nhwc_pad(x, B, h->cin, 224, 224, h->cin_p, x8, st);
    if (rc) return rc;
    // encoder levels: conv1 -> t[l]; conv2 -> upper half of cat[l] (the skip); pool -> p[l]
    const float* in = x8; int ld = h->cin_p;
    for (int l = 0; l < 4; ++l) {
        int S = 224 >> l, c = f << l;
        if ((rc = conv3(h->enc[l][0], in, ld, B, S, 1, true, nullptr, t[l], c, st))) return rc;
        if ((rc = conv3(h->enc[l][1], t[l], c, B, S, 1, true, nullptr, cat[l] + c, 2 * c, st))) return rc;
        if ((rc = smk::maxpool2x2(cat[l] + c, 2 * c, B, S, S, c, p[l], st))) return rc;
        in = p[l]; ld = c;
    }
    const int Sb = 14, cb = 16 * f;
    if ((rc = conv3(h->enc[4][0], p[3], 8 * f, B, Sb, 1, true, nullptr, tb, cb, st))) return rc;
    if ((rc = conv3(h->enc[4][1], tb, cb, B, Sb, 1, true, nullptr, b0, cb, st))) return rc;
    float *cur = b0, *nxt = b1;
    for (int r = 0; r < h->nres; ++r) {            // x + BN(conv(reflpad(ReLU(BN(conv(reflpad(x)))))))
        if ((rc = conv3(h->res[2 * r], cur, cb, B, Sb, 2, true, nullptr, tb, cb, st))) return rc;
        if ((rc = conv3(h->res[2 * r + 1], tb, cb, B, Sb, 2, false, cur, nxt, cb, st))) return rc;
        std::swap(cur, nxt);
    }
    // decoder levels 4..1: upconv -> lower half of cat; conv1 over the concat; conv2
    const float* din = cur; int dS = Sb;
    for (int l = 0; l < 4; ++l) {
        int lvl = 3 - l;                            // index into cat/t/d (3 = 28x28 ... 0 = 224x224)
        const UpConv& u = h->up[l];
        ConvProblem q{};
        q.in = din; q.ld_in = u.cin; q.B = B; q.H = dS; q.W = dS; q.Cin = u.cin; q.w = u.w; q.scale = u.scale; q.bias = u.bias;
        q.N = 4 * u.cout; q.K = u.cin; q.mode = 0; q.relu = 0; q.res = nullptr; q.ld_res = 0;
        q.out = cat[lvl]; q.ld_out = 2 * u.cout; q.shuffle = 1;
        if ((rc = smk::conv_gemm(q, st))) return rc;
        dS *= 2;
        if ((rc = conv3(h->dec[l][0], cat[lvl], 2 * u.cout, B, dS, 1, true, nullptr, t[lvl], u.cout, st))) return rc;
        if ((rc = conv3(h->dec[l][1], t[lvl], u.cout, B, dS, 1, true, nullptr, d[lvl], u.cout, st))) return rc;
        din = d[lvl];
    }
    return smk::conv1x1_sigmoid_nchw(d[0], B, 224 * 224, f, h->fw, h->fb, h->cout, y, st);
}
