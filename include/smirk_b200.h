/*
 * smirk_b200 — C ABI of the B200-native SMIRK hot path (encode -> FLAME -> render -> generator).
 *
 * The reference (georgeretsi/smirk) is pure Python and has no FFI of its own; its only native seams
 * on this path are third-party: `timm.create_model` (src/smirk_encoder.py:7-12), ATen/cuDNN ops, and
 * `pytorch3d.renderer.mesh.rasterize_meshes` (src/renderer/renderer.py:185-193).  Each entry point
 * below replaces the body of one reference nn.Module.forward; the Python classes in smirk_b200/ keep
 * the reference signatures and call these through ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - return 0 = OK, <0 = argument/shape error, >0 = cudaError_t; message via smk_last_error()
 *     (thread-local).  No exceptions, no exit().
 *   - `*_create` take HOST pointers to fp32 / int32 constant arrays, fold/pack them and upload to the
 *     CURRENT CUDA device; handles are immutable afterwards.  `*_forward` take DEVICE pointers
 *     (contiguous, 16-byte aligned), never allocate, never synchronise, and order all work after / before
 *     the caller's stream (cudaStream_t passed as void*) — CUDA-graph capturable.  smk_encoder_forward
 *     runs its backbones as parallel branches that fork from and join back into that stream, using
 *     per-call fork/join events and side streams taken round-robin from a pool of 8 inside the handle:
 *     forwards are re-entrant (up to 8 in flight per encoder handle) given distinct workspaces.
 *   - the caller owns inputs, outputs and the workspace (size from `*_workspace_bytes`).
 */
#ifndef SMIRK_B200_H
#define SMIRK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMK_VERSION 100

int smk_version(void);
const char* smk_last_error(void);

/* Launch accounting and the built-in event profiler (used by bench.py for `gpu_launches` and the
 * per-kernel roofline): smk_launch_count() = kernels launched by this library so far in the process;
 * with the profiler enabled every launch is bracketed by CUDA events on its own stream (do not enable
 * during CUDA-graph capture).  smk_profiler_report writes "tag launches total_ms bytes flops" lines
 * (algorithmic bytes / FLOPs summed over launches), returns the number of lines or -1 if buf is small. */
unsigned long long smk_launch_count(void);
void smk_profiler_enable(int on);
void smk_profiler_reset(void);
int smk_profiler_report(char* buf, size_t n);

/* ------------------------------------------------------------------------------------------------
 * FLAME  — replaces FLAME.forward (src/FLAME/FLAME.py:232-315) and lbs() (src/FLAME/lbs.py:140-227).
 * ---------------------------------------------------------------------------------------------- */
typedef struct SmkFlame SmkFlame;

typedef struct {
    int n_verts;            /* 5023 */
    int n_faces;            /* 9976 */
    int n_betas;            /* n_shape + n_exp = 350 */
    int n_joints;           /* 5, kinematic parents fixed to [-1,0,1,1,1] (FLAME.py:76-77) */
    const float* v_template;        /* [V,3]                FLAME.py:64 */
    const float* shapedirs;         /* [V,3,n_betas]        FLAME.py:67-69 */
    const float* posedirs;          /* [(J-1)*9, V*3]       FLAME.py:71-73 */
    const float* J_regressor;       /* [J,V]                FLAME.py:75 */
    const float* lbs_weights;       /* [V,J]                FLAME.py:78 */
    const float* l_eyelid;          /* [V,3]                FLAME.py:81 */
    const float* r_eyelid;          /* [V,3]                FLAME.py:82 */
    const int32_t* faces;           /* [F,3]                FLAME.py:61 */
    /* landmark embeddings (FLAME.py:94-113) */
    int n_static;  const int32_t* static_faces;  const float* static_bary;     /* 51 */
    int n_dyn_rows; int n_dyn; const int32_t* dyn_faces; const float* dyn_bary; /* 79 x 17 */
    int n_full;    const int32_t* full_faces;    const float* full_bary;       /* 68 */
    int n_mp;      const int32_t* mp_faces;      const float* mp_bary;         /* 105 */
} SmkFlameDesc;

int smk_flame_create(const SmkFlameDesc* desc, SmkFlame** out);
void smk_flame_destroy(SmkFlame* h);
size_t smk_flame_workspace_bytes(const SmkFlame* h, int B);
/* betas [B,n_betas] = cat(shape, expression); full_pose [B,15] = cat(global, neck, jaw, eyes(6));
 * eyelid [B,2] or NULL.  Outputs: verts [B,V,3]; lmk_fan [B,68,3] (17 dynamic-contour + 51 static);
 * lmk_fan3d [B,68,3]; lmk_mp [B,105,3]; joints [B,J,3] (posed joints, may be NULL);
 * dyn_idx int32 [B] (selected contour LUT row, may be NULL).                                     */
int smk_flame_forward(const SmkFlame* h, const float* betas, const float* full_pose, const float* eyelid,
                      int B, float* verts, float* lmk_fan, float* lmk_fan3d, float* lmk_mp,
                      float* joints, int32_t* dyn_idx, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Renderer — replaces Renderer.forward/render/rasterize (src/renderer/renderer.py:100-207,239-250),
 * util.vertex_normals/face_vertices/batch_orth_proj (src/renderer/util.py) and the third-party
 * pytorch3d rasterize_meshes call (renderer.py:185-193; blur 0, K=1, no perspective correction).
 * ---------------------------------------------------------------------------------------------- */
typedef struct SmkRenderer SmkRenderer;

typedef struct {
    int n_verts;               /* vertices of the incoming mesh (5023) */
    int n_mask;                /* rendered subset (1787 for the FLAME `face` mask; = n_verts for full head) */
    const int32_t* mask_ids;   /* [n_mask] vertex ids, order defines the sub-mesh numbering (renderer.py:71) */
    int n_faces;               /* 3408; must be < 65536 and a multiple of 4 (packed tile ranges are read 4 at a time) */
    const int32_t* faces;      /* [n_faces,3] indices into the sub-mesh (renderer.py:74) */
    int image_size;            /* 224 */
} SmkRendererDesc;

int smk_renderer_create(const SmkRendererDesc* desc, SmkRenderer** out);
void smk_renderer_destroy(SmkRenderer* h);
size_t smk_renderer_workspace_bytes(const SmkRenderer* h, int B);
/* verts [B,n_verts,3], cam [B,3] = (scale, tx, ty).  Outputs: rendered [B,3,S,S]; tverts [B,n_verts,3];
 * optional (NULL to skip): pix_to_face int64 [B,S,S] (packed b*n_faces+f, -1 = background),
 * bary [B,S,S,3], zbuf [B,S,S] (both -1 on background), normals [B,n_mask,3].                       */
int smk_renderer_forward(const SmkRenderer* h, const float* verts, const float* cam, int B,
                         float* rendered, float* tverts, int64_t* pix_to_face, float* bary, float* zbuf,
                         float* normals, void* ws, size_t ws_bytes, void* stream);
/* Orthographic projection of landmark sets (renderer.py:104-108): pts [B,L,3] -> out [B,L,2]. */
int smk_project_points(const float* pts, const float* cam, int B, int L, float* out_xy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SmirkEncoder — replaces SmirkEncoder.forward (src/smirk_encoder.py:123-133): three timm
 * tf_mobilenetv3_{small,large,large}_minimal_100 backbones (features_only, last feature) -> global
 * average pool -> Linear heads -> split/clamps (:34-45,66-73,95-110).
 * ---------------------------------------------------------------------------------------------- */
typedef struct SmkEncoder SmkEncoder;

typedef struct {
    /* Per backbone (0 = pose/small, 1 = shape/large, 2 = expression/large): the fp32 tensors of its
     * `encoder.*` state_dict in state_dict order with `num_batches_tracked` entries removed, i.e.
     * conv weight [Cout,Cin/groups,k,k] followed by BN weight, bias, running_mean, running_var.       */
    const float* const* tensors[3];
    int n_tensors[3];
    const float* head_w[3];    /* [6,576], [n_shape,960], [n_exp+5,960] */
    const float* head_b[3];
    int n_shape;               /* 300 */
    int n_exp;                 /* 50 */
    int precision;             /* 0 = fp32 CUDA-core GEMMs, 1 = TF32 tcgen05 GEMMs for the 1x1 convs,
                                  2 = 1 + inverted-residual blocks run expand-1x1 + depthwise-3x3 as one fused kernel,
                                  3 = 2 with error-compensated "3xTF32" tensor-core arithmetic (operands split into TF32
                                      head + tail, three products per term): fp32-equivalent results, the parity path */
} SmkEncoderDesc;

int smk_encoder_create(const SmkEncoderDesc* desc, SmkEncoder** out);
void smk_encoder_destroy(SmkEncoder* h);
size_t smk_encoder_workspace_bytes(const SmkEncoder* h, int B);
/* img [B,3,224,224] NCHW in [0,1].  Outputs: pose_cam [B,6], shape [B,n_shape], expr [B,n_exp+5]
 * with the clamps of smirk_encoder.py:105-108 already applied to columns n_exp..n_exp+4.           */
int smk_encoder_forward(const SmkEncoder* h, const float* img, int B, float* pose_cam, float* shape,
                        float* expr, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SmirkGenerator — replaces SmirkGenerator.forward (src/smirk_generator.py:51-86), eval-mode BN.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SmkGenerator SmkGenerator;

typedef struct {
    int in_channels, out_channels, init_features, res_blocks;    /* 6, 3, 32, 5 (demo.py:63) */
    /* fp32 tensors of the module's state_dict in state_dict order, `num_batches_tracked` removed.    */
    const float* const* tensors;
    int n_tensors;
    int precision;             /* 0 = fp32 CUDA-core implicit GEMM, 1 = TF32 tcgen05 implicit GEMM */
} SmkGeneratorDesc;

int smk_generator_create(const SmkGeneratorDesc* desc, SmkGenerator** out);
void smk_generator_destroy(SmkGenerator* h);
size_t smk_generator_workspace_bytes(const SmkGenerator* h, int B);
/* x [B,in_channels,224,224] NCHW -> y [B,out_channels,224,224] NCHW in (0,1). */
int smk_generator_forward(const SmkGenerator* h, const float* x, int B, float* y,
                          void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Crop / warp front and back end — replaces the CPU `skimage.transform.warp` calls of
 * demo.py:97 and demo_video.py:128,149 (+ BGR->RGB, /255, HWC->CHW of demo.py:103-105) for frames that are
 * already on the device.  Bilinear (order 1), mode 'constant', cval 0, clip to the source range, float64
 * arithmetic, truncation to uint8 — skimage's `_warp_fast` semantics.  m / minv: [B][9] row-major float64 3x3
 * maps from OUTPUT pixel (col,row,1) to INPUT (x,y,1) (affine rows only); ws >= smk_warp_workspace_bytes(B).
 *   smk_crop_warp      : frames uint8 [B,H,W,3] -> out float32 [B,3,S,S] = uint8 result / 255; swap_rb reverses the
 *                        channel order (BGR frame -> RGB tensor).
 *   smk_warp_u8        : src uint8 [B,Hs,Ws,3] -> dst uint8 [B,Hd,Wd,3].
 *   smk_f32chw_to_u8hwc: (x * 255.0f).astype(uint8) of a [B,3,S,S] float image -> [B,S,S,3] (demo_video.py:148).  */
size_t smk_warp_workspace_bytes(int B);
int smk_crop_warp(const uint8_t* frames, int B, int H, int W, const double* minv, int S, int swap_rb, float* out,
                  void* ws, size_t ws_bytes, void* stream);
int smk_warp_u8(const uint8_t* src, int B, int Hs, int Ws, const double* m, int Hd, int Wd, uint8_t* dst,
                void* ws, size_t ws_bytes, void* stream);
int smk_f32chw_to_u8hwc(const float* in, int B, int S, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masking between Renderer and SmirkGenerator (src/utils/masking.py, demo.py:138-167).
 * Random draws stay with the caller (torch); these entry points are the deterministic parts.
 *   face_weights: trans_verts [B,V,3], base_prob [F] -> weights [B,F]               (masking.py:146-160)
 *   points      : face_idx int64 [B,N], bary [B,N,3] -> npoints int64 [B,N,2] (x,y)  (masking.py:166-174)
 *   compose     : img [B,3,S,S], hull [B,1,S,S], npoints/rbound (first rbound[b] points are kept), optional
 *                 rendered_mask [B,1,S,S], noise_mult [B,3,S,S], random_centres [B,1,S,S] -> masked [B,3,S,S]  */
typedef struct SmkMasking SmkMasking;
typedef struct { int n_verts; int n_faces; const int32_t* faces; } SmkMaskingDesc;
int smk_masking_create(const SmkMaskingDesc* desc, SmkMasking** out);
void smk_masking_destroy(SmkMasking* h);
size_t smk_masking_workspace_bytes(const SmkMasking* h, int B, int S);
int smk_masking_face_weights(const SmkMasking* h, const float* trans_verts, const float* base_prob, int B,
                             float* weights, void* ws, size_t ws_bytes, void* stream);
int smk_masking_points(const SmkMasking* h, const float* trans_verts, const int64_t* face_idx, const float* bary,
                       int B, int N, int image_size, int64_t* npoints, void* stream);
/* extra_points [B,3,S,S] (nullable): masking()'s third argument given explicitly (masking.py:71; the trainer passes the
 * result of transfer_pixels) instead of img * point-mask(npoints, rbound).                                                 */
int smk_masking_compose(const SmkMasking* h, const float* img, const float* hull, const int64_t* npoints, const int64_t* rbound,
                        int N, const float* extra_points, const float* rendered_mask, const float* noise_mult, const float* random_centres,
                        int wr, int B, int S, float* masked, void* ws, size_t ws_bytes, void* stream);
/* transfer_pixels (masking.py:116-129): points int64 [B,N,2] (x, y); rbound int64 [B] or NULL; ws >= B*S*S*4 bytes.       */
int smk_masking_transfer_pixels(const float* img, const int64_t* points1, const int64_t* points2, const int64_t* rbound,
                                int B, int N, int S, float* out, void* ws, size_t ws_bytes, void* stream);
/* The whole step of demo.py:138-165 with the random draws made on the device (Philox4x32-10, keyed by rng_state[0] = seed
 * and rng_state[1] = call counter, which the call increments on the stream — graph replays draw fresh samples):
 *   face weights -> N = int(mask_ratio * ratio_mul * S * S) faces by inverse CDF (multinomial with replacement) -> uniform
 *   barycentrics -> pixel coordinates -> per-image budget rbound = N / ratio_mul * U(1, ratio_mul)^(+-1) -> point mask ->
 *   masking(img, hull, img * pmask, wr, rendered_mask = any(rendered != 0), extra_noise, random_mask = p_centre).
 * rendered [B,3,S,S] is the Renderer's output; hull [B,1,S,S] the landmark hull mask (1 outside the face).  Optional
 * debug outputs (NULL to skip) export the draws so the result can be checked against the reference's functions:
 * dbg_face_idx int64 [B,N], dbg_bary [B,N,3], dbg_npoints int64 [B,N,2], dbg_rbound int64 [B], dbg_noise [B,3,S,S],
 * dbg_centres [B,1,S,S].  ws >= smk_masking_forward_workspace_bytes(h, B, S, N).                                          */
size_t smk_masking_forward_workspace_bytes(const SmkMasking* h, int B, int S, int N);
int smk_masking_forward(const SmkMasking* h, const float* img, const float* hull, const float* trans_verts, const float* rendered,
                        const float* base_prob, int B, int S, int N, int wr, float ratio_mul, float p_centre, int extra_noise,
                        uint64_t* rng_state, float* masked, int64_t* dbg_face_idx, float* dbg_bary, int64_t* dbg_npoints,
                        int64_t* dbg_rbound, float* dbg_noise, float* dbg_centres, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Peer-mapped gather buffers: the all-gather of the final outputs across the GPUs of one node (the
 * reference has no multi-GPU code) as copy-engine pushes over NVLink, which take no SM from the persistent compute
 * kernels.  Each rank: smk_peer_alloc a [world][shard] buffer, exchange the 64-byte handles (any host channel),
 * smk_peer_open every peer's handle, and after a batch smk_peer_push its packed shard into slot `rank` of every buffer
 * on a communication stream.  Allocation and mapping are set-up calls; a forward never allocates.
 * ---------------------------------------------------------------------------------------------- */
int smk_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int smk_peer_free(void* ptr);
int smk_peer_open(const unsigned char* handle64, void** ptr);
int smk_peer_close(void* ptr);
int smk_peer_push(void* dst, const void* src, size_t bytes, void* stream);
/* The same copy to n destinations, spread over the fan's own streams (several copy engines / NVLink ports at once);
 * ordered after the work already on `stream`, which resumes only after every copy.                                        */
typedef struct SmkPeerFan SmkPeerFan;
int smk_peer_fan_create(int n_streams, SmkPeerFan** out);
void smk_peer_fan_destroy(SmkPeerFan* f);
int smk_peer_fan_push(SmkPeerFan* f, void* const* dsts, int n, const void* src, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Kernel-level test entry points (used by tests/ to check single convolution kernels against torch;
 * not part of the drop-in surface).  All pointers are device pointers.
 *   smk_debug_conv_f32: fp32 CUDA-core implicit GEMM.  w_kn is [K][N]; mode 0 = 1x1, 1 = 3x3 zero pad,
 *                       2 = 3x3 reflection pad; shuffle = 1 stores ConvTranspose2d(k2,s2) pixel-shuffled.
 *   smk_debug_conv_tc : TF32 tcgen05 implicit GEMM.  wt is [N][K]; mode 2 expects `in` to be a
 *                       [B,H+2,W+2,*] buffer whose halo was filled by smk_debug_reflect_halo;
 *                       store 0 plain, 1 pixel-shuffle, 2 interior of a padded [B,H+2,W+2,*] buffer.
 * ---------------------------------------------------------------------------------------------- */
int smk_debug_conv_f32(const float* in, int ld_in, int B, int H, int W, int Cin, const float* w_kn, const float* scale,
                       const float* bias, int N, int K, int mode, int relu, const float* res, int ld_res,
                       float* out, int ld_out, int shuffle, void* stream);
int smk_debug_conv_tc(const float* in, int ld_in, int B, int H, int W, int Cin, const float* wt, const float* scale,
                      const float* bias, int N, int K, int mode, int relu, const float* res, int ld_res, int res_pad,
                      float* out, int ld_out, int store, void* stream);
int smk_debug_reflect_halo(float* buf, int B, int H, int W, int C, void* stream);
/*   smk_debug_xdw: fused expand-1x1 (TF32 tcgen05) + BN + ReLU + depthwise-3x3 (fp32) + BN + ReLU of a
 *                  MobileNetV3 inverted-residual block.  x [B,H,W,Cin] NHWC; w1t [mid][Cin]; wdw [9][mid];
 *                  out [B,ceil(H/stride),ceil(W/stride),mid]; TF-SAME padding.                            */
int smk_debug_xdw(const float* x, int B, int H, int W, int Cin, const float* w1t, const float* scale1, const float* bias1,
                  int mid, const float* wdw, const float* scale2, const float* bias2, int stride, int round_out,
                  float* out, void* stream);

/*   smk_debug_conv3_win: the persistent windowed TF32 tcgen05 3x3 convolution of the high-resolution narrow layers
 *                  (zero padding 1, N in {32, 64}, Cin % 32 == 0, W >= 56, resident weights <= 72 KB); wt is [N][9*Cin].      */
int smk_debug_conv3_win(const float* in, int ld_in, int B, int H, int W, int Cin, const float* wt, const float* scale,
                        const float* bias, int N, int relu, float* out, int ld_out, void* stream);
/*   smk_debug_gemm_tc3x / smk_debug_xdw3x: the error-compensated 3xTF32 variants of the two tensor-core encoder kernels
 *                  (encoder precision 3).  wt_hi / wt_lo (w1t_hi / w1t_lo) are the TF32 heads and tails of the weights:
 *                  hi = tf32(w), lo = tf32(w - hi).  in is [M, ld_in] row-major; out [M, ld_out].                       */
int smk_debug_gemm_tc3x(const float* in, int ld_in, int M, const float* wt_hi, const float* wt_lo, const float* scale,
                        const float* bias, int N, int K, int relu, const float* res, int ld_res, float* out, int ld_out, void* stream);
int smk_debug_xdw3x(const float* x, int B, int H, int W, int Cin, const float* w1t_hi, const float* w1t_lo, const float* scale1,
                    const float* bias1, int mid, const float* wdw, const float* scale2, const float* bias2, int stride,
                    float* out, void* stream);

/*   smk_debug_stem_ds: fused stem conv (3x3 s2, 3 -> 16) + BN + ReLU + depthwise-separable block 0 (fp32 CUDA cores).
 *                      img [B,3,H,W] NCHW; stem_w [27][16] (k = (c*3+ky)*3+kx); dw_w [9][16]; pw_w [16 ci][16 co];
 *                      out [B,H/2/stride,W/2/stride,16] NHWC; the skip connection is added when stride == 1.          */
int smk_debug_stem_ds(const float* img, int B, int H, int W, const float* stem_w, const float* stem_s, const float* stem_b,
                      const float* dw_w, const float* dw_s, const float* dw_b, const float* pw_w, const float* pw_s,
                      const float* pw_b, int stride, int round_out, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMIRK_B200_H */
