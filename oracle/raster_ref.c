/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product path.
 *
 * CPU restatement of the naive (bin_size = 0) path of pytorch3d's `rasterize_meshes`, the
 * third-party call the reference makes at src/renderer/renderer.py:184-193 with
 *   image_size=224, blur_radius=0.0, faces_per_pixel=1, bin_size=None, max_faces_per_bin=None,
 *   perspective_correct=False   (clip_barycentric_coords / cull_backfaces default False).
 * pytorch3d is NOT vendored in /root/reference and its wheel is not pinned (readme.md:37 points at
 * the py39_cu117_pyt201 wheel index; it is absent from requirements.txt), so this file restates the
 * published algorithm of pytorch3d 0.7.x (csrc/rasterize_meshes/rasterize_meshes_cpu.cpp,
 * csrc/utils/geometry_utils.h) from its documented semantics: PARITY UNPINNED (no golden vectors
 * exist upstream).
 *
 * Semantics restated:
 *   - output row yi / col xi sample NDC (xf, yf) with yf = ndc(H-1-yi), xf = ndc(W-1-xi),
 *     ndc(i) = -1 + (2*i + 1)/S   computed in fp32 as  -offset + (range*i + offset)/S
 *   - per (pixel, face): skip when |area(v0,v1,v2)| <= 1e-8, when the pixel is outside the face's
 *     xy bounding box (blur 0), when pz < 0; bary = edge functions / (edge(v2,v0,v1) + 1e-8);
 *     inside <=> all three bary strictly > 0; with blur_radius 0 only inside pixels are kept
 *   - K = 1: keep the lexicographically smallest (pz, face index)
 *   - all outputs initialised to -1; pix_to_face holds the PACKED index n*F + f
 *   - dists = -(squared distance to the closest edge) for inside pixels
 * Arithmetic is plain fp32 with no FMA contraction (build with -ffp-contract=off), matching the
 * pytorch3d CPU build; the product kernel uses explicit non-fused intrinsics to agree bit for bit.
 */
#include <stdint.h>
#include <math.h>

#define K_EPS 1e-8f

static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    /* EdgeFunctionForward(p, a, b) = (p.x-a.x)*(b.y-a.y) - (p.y-a.y)*(b.x-a.x) */
    float t0 = (px - ax) * (by - ay);
    float t1 = (py - ay) * (bx - ax);
    return t0 - t1;
}

static inline float pix_to_ndc(int i, int S) {
    const float range = 2.0f, offset = 1.0f;
    return -offset + (range * (float)i + offset) / (float)S;
}

static inline float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    float bax = bx - ax, bay = by - ay;
    float l2 = bax * bax + bay * bay;
    if (l2 <= K_EPS) return (px - bx) * (px - bx) + (py - by) * (py - by);
    float t = (bax * (px - ax) + bay * (py - ay)) / l2;
    t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    float qx = ax + t * bax, qy = ay + t * bay;
    return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

/* face_verts: [N*F, 3, 3] (x, y, z per corner), already in pytorch3d NDC.
 * pix_to_face: int64 [N,H,W]; zbuf, dists: float [N,H,W]; bary: float [N,H,W,3].           */
void smk_oracle_rasterize(const float* face_verts, int N, int F, int H, int W,
                          int64_t* pix_to_face, float* zbuf, float* bary, float* dists) {
    for (long i = 0; i < (long)N * H * W; ++i) {
        pix_to_face[i] = -1; zbuf[i] = -1.f; dists[i] = -1.f;
        bary[3 * i] = bary[3 * i + 1] = bary[3 * i + 2] = -1.f;
    }
    for (int n = 0; n < N; ++n) {
        for (int f = 0; f < F; ++f) {
            const float* v = face_verts + ((long)n * F + f) * 9;
            float x0 = v[0], y0 = v[1], z0 = v[2];
            float x1 = v[3], y1 = v[4], z1 = v[5];
            float x2 = v[6], y2 = v[7], z2 = v[8];
            float face_area = edge_fn(x0, y0, x1, y1, x2, y2);
            if (face_area <= K_EPS && face_area >= -K_EPS) continue;
            float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
            float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
            /* conservative pixel range (the exact float bbox test is applied per pixel below) */
            int xi_lo = (int)floorf((1.f - xmax) * 0.5f * W) - 2, xi_hi = (int)ceilf((1.f - xmin) * 0.5f * W) + 2;
            int yi_lo = (int)floorf((1.f - ymax) * 0.5f * H) - 2, yi_hi = (int)ceilf((1.f - ymin) * 0.5f * H) + 2;
            if (xi_lo < 0) xi_lo = 0; if (yi_lo < 0) yi_lo = 0;
            if (xi_hi > W - 1) xi_hi = W - 1; if (yi_hi > H - 1) yi_hi = H - 1;
            float denom = edge_fn(x2, y2, x0, y0, x1, y1) + K_EPS;
            for (int yi = yi_lo; yi <= yi_hi; ++yi) {
                float yf = pix_to_ndc(H - 1 - yi, H);
                if (yf > ymax || yf < ymin) continue;
                for (int xi = xi_lo; xi <= xi_hi; ++xi) {
                    float xf = pix_to_ndc(W - 1 - xi, W);
                    if (xf > xmax || xf < xmin) continue;
                    float w0 = edge_fn(xf, yf, x1, y1, x2, y2) / denom;
                    float w1 = edge_fn(xf, yf, x2, y2, x0, y0) / denom;
                    float w2 = edge_fn(xf, yf, x0, y0, x1, y1) / denom;
                    float pz = (w0 * z0 + w1 * z1) + w2 * z2;
                    if (pz < 0.f) continue;
                    if (!(w0 > 0.f && w1 > 0.f && w2 > 0.f)) continue;   /* blur 0: inside only */
                    long o = ((long)n * H + yi) * W + xi;
                    /* f ascends, so strict < keeps the smallest (pz, f) */
                    if (pix_to_face[o] < 0 || pz < zbuf[o]) {
                        float d = fminf(seg_dist2(xf, yf, x0, y0, x1, y1),
                                        fminf(seg_dist2(xf, yf, x1, y1, x2, y2), seg_dist2(xf, yf, x2, y2, x0, y0)));
                        pix_to_face[o] = (int64_t)n * F + f;
                        zbuf[o] = pz; dists[o] = -d;
                        bary[3 * o] = w0; bary[3 * o + 1] = w1; bary[3 * o + 2] = w2;
                    }
                }
            }
        }
    }
}

/* Brute-force variant (every face tested at every pixel, exactly the naive loop nest); used by the
 * tests to show the bbox-driven traversal above visits the same (pixel, face) pairs.               */
void smk_oracle_rasterize_bruteforce(const float* face_verts, int N, int F, int H, int W,
                                     int64_t* pix_to_face, float* zbuf, float* bary) {
    for (long i = 0; i < (long)N * H * W; ++i) {
        pix_to_face[i] = -1; zbuf[i] = -1.f;
        bary[3 * i] = bary[3 * i + 1] = bary[3 * i + 2] = -1.f;
    }
    for (int n = 0; n < N; ++n)
        for (int yi = 0; yi < H; ++yi) {
            float yf = pix_to_ndc(H - 1 - yi, H);
            for (int xi = 0; xi < W; ++xi) {
                float xf = pix_to_ndc(W - 1 - xi, W);
                long o = ((long)n * H + yi) * W + xi;
                for (int f = 0; f < F; ++f) {
                    const float* v = face_verts + ((long)n * F + f) * 9;
                    float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
                    float face_area = edge_fn(x0, y0, x1, y1, x2, y2);
                    if (face_area <= K_EPS && face_area >= -K_EPS) continue;
                    float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
                    float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
                    if (xf > xmax || xf < xmin || yf > ymax || yf < ymin) continue;
                    float denom = edge_fn(x2, y2, x0, y0, x1, y1) + K_EPS;
                    float w0 = edge_fn(xf, yf, x1, y1, x2, y2) / denom;
                    float w1 = edge_fn(xf, yf, x2, y2, x0, y0) / denom;
                    float w2 = edge_fn(xf, yf, x0, y0, x1, y1) / denom;
                    float pz = (w0 * z0 + w1 * z1) + w2 * z2;
                    if (pz < 0.f) continue;
                    if (!(w0 > 0.f && w1 > 0.f && w2 > 0.f)) continue;
                    if (pix_to_face[o] < 0 || pz < zbuf[o]) {
                        pix_to_face[o] = (int64_t)n * F + f; zbuf[o] = pz;
                        bary[3 * o] = w0; bary[3 * o + 1] = w1; bary[3 * o + 2] = w2;
                    }
                }
            }
        }
}
