// f32 rectangle / view algebra shared by host glue (api.cpp) and device kernels
// (kernels_exact.cu).  Every function restates one reference function op-for-op; both
// translation units are compiled with FMA contraction OFF (-fmad=false / -ffp-contract=off)
// so each `*`, `+`, `-`, `/` rounds exactly like the reference's scalar f32 code.
//
//   Rect / RotatedRect ......... crates/zaru-image/src/rect.rs:11-237, :269-424
//   AspectRatio::as_f32 ........ crates/zaru-image/src/resolution.rs:128-162
//   ViewData::view ............. crates/zaru/src/image/mod.rs:201-210
//   Mat2 * Vec2 ................ crates/zaru-linalg/src/matrix/ops.rs:74-76, matrix.rs:571-579
#pragma once

#if defined(__CUDACC__)
#define ZB_HD __host__ __device__ __forceinline__
#else
#define ZB_HD inline
#endif

namespace zb {

struct RectF {
    float cx, cy, w, h;
};

struct RRectF {          // RotatedRect with its rotation matrix entries precomputed
    RectF r;
    float rad;
    float c, s;          // cos(rad), sin(rad) as f32
};

ZB_HD RectF rect_from_top_left(float x, float y, float w, float h) {
    RectF r;
    r.cx = x + w * 0.5f;
    r.cy = y + h * 0.5f;
    r.w = w;
    r.h = h;
    return r;
}
ZB_HD float rect_x(const RectF &r) { return r.cx - r.w * 0.5f; }
ZB_HD float rect_y(const RectF &r) { return r.cy - r.h * 0.5f; }

// AspectRatio::new(w,h).as_f32(): gcd-reduced w/h.
ZB_HD float aspect_as_f32(unsigned w, unsigned h) {
    unsigned a = w, b = h;
    while (b > 0) {
        unsigned t = b;
        b = a % b;
        a = t;
    }
    return (float)(w / a) / (float)(h / a);
}

// Rect::grow_to_fit_aspect (rect.rs:104-117)
ZB_HD RectF grow_to_fit_aspect(RectF r, float aspect) {
    float target_width = r.h * aspect;
    if (target_width >= r.w) {
        float inc_w = target_width - r.w;
        r.w = r.w + inc_w;
    } else {
        float target_height = r.w / aspect;
        float inc_h = target_height - r.h;
        r.h = r.h + inc_h;
    }
    return r;
}

// Rect::grow_rel (rect.rs:84-94)
ZB_HD RectF grow_rel(RectF r, float amount) {
    float left = r.w * amount, right = r.w * amount;
    float top = r.h * amount, bottom = r.h * amount;
    r.w = r.w + left + right;
    r.h = r.h + top + bottom;
    return r;
}

// Mat2::rotation_counterclockwise(rad) * (x, y): rows (c, -s), (s, c); fold from 0.
ZB_HD void rot_ccw_apply(float c, float s, float x, float y, float &ox, float &oy) {
    float ns = -s;
    ox = (0.0f + c * x) + ns * y;
    oy = (0.0f + s * x) + c * y;
}

// RotatedRect::transform_out (rect.rs:417-423)
ZB_HD void transform_out(const RRectF &rr, float px, float py, float &ox, float &oy) {
    float hx = rr.r.w * 0.5f, hy = rr.r.h * 0.5f;
    float rx, ry;
    rot_ccw_apply(rr.c, rr.s, px - hx, py - hy, rx, ry);
    ox = rx + hx + rect_x(rr.r);
    oy = ry + hy + rect_y(rr.r);
}

// ViewData::view (image/mod.rs:201-210): compose `child` (in parent's coordinates) with `parent`.
// The caller supplies cos/sin of the summed angle (host: glibc cosf/sinf; device: see DESIGN.md).
ZB_HD RRectF view_compose(const RRectF &parent, const RectF &child, float child_rad, float c_sum, float s_sum) {
    RRectF out;
    out.rad = parent.rad + child_rad;
    float tx, ty;
    transform_out(parent, child.cx, child.cy, tx, ty);
    float px = tx - child.w * 0.5f;
    float py = ty - child.h * 0.5f;
    out.r = rect_from_top_left(px, py, child.w, child.h);   // Rect::move_to
    out.c = c_sum;
    out.s = s_sum;
    return out;
}

// The view of a whole image: ViewData::full(image) (image/mod.rs:194-199).
ZB_HD RRectF full_view(int width, int height) {
    RRectF v;
    v.r = rect_from_top_left(0.0f, 0.0f, (float)width, (float)height);
    v.rad = 0.0f;
    v.c = 1.0f;
    v.s = 0.0f;
    return v;
}

}  // namespace zb
