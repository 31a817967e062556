// host_view.cpp -- host-side mirror of the reference's PER-VIEW math.
//
// These are the O(views) computations that stay on the CPU in the reference
// too: building a Frustum from a camera, and the
// per-view prologue of assign_objects_to_clusters.  In a real Bevy App the
// Rust shim computes them with glam and passes them through the C ABI; this
// C++ statement exists so that C/C++/Python hosts (bench.py, the tests) get
// the same constants without Rust.  Float ops follow glam's x86-64/SSE2 order
// with no FMA contraction (compiled with -ffp-contract=off).
//
// Product code: does not include, link or call anything under oracle/.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/b200vis.h"
#include "host_view.hpp"

namespace b200vis { namespace host {

namespace {

struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };
struct Cols3 { F3 x, y, z; };               // glam Mat3A (columns)
struct Affine { Cols3 m; F3 t; };           // glam Affine3A
struct Cols4 { F4 x, y, z, w; };            // glam Mat4 (columns)

inline F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline F3 operator-(F3 a) { return {-a.x, -a.y, -a.z}; }
inline F4 operator+(F4 a, F4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline F4 operator-(F4 a, F4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline F4 operator*(F4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline F4 hadamard(F4 a, F4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline float dot(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// SSE2 dot4: (x*x' + z*z') + (y*y' + w*w')
inline float dot(F4 a, F4 b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y + a.w * b.w); }
inline float length(F3 a) { return std::sqrt(dot(a, a)); }
inline F3 cross(F3 a, F3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline F3 xyz(F4 a) { return {a.x, a.y, a.z}; }
inline F4 extend(F3 a, float w) { return {a.x, a.y, a.z, w}; }

inline F3 mul(const Cols3 &m, F3 v) { return (m.x * v.x + m.y * v.y) + m.z * v.z; }
inline F4 mul(const Cols4 &m, F4 v) { return ((m.x * v.x + m.y * v.y) + m.z * v.z) + m.w * v.w; }
inline Cols4 mul(const Cols4 &a, const Cols4 &b) { return {mul(a, b.x), mul(a, b.y), mul(a, b.z), mul(a, b.w)}; }
inline F4 row(const Cols4 &m, int i) {
    const float *x = &m.x.x, *y = &m.y.x, *z = &m.z.x, *w = &m.w.x;
    return {x[i], y[i], z[i], w[i]};
}
inline Affine load_affine(const float *g) {
    return {{{g[0], g[1], g[2]}, {g[3], g[4], g[5]}, {g[6], g[7], g[8]}}, {g[9], g[10], g[11]}};
}
inline Cols4 load_mat4(const float *p) {
    return {{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6], p[7]}, {p[8], p[9], p[10], p[11]}, {p[12], p[13], p[14], p[15]}};
}
// glam Mat3A::inverse
inline Cols3 inverse(const Cols3 &m) {
    F3 t0 = cross(m.y, m.z), t1 = cross(m.z, m.x), t2 = cross(m.x, m.y);
    float inv = 1.0f / dot(m.z, t2);
    F3 c0 = t0 * inv, c1 = t1 * inv, c2 = t2 * inv;
    return {{c0.x, c1.x, c2.x}, {c0.y, c1.y, c2.y}, {c0.z, c1.z, c2.z}};
}
// glam Affine3A::inverse
inline Affine inverse(const Affine &a) {
    Affine r; r.m = inverse(a.m); r.t = -mul(r.m, a.t); return r;
}
inline Cols4 to_mat4(const Affine &a) {
    return {extend(a.m.x, 0.0f), extend(a.m.y, 0.0f), extend(a.m.z, 0.0f), extend(a.t, 1.0f)};
}
// glam Mat4::inverse (cofactor expansion)
Cols4 inverse(const Cols4 &m) {
    const float m00 = m.x.x, m01 = m.x.y, m02 = m.x.z, m03 = m.x.w;
    const float m10 = m.y.x, m11 = m.y.y, m12 = m.y.z, m13 = m.y.w;
    const float m20 = m.z.x, m21 = m.z.y, m22 = m.z.z, m23 = m.z.w;
    const float m30 = m.w.x, m31 = m.w.y, m32 = m.w.z, m33 = m.w.w;
    const float c00 = m22 * m33 - m32 * m23, c02 = m12 * m33 - m32 * m13, c03 = m12 * m23 - m22 * m13;
    const float c04 = m21 * m33 - m31 * m23, c06 = m11 * m33 - m31 * m13, c07 = m11 * m23 - m21 * m13;
    const float c08 = m21 * m32 - m31 * m22, c10 = m11 * m32 - m31 * m12, c11 = m11 * m22 - m21 * m12;
    const float c12 = m20 * m33 - m30 * m23, c14 = m10 * m33 - m30 * m13, c15 = m10 * m23 - m20 * m13;
    const float c16 = m20 * m32 - m30 * m22, c18 = m10 * m32 - m30 * m12, c19 = m10 * m22 - m20 * m12;
    const float c20 = m20 * m31 - m30 * m21, c22 = m10 * m31 - m30 * m11, c23 = m10 * m21 - m20 * m11;
    const F4 f0{c00, c00, c02, c03}, f1{c04, c04, c06, c07}, f2{c08, c08, c10, c11};
    const F4 f3{c12, c12, c14, c15}, f4{c16, c16, c18, c19}, f5{c20, c20, c22, c23};
    const F4 v0{m10, m00, m00, m00}, v1{m11, m01, m01, m01}, v2{m12, m02, m02, m02}, v3{m13, m03, m03, m03};
    const F4 i0 = (hadamard(v1, f0) - hadamard(v2, f1)) + hadamard(v3, f2);
    const F4 i1 = (hadamard(v0, f0) - hadamard(v2, f3)) + hadamard(v3, f4);
    const F4 i2 = (hadamard(v0, f1) - hadamard(v1, f3)) + hadamard(v3, f5);
    const F4 i3 = (hadamard(v0, f2) - hadamard(v1, f4)) + hadamard(v2, f5);
    const F4 sa{1.0f, -1.0f, 1.0f, -1.0f}, sb{-1.0f, 1.0f, -1.0f, 1.0f};
    const F4 a0 = hadamard(i0, sa), a1 = hadamard(i1, sb), a2 = hadamard(i2, sa), a3 = hadamard(i3, sb);
    const float det = dot(m.x, F4{a0.x, a1.x, a2.x, a3.x});
    const float rcp = 1.0f / det;
    return {a0 * rcp, a1 * rcp, a2 * rcp, a3 * rcp};
}
// HalfSpace::new (crates/bevy_math/src/primitives/half_space.rs:53-57)
inline F4 half_space(F4 nd) { return nd * (1.0f / length(xyz(nd))); }
inline void store(float *dst, F4 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }

// Rust `as u32` (saturating, NaN -> 0)
inline uint32_t as_u32(float f) {
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return static_cast<uint32_t>(f);
}
inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// clip_to_view (assign.rs:1064-1067)
inline F4 clip_to_view(const Cols4 &view_from_clip, F4 clip) {
    F4 v = mul(view_from_clip, clip);
    return {v.x / v.w, v.y / v.w, v.z / v.w, v.w / v.w};
}

}  // namespace

void perspective_infinite_reverse_rh(float fov_y, float aspect, float near_z, float *out16) {
    const float f = 1.0f / std::tan(0.5f * fov_y);
    const float m[16] = {f / aspect, 0, 0, 0, 0, f, 0, 0, 0, 0, 0, -1.0f, 0, 0, near_z, 0};
    std::memcpy(out16, m, sizeof m);
}

// CameraProjection::compute_frustum + ViewFrustum::from_clip_from_world_custom_far
// (crates/bevy_camera/src/projection.rs:72-80, crates/bevy_math/src/primitives/view_frustum.rs:51-107)
void compute_frustum(const float *clip_from_view16, const float *camera_gt12, float far_z, float hs[6][4]) {
    const Cols4 cfv = load_mat4(clip_from_view16);
    const Affine cam = load_affine(camera_gt12);
    const Cols4 cfw = mul(cfv, to_mat4(inverse(cam)));
    const F4 r0 = row(cfw, 0), r1 = row(cfw, 1), r2 = row(cfw, 2), r3 = row(cfw, 3);
    store(hs[0], half_space(r3 + r0));
    store(hs[1], half_space(r3 - r0));
    store(hs[2], half_space(r3 + r1));
    store(hs[3], half_space(r3 - r1));
    store(hs[4], half_space(r3 + r2));
    // GlobalTransform::back(): (matrix3 * Vec3::Z).normalize()
    const F3 zaxis = mul(cam.m, F3{0.0f, 0.0f, 1.0f});
    const F3 back = zaxis * (1.0f / length(zaxis));
    const F3 far_center = cam.t - back * far_z;
    store(hs[5], half_space(extend(back, -dot(back, far_center))));
}

// ---- CubemapFrusta of a point light --------------------------------------------------------------
// update_point_light_frusta (crates/bevy_light/src/point_light.rs:212-265): six 90-degree views along the world axes
// (CUBE_MAP_FACES, bevy_camera/src/primitives.rs:348-379) at the light's translation; every face shares one far
// plane, `range` behind the light along the LIGHT's back direction.  For hosts without glam; a Rust shim passes the
// CubemapFrusta component instead.
namespace {
// glam Quat::from_rotation_axes on the columns (right, up, back) that Transform::look_to builds (transform.rs:475-484)
F4 rotation_from_axes(F3 xa, F3 ya, F3 za) {
    if (za.z <= 0.0f) {
        const float dif10 = ya.y - xa.x, omm22 = 1.0f - za.z;
        if (dif10 <= 0.0f) {
            const float four_xsq = omm22 - dif10, inv = 0.5f / std::sqrt(four_xsq);
            return {four_xsq * inv, (xa.y + ya.x) * inv, (xa.z + za.x) * inv, (ya.z - za.y) * inv};
        }
        const float four_ysq = omm22 + dif10, inv = 0.5f / std::sqrt(four_ysq);
        return {(xa.y + ya.x) * inv, four_ysq * inv, (ya.z + za.y) * inv, (za.x - xa.z) * inv};
    }
    const float sum10 = ya.y + xa.x, opm22 = 1.0f + za.z;
    if (sum10 <= 0.0f) {
        const float four_zsq = opm22 - sum10, inv = 0.5f / std::sqrt(four_zsq);
        return {(xa.z + za.x) * inv, (ya.z + za.y) * inv, four_zsq * inv, (xa.y - ya.x) * inv};
    }
    const float four_wsq = opm22 + sum10, inv = 0.5f / std::sqrt(four_wsq);
    return {(ya.z - za.y) * inv, (za.x - xa.z) * inv, (xa.y - ya.x) * inv, four_wsq * inv};
}
// Affine3A::from_scale_rotation_translation(ONE, q, t)  (Mat3A::from_quat, each column * 1.0)
Affine affine_from_rotation_translation(F4 q, F3 t) {
    const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    const float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2, yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
    const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    Affine a;
    a.m.x = F3{1.0f - (yy + zz), xy + wz, xz - wy} * 1.0f;
    a.m.y = F3{xy - wz, 1.0f - (xx + zz), yz + wx} * 1.0f;
    a.m.z = F3{xz + wy, yz - wx, 1.0f - (xx + yy)} * 1.0f;
    a.t = t;
    return a;
}
}  // namespace
void point_light_frusta(const float *light_gt12, float range, float shadow_map_near_z, float hs[6][6][4]) {
    struct Face { F3 target, up; };
    static const Face kFaces[6] = {{{1, 0, 0}, {0, 1, 0}},  {{-1, 0, 0}, {0, 1, 0}}, {{0, 1, 0}, {0, 0, 1}},
                                   {{0, -1, 0}, {0, 0, -1}}, {{0, 0, -1}, {0, 1, 0}}, {{0, 0, 1}, {0, 1, 0}}};
    const Affine light = load_affine(light_gt12);
    float cfv16[16];
    perspective_infinite_reverse_rh(1.57079632679489661923f, 1.0f, shadow_map_near_z, cfv16);
    const Cols4 cfv = load_mat4(cfv16);
    const F3 zaxis = mul(light.m, F3{0.0f, 0.0f, 1.0f});
    const F3 view_backward = zaxis * (1.0f / length(zaxis));              // GlobalTransform::back()
    const F3 far_center = light.t - view_backward * range;
    const F4 far_plane = half_space(extend(view_backward, -dot(view_backward, far_center)));
    for (int f = 0; f < 6; ++f) {
        const F3 dir = kFaces[f].target * (1.0f / length(kFaces[f].target));   // Dir3::new (unit axes: exact)
        const F3 back = -dir;
        const F3 up0 = kFaces[f].up * (1.0f / length(kFaces[f].up));
        F3 right = cross(up0, back);
        right = right * (1.0f / length(right));
        const F3 up = cross(back, right);
        const Affine world_from_view = affine_from_rotation_translation(rotation_from_axes(right, up, back), light.t);
        const Cols4 cfw = mul(cfv, to_mat4(inverse(world_from_view)));
        const F4 r0 = row(cfw, 0), r1 = row(cfw, 1), r2 = row(cfw, 2), r3 = row(cfw, 3);
        store(hs[f][0], half_space(r3 + r0));
        store(hs[f][1], half_space(r3 - r0));
        store(hs[f][2], half_space(r3 + r1));
        store(hs[f][3], half_space(r3 - r1));
        store(hs[f][4], half_space(r3 + r2));
        store(hs[f][5], far_plane);
    }
}

void default_cluster_config(b200vis_cluster_config *c, uint32_t w, uint32_t h) {
    // ClusterConfig::default() + ClusterZConfig::default() (cluster/mod.rs:288-307)
    std::memset(c, 0, sizeof *c);
    c->kind = 3; c->total = 4096; c->z_slices = 24; c->first_slice_depth = 5.0f;
    c->far_z_mode = 0; c->dynamic_resizing = 1; c->screen_w = w; c->screen_h = h;
    c->view_cluster_bindings_max_indices = 16384;   // the GlobalClusterSettings of the CPU harness
}

// ClusterConfig::dimensions_for_screen_size (cluster/mod.rs:311-347)
static void dimensions_for_screen_size(const b200vis_cluster_config &c, uint32_t out[3]) {
    switch (c.kind) {
    case 0: out[0] = out[1] = out[2] = 0; return;
    case 1: out[0] = out[1] = out[2] = 1; return;
    case 2: out[0] = c.dims[0]; out[1] = c.dims[1]; out[2] = c.dims[2]; return;
    default: break;
    }
    const float aspect = static_cast<float>(c.screen_w) / static_cast<float>(c.screen_h);
    uint32_t z_slices = c.z_slices;
    if (c.total < z_slices) z_slices = c.total;
    const float per_layer = static_cast<float>(c.total) / static_cast<float>(z_slices);
    const float yf = std::sqrt(per_layer / aspect);
    uint32_t x = as_u32(yf * aspect), y = as_u32(yf);
    if (x == 0) { x = 1; y = as_u32(per_layer); }
    if (y == 0) { x = as_u32(per_layer); y = 1; }
    out[0] = x; out[1] = y; out[2] = z_slices;
}

// z_slice_to_view_z (assign.rs:903-920)
static float z_slice_to_view_z(float near_z, float far_z, uint32_t z_slices, uint32_t z, bool ortho) {
    if (ortho) return -near_z - (far_z - near_z) * static_cast<float>(z) / static_cast<float>(z_slices);
    if (z == 0) return 0.0f;
    return -near_z * std::pow(far_z / near_z, static_cast<float>(z - 1) / static_cast<float>(z_slices - 1));
}

int32_t cluster_view_setup(const b200vis_cluster_config *cfg, const float *camera_gt12,
                           const float *clip_from_view16, const float frustum[6][4], uint64_t layer_mask,
                           const b200vis_cluster_feedback *fb, float *scratch, b200vis_cluster_view *out) {
    std::memset(out, 0, sizeof *out);
    out->layer_mask = layer_mask;
    std::memcpy(out->frustum, frustum, sizeof out->frustum);
    std::memcpy(out->clip_from_view, clip_from_view16, sizeof out->clip_from_view);
    // ClusterConfig::None or a zero-sized viewport => clusters.clear() (assign.rs:329-340)
    if (cfg->kind == 0 || cfg->screen_w == 0 || cfg->screen_h == 0) {
        out->enabled = 0; out->tile_size[0] = out->tile_size[1] = 1;
        return B200VIS_OK;
    }
    uint32_t req[3];
    dimensions_for_screen_size(*cfg, req);
    const Affine cam = load_affine(camera_gt12);
    const Cols4 cfv = load_mat4(clip_from_view16);
    // camera_transform.compute_transform().scale.recip()  (glam to_scale_rotation_translation)
    const float det = dot(cam.m.z, cross(cam.m.x, cam.m.y));
    const F3 scale{length(cam.m.x) * std::copysign(1.0f, det), length(cam.m.y), length(cam.m.z)};
    const F3 inv_scale{1.0f / scale.x, 1.0f / scale.y, 1.0f / scale.z};
    float smax = std::fabs(inv_scale.x);
    if (std::fabs(inv_scale.y) > smax) smax = std::fabs(inv_scale.y);
    if (std::fabs(inv_scale.z) > smax) smax = std::fabs(inv_scale.z);
    const Cols4 vfw = to_mat4(inverse(cam));
    const bool ortho = cfv.w.w == 1.0f;
    const float cfg_first = cfg->kind == 1 ? 0.0f : cfg->first_slice_depth;
    const uint32_t far_mode = cfg->kind == 1 ? 0u : cfg->far_z_mode;
    float far_z = far_mode == 0 ? ((fb && fb->has_farthest_z) ? fb->farthest_z : 1000.0f) : cfg->far_z_constant;
    float first;
    if (ortho) first = (cfv.w.z - 1.0f) / cfv.z.z;
    else if (req[2] == 1) first = std::fmax(cfg_first, far_z);
    else first = cfg_first;
    first = first * inv_scale.z;
    far_z = std::fmax(far_z, first);
    // calculate_cluster_factors (assign.rs:817-832)
    const float zs = static_cast<float>(req[2]);
    if (ortho) { out->cluster_factors[0] = -first; out->cluster_factors[1] = zs / (-far_z - -first); }
    else {
        const float k = (zs - 1.0f) / std::log(far_z / first);
        out->cluster_factors[0] = k; out->cluster_factors[1] = std::log(first) * k;
    }
    const bool dyn = cfg->kind >= 2 && cfg->dynamic_resizing;
    if (dyn && fb && fb->has_index_count && fb->index_count > cfg->view_cluster_bindings_max_indices) {
        const float ratio = static_cast<float>(cfg->view_cluster_bindings_max_indices) / static_cast<float>(fb->index_count);
        const float xy = std::sqrt(ratio);
        req[0] = umax(as_u32(std::floor(static_cast<float>(req[0]) * xy)), 1u);
        req[1] = umax(as_u32(std::floor(static_cast<float>(req[1]) * xy)), 1u);
    }
    // Clusters::update (cluster/mod.rs:398-416)
    const float w = static_cast<float>(cfg->screen_w), h = static_cast<float>(cfg->screen_h);
    out->tile_size[0] = umax(as_u32(std::ceil(w / static_cast<float>(req[0]))), 1u);
    out->tile_size[1] = umax(as_u32(std::ceil(h / static_cast<float>(req[1]))), 1u);
    out->dims[0] = umax(as_u32(std::ceil(w / static_cast<float>(out->tile_size[0]))), 1u);
    out->dims[1] = umax(as_u32(std::ceil(h / static_cast<float>(out->tile_size[1]))), 1u);
    out->dims[2] = umax(req[2], 1u);
    if (static_cast<uint64_t>(out->dims[0]) * out->dims[1] * out->dims[2] > B200VIS_MAX_CLUSTERS)
        return B200VIS_ERR_CAPACITY;
    out->enabled = 1;
    out->is_orthographic = ortho ? 1u : 0u;
    out->near_z = first; out->far_z = far_z;
    std::memcpy(out->view_from_world, &vfw, sizeof out->view_from_world);
    out->view_from_world_scale[0] = inv_scale.x; out->view_from_world_scale[1] = inv_scale.y;
    out->view_from_world_scale[2] = inv_scale.z; out->view_from_world_scale_max = smax;
    // plane tables (assign.rs:429-485)
    const Cols4 vfc = inverse(cfv);
    float *xp = scratch, *yp = scratch + 4097 * 4, *zp = scratch + 2 * 4097 * 4;
    for (uint32_t x = 0; x <= out->dims[0]; ++x) {
        const float x_pos = (static_cast<float>(x) / static_cast<float>(out->dims[0])) * 2.0f - 1.0f;
        if (ortho) {
            const float view_x = clip_to_view(vfc, {x_pos, 0.0f, 1.0f, 1.0f}).x;
            store(xp + 4 * x, half_space({1.0f, 0.0f, 0.0f, view_x * 1.0f}));
        } else {
            const F3 nb = xyz(clip_to_view(vfc, {x_pos, -1.0f, 1.0f, 1.0f}));
            const F3 nt = xyz(clip_to_view(vfc, {x_pos, 1.0f, 1.0f, 1.0f}));
            const F3 normal = cross(nb, nt);
            store(xp + 4 * x, half_space(extend(normal, dot(nb, normal))));
        }
    }
    for (uint32_t y = 0; y <= out->dims[1]; ++y) {
        const float y_pos = (1.0f - static_cast<float>(y) / static_cast<float>(out->dims[1])) * 2.0f - 1.0f;
        if (ortho) {
            const float view_y = clip_to_view(vfc, {0.0f, y_pos, 1.0f, 1.0f}).y;
            store(yp + 4 * y, half_space({0.0f, 1.0f, 0.0f, view_y * 1.0f}));
        } else {
            const F3 nl = xyz(clip_to_view(vfc, {-1.0f, y_pos, 1.0f, 1.0f}));
            const F3 nr = xyz(clip_to_view(vfc, {1.0f, y_pos, 1.0f, 1.0f}));
            const F3 normal = cross(nr, nl);
            store(yp + 4 * y, half_space(extend(normal, dot(nr, normal))));
        }
    }
    for (uint32_t z = 0; z <= out->dims[2]; ++z) {
        const float view_z = z_slice_to_view_z(first, far_z, out->dims[2], z, ortho);
        store(zp + 4 * z, half_space({-0.0f, -0.0f, -1.0f, view_z * -1.0f}));
    }
    out->x_planes = xp; out->y_planes = yp; out->z_planes = zp;
    return B200VIS_OK;
}

// ---- z-slice thresholds ------------------------------------------------------------
// view_z_to_z_slice (assign.rs:1046-1062) needs ops::ln per light, i.e. the HOST's libm logf.
// Rather than re-implementing a libm on the device (1-ulp differences move lights across slice
// boundaries), the host finds, with its own libm, the exact float thresholds at which the slice
// index steps; the device then only compares.  g(u), u = -view_z, is evaluated with exactly the
// reference's float expression; it is monotone non-decreasing in u, so bisection over the ordered
// float bit patterns finds t_k = min{u : g(u) >= k}.
static inline uint32_t slice_of(float u, const float f[2], bool ortho) {
    const float view_z = -u;
    if (ortho) return as_u32(std::floor((view_z - f[0]) * f[1]));
    return as_u32(std::log(-view_z) * f[0] - f[1] + 1.0f);
}
// order-preserving map float <-> uint32 (total order, -inf .. +inf; NaNs excluded)
static inline uint32_t f2key(float f) {
    uint32_t b; std::memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static inline float key2f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float f; std::memcpy(&f, &b, 4); return f;
}

void z_slice_thresholds(const float factors[2], uint32_t z_slices, bool ortho, float *thresholds /* z_slices-1 */) {
    const uint32_t lo_key = f2key(-INFINITY), hi_key = f2key(INFINITY);
    for (uint32_t k = 1; k < z_slices; ++k) {
        // smallest key in [lo_key, hi_key] with slice_of >= k; NaN threshold if none
        if (slice_of(INFINITY, factors, ortho) < k) { thresholds[k - 1] = NAN; continue; }
        uint32_t lo = lo_key, hi = hi_key;   // invariant: slice_of(lo) < k <= slice_of(hi)
        if (slice_of(key2f(lo), factors, ortho) >= k) { thresholds[k - 1] = -INFINITY; continue; }
        // Start from the analytic inverse of the slice formula and bracket it tightly: the exact step lies within a
        // few hundred ulps of it, so ~10 evaluations of the host libm replace a 32-step bisection over all floats.
        // (Any failure of the bracket falls back to the full range: the result is the same exact threshold.)
        {
            const float guess = ortho ? -(static_cast<float>(k) / factors[1] + factors[0])
                                      : std::exp((static_cast<float>(k) - 1.0f + factors[1]) / factors[0]);
            if (std::isfinite(guess)) {
                const uint32_t g = f2key(guess);
                const uint32_t span = 512;
                const uint32_t blo = g > lo_key + span ? g - span : lo_key, bhi = g < hi_key - span ? g + span : hi_key;
                if (slice_of(key2f(blo), factors, ortho) < k && slice_of(key2f(bhi), factors, ortho) >= k) { lo = blo; hi = bhi; }
            }
        }
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (slice_of(key2f(mid), factors, ortho) >= k) hi = mid; else lo = mid;
        }
        thresholds[k - 1] = key2f(hi);
    }
}

}}  // namespace b200vis::host

// ---- C ABI wrappers ----------------------------------------------------------------
extern "C" {
void b200vis_host_perspective(float fov_y, float aspect, float near_z, float *out16) {
    b200vis::host::perspective_infinite_reverse_rh(fov_y, aspect, near_z, out16);
}
void b200vis_host_compute_frustum(const float *cfv16, const float *cam12, float far_z, float hs[6][4]) {
    b200vis::host::compute_frustum(cfv16, cam12, far_z, hs);
}
void b200vis_host_point_light_frusta(const float *light_gt12, float range, float shadow_map_near_z, float hs[6][6][4]) {
    b200vis::host::point_light_frusta(light_gt12, range, shadow_map_near_z, hs);
}
void b200vis_host_z_slice_thresholds(const float f[2], uint32_t z_slices, uint32_t ortho, float *thr) {
    b200vis::host::z_slice_thresholds(f, z_slices, ortho != 0, thr);
}
void b200vis_host_default_cluster_config(b200vis_cluster_config *cfg, uint32_t w, uint32_t h) {
    b200vis::host::default_cluster_config(cfg, w, h);
}
int32_t b200vis_host_cluster_view_setup(const b200vis_cluster_config *cfg, const float *camera_gt12,
                                        const float *clip_from_view16, const float frustum[6][4],
                                        uint64_t layer_mask, const b200vis_cluster_feedback *feedback,
                                        float *planes_scratch, b200vis_cluster_view *out) {
    if (!cfg || !camera_gt12 || !clip_from_view16 || !frustum || !planes_scratch || !out) return B200VIS_ERR_INVALID_ARG;
    return b200vis::host::cluster_view_setup(cfg, camera_gt12, clip_from_view16, frustum, layer_mask, feedback,
                                             planes_scratch, out);
}
}
