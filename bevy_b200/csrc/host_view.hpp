// host_view.hpp -- host-side mirror of the reference's per-view math (see host_view.cpp).
#pragma once
#include <cstdint>
#include "../../include/b200vis.h"

namespace b200vis { namespace host {
void perspective_infinite_reverse_rh(float fov_y, float aspect, float near_z, float *out16);
void compute_frustum(const float *clip_from_view16, const float *camera_gt12, float far_z, float hs[6][4]);
void point_light_frusta(const float *light_gt12, float range, float shadow_map_near_z, float hs[6][6][4]);
void default_cluster_config(b200vis_cluster_config *c, uint32_t w, uint32_t h);
int32_t cluster_view_setup(const b200vis_cluster_config *cfg, const float *camera_gt12,
                           const float *clip_from_view16, const float frustum[6][4], uint64_t layer_mask,
                           const b200vis_cluster_feedback *fb, float *scratch, b200vis_cluster_view *out);
// thresholds[k-1] = smallest u = -view_z whose z-slice (assign.rs:1046-1062, host libm) is >= k
void z_slice_thresholds(const float factors[2], uint32_t z_slices, bool ortho, float *thresholds);
}}
