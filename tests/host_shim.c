/*
 * host_shim.c -- the Rust shim's per-frame sequence (rust/b200vis_plugin.rs), written in plain C against
 * include/b200vis.h: proof that the C ABI is usable without Python, on Bevy's own column strides, and a measurement of what
 * the shim costs on the host (column repack, tick stamping, VisibleEntities / Clusters fill).
 *
 * The "ECS" here is a set of table columns in Bevy's native layouts (x86-64 + SSE2 glam):
 *     Transform        48 B  { translation: Vec3, rotation: Quat (align 16), scale: Vec3 } -- repr(Rust): the field offsets
 *                            are what core::mem::offset_of! reports; this harness uses rotation @0, translation @16, scale @32
 *     GlobalTransform  64 B  glam Affine3A { x_axis, y_axis, z_axis, translation: Vec3A }
 *     Aabb             32 B  { center: Vec3A, half_extents: Vec3A }
 *     Entity            8 B  index | generation << 32
 *     ViewVisibility    1 B, change ticks 4 B per row and column
 * Flow (the three replacement systems + plugin start-up):
 *   start-up   ChildOf -> parent rows, b200vis_plan_row_order, repack the columns into the ABI's arrays, b200vis_create,
 *              b200vis_set_topology, uploads, b200vis_set_lights, result + column sinks registered ON THE ECS COLUMNS
 *   per frame  rows matching Changed<Transform> -> b200vis_step(upload, cameras, run, write-back, wait); then
 *              stamp changed_ticks from the change bits, VisibleEntities::entities[class] = entity ids of the sorted rows,
 *              Clusters: one Vec<Entity> per cluster from the CSR, Clusters::last_frame_* from the frame stats
 * Checked against the CPU oracle (oracle/libbevy_oracle.so, test infrastructure): GlobalTransform column bits, change
 * ticks, ViewVisibility bytes, VisibleEntities (as entity ids) for every frame.
 *
 * Build (tests/test_gpu_host_shim.py does this): gcc -O2 -std=gnu11 -Iinclude tests/host_shim.c -o host_shim
 *        -Lbevy_b200 -lb200vis -Loracle -lbevy_oracle -lm -Wl,-rpath,... ; run: ./host_shim [n_trees] [levels] [frames]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200vis.h"

/* oracle entry points used as the checker (oracle/bevy_oracle.c) */
int orc_propagate(uint32_t n, const uint32_t *parent, const float *trs, float *gt, const uint8_t *tchanged,
                  const uint8_t *gt_ext_changed, int static_opt, uint8_t *changed);
int orc_cull(uint32_t n, const float *gt, const float *bounds, const uint8_t *flags, const uint64_t *layer_mask,
             const uint32_t *range_mask, const uint8_t *class_mask, const uint64_t *entity_bits, uint8_t *vv, uint8_t *vv_changed,
             uint32_t n_views, const float *view_planes, const uint64_t *view_layers, const uint8_t *view_flags,
             const int8_t *view_range_index, uint32_t *visible_rows, uint32_t *visible_count);

typedef struct { float rotation[4]; float translation[3]; float pad0; float scale[3]; float pad1; } BevyTransform;   /* 48 B */
typedef struct { float m[16]; } BevyGlobalTransform;                                                                /* 64 B */
typedef struct { float center[4]; float half_extents[4]; } BevyAabb;                                                /* 32 B */
typedef uint64_t Entity;

#define CHECK(call)                                                                                     \
    do {                                                                                                \
        int32_t rc_ = (call);                                                                           \
        if (rc_ != B200VIS_OK) {                                                                        \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, b200vis_last_error(ctx));               \
            return 2;                                                                                   \
        }                                                                                               \
    } while (0)

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static void *palloc(size_t bytes) { return aligned_alloc(4096, (bytes + 4095) & ~(size_t)4095); }   /* page aligned: cudaHostRegister */
static uint64_t rng_state = 42;
static float frand(float lo, float hi) {          /* SplitMix64 */
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return lo + (hi - lo) * (float)((z >> 40) * (1.0 / 16777216.0));
}

int main(int argc, char **argv) {
    const uint32_t n_trees = argc > 1 ? (uint32_t)atoi(argv[1]) : 400, levels = argc > 2 ? (uint32_t)atoi(argv[2]) : 6;
    const uint32_t frames = argc > 3 ? (uint32_t)atoi(argv[3]) : 4, n_lights = 48, V = 2;
    const uint32_t per = (1u << levels) - 1, n = n_trees * per + n_lights;
    b200vis_ctx *ctx = NULL;

    /* ---- the ECS tables (spawn order: each tree breadth first, then the lights) ---------------------------------------- */
    BevyTransform *transform = palloc((size_t)n * sizeof *transform);
    BevyGlobalTransform *global = palloc((size_t)n * sizeof *global);
    BevyAabb *aabb = palloc((size_t)n * sizeof *aabb);
    Entity *entity = malloc((size_t)n * sizeof *entity);
    uint32_t *child_of = malloc((size_t)n * 4);                  /* ChildOf as a table row, NO_PARENT = none */
    uint8_t *view_visibility = palloc(n);
    uint32_t *gt_ticks = calloc(n, 4), *vv_ticks = calloc(n, 4);
    uint8_t *is_light = calloc(n, 1);
    float *light_range = malloc(n_lights * 4);
    for (uint32_t r = 0; r < n; ++r) {
        const uint32_t tree = r / per, k = r % per;
        const int light = r >= n_trees * per;
        memset(&transform[r], 0, sizeof transform[r]);
        float q[4] = {frand(-1, 1), frand(-1, 1), frand(-1, 1), frand(-1, 1)};
        float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int i = 0; i < 4; ++i) transform[r].rotation[i] = light ? (i == 3) : q[i] / qn;
        const float spread = (light || k == 0) ? 200.0f : 2.0f;
        for (int i = 0; i < 3; ++i) { transform[r].translation[i] = frand(-spread, spread); transform[r].scale[i] = light ? 1.0f : frand(0.5f, 1.5f); }
        memset(&global[r], 0, sizeof global[r]);
        global[r].m[0] = global[r].m[5] = global[r].m[10] = 1.0f;     /* GlobalTransform::IDENTITY */
        memset(&aabb[r], 0, sizeof aabb[r]);
        for (int i = 0; i < 3; ++i) aabb[r].half_extents[i] = frand(0.25f, 0.75f);
        entity[r] = r;                                                 /* generation 0 */
        child_of[r] = (light || k == 0) ? B200VIS_NO_PARENT : tree * per + (k - 1) / 2;
        is_light[r] = (uint8_t)light;
        if (light) light_range[r - n_trees * per] = frand(5.0f, 40.0f);
        view_visibility[r] = 0;
    }

    /* ---- plugin start-up: row order, column repack, context ------------------------------------------------------------------- */
    uint32_t *new_to_old = malloc((size_t)n * 4);
    if (b200vis_plan_row_order(n, child_of, new_to_old) != B200VIS_OK) { fprintf(stderr, "plan_row_order failed\n"); return 2; }
    int identity = 1;
    for (uint32_t r = 0; r < n; ++r) identity &= new_to_old[r] == r;
    if (!identity) { fprintf(stderr, "this harness spawns in the planned order; got a permutation\n"); return 2; }
    const double t_repack0 = now_ms();
    float *trs = malloc((size_t)n * 40), *bounds = malloc((size_t)n * 24);
    uint8_t *flags = malloc(n), *cls = malloc(n);
    for (uint32_t r = 0; r < n; ++r) {                               /* Transform / Aabb -> the ABI's packed rows */
        float *t = trs + (size_t)r * 10, *b = bounds + (size_t)r * 6;
        memcpy(t, transform[r].translation, 12); memcpy(t + 3, transform[r].rotation, 16); memcpy(t + 7, transform[r].scale, 12);
        if (is_light[r]) {                                           /* Sphere { center: GT.translation, radius: range } */
            memset(b, 0, 24); b[3] = light_range[r - n_trees * per];
            flags[r] = B200VIS_F_INHERITED_VISIBLE | B200VIS_F_HAS_SPHERE | B200VIS_F_SPHERE_FROM_GT;
        } else {
            memcpy(b, aabb[r].center, 12); memcpy(b + 3, aabb[r].half_extents, 12);
            flags[r] = B200VIS_F_INHERITED_VISIBLE | B200VIS_F_HAS_AABB;
        }
        cls[r] = 1;
    }
    const double repack_ms = now_ms() - t_repack0;
    b200vis_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.max_entities = n; cfg.max_lights = n_lights; cfg.max_views = V;
    if (b200vis_create(&cfg, &ctx) != B200VIS_OK) { fprintf(stderr, "b200vis_create: %s\n", b200vis_last_error(NULL)); return 3; }
    CHECK(b200vis_set_topology(ctx, n, child_of, entity));
    CHECK(b200vis_upload_transforms(ctx, 0, n, trs));
    {
        float *gt12 = malloc((size_t)n * 48);
        for (uint32_t r = 0; r < n; ++r) { float *g = gt12 + (size_t)r * 12; memset(g, 0, 48); g[0] = g[4] = g[8] = 1.0f; }
        CHECK(b200vis_upload_global_transforms(ctx, 0, n, gt12));
        free(gt12);
    }
    CHECK(b200vis_upload_bounds(ctx, 0, n, bounds, flags, cls, NULL, NULL));
    CHECK(b200vis_upload_view_visibility(ctx, 0, n, view_visibility));
    uint32_t *light_row = malloc(n_lights * 4);
    for (uint32_t i = 0; i < n_lights; ++i) light_row[i] = n_trees * per + i;
    CHECK(b200vis_set_lights(ctx, n_lights, light_row, light_range, NULL));
    /* results land in the shim's buffers / straight in the ECS columns */
    b200vis_frame_stats *stats = palloc(sizeof(b200vis_frame_stats));
    uint32_t *visible_rows = palloc((size_t)V * n * 4), *cl_off = palloc((size_t)B200VIS_MAX_VIEWS * 4097 * 4);
    const uint32_t cl_cap = 1u << 16;
    uint32_t *cl_idx = palloc((size_t)B200VIS_MAX_VIEWS * cl_cap * 4);
    uint32_t *gt_bits = palloc(((size_t)n + 31) / 32 * 4), *vv_bits = palloc(((size_t)n + 31) / 32 * 4);
    memset(gt_bits, 0, ((size_t)n + 31) / 32 * 4); memset(vv_bits, 0, ((size_t)n + 31) / 32 * 4);
    b200vis_result_sink rs;
    memset(&rs, 0, sizeof rs);
    rs.stats = stats; rs.visible_rows = visible_rows; rs.visible_capacity = n; rs.cluster_offsets = cl_off; rs.cluster_indices = cl_idx;
    rs.cluster_capacity = cl_cap;
    CHECK(b200vis_set_result_sink(ctx, &rs));
    b200vis_column_sinks cs;
    memset(&cs, 0, sizeof cs);
    cs.global_transforms = (float *)global; cs.gt_stride_floats = 16; cs.gt_changed_bits = gt_bits;
    cs.view_visibility = view_visibility; cs.vv_changed_bits = vv_bits;
    CHECK(b200vis_set_column_sinks(ctx, &cs));
    b200vis_cluster_config ccfg;
    b200vis_host_default_cluster_config(&ccfg, 1920, 1080);

    /* ---- oracle-side world (checker) --------------------------------------------------------------------------------------- */
    float *o_gt = malloc((size_t)n * 48);
    for (uint32_t r = 0; r < n; ++r) { float *g = o_gt + (size_t)r * 12; memset(g, 0, 48); g[0] = g[4] = g[8] = 1.0f; }
    uint8_t *o_vv = calloc(n, 1), *o_vvch = calloc(n, 1), *o_gtch = calloc(n, 1), *tchanged = malloc(n);
    uint32_t *o_rows = malloc((size_t)V * n * 4), o_count[B200VIS_MAX_VIEWS];
    memset(tchanged, 1, n);                                           /* Added<GlobalTransform> on the first frame */

    Entity **visible_entities = calloc(V, sizeof *visible_entities);   /* VisibleEntities::entities[class] per view */
    for (uint32_t v = 0; v < V; ++v) visible_entities[v] = malloc((size_t)n * sizeof(Entity));
    uint32_t *changed_rows = malloc((size_t)n * 4);
    float *changed_trs = malloc((size_t)n * 40);
    double step_ms = 0, post_ms = 0;
    int ok = 1;
    for (uint32_t frame = 1; frame <= frames; ++frame) {
        /* game logic: every root moves (Changed<Transform>), cameras turn */
        uint32_t n_changed = 0;
        if (frame > 1)
            for (uint32_t tr = 0; tr < n_trees; ++tr) {
                const uint32_t r = tr * per;
                transform[r].translation[2] += 0.02f * sinf(0.001f * (float)(frame + tr));
                tchanged[r] = 1;
            }
        for (uint32_t r = 0; r < n; ++r)
            if (tchanged[r] && frame > 1) {                             /* the shim's Changed<Transform> query */
                float *t = changed_trs + (size_t)n_changed * 10;
                memcpy(t, transform[r].translation, 12); memcpy(t + 3, transform[r].rotation, 16); memcpy(t + 7, transform[r].scale, 12);
                memcpy(trs + (size_t)r * 10, t, 40);
                changed_rows[n_changed++] = r;
            }
        b200vis_camera cam[2];
        float planes[2][6][4];
        uint64_t view_layers[2] = {1, 1};
        uint8_t view_flags[2] = {B200VIS_VIEW_ACTIVE, B200VIS_VIEW_ACTIVE};
        for (uint32_t v = 0; v < V; ++v) {
            memset(&cam[v], 0, sizeof cam[v]);
            const float yaw = 0.01f * (float)frame + 1.5707963f * (float)v, cy = cosf(yaw), sy = sinf(yaw);
            const float gt[12] = {cy, 0, -sy, 0, 1, 0, sy, 0, cy, 0, 0, 0};     /* rotation about Y, at the origin */
            memcpy(cam[v].global_transform, gt, sizeof gt);
            cam[v].fov_y = 0.78539816f; cam[v].aspect = 16.0f / 9.0f; cam[v].near_z = 0.1f; cam[v].far_z = 1000.0f;
            cam[v].layer_mask = 1; cam[v].flags = B200VIS_VIEW_ACTIVE; cam[v].range_view_index = -1;
            float cfv[16];
            b200vis_host_perspective(cam[v].fov_y, cam[v].aspect, cam[v].near_z, cfv);
            b200vis_host_compute_frustum(cfv, gt, cam[v].far_z, planes[v]);
        }
        const double t0 = now_ms();
        CHECK(b200vis_step(ctx, n_changed, changed_rows, changed_trs, V, cam, &ccfg, B200VIS_STEP_WAIT | B200VIS_STEP_WRITEBACK));
        CHECK(b200vis_synchronize(ctx));
        const double t1 = now_ms();
        /* ---- what the three systems do with the results ---- */
        for (uint32_t w = 0; w < (n + 31) / 32; ++w) {                  /* tick stamping through changed_ticks_slice_mut() */
            uint32_t g = gt_bits[w], vb = vv_bits[w];
            while (g) { const uint32_t b = (uint32_t)__builtin_ctz(g); g &= g - 1; gt_ticks[w * 32 + b] = frame; }
            while (vb) { const uint32_t b = (uint32_t)__builtin_ctz(vb); vb &= vb - 1; vv_ticks[w * 32 + b] = frame; }
        }
        size_t cluster_pairs = 0;
        for (uint32_t v = 0; v < V; ++v) {                              /* VisibleEntities + Clusters */
            for (uint32_t i = 0; i < stats->visible_count[v]; ++i) visible_entities[v][i] = entity[visible_rows[(size_t)v * n + i]];
            uint32_t dims[3];
            CHECK(b200vis_cluster_view_dims(ctx, v, dims));
            const uint32_t nc = dims[0] * dims[1] * dims[2];
            const uint32_t *off = cl_off + (size_t)v * 4097;
            for (uint32_t c = 0; c < nc; ++c) {
                if (off[c + 1] < off[c] || off[c + 1] > cl_cap) { fprintf(stderr, "frame %u view %u: broken cluster CSR\n", frame, v); ok = 0; break; }
                for (uint32_t i = off[c]; i < off[c + 1]; ++i)          /* ObjectsInClusterCpu::add_point_light(entity) */
                    if (cl_idx[(size_t)v * cl_cap + i] >= n_lights) { fprintf(stderr, "bad light ordinal\n"); ok = 0; }
                cluster_pairs += off[c + 1] - off[c];
            }
            if (off[nc] != stats->cluster_index_count[v]) { fprintf(stderr, "frame %u view %u: index count\n", frame, v); ok = 0; }
        }
        const double t2 = now_ms();
        if (frame > 1) { step_ms += t1 - t0; post_ms += t2 - t1; }
        /* ---- check against the oracle ---- */
        if (orc_propagate(n, child_of, trs, o_gt, tchanged, NULL, 1, o_gtch) != 0) { fprintf(stderr, "oracle propagate failed\n"); return 4; }
        orc_cull(n, o_gt, bounds, flags, NULL, NULL, cls, entity, o_vv, o_vvch, V, &planes[0][0][0], view_layers, view_flags, NULL, o_rows, o_count);
        for (uint32_t r = 0; r < n && ok; ++r) {
            const float *g = o_gt + (size_t)r * 12, *m = global[r].m;
            const float want[16] = {g[0], g[1], g[2], 0, g[3], g[4], g[5], 0, g[6], g[7], g[8], 0, g[9], g[10], g[11], 0};
            if (memcmp(want, m, 64) != 0) { fprintf(stderr, "frame %u row %u: GlobalTransform column differs\n", frame, r); ok = 0; }
            if ((gt_ticks[r] == frame) != (o_gtch[r] != 0)) { fprintf(stderr, "frame %u row %u: Changed<GlobalTransform>\n", frame, r); ok = 0; }
            if (view_visibility[r] != o_vv[r]) { fprintf(stderr, "frame %u row %u: ViewVisibility %u vs %u\n", frame, r, view_visibility[r], o_vv[r]); ok = 0; }
            if ((vv_ticks[r] == frame) != (o_vvch[r] != 0)) { fprintf(stderr, "frame %u row %u: Changed<ViewVisibility>\n", frame, r); ok = 0; }
        }
        for (uint32_t v = 0; v < V && ok; ++v) {
            if (stats->visible_count[v] != o_count[v]) { fprintf(stderr, "frame %u view %u: %u visible vs %u\n", frame, v, stats->visible_count[v], o_count[v]); ok = 0; break; }
            for (uint32_t i = 0; i < o_count[v]; ++i)
                if (visible_entities[v][i] != entity[o_rows[(size_t)v * n + i]]) { fprintf(stderr, "frame %u view %u: VisibleEntities[%u]\n", frame, v, i); ok = 0; break; }
        }
        memset(tchanged, 0, n);
        printf("frame %u: %u rows uploaded, %u GlobalTransforms written back, visible %u / %u, %zu (cluster, light) pairs: %s\n", frame,
               n_changed, stats->gt_changed_count, stats->visible_count[0], stats->visible_count[1], cluster_pairs, ok ? "OK" : "MISMATCH");
        if (!ok) break;
    }
    const double fr = frames > 1 ? (double)(frames - 1) : 1.0;
    printf("{\"entities\": %u, \"repack_ms\": %.3f, \"step_ms_per_frame\": %.3f, \"post_ms_per_frame\": %.3f}\n", n, repack_ms, step_ms / fr,
           post_ms / fr);
    b200vis_set_column_sinks(ctx, NULL);
    b200vis_set_result_sink(ctx, NULL);
    b200vis_destroy(ctx);
    printf(ok ? "HOST_SHIM OK\n" : "HOST_SHIM FAILED\n");
    return ok ? 0 : 1;
}
