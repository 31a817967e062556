//! b200vis_plugin.rs — the Bevy-side shim for libb200vis (SOURCE ONLY: there is no Rust toolchain in the
//! build image; compile it in a crate that depends on bevy 0.20 and links `b200vis`).
//!
//! It removes the three reference system sets from `PostUpdate` and adds replacements with the same query
//! signatures that call the C ABI of `include/b200vis.h`:
//!   propagate  <- mark_dirty_trees / propagate_parent_transforms / sync_simple_transforms
//!                 (crates/bevy_transform/src/systems.rs:42, 111, 506; registered at plugins.rs:37-47)
//!   cull       <- check_visibility_cpu_culling (crates/bevy_camera/src/visibility/mod.rs:748)
//!   cluster    <- assign_objects_to_clusters   (crates/bevy_light/src/cluster/assign.rs:137, set
//!                 SimulationLightSystems::AssignLightsToClusters, crates/bevy_light/src/lib.rs:187-191)
//! `reset_view_visibility` and `mark_newly_hidden_entities_invisible` are private and share their sets with
//! systems that must stay, so they remain on the CPU; the GPU stage's "visible in >= 1 view" bit is turned
//! into `set_visible()` calls.
#![allow(non_camel_case_types)]
use bevy::prelude::*;
use bevy::camera::primitives::{Aabb, Frustum, Sphere};
use bevy::camera::visibility::*;
use bevy::light::{cluster::*, PointLight, SimulationLightSystems};
use bevy::transform::{systems::*, TransformSystems};
use core::ffi::{c_char, c_void};

// ---- FFI (mirrors include/b200vis.h) -------------------------------------------------------------------
#[repr(C)] pub struct b200vis_ctx { _p: [u8; 0] }
#[repr(C)] pub struct b200vis_config { device: i32, max_entities: u32, max_lights: u32, max_views: u32,
                                        max_cluster_indices: u32, world_size: u32, rank: u32, reserved: u32 }
#[repr(C)] pub struct b200vis_view { half_spaces: [[f32; 4]; 6], layer_mask: u64, flags: u8, range_view_index: i8, pad: [u8; 6] }
#[repr(C)] pub struct b200vis_frame_stats { visible_count: [u32; 8], cluster_index_count: [u32; 8], cluster_farthest_z: [f32; 8],
                                             cluster_index_overflow: [u32; 8], gt_changed_count: u32, vv_changed_count: u32, frame: u32, pad: u32 }
#[repr(C)] pub struct b200vis_cluster_view { enabled: u32, dims: [u32; 3], tile_size: [u32; 2], is_orthographic: u32, near_z: f32, far_z: f32,
    cluster_factors: [f32; 2], view_from_world: [f32; 16], clip_from_view: [f32; 16], view_from_world_scale: [f32; 3],
    view_from_world_scale_max: f32, frustum: [[f32; 4]; 6], layer_mask: u64, x_planes: *const f32, y_planes: *const f32, z_planes: *const f32 }

#[link(name = "b200vis")]
extern "C" {
    fn b200vis_create(cfg: *const b200vis_config, out: *mut *mut b200vis_ctx) -> i32;
    fn b200vis_destroy(ctx: *mut b200vis_ctx);
    fn b200vis_last_error(ctx: *const b200vis_ctx) -> *const c_char;
    fn b200vis_set_topology(ctx: *mut b200vis_ctx, n: u32, parent_row: *const u32, entity_bits: *const u64) -> i32;
    fn b200vis_plan_row_order(n: u32, parent_row: *const u32, new_to_old: *mut u32) -> i32;
    fn b200vis_upload_transforms_scattered(ctx: *mut b200vis_ctx, count: u32, rows: *const u32, trs: *const f32) -> i32;
    fn b200vis_mark_transforms_changed(ctx: *mut b200vis_ctx, first: u32, count: u32) -> i32;
    fn b200vis_upload_global_transforms(ctx: *mut b200vis_ctx, first: u32, count: u32, gt: *const f32) -> i32;
    fn b200vis_upload_bounds(ctx: *mut b200vis_ctx, first: u32, count: u32, bounds: *const f32, flags: *const u8,
                             class_mask: *const u8, layer_mask: *const u64, range_mask: *const u32) -> i32;
    fn b200vis_upload_view_visibility(ctx: *mut b200vis_ctx, first: u32, count: u32, vv: *const u8) -> i32;
    fn b200vis_set_static_transform_optimizations(ctx: *mut b200vis_ctx, enabled: i32) -> i32;
    fn b200vis_set_views(ctx: *mut b200vis_ctx, n: u32, views: *const b200vis_view) -> i32;
    fn b200vis_set_lights(ctx: *mut b200vis_ctx, n: u32, light_row: *const u32, range: *const f32, layers: *const u64) -> i32;
    fn b200vis_set_cluster_view(ctx: *mut b200vis_ctx, view: u32, p: *const b200vis_cluster_view) -> i32;
    fn b200vis_run(ctx: *mut b200vis_ctx, stages: u32) -> i32;
    fn b200vis_download_frame(ctx: *mut b200vis_ctx, stats: *mut b200vis_frame_stats, visible_rows: *mut u32, visible_cap: u32,
                              cluster_offsets: *mut u32, cluster_indices: *mut u32, cluster_cap: u32) -> i32;
    // the systems either side of the path
    fn b200vis_upload_visibility(ctx: *mut b200vis_ctx, first: u32, count: u32, visibility: *const u8) -> i32;
    fn b200vis_propagate_visibility(ctx: *mut b200vis_ctx) -> i32;
    fn b200vis_download_inherited_visibility(ctx: *mut b200vis_ctx, first: u32, count: u32, inherited: *mut u8, changed: *mut u8) -> i32;
    fn b200vis_upload_visibility_ranges(ctx: *mut b200vis_ctx, first: u32, count: u32, start_end: *const f32, use_aabb: *const u8) -> i32;
    fn b200vis_set_visibility_range_views(ctx: *mut b200vis_ctx, n_views: u32, positions: *const f32) -> i32;
    fn b200vis_download_visibility_ranges(ctx: *mut b200vis_ctx, first: u32, count: u32, mask: *mut u32) -> i32;
    fn b200vis_upload_shadow_casters(ctx: *mut b200vis_ctx, first: u32, count: u32, caster: *const u8) -> i32;
    fn b200vis_set_shadow_lights(ctx: *mut b200vis_ctx, n: u32, light_ordinals: *const u32, frusta: *const f32, layers: *const u64,
                                 lod_origin_range_index: i32, list_capacity: u32) -> i32;
    fn b200vis_run_shadow_culling(ctx: *mut b200vis_ctx) -> i32;
    fn b200vis_download_shadow_visible(ctx: *mut b200vis_ctx, shadow_light: u32, face: u32, rows: *mut u32, cap: u32, count: *mut u32) -> i32;
    fn b200vis_enable_visible_diff(ctx: *mut b200vis_ctx, enabled: i32) -> i32;
    fn b200vis_download_visible_diff(ctx: *mut b200vis_ctx, view: u32, added: *mut u32, added_cap: u32, n_added: *mut u32,
                                     removed: *mut u32, removed_cap: u32, n_removed: *mut u32) -> i32;
    fn b200vis_set_cluster_bindings(ctx: *mut b200vis_ctx, mode: u32, gpu_index_of_light: *const u32, n_map: u32) -> i32;
    fn b200vis_download_cluster_bindings(ctx: *mut b200vis_ctx, view: u32, offsets_and_counts: *mut u32, oc_cap: u32,
                                         index_lists: *mut u32, il_cap: u32, n_offsets: *mut u32, n_indices: *mut u32) -> i32;
    fn b200vis_download_global_transforms(ctx: *mut b200vis_ctx, first: u32, count: u32, gt: *mut f32, stride: u32, changed: *mut u8) -> i32;
    fn b200vis_download_view_visibility(ctx: *mut b200vis_ctx, first: u32, count: u32, vv: *mut u8, changed: *mut u8) -> i32;
}
const STAGE_PROPAGATE: u32 = 1; const STAGE_CULL: u32 = 2; const STAGE_CLUSTER: u32 = 12;
const ERR_HIERARCHY_CYCLE: i32 = 4;

/// Device context + the entity <-> row map the mirror needs.  `Send + Sync`: one system touches it at a time.
#[derive(Resource)]
pub struct B200Vis { ctx: *mut b200vis_ctx, row_of: bevy::ecs::entity::EntityHashMap<u32>, entity_of: Vec<Entity>, topology_dirty: bool }
unsafe impl Send for B200Vis {}
unsafe impl Sync for B200Vis {}
impl Drop for B200Vis { fn drop(&mut self) { unsafe { b200vis_destroy(self.ctx) } } }

fn check(vis: &B200Vis, rc: i32) -> Result<(), BevyError> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(b200vis_last_error(vis.ctx)) }.to_string_lossy().into_owned();
    // crates/bevy_transform/src/systems.rs:715 panics on a malformed hierarchy; keep that behaviour
    if rc == ERR_HIERARCHY_CYCLE { panic!("Malformed hierarchy: {msg}"); }
    Err(format!("b200vis error {rc}: {msg}").into())
}

pub struct B200VisibilityPlugin { pub max_entities: u32, pub max_lights: u32 }

impl Plugin for B200VisibilityPlugin {
    fn build(&self, app: &mut App) {
        let cfg = b200vis_config { device: 0, max_entities: self.max_entities, max_lights: self.max_lights, max_views: 8,
                                   max_cluster_indices: 0, world_size: 1, rank: 0, reserved: 0 };
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { b200vis_create(&cfg, &mut ctx) };
        assert_eq!(rc, 0, "b200vis_create failed: there is no CPU fallback");
        app.insert_resource(B200Vis { ctx, row_of: Default::default(), entity_of: Vec::new(), topology_dirty: true });
        // CPU clustering mode, so `Clusters` holds `ClusterableObjects::Cpu` that we fill
        app.insert_resource(GlobalClusterSettings { gpu_clustering: None, supports_storage_buffers: true,
            clustered_decals_are_usable: false, max_uniform_buffer_clusterable_objects: 204, view_cluster_bindings_max_indices: 16384 });
    }
    fn finish(&self, app: &mut App) {
        use bevy::ecs::schedule::ScheduleCleanupPolicy::RemoveSystemsOnly;
        for schedule in [PostStartup.intern(), PostUpdate.intern()] {
            app.remove_systems_in_set(schedule, mark_dirty_trees, RemoveSystemsOnly);
            app.remove_systems_in_set(schedule, propagate_parent_transforms, RemoveSystemsOnly);
            app.remove_systems_in_set(schedule, sync_simple_transforms, RemoveSystemsOnly);
        }
        app.remove_systems_in_set(PostUpdate, check_visibility_cpu_culling, RemoveSystemsOnly);
        app.remove_systems_in_set(PostUpdate, SimulationLightSystems::AssignLightsToClusters, RemoveSystemsOnly);
        app.add_systems(PostUpdate, (
            b200_propagate.in_set(TransformSystems::Propagate),
            b200_check_visibility.in_set(VisibilitySystems::CheckVisibility),
            b200_assign_lights_to_clusters.in_set(SimulationLightSystems::AssignLightsToClusters)
                .after(TransformSystems::Propagate).after(VisibilitySystems::CheckVisibility),
        ));
    }
}

/// propagate: same data as propagate_parent_transforms' queries (systems.rs:506-520) and
/// sync_simple_transforms' (systems.rs:42-55).
fn b200_propagate(
    mut vis: ResMut<B200Vis>,
    mut q: Query<(Entity, Ref<Transform>, &mut GlobalTransform, Option<&Children>, Option<Ref<ChildOf>>)>,
    mut orphaned: RemovedComponents<ChildOf>,
    opts: Res<StaticTransformOptimizations>,
) -> Result<(), BevyError> {
    // (1) hierarchy changed (spawn/despawn/Changed<ChildOf>): rebuild rows in b200vis_plan_row_order order,
    //     b200vis_set_topology(parent_row, Entity::to_bits()), re-upload every column.
    // (2) steady state: collect rows with Changed<Transform> -> b200vis_upload_transforms_scattered;
    //     Changed<ChildOf> | Added<GlobalTransform> | orphaned -> b200vis_mark_transforms_changed.
    // (3) b200vis_run(STAGE_PROPAGATE); b200vis_download_global_transforms(.., changed) and for every row whose
    //     changed byte is 1: *global_transform = new value (Mut deref stamps the change tick); all other rows
    //     are written through bypass_change_detection() only if their bits differ (they do not).
    let _ = (&mut q, &mut orphaned);
    unsafe { check(&vis, b200vis_set_static_transform_optimizations(vis.ctx, opts.is_enabled() as i32))?; }
    unsafe { check(&vis, b200vis_run(vis.ctx, STAGE_PROPAGATE))?; }
    vis.topology_dirty = false;
    Ok(())
}

/// cull: same parameter list as check_visibility_cpu_culling (visibility/mod.rs:748-774).
fn b200_check_visibility(
    vis: Res<B200Vis>,
    mut view_query: Query<(Entity, &mut VisibleEntities, &Frustum, Option<&RenderLayers>, &Camera, Has<NoCpuCulling>)>,
    mut visible_aabb_query: Query<(Entity, &InheritedVisibility, &mut ViewVisibility, Option<&VisibilityClass>, Option<&RenderLayers>,
                                   Option<&Aabb>, Option<&Sphere>, &GlobalTransform, Has<NoFrustumCulling>, Has<VisibilityRange>),
                                  Without<NoCpuCulling>>,
    visible_entity_ranges: Option<Res<VisibleEntityRanges>>,
) -> Result<(), BevyError> {
    // views: one b200vis_view per camera straight from `Frustum.half_spaces` (bit-identical by construction),
    // layer mask = RenderLayers::bits()[0], flags = is_active | NoCpuCulling.
    // rows: bounds/flags/class/layers/range columns are uploaded when they change (Changed<Aabb> etc.).
    // b200vis_run(STAGE_CULL); b200vis_download_frame(): for every view, VisibleEntities.clear_all() then push
    // entity_of[row] for each returned row into the class lists (already sorted by Entity::to_bits(), so the
    // reference's sort_unstable (visibility/mod.rs:870-874) is a no-op); for every row with bit0 set call
    // view_visibility.set_visible() (the bracket systems stay on the CPU and fire Changed<ViewVisibility>).
    let _ = (&mut view_query, &mut visible_aabb_query, &visible_entity_ranges);
    unsafe { check(&vis, b200vis_run(vis.ctx, STAGE_CULL)) }
}

/// cluster: same queries as assign_objects_to_clusters for point lights (assign.rs:137-153).
fn b200_assign_lights_to_clusters(
    vis: Res<B200Vis>,
    mut views: Query<(&GlobalTransform, &Camera, &Frustum, Option<&ClusterConfig>, &mut Clusters, Option<&RenderLayers>)>,
    point_lights_query: Query<(Entity, &GlobalTransform, &ViewVisibility, &PointLight, Option<&RenderLayers>)>,
    settings: Option<Res<GlobalClusterSettings>>,
) -> Result<(), BevyError> {
    // per view: restate the prologue of assign.rs:324-485 with glam (dims via ClusterConfig, Clusters::update,
    // far_z / cluster_factors from clusters.last_frame_*, x/y/z HalfSpace tables) -> b200vis_set_cluster_view.
    // lights: b200vis_set_lights(row_of[entity], range, layers) in query order.
    // b200vis_run(STAGE_CLUSTER); b200vis_download_frame(): clusters.clusterable_objects =
    // Cpu(vec![ObjectsInClusterCpu; n]) filled with add_point_light(light_entity[idx]) in list order;
    // clusters.last_frame_total_cluster_index_count / last_frame_farthest_z from the stats block.
    let _ = (&mut views, &point_lights_query, &settings);
    unsafe { check(&vis, b200vis_run(vis.ctx, STAGE_CLUSTER)) }
}
