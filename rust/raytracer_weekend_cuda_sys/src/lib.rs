//! `extern "C"` surface of include/rtw_cuda.h (RTW_ABI_VERSION 3), declaration for declaration.
//! Every function returns an int: >= 0 on success (ids / counts), < 0 = RTW_ERR_*; the message of the calling
//! thread's last error is `rtw_last_error()`.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const RTW_ABI_VERSION: c_int = 3;
pub const RTW_OK: c_int = 0;
pub const RTW_ERR_INVALID: c_int = -1;
pub const RTW_ERR_CUDA: c_int = -2;
pub const RTW_ERR_NOMEM: c_int = -3;
pub const RTW_ERR_UNSUPPORTED: c_int = -4;
pub const RTW_ERR_STATE: c_int = -5;
pub const RTW_RENDER_COUNT_TRAVERSAL: u32 = 1;
pub const RTW_RENDER_TIME_KERNELS: u32 = 2;
pub const RTW_TRACE_BVH: c_int = 0;
pub const RTW_TRACE_BRUTE: c_int = 1;

/// Opaque scene handle: owns all device memory; single owner (`!Sync`).
#[repr(C)]
pub struct rtw_scene {
    _private: [u8; 0],
}

/// ray.rs:6-10 plus the (t_min, t_max) window of `Hittable::hit` (hittable/mod.rs:52). 36 bytes.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_ray {
    pub origin: [f32; 3],
    pub direction: [f32; 3],
    pub time: f32,
    pub t_min: f32,
    pub t_max: f32,
}

/// HitRecord (hittable/mod.rs:22-29) + canonical primitive id + material id. `prim_id < 0` = miss.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_hit {
    pub prim_id: i32,
    pub material_id: i32,
    pub t: f32,
    pub p: [f32; 3],
    pub normal: [f32; 3],
    pub u: f32,
    pub v: f32,
    pub front_face: i32,
}

/// The fields of camera.rs:8-19 after `Camera::new`.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_camera {
    pub origin: [f32; 3],
    pub lower_left_corner: [f32; 3],
    pub horizontal: [f32; 3],
    pub vertical: [f32; 3],
    pub u: [f32; 3],
    pub v: [f32; 3],
    pub w: [f32; 3],
    pub lens_radius: f32,
    pub time0: f32,
    pub time1: f32,
}

/// `Raytracer::new` arguments (lib.rs:41-48) + MAX_DEPTH + the stream seed + the multi-GPU partition.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_render_params {
    pub width: u32,
    pub height: u32,
    pub spp: u32,
    pub max_depth: u32,
    pub background: [f32; 3],
    pub sample_begin: u32,
    pub sample_end: u32,
    pub seed: u64,
    pub tile_size: u32,
    pub part_rank: u32,
    pub part_count: u32,
    pub pool_size: u32,
    pub slices: u32,
    pub flags: u32,
    /// rtw_render / rtw_render_frames: spread the frame over this many devices (0 or 1 = the scene's device)
    pub gpus: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_render_stats {
    pub segments: u64,
    pub paths: u64,
    pub node_visits: u64,
    pub prim_tests: u64,
    pub prim_bytes: u64,
    pub iterations: u32,
    pub launches: u32,
    pub pool_size: u32,
    pub slices: u32,
    pub ms_render: f32,
    pub ms_traverse: f32,
    pub ms_shade: f32,
    pub node_record_bytes: f32,
    /// 1: one-leaf scene rendered by the fused persistent kernel
    pub fused: u32,
    /// devices that rendered the frame
    pub gpus: u32,
    /// summed CUDA-event time of the ray-reordering passes (if timed)
    pub ms_sort: f32,
    /// 1: the rays of every iteration were traced in scene-cell order
    pub ray_sort: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_build_stats {
    pub num_prims: u32,
    pub num_nodes: u32,
    pub max_depth: u32,
    pub num_instances: u32,
    pub ms_build: f32,
    pub ms_upload: f32,
    pub device_bytes: u64,
}

/// 32-byte BVH child record; two of them form the 64-byte pair one traversal step fetches.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtw_bvh_node {
    pub bmin: [f32; 3],
    pub link: i32,
    pub bmax: [f32; 3],
    pub meta: u32,
}

/// Called on a helper thread, one frame at a time, in frame order; non-zero return stops the animation.
pub type rtw_frame_callback =
    Option<unsafe extern "C" fn(user: *mut c_void, frame: u32, accum_rgb: *const f32, stats: *const rtw_render_stats) -> c_int>;

extern "C" {
    // ---- life cycle
    pub fn rtw_abi_version() -> c_int;
    pub fn rtw_last_error() -> *const c_char;
    pub fn rtw_device_count() -> c_int;
    pub fn rtw_scene_create(device: c_int, out: *mut *mut rtw_scene) -> c_int;
    pub fn rtw_scene_destroy(s: *mut rtw_scene) -> c_int;
    pub fn rtw_scene_clone(src: *const rtw_scene, device: c_int, out: *mut *mut rtw_scene) -> c_int;
    pub fn rtw_debug_live_handles() -> c_int;
    pub fn rtw_trim_memory() -> c_int;
    // ---- textures (texture.rs, image_texture.rs)
    pub fn rtw_add_texture_solid(s: *mut rtw_scene, r: f32, g: f32, b: f32) -> c_int;
    pub fn rtw_add_texture_checker(s: *mut rtw_scene, odd: c_int, even: c_int, frequency: f32) -> c_int;
    pub fn rtw_add_texture_noise(s: *mut rtw_scene, gradients_256x3: *const f32, perm_x_256: *const i32,
                                 perm_y_256: *const i32, perm_z_256: *const i32, scale: f32) -> c_int;
    pub fn rtw_add_texture_uvdebug(s: *mut rtw_scene) -> c_int;
    pub fn rtw_add_texture_image(s: *mut rtw_scene, rgb8: *const u8, width: u32, height: u32) -> c_int;
    // ---- materials (material.rs, light_source.rs)
    pub fn rtw_add_material_lambertian(s: *mut rtw_scene, albedo_texture: c_int) -> c_int;
    pub fn rtw_add_material_metal(s: *mut rtw_scene, r: f32, g: f32, b: f32, fuzz: f32) -> c_int;
    pub fn rtw_add_material_dielectric(s: *mut rtw_scene, ir: f32) -> c_int;
    pub fn rtw_add_material_diffuse_light(s: *mut rtw_scene, emit_texture: c_int) -> c_int;
    // ---- instance wrappers (hittable/transformations.rs), groups (bvh.rs), media (hittable/volumes.rs)
    pub fn rtw_push_translation(s: *mut rtw_scene, offset: *const f32) -> c_int;
    pub fn rtw_push_rotation_y(s: *mut rtw_scene, angle_degrees: f32) -> c_int;
    pub fn rtw_push_rotation_y_sincos(s: *mut rtw_scene, sin_theta: f32, cos_theta: f32) -> c_int;
    pub fn rtw_pop_transform(s: *mut rtw_scene) -> c_int;
    pub fn rtw_begin_group(s: *mut rtw_scene) -> c_int;
    pub fn rtw_end_group(s: *mut rtw_scene) -> c_int;
    pub fn rtw_begin_medium(s: *mut rtw_scene, density: f32, texture: c_int) -> c_int;
    pub fn rtw_end_medium(s: *mut rtw_scene) -> c_int;
    // ---- primitives: return the canonical id of the first primitive they emit
    pub fn rtw_add_sphere(s: *mut rtw_scene, center: *const f32, radius: f32, material: c_int) -> c_int;
    pub fn rtw_add_moving_sphere(s: *mut rtw_scene, center0: *const f32, time0: f32, center1: *const f32, time1: f32,
                                 radius: f32, material: c_int) -> c_int;
    pub fn rtw_add_xy_rect(s: *mut rtw_scene, x0: f32, x1: f32, y0: f32, y1: f32, k: f32, material: c_int) -> c_int;
    pub fn rtw_add_xz_rect(s: *mut rtw_scene, x0: f32, x1: f32, z0: f32, z1: f32, k: f32, material: c_int) -> c_int;
    pub fn rtw_add_yz_rect(s: *mut rtw_scene, y0: f32, y1: f32, z0: f32, z1: f32, k: f32, material: c_int) -> c_int;
    pub fn rtw_add_cuboid(s: *mut rtw_scene, p0: *const f32, p1: *const f32, material: c_int) -> c_int;
    pub fn rtw_add_triangles(s: *mut rtw_scene, n: u32, vertices: *const f32, normals: *const f32, uvs: *const f32,
                             material_ids: *const i32, material: c_int) -> c_int;
    // ---- build + introspection
    pub fn rtw_build(s: *mut rtw_scene, time0: f32, time1: f32, stats: *mut rtw_build_stats) -> c_int;
    pub fn rtw_scene_num_prims(s: *const rtw_scene) -> c_int;
    pub fn rtw_scene_num_nodes(s: *const rtw_scene) -> c_int;
    pub fn rtw_scene_num_instances(s: *const rtw_scene) -> c_int;
    pub fn rtw_scene_prim_info(s: *const rtw_scene, prim_id: c_int, ty: *mut i32, instance: *mut i32, material: *mut i32) -> c_int;
    pub fn rtw_scene_instance_ops(s: *const rtw_scene, instance: c_int, max_ops: c_int, kinds: *mut i32, abc: *mut f32) -> c_int;
    pub fn rtw_get_bvh(s: *const rtw_scene, nodes: *mut rtw_bvh_node, slot_prim_ids: *mut i32, root_box: *mut f32) -> c_int;
    // ---- closest hit of a ray batch (the parity entry point)
    pub fn rtw_trace_closest(s: *mut rtw_scene, rays: *const rtw_ray, n: u64, hits: *mut rtw_hit, mode: c_int) -> c_int;
    pub fn rtw_trace_closest_device(s: *mut rtw_scene, d_rays: *const rtw_ray, n: u64, d_hits: *mut rtw_hit, mode: c_int,
                                    stream: *mut c_void) -> c_int;
    // ---- Raytracer::render (lib.rs:57-76)
    pub fn rtw_render(s: *mut rtw_scene, cam: *const rtw_camera, params: *const rtw_render_params, accum_rgb: *mut f32,
                      stats: *mut rtw_render_stats) -> c_int;
    pub fn rtw_render_device(s: *mut rtw_scene, cam: *const rtw_camera, params: *const rtw_render_params,
                             d_accum_rgb: *mut f32, stream: *mut c_void, stats: *mut rtw_render_stats) -> c_int;
    pub fn rtw_render_frames(s: *mut rtw_scene, cameras: *const rtw_camera, n_frames: u32, params: *const rtw_render_params,
                             on_frame: rtw_frame_callback, user: *mut c_void) -> c_int;
    pub fn rtw_resolve_rgb8(s: *mut rtw_scene, accum_rgb: *const f32, width: u32, height: u32, spp: u32, rgb8: *mut u8) -> c_int;
}
