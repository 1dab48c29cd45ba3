//! Safe wrapper over the C ABI (include/rtw_cuda.h): `CudaScene` owns the opaque handle, implements the lib's
//! `SceneSink` (patches/flatten.rs) so that `world.flatten(&mut scene)` emits the scene in canonical order, and
//! renders frames into the same `Pixel` sequence `Raytracer::render` yields (lib.rs:57-76).
//!
//! NOT COMPILED in the repository that ships it (no Rust toolchain there); the executable specification of this
//! file is the C++ mirror raytracer-weekend_b200/host/rtw_host.cpp (`Flattener`, `Raytracer::render`,
//! `render_animation`), which the test suite runs against the same ABI.
use std::ffi::{c_int, c_void, CStr};
use std::fmt;

use raytracer_weekend_cuda_sys as sys;
use raytracer_weekend_lib::{
    camera::Camera,
    flatten::{FlattenError, SceneSink},
    vec3::{Color, Point3, Vec3},
    Pixel,
};

#[derive(Debug)]
pub struct CudaError {
    pub code: i32,
    pub message: String,
}

impl fmt::Display for CudaError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "rtw_cuda error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for CudaError {}

fn check(rc: c_int) -> Result<i32, CudaError> {
    if rc >= 0 {
        Ok(rc)
    } else {
        let message = unsafe { CStr::from_ptr(sys::rtw_last_error()) }.to_string_lossy().into_owned();
        Err(CudaError { code: rc, message })
    }
}

/// Owns an `rtw_scene*`. `!Sync` (raw pointer): the handle is single-owner, calls block (SURVEY §8b "Threading").
pub struct CudaScene {
    raw: *mut sys::rtw_scene,
    /// first error of a `SceneSink` call (the trait's emit methods return ids, not Results)
    sink_error: Option<CudaError>,
}

unsafe impl Send for CudaScene {}

impl CudaScene {
    pub fn new(device: i32) -> Result<Self, CudaError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::rtw_scene_create(device, &mut raw) })?;
        Ok(Self { raw, sink_error: None })
    }

    fn id(&mut self, rc: c_int) -> i32 {
        match check(rc) {
            Ok(v) => v,
            Err(e) => {
                self.sink_error.get_or_insert(e);
                -1
            }
        }
    }

    /// Upload + LBVH build on the GPU. `[time0, time1]` = the interval moving primitives' boxes must cover
    /// (bvh.rs:22-23; every scene of the reference passes 0, 1).
    pub fn build(&mut self, time0: f32, time1: f32) -> Result<sys::rtw_build_stats, CudaError> {
        if let Some(e) = self.sink_error.take() {
            return Err(e);
        }
        let mut st = sys::rtw_build_stats::default();
        check(unsafe { sys::rtw_build(self.raw, time0, time1, &mut st) })?;
        Ok(st)
    }

    /// `Raytracer::render` (lib.rs:57-76): the un-normalised per-pixel sums in the reference's yield order
    /// `(0..h).rev() x (0..w)` (lib.rs:58).
    pub fn render(&mut self, cam: &Camera, params: &sys::rtw_render_params) -> Result<(Vec<Pixel>, sys::rtw_render_stats), CudaError> {
        let (w, h) = (params.width as usize, params.height as usize);
        let mut accum = vec![0f32; w * h * 3];
        let mut st = sys::rtw_render_stats::default();
        check(unsafe { sys::rtw_render(self.raw, &camera_to_ffi(cam), params, accum.as_mut_ptr(), &mut st) })?;
        Ok((pixels_from_accum(&accum, params.width, params.height), st))
    }

    /// main.rs:48-95 over a resident scene: frame i = `cams[i]`, stream seed `params.seed + i`; `on_frame` runs on a
    /// helper thread (hence `Send`) while the next frame renders; return `false` to stop.
    pub fn render_frames<F>(&mut self, cams: &[Camera], params: &sys::rtw_render_params, mut on_frame: F) -> Result<u32, CudaError>
    where
        F: FnMut(u32, Vec<Pixel>, &sys::rtw_render_stats) -> bool + Send,
    {
        struct Ctx<'a> {
            f: &'a mut (dyn FnMut(u32, Vec<Pixel>, &sys::rtw_render_stats) -> bool + Send),
            w: u32,
            h: u32,
        }
        unsafe extern "C" fn tramp(user: *mut c_void, frame: u32, accum: *const f32, st: *const sys::rtw_render_stats) -> c_int {
            let ctx = &mut *(user as *mut Ctx);
            let n = ctx.w as usize * ctx.h as usize * 3;
            let px = pixels_from_accum(std::slice::from_raw_parts(accum, n), ctx.w, ctx.h);
            // a panic must not unwind into C
            match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| (ctx.f)(frame, px, &*st))) {
                Ok(true) => 0,
                _ => 1,
            }
        }
        let ffi: Vec<sys::rtw_camera> = cams.iter().map(camera_to_ffi).collect();
        let mut ctx = Ctx { f: &mut on_frame, w: params.width, h: params.height };
        let n = check(unsafe {
            sys::rtw_render_frames(self.raw, ffi.as_ptr(), ffi.len() as u32, params, Some(tramp), &mut ctx as *mut Ctx as *mut c_void)
        })?;
        Ok(n as u32)
    }
}

impl Drop for CudaScene {
    fn drop(&mut self) {
        unsafe { sys::rtw_scene_destroy(self.raw) };
    }
}

/// `Raytracer::new(world, cam, background, w, h, spp)` argument order (lib.rs:41-48) -> `rtw_render_params`
pub fn render_params(background: Color, image_width: u32, image_height: u32, samples_per_pixel: u32, seed: u64) -> sys::rtw_render_params {
    sys::rtw_render_params {
        width: image_width,
        height: image_height,
        spp: samples_per_pixel,
        max_depth: 50, // MAX_DEPTH, lib.rs:32
        background: [background.x(), background.y(), background.z()],
        seed,
        ..Default::default()
    }
}

/// accumulation buffer -> the `Pixel`s of lib.rs:120-126 in the order of lib.rs:58
pub fn pixels_from_accum(accum: &[f32], w: u32, h: u32) -> Vec<Pixel> {
    let mut out = Vec::with_capacity(w as usize * h as usize);
    let mut it = accum.chunks_exact(3);
    for row in (0..h).rev() {
        for column in 0..w {
            let c = it.next().expect("accumulation buffer too short");
            out.push(Pixel { row, column, color: Color::new(c[0], c[1], c[2]) });
        }
    }
    out
}

/// camera.rs:8-19 -> rtw_camera. Needs `pub(crate)`-style accessors on `Camera` (patches/flatten.rs adds `Camera::raw()`).
fn camera_to_ffi(cam: &Camera) -> sys::rtw_camera {
    let r = cam.raw();
    let a = |v: Vec3| [v.x(), v.y(), v.z()];
    sys::rtw_camera {
        origin: a(r.origin),
        lower_left_corner: a(r.lower_left_corner),
        horizontal: a(r.horizontal),
        vertical: a(r.vertical),
        u: a(r.u),
        v: a(r.v),
        w: a(r.w),
        lens_radius: r.lens_radius,
        time0: r.time0,
        time1: r.time1,
    }
}

fn p3(p: Point3) -> [f32; 3] {
    [p.x(), p.y(), p.z()]
}

/// One method per emit call of include/rtw_cuda.h; the call order defines the canonical primitive ids.
impl SceneSink for CudaScene {
    fn texture_solid(&mut self, c: Color) -> i32 {
        let rc = unsafe { sys::rtw_add_texture_solid(self.raw, c.x(), c.y(), c.z()) };
        self.id(rc)
    }
    fn texture_checker(&mut self, odd: i32, even: i32, frequency: f32) -> i32 {
        let rc = unsafe { sys::rtw_add_texture_checker(self.raw, odd, even, frequency) };
        self.id(rc)
    }
    fn texture_noise(&mut self, gradients: &[[f32; 3]; 256], perms: &[[i32; 256]; 3], scale: f32) -> i32 {
        let rc = unsafe {
            sys::rtw_add_texture_noise(self.raw, gradients.as_ptr() as *const f32, perms[0].as_ptr(), perms[1].as_ptr(), perms[2].as_ptr(), scale)
        };
        self.id(rc)
    }
    fn texture_uvdebug(&mut self) -> i32 {
        let rc = unsafe { sys::rtw_add_texture_uvdebug(self.raw) };
        self.id(rc)
    }
    fn texture_image(&mut self, rgb8: &[u8], width: u32, height: u32) -> i32 {
        assert_eq!(rgb8.len(), width as usize * height as usize * 3);
        let rc = unsafe { sys::rtw_add_texture_image(self.raw, rgb8.as_ptr(), width, height) };
        self.id(rc)
    }
    fn material_lambertian(&mut self, tex: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_material_lambertian(self.raw, tex) };
        self.id(rc)
    }
    fn material_metal(&mut self, albedo: Color, fuzz: f32) -> i32 {
        let rc = unsafe { sys::rtw_add_material_metal(self.raw, albedo.x(), albedo.y(), albedo.z(), fuzz) };
        self.id(rc)
    }
    fn material_dielectric(&mut self, ir: f32) -> i32 {
        let rc = unsafe { sys::rtw_add_material_dielectric(self.raw, ir) };
        self.id(rc)
    }
    fn material_diffuse_light(&mut self, tex: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_material_diffuse_light(self.raw, tex) };
        self.id(rc)
    }
    fn push_translation(&mut self, offset: Vec3) {
        let rc = unsafe { sys::rtw_push_translation(self.raw, p3(offset).as_ptr()) };
        self.id(rc);
    }
    fn push_rotation_y(&mut self, sin_theta: f32, cos_theta: f32) {
        let rc = unsafe { sys::rtw_push_rotation_y_sincos(self.raw, sin_theta, cos_theta) };
        self.id(rc);
    }
    fn pop_transform(&mut self) {
        let rc = unsafe { sys::rtw_pop_transform(self.raw) };
        self.id(rc);
    }
    fn begin_group(&mut self) {
        let rc = unsafe { sys::rtw_begin_group(self.raw) };
        self.id(rc);
    }
    fn end_group(&mut self) {
        let rc = unsafe { sys::rtw_end_group(self.raw) };
        self.id(rc);
    }
    fn begin_medium(&mut self, density: f32, tex: i32) {
        let rc = unsafe { sys::rtw_begin_medium(self.raw, density, tex) };
        self.id(rc);
    }
    fn end_medium(&mut self) -> i32 {
        let rc = unsafe { sys::rtw_end_medium(self.raw) };
        self.id(rc)
    }
    fn sphere(&mut self, center: Point3, radius: f32, material: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_sphere(self.raw, p3(center).as_ptr(), radius, material) };
        self.id(rc)
    }
    fn moving_sphere(&mut self, c0: Point3, t0: f32, c1: Point3, t1: f32, radius: f32, material: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_moving_sphere(self.raw, p3(c0).as_ptr(), t0, p3(c1).as_ptr(), t1, radius, material) };
        self.id(rc)
    }
    fn xy_rect(&mut self, x0: f32, x1: f32, y0: f32, y1: f32, k: f32, material: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_xy_rect(self.raw, x0, x1, y0, y1, k, material) };
        self.id(rc)
    }
    fn xz_rect(&mut self, x0: f32, x1: f32, z0: f32, z1: f32, k: f32, material: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_xz_rect(self.raw, x0, x1, z0, z1, k, material) };
        self.id(rc)
    }
    fn yz_rect(&mut self, y0: f32, y1: f32, z0: f32, z1: f32, k: f32, material: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_yz_rect(self.raw, y0, y1, z0, z1, k, material) };
        self.id(rc)
    }
    fn cuboid(&mut self, p0: Point3, p1: Point3, material: i32) -> i32 {
        let rc = unsafe { sys::rtw_add_cuboid(self.raw, p3(p0).as_ptr(), p3(p1).as_ptr(), material) };
        self.id(rc)
    }
    fn triangles(&mut self, vertices: &[f32], normals: Option<&[f32]>, uvs: Option<&[f32]>, material: i32) -> i32 {
        let n = (vertices.len() / 9) as u32;
        let rc = unsafe {
            sys::rtw_add_triangles(
                self.raw,
                n,
                vertices.as_ptr(),
                normals.map_or(std::ptr::null(), |x| x.as_ptr()),
                uvs.map_or(std::ptr::null(), |x| x.as_ptr()),
                std::ptr::null(),
                material,
            )
        };
        self.id(rc)
    }
}

/// `world.as_slice().flatten(&mut scene)` + `scene.build(0, 1)`; surfaces `FlattenError::Unsupported` of a type
/// that has no `flatten` impl.
pub fn upload_world(world: &[Box<dyn raytracer_weekend_lib::hittable::Hittable>], device: i32) -> Result<CudaScene, Box<dyn std::error::Error>> {
    let mut scene = CudaScene::new(device)?;
    for object in world {
        object.flatten(&mut scene).map_err(|e: FlattenError| Box::new(e) as Box<dyn std::error::Error>)?;
    }
    scene.build(0.0, 1.0)?;
    Ok(scene)
}
