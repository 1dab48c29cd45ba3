//! raytracer_weekend_lib/src/flatten.rs — NEW module (add `pub mod flatten;` to lib.rs).
//!
//! The scene graph of the lib is opaque from outside: struct fields are private, the traits cannot be downcast
//! and the `aabb` / `ray` modules are private (lib.rs:6,14).  A back end therefore cannot walk a
//! `Vec<Box<dyn Hittable>>`; every type EMITS ITSELF into a `SceneSink` instead.  The emit order is the
//! canonical primitive order of the back end (world-Vec order, depth first; `Cuboid` = its six sides in
//! rectangular.rs:177-234 order; a mesh = face order), which is what makes closest-hit ids comparable.
//!
//! Executable specification: `Hittable::flatten` / `Material::flatten` / `Texture::flatten` of the C++ mirror
//! (raytracer-weekend_b200/host/rtw_host.hpp), exercised by the test suite against the same C ABI.
//! NOT COMPILED in the repository that ships it (no Rust toolchain there).
use crate::{
    bvh::BvhNode,
    camera::Camera,
    hittable::{
        rectangular::{Cuboid, XYRectangle, XZRectangle, YZRectangle},
        spherical::{MovingSphere, Sphere},
        transformations::{Translation, YRotation},
        triangular::Triangle,
        volumes::ConstantMedium,
        Hittable,
    },
    image_texture::ImageTexture,
    light_source::DiffuseLight,
    material::{Dielectric, Isotropic, Lambertian, Material, Metal},
    perlin::Perlin,
    texture::{Checker, Noise, SolidColor, Texture, UVDebug},
    vec3::{Color, Point3, Vec3},
};

#[derive(Debug)]
pub enum FlattenError {
    /// the type has no `flatten` impl (default of the trait methods)
    Unsupported(&'static str),
}
impl core::fmt::Display for FlattenError {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        match self {
            FlattenError::Unsupported(what) => write!(f, "{what} cannot be flattened for the cuda backend"),
        }
    }
}
impl std::error::Error for FlattenError {}

/// One method per emit call of include/rtw_cuda.h.  Ids returned by texture_* / material_* are handles for later
/// calls; primitive calls return the canonical id of the first primitive they emit.
pub trait SceneSink {
    fn texture_solid(&mut self, c: Color) -> i32;
    fn texture_checker(&mut self, odd: i32, even: i32, frequency: f32) -> i32;
    fn texture_noise(&mut self, gradients: &[[f32; 3]; 256], perms: &[[i32; 256]; 3], scale: f32) -> i32;
    fn texture_uvdebug(&mut self) -> i32;
    fn texture_image(&mut self, rgb8: &[u8], width: u32, height: u32) -> i32;
    fn material_lambertian(&mut self, tex: i32) -> i32;
    fn material_metal(&mut self, albedo: Color, fuzz: f32) -> i32;
    fn material_dielectric(&mut self, ir: f32) -> i32;
    fn material_diffuse_light(&mut self, tex: i32) -> i32;
    fn push_translation(&mut self, offset: Vec3);
    fn push_rotation_y(&mut self, sin_theta: f32, cos_theta: f32);
    fn pop_transform(&mut self);
    fn begin_group(&mut self);
    fn end_group(&mut self);
    fn begin_medium(&mut self, density: f32, tex: i32);
    fn end_medium(&mut self) -> i32;
    fn sphere(&mut self, center: Point3, radius: f32, material: i32) -> i32;
    fn moving_sphere(&mut self, c0: Point3, t0: f32, c1: Point3, t1: f32, radius: f32, material: i32) -> i32;
    fn xy_rect(&mut self, x0: f32, x1: f32, y0: f32, y1: f32, k: f32, material: i32) -> i32;
    fn xz_rect(&mut self, x0: f32, x1: f32, z0: f32, z1: f32, k: f32, material: i32) -> i32;
    fn yz_rect(&mut self, y0: f32, y1: f32, z0: f32, z1: f32, k: f32, material: i32) -> i32;
    fn cuboid(&mut self, p0: Point3, p1: Point3, material: i32) -> i32;
    fn triangles(&mut self, vertices: &[f32], normals: Option<&[f32]>, uvs: Option<&[f32]>, material: i32) -> i32;
}

// ---- the three trait hooks (ADD to the trait definitions; default = unsupported) ---------------------------
//
//   hittable/mod.rs:51-54   pub trait Hittable: Sync + Send + Debug {
//                               fn hit(..) -> Option<HitRecord>;
//                               fn bounding_box(..) -> Option<Aabb>;
//   +                           fn flatten(&self, _sink: &mut dyn SceneSink) -> Result<(), FlattenError> {
//   +                               Err(FlattenError::Unsupported(core::any::type_name::<Self>()))
//   +                           }
//                           }
//   material.rs:23-26       pub trait Material { .. + fn flatten(&self, _: &mut dyn SceneSink) -> Result<i32, FlattenError> }
//   texture.rs:41-43        pub trait Texture  { .. + fn flatten(&self, _: &mut dyn SceneSink) -> Result<i32, FlattenError> }
//
// Below: the body of `flatten` for every concrete type, to be placed inside its existing `impl` block
// (the fields are private to those modules; shown here together for review).

// ---- textures (texture.rs, image_texture.rs) -------------------------------------------------------------
impl SolidColor {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        Ok(s.texture_solid(self.color_value)) // texture.rs:45-60
    }
}
impl<E: Texture, O: Texture> Checker<E, O> {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        // NOTE the constructor's argument order is new(odd, even, frequency) (texture.rs:62-68)
        let odd = self.odd.flatten(s)?;
        let even = self.even.flatten(s)?;
        Ok(s.texture_checker(odd, even, self.frequency))
    }
}
impl Noise {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        let (g, p) = self.noise.tables(); // perlin.rs:9-13, see `Perlin::tables` below
        Ok(s.texture_noise(&g, &p, self.scale))
    }
}
impl Perlin {
    /// 256 gradients + the three permutations as the back end wants them (perlin.rs:9-13)
    pub(crate) fn tables(&self) -> ([[f32; 3]; 256], [[i32; 256]; 3]) {
        let mut g = [[0f32; 3]; 256];
        let mut p = [[0i32; 256]; 3];
        for i in 0..256 {
            g[i] = [self.gradients[i].x(), self.gradients[i].y(), self.gradients[i].z()];
            for a in 0..3 {
                p[a][i] = self.permutations[a][i] as i32;
            }
        }
        (g, p)
    }
}
impl UVDebug {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        Ok(s.texture_uvdebug())
    }
}
impl ImageTexture {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        // the DECODED pixels (image_texture.rs:24): both back ends then read identical texels
        let rgb = self.image.to_rgb8();
        Ok(s.texture_image(rgb.as_raw(), rgb.width(), rgb.height()))
    }
}

// ---- materials (material.rs, light_source.rs) ------------------------------------------------------------
impl<T: Texture> Lambertian<T> {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        let t = self.albedo.flatten(s)?;
        Ok(s.material_lambertian(t))
    }
}
impl Metal {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        Ok(s.material_metal(self.albedo, self.fuzz))
    }
}
impl Dielectric {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        Ok(s.material_dielectric(self.ir))
    }
}
impl<T: Texture> DiffuseLight<T> {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<i32, FlattenError> {
        let t = self.emit.flatten(s)?;
        Ok(s.material_diffuse_light(t))
    }
}
// Isotropic is only ever the phase function of a ConstantMedium (volumes.rs:26-31): emitted by begin_medium.

// ---- primitives --------------------------------------------------------------------------------------------
impl Sphere {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        let m = self.material.flatten(s)?; // spherical.rs:80-84
        s.sphere(self.center, self.radius, m);
        Ok(())
    }
}
impl MovingSphere {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        let m = self.material.flatten(s)?; // spherical.rs:107-115
        s.moving_sphere(self.center0, self.time0, self.center1, self.time1, self.radius, m);
        Ok(())
    }
}
impl XYRectangle {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        let m = self.material.flatten(s)?;
        s.xy_rect(self.x0, self.x1, self.y0, self.y1, self.k, m);
        Ok(())
    }
}
impl XZRectangle {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        let m = self.material.flatten(s)?;
        s.xz_rect(self.x0, self.x1, self.z0, self.z1, self.k, m);
        Ok(())
    }
}
impl YZRectangle {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        let m = self.material.flatten(s)?;
        s.yz_rect(self.y0, self.y1, self.z0, self.z1, self.k, m);
        Ok(())
    }
}
impl Cuboid {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        // all six sides share clones of one material (rectangular.rs:177-234): emit it once.  `Cuboid` must keep
        // that material (add a field) or take it from side 0 through a crate-private accessor.
        let m = self.material().flatten(s)?;
        s.cuboid(self.box_min, self.box_max, m);
        Ok(())
    }
}
impl Triangle {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        let m = self.material.flatten(s)?; // one call per triangle; a mesh loader batches (see load_wavefront_obj)
        let mut v = [0f32; 9];
        let mut n = [0f32; 9];
        let mut uv = [0f32; 6];
        for i in 0..3 {
            v[3 * i..3 * i + 3].copy_from_slice(&[self.vertices[i].x(), self.vertices[i].y(), self.vertices[i].z()]);
            n[3 * i..3 * i + 3].copy_from_slice(&[self.normals[i].x(), self.normals[i].y(), self.normals[i].z()]);
            uv[2 * i..2 * i + 2].copy_from_slice(&[self.texture_uv[i].u, self.texture_uv[i].v]);
        }
        // normals / uvs are always materialised by Triangle::new (triangular.rs:53-65, incl. the un-normalised
        // face normal): pass them as they are stored
        s.triangles(&v, Some(&n), Some(&uv), m);
        Ok(())
    }
}

// ---- wrappers ----------------------------------------------------------------------------------------------
impl<T: Hittable> Translation<T> {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        s.push_translation(self.offset); // transformations.rs:16-20
        let r = self.inner.flatten(s);
        s.pop_transform();
        r
    }
}
impl<T: Hittable> YRotation<T> {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        s.push_rotation_y(self.sin_theta, self.cos_theta); // the stored values (transformations.rs:51-56)
        let r = self.inner.flatten(s);
        s.pop_transform();
        r
    }
}
impl BvhNode {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        // an acceleration hint only: the back end builds its own LBVH over every primitive.  Canonical ids are
        // then "tree order"; ids only matter for exact-t ties, where the reference's own winner already depends
        // on this (random) topology (bvh.rs:25,101-120).
        s.begin_group();
        let mut r = self.left.flatten(s);
        if let (Ok(()), Some(right)) = (&r, &self.right) {
            r = right.flatten(s);
        }
        s.end_group();
        r
    }
}
impl<H: Hittable, T: Texture> ConstantMedium<H, T> {
    pub(crate) fn flatten_impl(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        // volumes.rs:24-35: density = -1 / neg_inv_density; the boundary is emitted between begin / end
        let tex = self.phase_function.albedo().flatten(s)?;
        s.begin_medium(-1.0 / self.neg_inv_density, tex);
        let r = self.boundary.flatten(s);
        s.end_medium();
        r
    }
}
impl<T: Texture> Isotropic<T> {
    pub(crate) fn albedo(&self) -> &T {
        &self.albedo
    }
}
impl Hittable for [Box<dyn Hittable>] {
    // hittable/mod.rs:56-88 — ADD to the existing impl
    fn flatten(&self, s: &mut dyn SceneSink) -> Result<(), FlattenError> {
        for o in self {
            o.flatten(s)?;
        }
        Ok(())
    }
}

// ---- camera (camera.rs:8-19): the fields are private; the back end needs all of them -----------------------
pub struct CameraRaw {
    pub origin: Point3,
    pub lower_left_corner: Point3,
    pub horizontal: Vec3,
    pub vertical: Vec3,
    pub u: Vec3,
    pub v: Vec3,
    pub w: Vec3,
    pub lens_radius: f32,
    pub time0: f32,
    pub time1: f32,
}
impl Camera {
    pub fn raw(&self) -> CameraRaw {
        CameraRaw {
            origin: self.origin,
            lower_left_corner: self.lower_left_corner,
            horizontal: self.horizontal,
            vertical: self.vertical,
            u: self.u,
            v: self.v,
            w: self._w,
            lens_radius: self.lens_radius,
            time0: self.time0,
            time1: self.time1,
        }
    }
}
