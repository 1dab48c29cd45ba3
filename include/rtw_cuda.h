/*
 * rtw_cuda.h — C ABI of the B200-native path-tracing backend for raytracer_weekend_lib.
 *
 * This is the drop-in boundary: everything the reference's render hot path
 *   Raytracer::render -> sample_pixel -> sample_ray -> world.hit -> material.scatter
 *   (raytracer_weekend_lib/src/lib.rs:57-117)
 * needs, expressed as plain pointers and sizes so that a Rust `-sys` crate (or cgo / ctypes)
 * can bind it.  No torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every function returns an int: >= 0 on success (ids / counts), < 0 = RTW_ERR_* ;
 *     the message for the last error of the calling thread is rtw_last_error().
 *   - a scene is built by "emit" calls that mirror the reference constructors one to one; the
 *     order of the primitive emit calls defines the CANONICAL PRIMITIVE ID (world-Vec order,
 *     depth first; Cuboid = its 6 sides in rectangular.rs:177-234 order; mesh = face order).
 *   - all scalars are IEEE f32 (vec3.rs:354).
 *   - the library never falls back to the CPU: without a CUDA device every compute call fails
 *     with RTW_ERR_CUDA.
 */
#ifndef RTW_CUDA_H
#define RTW_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTW_ABI_VERSION 3 /* 2: rtw_render_params.gpus, rtw_render_stats.fused / gpus, rtw_scene_clone; 3: rtw_render_stats.ms_sort / ray_sort */

/* error codes */
#define RTW_OK 0
#define RTW_ERR_INVALID (-1)   /* bad argument / bad handle / wrong state      */
#define RTW_ERR_CUDA (-2)      /* CUDA runtime error or no device              */
#define RTW_ERR_NOMEM (-3)     /* host or device allocation failed             */
#define RTW_ERR_UNSUPPORTED (-4)
#define RTW_ERR_STATE (-5)     /* e.g. render before rtw_build                 */

typedef struct rtw_scene rtw_scene; /* opaque; owns all device memory; single-owner (!Sync) */

/* A ray batch element: ray.rs:6-10 plus the (t_min, t_max) window of Hittable::hit
 * (hittable/mod.rs:52).  36 bytes, tightly packed. */
typedef struct rtw_ray {
  float origin[3];
  float direction[3];
  float time;
  float t_min;
  float t_max;
} rtw_ray;

/* Closest-hit result: HitRecord (hittable/mod.rs:22-29) plus the canonical primitive id and
 * the material id that the reference keeps as a `&dyn Material`.  prim_id < 0 = miss (all the
 * other fields are then 0). */
typedef struct rtw_hit {
  int32_t prim_id;
  int32_t material_id;
  float t;
  float p[3];
  float normal[3];
  float u, v;
  int32_t front_face;
} rtw_hit;

/* The fields of camera.rs:8-19 after Camera::new (camera.rs:25-64). */
typedef struct rtw_camera {
  float origin[3];
  float lower_left_corner[3];
  float horizontal[3];
  float vertical[3];
  float u[3];
  float v[3];
  float w[3];
  float lens_radius;
  float time0;
  float time1;
} rtw_camera;

/* Raytracer::new arguments (lib.rs:41-48) + what the reference hard-codes (MAX_DEPTH lib.rs:32,
 * the RNG) + how the frame is partitioned over GPUs. */
typedef struct rtw_render_params {
  uint32_t width;         /* image_width                                                  */
  uint32_t height;        /* image_height                                                 */
  uint32_t spp;           /* samples_per_pixel of the whole frame                         */
  uint32_t max_depth;     /* MAX_DEPTH; 0 -> 50                                           */
  float background[3];
  uint32_t sample_begin;  /* this call renders samples [sample_begin, sample_end) of each */
  uint32_t sample_end;    /*   pixel; 0,0 -> [0, spp)                                      */
  uint64_t seed;          /* Philox seed; the stream is keyed by (seed, pixel, sample)     */
  uint32_t tile_size;     /* multi-GPU tile partition: tile k (row-major over the tile     */
  uint32_t part_rank;     /*   grid) is rendered iff k % part_count == part_rank;          */
  uint32_t part_count;    /*   part_count 0 or 1 -> every pixel. tile_size 0 -> 32         */
  uint32_t pool_size;     /* path slots in flight; 0 -> auto                              */
  uint32_t slices;        /* sample slices per pixel (summation tree); 0 -> auto          */
  uint32_t flags;         /* RTW_RENDER_*                                                 */
  uint32_t gpus;          /* rtw_render / rtw_render_frames: spread the frame over this many */
                          /*   devices of the box (scene's device first, then the following  */
                          /*   ones; replicas are made on first use); 0 or 1 -> one device.  */
                          /*   The frame has the same bits for every value.                  */
  uint32_t reserved;      /* 0                                                             */
} rtw_render_params;

#define RTW_RENDER_COUNT_TRAVERSAL 1u /* also count BVH node visits / primitive tests (slower) */
#define RTW_RENDER_TIME_KERNELS 2u    /* CUDA-event pair around every kernel -> ms_traverse / ms_shade */

typedef struct rtw_render_stats {
  uint64_t segments;      /* path segments = world.hit() queries (lib.rs:102)              */
  uint64_t paths;         /* camera paths started                                          */
  uint64_t node_visits;   /* child-pair fetches (only with COUNT_TRAVERSAL)                */
  uint64_t prim_tests;    /* primitive intersection tests (only with COUNT_TRAVERSAL)      */
  uint64_t prim_bytes;    /* bytes those tests fetched: 4 (slot meta) + geometry: sphere   */
                          /*   16, rect 32, moving sphere / triangle 48 (COUNT_TRAVERSAL)  */
  uint32_t iterations;    /* wavefront iterations (one traverse + one shade launch each)   */
  uint32_t launches;      /* kernels launched by this call                                 */
  uint32_t pool_size;     /* slots actually used                                           */
  uint32_t slices;        /* slices actually used                                          */
  float ms_render;        /* CUDA-event time of the render, first launch to last           */
  float ms_traverse;      /* summed CUDA-event time of the traversal kernel (if timed)     */
  float ms_shade;         /* summed CUDA-event time of the shade kernel (if timed)         */
  float node_record_bytes;/* bytes one node_visit fetches: 64 (fp32 pair) or 32 (compact    */
                          /*   pair, hierarchies beyond the caches); 0 = not applicable      */
  uint32_t fused;         /* 1: one-leaf scene rendered by the fused persistent kernel (no   */
                          /*   wavefront, one launch per frame)                              */
  uint32_t gpus;          /* devices that rendered this frame                               */
  float ms_sort;          /* summed CUDA-event time of the ray-reordering passes (if timed)  */
  uint32_t ray_sort;      /* 1: the rays of every iteration were traced in scene-cell order  */
                          /*   (hierarchies that do not fit the caches; same frame bits)      */
} rtw_render_stats;

typedef struct rtw_build_stats {
  uint32_t num_prims;
  uint32_t num_nodes;     /* internal nodes (= 64-byte child pairs)                        */
  uint32_t max_depth;     /* deepest leaf                                                  */
  uint32_t num_instances; /* transform chains (incl. the identity)                         */
  float ms_build;         /* bounds + morton + sort + hierarchy + refit on the GPU         */
  float ms_upload;
  uint64_t device_bytes;
} rtw_build_stats;

/* 32-byte BVH child record as it lies in HBM; two of them (left, right) form the 64-byte
 * "pair" that one traversal step fetches with 4 LDG.128.  link >= 0: index of the child's own
 * pair; link < 0: leaf = the contiguous primitive-slot range [~link, ~link + meta)
 * (meta = primitive count; the SAH collapse of the build decides how many). */
typedef struct rtw_bvh_node {
  float bmin[3];
  int32_t link;
  float bmax[3];
  uint32_t meta;
} rtw_bvh_node;

/* ---- life cycle ------------------------------------------------------------------------ */
int rtw_abi_version(void);
const char *rtw_last_error(void);
int rtw_device_count(void);              /* number of CUDA devices, < 0 on error             */
int rtw_scene_create(int device, rtw_scene **out);
int rtw_scene_destroy(rtw_scene *s);
/* A replica of a BUILT scene on another device: every device buffer (geometry, LBVH, materials, textures) is copied
 * device to device (NVLink when the devices are peers) — the scene is neither re-flattened nor rebuilt, so all
 * replicas hold bit-identical data.  SURVEY.md 8(e): "build on GPU 0 and broadcast the node / primitive arrays". */
int rtw_scene_clone(const rtw_scene *src, int device, rtw_scene **out);
/* events / streams / graphs this library holds right now (0 once every scene is destroyed): leak check for tests */
int rtw_debug_live_handles(void);
/* The library keeps freed device and pinned blocks for the next scene (a front end that builds a scene per frame would
 * otherwise pay tens to hundreds of ms of cudaMalloc / cudaFree per scene).  This gives every cached block back to the
 * driver; returns the number of bytes released as a count of MiB (>= 0).  RTW_MEM_CACHE=0 disables the cache. */
int rtw_trim_memory(void);

/* ---- textures: texture.rs, image_texture.rs ---------------------------------------------- */
int rtw_add_texture_solid(rtw_scene *s, float r, float g, float b);            /* texture.rs:45-60  */
int rtw_add_texture_checker(rtw_scene *s, int odd, int even, float frequency); /* texture.rs:62-81  */
/* Noise{Perlin, scale} (texture.rs:83-95). The tables are what Perlin::new (perlin.rs:15-29)
 * produced: 256 unit gradients (xyz) and the x/y/z permutations (values 0..255). */
int rtw_add_texture_noise(rtw_scene *s, const float *gradients_256x3, const int32_t *perm_x_256,
                          const int32_t *perm_y_256, const int32_t *perm_z_256, float scale);
int rtw_add_texture_uvdebug(rtw_scene *s);                                     /* texture.rs:97-104 */
/* ImageTexture (image_texture.rs:17-51): decoded, tightly packed RGB8, row 0 = top. */
int rtw_add_texture_image(rtw_scene *s, const uint8_t *rgb8, uint32_t width, uint32_t height);

/* ---- materials: material.rs, light_source.rs --------------------------------------------- */
int rtw_add_material_lambertian(rtw_scene *s, int albedo_texture);             /* material.rs:30-61  */
int rtw_add_material_metal(rtw_scene *s, float r, float g, float b, float fuzz); /* material.rs:63-100; fuzz > 1 -> error like the assert at :71 */
int rtw_add_material_dielectric(rtw_scene *s, float ir);                       /* material.rs:102-147 */
int rtw_add_material_diffuse_light(rtw_scene *s, int emit_texture);            /* light_source.rs:13-24 */

/* ---- instance wrappers: hittable/transformations.rs -------------------------------------- */
/* push = "everything emitted until the matching pop is `inner`".  Nesting mirrors the reference:
 *   cuboid.rotate_y(15).translate(v)  ==  push_translation(v); push_rotation_y(15); add_cuboid; pop; pop */
int rtw_push_translation(rtw_scene *s, const float offset[3]);   /* Translation  :16-48   */
int rtw_push_rotation_y(rtw_scene *s, float angle_degrees);      /* YRotation    :50-153  */
/* the same wrapper from the two values YRotation actually stores (sin_theta, cos_theta: transformations.rs:51-56,
 * computed at :60-63) — what the Rust `flatten` hook passes, so that no libm call is repeated on this side */
int rtw_push_rotation_y_sincos(rtw_scene *s, float sin_theta, float cos_theta);
int rtw_pop_transform(rtw_scene *s);
/* BvhNode::new(objects, ..) (bvh.rs:19-74) is an acceleration hint with no effect on results:
 * the backend always builds one LBVH over every primitive.  begin/end_group keep the call
 * structure of the reference visible to sinks that want it. */
int rtw_begin_group(rtw_scene *s);
int rtw_end_group(rtw_scene *s);

/* ---- participating media: hittable/volumes.rs ----------------------------------------------- */
/* ConstantMedium::new(boundary, density, texture) (volumes.rs:24-35): everything emitted until the
 * matching rtw_end_medium is the BOUNDARY — exactly one sphere or cuboid, optionally inside
 * push/pop transforms (the two shapes the reference's scenes use; anything else is
 * RTW_ERR_UNSUPPORTED).  The medium is ONE canonical primitive (returned id); its boundary gets no
 * ids.  Its material is Isotropic{texture} (material.rs:149-168).  The random draw of
 * ConstantMedium::hit (volumes.rs:58) is keyed by (pixel, sample, bounce, medium id) — see DESIGN.md. */
int rtw_begin_medium(rtw_scene *s, float density, int texture);
int rtw_end_medium(rtw_scene *s);

/* ---- primitives (return the canonical id of the first primitive they emit) ---------------- */
int rtw_add_sphere(rtw_scene *s, const float center[3], float radius, int material); /* spherical.rs:80-105 */
int rtw_add_moving_sphere(rtw_scene *s, const float center0[3], float time0, const float center1[3],
                          float time1, float radius, int material);                   /* spherical.rs:107-151 */
int rtw_add_xy_rect(rtw_scene *s, float x0, float x1, float y0, float y1, float k, int material); /* rectangular.rs:16-65   */
int rtw_add_xz_rect(rtw_scene *s, float x0, float x1, float z0, float z1, float k, int material); /* rectangular.rs:67-116  */
int rtw_add_yz_rect(rtw_scene *s, float y0, float y1, float z0, float z1, float k, int material); /* rectangular.rs:118-167 */
int rtw_add_cuboid(rtw_scene *s, const float p0[3], const float p1[3], int material);             /* rectangular.rs:170-245: 6 prims */
/* n triangles (triangular.rs:34-73).  vertices: n*9 floats (a,b,c).  normals: n*9 floats or NULL
 * (-> un-normalised face normal (b-a)x(c-a), triangular.rs:53-55).  uvs: n*6 floats or NULL
 * (-> (0,0),(1,0),(0,1), triangular.rs:57-65).  material_ids: n ints or NULL (-> `material`). */
int rtw_add_triangles(rtw_scene *s, uint32_t n, const float *vertices, const float *normals,
                      const float *uvs, const int32_t *material_ids, int material);

/* ---- build: flatten -> SoA upload -> LBVH on the GPU ---------------------------------------- */
/* [time0, time1] is the interval the boxes of moving primitives must cover (bvh.rs:22-23;
 * every scene of the reference passes 0, 1). */
int rtw_build(rtw_scene *s, float time0, float time1, rtw_build_stats *stats /* may be NULL */);
int rtw_scene_num_prims(const rtw_scene *s);
int rtw_scene_num_nodes(const rtw_scene *s);
int rtw_scene_num_instances(const rtw_scene *s); /* transform chains incl. the identity (index 0) */
/* Host-side view of the flattening (works before rtw_build and without a GPU):
 * type (0 sphere, 1 moving sphere, 2 yz-rect, 3 xz-rect, 4 xy-rect, 5 triangle, 6 sphere medium,
 * 7 cuboid medium), instance chain
 * index and material of a primitive; and the wrappers of a chain, outermost first
 * (kind 0 = Translation{a,b,c = offset}, kind 1 = YRotation{a = sin, b = cos}); returns the op count. */
int rtw_scene_prim_info(const rtw_scene *s, int prim_id, int32_t *type, int32_t *instance, int32_t *material);
int rtw_scene_instance_ops(const rtw_scene *s, int instance, int max_ops, int32_t *kinds, float *abc);
/* Copy the LBVH back to the host for inspection (tests / oracle cross-check).
 * nodes: 2*num_nodes records (pair i = nodes[2i], nodes[2i+1]); slot_prim_ids: num_prims ints
 * (primitive slot -> canonical id); root_box: 6 floats. Any pointer may be NULL. */
int rtw_get_bvh(const rtw_scene *s, rtw_bvh_node *nodes, int32_t *slot_prim_ids, float *root_box);

/* ---- the parity entry point: closest hit of a ray batch ------------------------------------- */
#define RTW_TRACE_BVH 0    /* persistent-thread LBVH traversal (the product path)  */
#define RTW_TRACE_BRUTE 1  /* every primitive against every ray, canonical order    */
int rtw_trace_closest(rtw_scene *s, const rtw_ray *rays, uint64_t n, rtw_hit *hits, int mode);
/* same with device pointers, asynchronous on `stream` (a cudaStream_t; NULL = default stream) */
int rtw_trace_closest_device(rtw_scene *s, const rtw_ray *d_rays, uint64_t n, rtw_hit *d_hits,
                             int mode, void *stream);

/* ---- the render entry point: Raytracer::render (lib.rs:57-76) ------------------------------- */
/* accum_rgb: width*height*3 floats, pixel (row, column) at ((height-1-row)*width + column)*3,
 * i.e. the order in which the reference yields its Pixels (row = bottom-up like Pixel.row,
 * lib.rs:58,120-126).  Value = un-normalised SUM over the rendered samples (lib.rs:82-94).
 * Pixels outside this call's tile partition are written as 0.
 * params->gpus > 1 (the `gpus` argument of SURVEY.md 8b): one process drives that many devices — the scene is
 * replicated with rtw_scene_clone on first use, device i renders the 32x32 tiles k with k % gpus == i (one host
 * thread per device) and stores its finished pixels straight into the frame on the scene's device through peer
 * memory (NVLink / NVSwitch; a staged peer copy when the devices are not peers), then ONE device-to-host copy.
 * part_rank / part_count must be 0 in that case. */
int rtw_render(rtw_scene *s, const rtw_camera *cam, const rtw_render_params *params,
               float *accum_rgb, rtw_render_stats *stats /* may be NULL */);
/* device-resident variant: d_accum_rgb is device memory of the scene's device; work is queued on
 * `stream`; the call returns after the stream has drained (stats need the final counters). */
int rtw_render_device(rtw_scene *s, const rtw_camera *cam, const rtw_render_params *params,
                      float *d_accum_rgb, void *stream, rtw_render_stats *stats);

/* ---- animation: one resident scene, many cameras (scenes.rs:622-667, main.rs:48-95) ------------ */
/* The reference renders `cams` one after the other over the same world (main.rs:48) and writes a PNG per
 * frame.  rtw_render_frames keeps the scene and its LBVH on the device and renders frame i with cameras[i]
 * and the stream seed params->seed + i.  A finished frame is copied to pinned host memory on a copy stream
 * and handed to `on_frame` on a helper thread while the device is already rendering the next one (two
 * buffers in flight), so PNG encoding / ProgressMessage streaming overlaps the rendering.
 * Callbacks arrive one at a time, in frame order; accum_rgb (layout as in rtw_render) and stats are valid
 * only during the call.  A non-zero return value stops the animation after the frames already in flight.
 * on_frame may be NULL (frames are rendered and dropped: benchmarking).  Returns the number of frames whose
 * callback ran (or that were rendered, when on_frame is NULL), < 0 on error. */
typedef int (*rtw_frame_callback)(void *user, uint32_t frame, const float *accum_rgb, const rtw_render_stats *stats);
int rtw_render_frames(rtw_scene *s, const rtw_camera *cameras, uint32_t n_frames, const rtw_render_params *params,
                      rtw_frame_callback on_frame, void *user);

/* console_app/src/main.rs:73-86: c = sqrt(sum/spp); (255.999 * clamp(c, 0, 0.999)) as u8.
 * accum_rgb as above, rgb8 = width*height*3 bytes, same pixel order (top row first). */
int rtw_resolve_rgb8(rtw_scene *s, const float *accum_rgb, uint32_t width, uint32_t height,
                     uint32_t spp, uint8_t *rgb8);

#ifdef __cplusplus
}
#endif
#endif /* RTW_CUDA_H */
