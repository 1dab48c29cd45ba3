/*
 * rtw_sink.h — the "SceneSink" the host front end flattens a scene into (SURVEY.md §8b, row
 * "Obstacle": `fn flatten(&self, b: &mut dyn SceneSink)` on Hittable / Material / Texture).
 *
 * It is nothing but the emit / build / trace / render entry points of include/rtw_cuda.h as a
 * table of function pointers plus the handle they act on, so that the front end depends on the
 * C ABI only and not on a particular shared library.  The product fills it from
 * librtw_cuda.so (prefix "rtw_"); the test-suite may fill a second one from the CPU oracle
 * (prefix "orc_"), which exports the same signatures — the front end never links the oracle.
 */
#ifndef RTW_SINK_H
#define RTW_SINK_H

#include "rtw_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtw_sink {
  void *lib;    /* dlopen handle (owned by the sink)            */
  void *scene;  /* rtw_scene* (or the oracle's equivalent)      */
  const char *(*last_error)(void);
  int (*scene_create)(int device, void **out);
  int (*scene_destroy)(void *s);
  int (*add_texture_solid)(void *s, float r, float g, float b);
  int (*add_texture_checker)(void *s, int odd, int even, float frequency);
  int (*add_texture_noise)(void *s, const float *g, const int32_t *px, const int32_t *py, const int32_t *pz, float scale);
  int (*add_texture_uvdebug)(void *s);
  int (*add_texture_image)(void *s, const uint8_t *rgb8, uint32_t w, uint32_t h);
  int (*add_material_lambertian)(void *s, int tex);
  int (*add_material_metal)(void *s, float r, float g, float b, float fuzz);
  int (*add_material_dielectric)(void *s, float ir);
  int (*add_material_diffuse_light)(void *s, int tex);
  int (*push_translation)(void *s, const float offset[3]);
  int (*push_rotation_y)(void *s, float angle_degrees);
  int (*pop_transform)(void *s);
  int (*begin_group)(void *s);
  int (*end_group)(void *s);
  int (*begin_medium)(void *s, float density, int texture);
  int (*end_medium)(void *s);
  int (*add_sphere)(void *s, const float c[3], float radius, int material);
  int (*add_moving_sphere)(void *s, const float c0[3], float t0, const float c1[3], float t1, float radius, int material);
  int (*add_xy_rect)(void *s, float x0, float x1, float y0, float y1, float k, int material);
  int (*add_xz_rect)(void *s, float x0, float x1, float z0, float z1, float k, int material);
  int (*add_yz_rect)(void *s, float y0, float y1, float z0, float z1, float k, int material);
  int (*add_cuboid)(void *s, const float p0[3], const float p1[3], int material);
  int (*add_triangles)(void *s, uint32_t n, const float *v, const float *nrm, const float *uv, const int32_t *mats, int material);
  int (*build)(void *s, float time0, float time1, rtw_build_stats *stats);
  int (*render)(void *s, const rtw_camera *cam, const rtw_render_params *p, float *accum_rgb, rtw_render_stats *stats);
  int (*render_frames)(void *s, const rtw_camera *cams, uint32_t n_frames, const rtw_render_params *p,
                       rtw_frame_callback on_frame, void *user);
} rtw_sink;

/* Fill `out` from the shared library at `path`, looking every entry point up as <prefix><name>
 * (e.g. "rtw_" + "add_sphere") and creating a scene on `device`.  Returns RTW_OK or RTW_ERR_*;
 * rtwh_last_error() has the message.  Fails loudly when the library or a symbol is missing. */
int rtwh_sink_open(const char *path, const char *prefix, int device, rtw_sink *out);
int rtwh_sink_close(rtw_sink *sink); /* destroys the scene and closes the library */
const char *rtwh_last_error(void);

/* ---- ProgressMessage stream (lib.rs:128-138) in the wire format of the untouched host receivers ------------
 * discovery_app serialises ProgressMessage with postcard 0.7.3 `to_vec_cobs` (discovery_app/src/bin/raytracer.rs:
 * 62,105,111) and discovery_host_receiver decodes it with `from_bytes_cobs` (src/main.rs:37).  postcard 0.7 (pinned
 * in Cargo.lock; the crate is not vendored, its published format is restated here): enum variant index = varint,
 * u32 / f32 = 4 bytes little endian, [f32; 3] = 12 bytes; then COBS framing + one 0x00 terminator.
 *   ImageStart{width, height, samples_per_pixel} -> 00 | w | h | spp          (13 bytes before COBS)
 *   Pixel(Pixel{row, column, color})             -> 01 | row | column | r g b  (21 bytes before COBS)
 *   ImageEnd                                     -> 02
 * Each function writes one framed message to `out` (capacity `cap`) and returns its length, or RTW_ERR_INVALID
 * when `cap` is too small (32 bytes always suffice). */
int rtwh_progress_image_start(uint32_t width, uint32_t height, uint32_t samples_per_pixel, uint8_t *out, size_t cap);
int rtwh_progress_pixel(uint32_t row, uint32_t column, const float color[3], uint8_t *out, size_t cap);
int rtwh_progress_image_end(uint8_t *out, size_t cap);
/* A whole frame as the reference would stream it: ImageStart, one Pixel per pixel in the order of lib.rs:58
 * ((0..h).rev() x (0..w), the order of accum_rgb), ImageEnd.  Returns the bytes written (size the buffer with
 * rtwh_progress_frame_bound) or RTW_ERR_INVALID. */
size_t rtwh_progress_frame_bound(uint32_t width, uint32_t height);
long long rtwh_progress_frame(const float *accum_rgb, uint32_t width, uint32_t height, uint32_t samples_per_pixel,
                              uint8_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* RTW_SINK_H */
