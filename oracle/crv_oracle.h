/* oracle/crv_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A scalar fp32 restatement of the reference's two compute shaders
 *     /root/reference/shaders/PathTracer.comp      (all of :104-127,139-184,197-206,220-361,366-425)
 *     /root/reference/shaders/ShadowVolume.comp    (:34-56,77-94)
 * with the sampler semantics of Initializers.cpp:408-418 (LINEAR, CLAMP_TO_EDGE, single mip) evaluated by
 * the Vulkan "texel filtering" rules.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (cloudrendering_vulkan_b200/) never does.
 *
 * PARITY STATUS: "parity unpinned" by the reference -- the reference ships no SPIR-V, its Tests.cpp holds no
 * expected values and is never run, and this image has no Vulkan loader / glslc /
 * lavapipe, so the shader TEXT is the only specification.  What *is* pinned against real reference code:
 * the five parameter structs and Grid3D (oracle/_ref, built from the reference's own headers).
 * Where GLSL leaves behaviour to the driver the oracle fixes it as: no FMA contraction, correctly rounded
 * + - * / sqrt, glibc libm for log/exp/pow/sin/cos/tan, constants folded in fp32, lowp evaluated as fp32,
 * dot products and matrix*vector summed left to right.
 *
 * Parameter blocks are passed as raw bytes and read at the std140 offsets the shaders declare
 * (PathTracer.comp:43-99), independently of the product's headers.
 */
#ifndef CRV_ORACLE_H
#define CRV_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG (PathTracer.comp:220-241, seed expression :368) ---- */
uint32_t orc_rng_seed_state(int32_t seed, uint32_t gid_x, uint32_t gid_y, uint32_t num_groups_x);
uint32_t orc_rng_warmup_count(uint32_t state0);
/* n draws after the warm-up: raw states and the floats generateRandomNumber returns */
void orc_rng_draws(int32_t seed, uint32_t gid_x, uint32_t gid_y, uint32_t num_groups_x, uint32_t n,
                   uint32_t* states, float* floats);
/* first n_draws raw states for every pixel of a WxH image, out[(y*W+x)*n_draws + i] */
void orc_rng_tap(int32_t seed, uint32_t width, uint32_t height, uint32_t n_draws, uint32_t* out);

/* ---- unit functions ---- */
void  orc_camera_ray(const void* camera68, int x, int y, float pos[3], float dir[3]);           /* :144-151 */
int   orc_intersect_cloud(const void* cloud60, const float pos[3], const float dir[3], float* tmax, float* tmin); /* :154-184 */
void  orc_orthonormal_basis(const float n[3], float xp[3], float yp[3]);                        /* :197-206 */
float orc_calculate_pdf(const void* params28, float costheta);                                   /* :247-250 */
float orc_invert_cdf(const void* params28, float xi);                                            /* :252-256 */
float orc_sample_phase(const void* params28, const float incoming[3], const float sample[3]);   /* :259-274 */
void  orc_background(const float dir[3], float rgba[4]);                                         /* :104-113 */
/* texture(sampler3D, uvw).x with LINEAR/CLAMP_TO_EDGE on an x-fastest nx*ny*nz float volume */
float orc_texture3d(const float* vol, uint32_t nx, uint32_t ny, uint32_t nz, const float uvw[3]);
float orc_sample_cloud(const float* vol, const void* cloud60, const float pos[3]);               /* :115-119 */
float orc_sample_shadow_volume(const float* sv, const void* svprops152, const float pos[3]);     /* :121-127 */
/* scatterRay with an explicit RNG state (in/out) */
void  orc_scatter_ray(const void* params28, uint32_t* rng_state, const float incoming[3], float out_dir[3]); /* :277-311 */

/* ---- ShadowVolume.comp:main (:77-94): out is x-fastest N*N*N ---- */
void orc_shadow_volume(const float* cloud, const void* cloud60, const void* svprops152, float* out, int num_threads);

/* ---- PathTracer.comp:main (:366-425): one dispatch = one progressive frame ----
 * image      : W*H*4 floats, read-modify-written with the running mean (:420-424)
 * sample     : optional W*H*4, the per-pixel `result` before accumulation
 * steps      : optional W*H u32, delta-tracking loop iterations (tentative collisions)
 * collisions : optional W*H u32, real collisions
 * absorption : extension, 0 = reference semantics (see crv_pt_extensions in include/crv_c_api.h)
 * accum_sum  : 0 = reference running mean, 1 = image += sample                                  */
void orc_path_trace_frame(const float* cloud, const float* sv, const void* camera68, const void* cloud60,
                          const void* params28, const void* svprops152, const void* push24, float absorption,
                          int accum_sum, float* image, float* sample, uint32_t* steps, uint32_t* collisions,
                          int num_threads);
/* same, restricted to the pixel rectangle [x0,x1) x [y0,y1) (bounded CPU-baseline samples) */
void orc_path_trace_rect(const float* cloud, const float* sv, const void* camera68, const void* cloud60,
                         const void* params28, const void* svprops152, const void* push24, float absorption,
                         int accum_sum, int x0, int y0, int x1, int y1, float* image, float* sample,
                         uint32_t* steps, uint32_t* collisions, int num_threads);

/* same, for rows 0, row_step, 2*row_step, ... of the image (bounded, unbiased sample of one frame) */
void orc_path_trace_rows(const float* cloud, const float* sv, const void* camera68, const void* cloud60,
                         const void* params28, const void* svprops152, const void* push24, float absorption,
                         int row_step, float* image, float* sample, uint32_t* steps, uint32_t* collisions,
                         int num_threads);

/* ---- synthetic fBm cloud (frozen generator, see oracle/crv_oracle.cpp) ---- */
void orc_fbm_generate(uint32_t n, uint32_t seed, float* out_xfastest, int num_threads);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

/* ---- scene setup that lives in main.cpp (used by the reference arm of bench.py) ---- */
#ifdef __cplusplus
extern "C" {
#endif
void orc_set_cloud_properties(const uint32_t counts[3], const double voxel_size[3], float majorant, float density_scaling, void* out60); /* main.cpp:144-163 */
void orc_msvc_rand(uint32_t seed0, uint32_t n, int32_t* out);   /* main.cpp:722,949 */
#ifdef __cplusplus
}
#endif
