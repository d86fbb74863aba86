// pt_wavefront.h -- kernel-parameter block of the path tracer (see pt_wavefront.cu)
#pragma once
#include "crv_internal.h"

// stats slots (u64 each)
enum { CRV_PT_SCHED_POOL = 0, CRV_PT_SCHED_LANES = 1 };
enum { PT_STAT_PATHS = 0, PT_STAT_STEPS = 1, PT_STAT_COLLISIONS = 2, PT_STAT_FETCHES = 3, PT_STAT_TRACKED = 4, PT_STAT_COUNT = 5 };

constexpr int PT_MAX_FRAMES_PER_LAUNCH = 256;
constexpr int PT_THREADS = 256;
// y-slot of a sample record that carries no background term (absorbed path): a quiet NaN with a payload that fp32
// arithmetic never generates (it produces the canonical 0x7fffffff)
constexpr int PT_NO_BACKGROUND_BITS = 0x7fc0dead;

struct PTConst {
    // ---- CameraProperties (PathTracer.comp:43-55), products precomputed on the host in the shader's order ----
    float cam_pos[3];
    float fwd_near[3];       // forward * nearPlane
    float right_psx[3];      // right * pixelSizeX
    float up_psy[3];         // up * pixelSizeY
    int half_w, half_h;
    int width, height;       // 2*half (CameraProperties::GetWidth/GetHeight)
    uint32_t row_stride;     // gl_WorkGroupSize.x * gl_NumWorkGroups.x = 32 * (W/32 + 1)  (:368, RenderTechniquePT.cpp:141)

    // ---- CloudProperties (:58-66) ----
    float cb0[3], cb1[3], cext[3];     // bounds[0], bounds[1], bounds[1]-bounds[0]
    float max_ext;                     // maxExtinction
    float sigma;                       // (maxExtinction * densityScaling) / baseScaling   (:338)
    float density_scaling;
    float r_sigma, r_max_ext, rcext[3]; // RN(1/x) of the divisors above (host IEEE division) for exact_div.cuh
    float neg_ln2_over_sigma;          // CRV_PT_MATH_DRIVER: t += lg2(zeta) * (-ln2 / sigma)
    int math_mode;                     // crv_pt_math_mode
    int use_wide;                      // 1: approximate brick pre-test enabled (its error bound holds for this geometry);
                                       // 0: asc8 = aof8 = 0 and pbias selects the +inf entry behind the table
    // approximate PADDED brick coordinate: q*asc8 + aof8 ~ (((q - cb0)/cext)*size - 0.5)/8 + 1, floored by a round-down
    // magic add; (pnbx, pnbxy) are the strides of the padded table and pbias removes the three magic exponents
    float asc8[3], aof8[3];
    uint32_t pnbx, pnbxy, pbias, pmax;

    // ---- Parameters (:68-78) ----
    float light_intensity;             // 1.0f * lightIntensity
    float two_g;                       // 2.0f * phaseG
    float one_minus_g;                 // 1.0f - phaseG
    float one_plus_g2, one_minus_g2, one_over_2g;
    float inv4pi_omg2;                 // INV_4Pi * phaseOneMinusG2   (:249)
    int isotropic;
    float absorption;                  // extension; 0 = reference semantics

    // ---- ShadowVolumeProperties (:81-91) ----
    float light[3];                    // lightDirection.xyz
    float sb0[4];                      // bounds[0]
    float m0[3], m1[3], m2[3], m3[3];  // basisChange columns (.xyz)
    float sden[3];                     // (basisChange * (bounds[1] - bounds[0])).xyz
    float rsden[3];                    // RN(1 / sden[i])

    // ---- resources ----
    VolView cloud, sv;
    const float* qtable;               // per 8^3 brick: upper bound of ext/maxExtinction over voxels [8b, 8b+8]^3
    const float* qtable_wide;          // same over voxels [8b-1, 8b+9]^3: valid for a brick index that is off by < 1 texel;
                                       // stored TIMES 2^32 with one brick of clamp-to-edge padding on every side
                                       // ((nbx+2)(nby+2)(nbz+2) entries) and one +inf entry behind them
    // Pool scheduler: the same bound at a coarser granularity, small enough to live in shared memory.  One byte per cell
    // of (8 << coarse_shift)^3 voxels (+ the one-texel dilation), padded by one cell on every side like qtable_wide:
    // thr = ceil(128 * q) clamped to 128, so that (u >> 25) >= thr implies float(u) * 2^-32 >= q for the raw draw u.
    const uint8_t* qcoarse;
    uint32_t qcoarse_bytes;            // padded cell count + 1 (the last byte, 128 = never reject, is the disabled-test target)
    float ascc[3], aofc[3];            // approximate padded CELL coordinate, like asc8/aof8
    uint32_t cnx, cnxy, cbias, cmax;   // like pnbx/pnbxy/pbias/pmax for the coarse table
    const uint32_t* hitlist;           // pixels whose path has to be tracked (pt_classify_kernel), x | y << 16, *nhit entries
    const uint32_t* nhit;
    const float* start;                // per tracked pixel, 32 bytes: entry point xyz, direction xyz, `lim`, `t_lo` (below)

    // ---- occupancy spans (proof-based skipping of whole ray sections; see pt_classify_kernel) ----
    // A brick b is OCCUPIED when occ_wide[b] != 0, i.e. some voxel of [8b-1, 8b+9]^3 is not +-0.  A sample position whose
    // brick (and that within one texel) is unoccupied blends eight zeros: ext == +-0 and the shader's acceptance test
    // `xi < ext / maxExt` (:351) is false for every draw.  So the part of a segment before its first occupied brick
    // (t < t_lo) only advances the RNG and t, and once a segment is past its last occupied brick (t >= lim) the path is
    // certain to leave the box without another collision: its sample is final.
    const float* occ_wide;             // == cloud volume's brick_absmax_wide (nbx*nby*nbz)
    const float* occ_aabb;             // device, 6 floats: world-space box around all occupied bricks (lo xyz, hi xyz), one
                                       // texel of margin; lo = +inf / hi = -inf when nothing is occupied
    float bsc[3], bof[3];              // brick coordinate of a point: v = q*bsc + bof = ((q - cb0)/cext * n - 0.5) / 8
    float span_margin;                 // safety margin of t_lo / lim in world units (a quarter texel)
    int skip_level;                    // 0: every step of the reference is executed (counters equal the oracle's);
                                       // 1: rays that meet no occupied brick are background pixels, segments end at `lim`

    // ---- job description ----
    uint32_t npix;                     // width*height
    uint32_t nframes;                  // frames in this launch
    // Job order (crv_pt_set_job_order).  0: frame-major, job = frame * nhit + k (the hit list as the classification left it).
    // 1 (default): pixel-major, job = k * nframes + frame, over a hit list sorted by descending expected path length: all
    // frames of the longest pixels start first, so what is still in flight when the queue runs dry are short paths.
    int job_order;
    uint32_t div_magic, div_shift;             // job / nframes without a division (U32Div below; host: make_u32_div)
    const uint32_t* hitlist_sorted;            // the sorted copy of hitlist / start (crv_launch_sort_hits) and, per entry, its
    const float* start_sorted;                 // index in the unsorted list: sample records stay indexed frame * nhit + that,
    const uint32_t* sorted_from;               // so the fold does not depend on the job order
    int32_t seeds[PT_MAX_FRAMES_PER_LAUNCH];   // pushConstants.seed of every frame of this launch (by value: no staging copy)
    uint32_t first_frame_count;        // pushConstants.frameCount of frame 0 of this launch
    const float* rcp_fc1;              // rcp_fc1[f] = RN(1 / float(first_frame_count + f + 1)) for the frames of this render CALL
                                       // (rcp_kernel, IEEE division): the running mean's divisor is frame-uniform, so its
                                       // division is the exact FMA division of exact_div.cuh
    int accum_mode;                    // crv_accum_mode
    float4* image;                     // W*H rgba32f result image (binding 0); written by pt_classify_kernel / fold_kernel
    float2* samples;                   // nframes * nhit records (S, dir.y), index = job
    unsigned int* job_counter;         // next unclaimed job of this launch
    unsigned long long* stats;         // PT_STAT_COUNT u64 (may be null)
    uint32_t* dbg_steps;               // per-pixel counters (single-frame renders, both null or both set)
    uint32_t* dbg_collisions;
    int scheduler;                     // CRV_PT_SCHED_POOL (default) or CRV_PT_SCHED_LANES
    int track_threshold;               // LANES: service collisions/regeneration once fewer lanes than this are tracking
    int steps_per_decision;            // LANES scheduler: tracking steps between two scheduling decisions (the POOL round is compile-time: crv_pt_pool_round_steps)
};

struct FoldConst {
    float4* image;
    const float2* samples;
    const uint32_t* hitlist;
    const uint32_t* nhit;
    uint32_t npix, nframes, first_frame_count, width;
    int accum_mode;
    const float* rcp_fc1;              // RN(1 / float(first_frame_count + f + 1)), f = 0 .. nframes-1 of THIS launch; null: computed in place
    const uint32_t* misslist;          // pixels whose camera ray misses the box (npix - *nhit entries): pixel index
    const float* missdy;               // and dir.y of that ray
    unsigned int* job_counter;         // reset for the next tracker launch
    unsigned long long* stats;         // PT_STAT_PATHS += the miss pixels' paths (may be null)
    // ---- push mode (crv_pt_render_batch_push, fold_push_kernel): CRV_ACCUM_SUM whose result leaves for the group members ----
    int fresh;                         // the image holds nothing yet: this launch's sums replace its contents (no clear needed)
    int last_round;                    // last launch of the render call: tracked sums go to the inboxes, the other pixels are finished
    uint32_t frames_total;             // frames of the whole render call (the other pixels hold frames_total identical samples)
    float others_scale;                // their final value: (frames_total * sampleBackground) * others_scale
    uint32_t push_n, push_rank, push_epoch;
    float4* push_inbox[16];            // member p's inbox image for this member's sums (own inbox included)
    uint32_t* push_flag[16];           // word in member p's flag page: "the sums of member push_rank, epoch push_epoch, have arrived"
    uint32_t* push_done;               // CTA counter in this member's flag page, zero between launches
};


int crv_launch_pt(crv_ctx* ctx, const PTConst& c, int cloud_layout, int sv_layout, int ctas_per_sm, cudaStream_t s);
int crv_launch_classify(crv_ctx* ctx, const PTConst& c, uint32_t* hitlist, float* start, uint32_t* counts, uint32_t* misslist, float* missdy, float* keys, cudaStream_t s);
// hit list + start records sorted by descending key (a counting sort over PT_SORT_BUCKETS key ranges; the order inside a
// range is arbitrary) into (hitlist_out, start_out), from_out[j] = index of entry j in the unsorted list;
// scratch = PT_SORT_BUCKETS + 1 u32, key_max = an upper bound of the keys
constexpr int PT_SORT_BUCKETS = 4096;
int crv_launch_sort_hits(crv_ctx* ctx, const uint32_t* counts, const float* keys, float key_max, const uint32_t* hitlist, const float* start,
                         uint32_t* hitlist_out, float* start_out, uint32_t* from_out, uint32_t* scratch, cudaStream_t s);
// unsigned 32-bit division by a launch-uniform divisor d >= 1: magic == 0 (d a power of two): n >> shift; otherwise
// q = umulhi(n, magic), n / d = (((n - q) >> 1) + q) >> shift  (Granlund-Montgomery, the round-up form that needs no 33rd bit)
struct U32Div { uint32_t magic, shift; };
U32Div make_u32_div(uint32_t d);
int crv_launch_fold(crv_ctx* ctx, const FoldConst& c, cudaStream_t s);
int crv_launch_fold_push(crv_ctx* ctx, const FoldConst& c, cudaStream_t s);
int crv_launch_occ_aabb(crv_ctx* ctx, const int* ext6, float* aabb6, const float* b0, const float* extent, int nx, int ny, int nz, cudaStream_t s);
int crv_launch_rcp_frames(crv_ctx* ctx, float* rcp, uint32_t first_frame_count, uint32_t n, cudaStream_t s);
int crv_launch_qtable(crv_ctx* ctx, const float* brick_max, float* qtable, long long n, float max_ext, cudaStream_t s);
int crv_launch_qcoarse(crv_ctx* ctx, const float* absmax_wide, uint8_t* qc, int nbx, int nby, int nbz, int shift, float max_ext, cudaStream_t s);
int crv_pt_pool_round_steps();     // tracking steps of the pool scheduler's unrolled round
int crv_pt_pool_smem_budget();     // bytes of shared memory the pool scheduler leaves for the coarse table
int crv_launch_qtable_padded(crv_ctx* ctx, const float* absmax_wide, float* qpad, int nbx, int nby, int nbz, float max_ext, cudaStream_t s);
int crv_launch_rng_tap(crv_ctx* ctx, int32_t seed, uint32_t w, uint32_t h, uint32_t n_draws, uint32_t* d_out, cudaStream_t s);
int crv_launch_log_selftest(crv_ctx* ctx, unsigned long long* d_mismatches, cudaStream_t s);
int crv_launch_scale(crv_ctx* ctx, float* data, size_t n, float scale, cudaStream_t s);

// ---- PPM photon tracer (PPM_PT.comp; RenderTechniquePPM's tracing pass) -------------------------------------------
constexpr uint32_t PPM_ELEMENTS_PER_CELL = 32;      // PPM_PT.comp:26
struct PPMConst {
    PTConst base;                  // cloud bounds / extinction / phase parameters / cloud view / narrow q-table, as the path tracer builds them
    // PhotonMapProperties (PPM_PT.comp:57-67)
    float pm_light[3];
    float pm_b0[3];
    int pm_count[3];
    float pm_voxel_size;
    float pm_absorption;
    int32_t seed;                  // pushConstants.seed
    uint32_t n_photons;            // 64 * groups_x invocations (= emittedPhotons, :459)
    crv_photon* photons;           // cells * 32 records
    uint32_t* collisions;          // cells uvec4, .x = photons deposited in the cell
    unsigned long long* stats;     // [0] tracking iterations, [1] deposits, [2] deposits dropped (cell already holds 32), [3] estimate march steps
    unsigned long long cells;
    // ---- estimate pass (PPM_PE.comp) ----
    float pm_step_size;            // PhotonMapProperties.stepSize
    float pm_radius;               // pushConstants.pmRadius
    float base_scaling;            // CloudProperties.baseScaling
    uint32_t frame_count;          // pushConstants.frameCount
    float4* image;                 // result image (binding 0)
};
int crv_launch_ppm_estimate(crv_ctx* ctx, const PPMConst& c, int cloud_layout, int sv_layout, cudaStream_t s);
int crv_launch_ppm_clear(crv_ctx* ctx, crv_photon* photons, uint32_t* collisions, unsigned long long cells, cudaStream_t s);
int crv_launch_ppm_trace(crv_ctx* ctx, const PPMConst& c, int cloud_layout, cudaStream_t s);
