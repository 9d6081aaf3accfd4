/*
 * ds_abi.h -- C ABI of libdeepestscatter_b200.so: the drop-in boundary for the radiance
 * estimation hot path of marsermd/DeepestScatter's DataGen subsystem on NVIDIA B200.
 *
 * Every entry point replaces a piece of the reference's host<->OptiX interface; the
 * reference interface it stands in for is cited as file:line with
 *   DG/ = DeepestScatter_DataGen/DeepestScatter_DataGen/src/      CU/ = DG/CUDA/
 *
 * Conventions
 *   - plain C types only; all functions return 0 (DS_OK) or a negative DS_ERR_* code and
 *     record a message readable with ds_last_error().  No exceptions cross the boundary
 *     (the reference throws optix::Exception / sutil::APIError, DG/main.cpp:73-76).
 *   - a DsContext owns one CUDA device, one stream and all device memory (the analogue of
 *     the per-task optix::Context, DG/ExecutionLoop/GuiExecutionLoop.cpp:93-97).  It is
 *     thread-compatible: one host thread at a time.  One context per GPU.
 *   - host pointers are caller-owned; `_device` variants take device pointers.
 *   - images are float4 / uchar4 [H][W], row 0 at the bottom (CU/cameraCommon.cuh:22);
 *     volumes are u8 [Nz][Ny][Nx], x fastest (DG/Util/Resources.cpp:127-141).
 *   - there is no CPU fallback: every call fails with DS_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef DS_ABI_H
#define DS_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_OK 0
#define DS_ERR_INVALID (-1) /* bad argument / call order */
#define DS_ERR_CUDA (-2)    /* CUDA runtime error (message in ds_last_error) */
#define DS_ERR_STATE (-3)   /* required state missing (no volume, no frame, ...) */
#define DS_ERR_IO (-4)      /* file system error (record writer) */

typedef struct DsContext DsContext;

/* Cloud::Rendering::Mode (DG/Scene/SceneDescription.h:41-46) -> closest-hit program
 * (DG/Scene/CloudMaterial.cpp:51-64) */
typedef enum DsMode {
    DS_MODE_SUN_AND_SKY_ALL_SCATTER = 0, /* totalRadiance,              CU/cloudRadianceMaterials.cu:9  */
    DS_MODE_SUN_MULTIPLE_SCATTER = 1,    /* multipleScatterSunRadiance, CU/cloudRadianceMaterials.cu:72 */
    DS_MODE_SUN_SINGLE_SCATTER = 2       /* singleScatterSunRadiance,   CU/cloudRadianceMaterials.cu:120 */
} DsMode;

/* Arithmetic flavour of the estimator kernels.
 *  EXACT: software fp32 trilinear + include/ds_detmath.h transcendental kernels, compiled
 *         -fmad=false; bit-identical to the host oracle (tests assert equality).
 *  FAST:  hardware 3-D texture filtering + MUFU intrinsics, i.e. what the reference itself
 *         runs (--use_fast_math, vcxproj:320; rtTex3D, CU/cloud.cuh:61); checked against
 *         the oracle statistically (3 sigma per pixel, <0.5 % relative RMSE). */
typedef enum DsPrecision { DS_PRECISION_EXACT = 0, DS_PRECISION_FAST = 1 } DsPrecision;

/* The OptiX context variables of the scene (SURVEY.md 8b "parameter surface"). */
typedef struct DsSceneParams {
    float cloud_size_m;         /* Cloud::Model::size, SceneDescription.h:76; "cloudSizeInMeters" VDBCloud.cpp:110 */
    float mean_free_path_m;     /* SceneDescription.h:80 (10 m) -> densityMultiplier = size / mfp, VDBCloud.cpp:109 */
    float sample_step;          /* "sampleStep", installers.cpp:86 (1/512) */
    float light_direction[3];   /* "lightDirection", Sun.cpp:15; normalised like DirectionalLight's ctor */
    float light_color[3];       /* "lightColor", Sun.cpp:16 */
    float light_intensity;      /* "lightIntensity", Sun.cpp:17 (1e6, installers.cpp:100) */
    float minimal_ray_distance; /* "minimalRayDistance", CloudMaterial.cpp:23 (1e-6) */
} DsSceneParams;

/* "eye", "U", "V", "W" (DG/Scene/Cameras/Camera.cpp:129-132) */
typedef struct DsCamera {
    float eye[3];
    float U[3];
    float V[3];
    float W[3];
} DsCamera;

/* Gpu::PointRadianceTask (CU/PointRadianceTask.h:70-77), 40 bytes, same field order */
typedef struct DsPointRadianceTask {
    int32_t id;
    uint32_t experiment_count;
    float radiance;
    float running_variance;
    float position[3];
    float direction[3];
} DsPointRadianceTask;

/* Work counters accumulated by the estimator kernels since the last reset. */
typedef struct DsCounters {
    uint64_t paths;        /* primary samples started (hit or miss) */
    uint64_t events;       /* scatter events with a next-event estimate (getInScattering calls) */
    uint64_t steps;        /* ray-march steps of the reference algorithm (CU/cloud.cuh:87-104 iterations) */
    uint64_t density_taps; /* density fetches actually issued (<= steps when empty space is skipped) */
    uint64_t nonfinite;    /* samples that were NaN/Inf (the reference paints an error colour, progressive.cu:36) */
    /* Part of `paths` / `steps` that NO kernel thread traced: samples of pixels whose primary ray never reaches an occupied
     * cell (their value is exactly 0; the FAST flavour's primary-ray cache settles them once per camera) and the march steps
     * the reference would have spent on them.  paths - untraced_paths and steps - untraced_steps are what the path-tracing
     * kernel itself counted. */
    uint64_t untraced_paths;
    uint64_t untraced_steps;
} DsCounters;

/* RadianceCollector constants (DG/Scene/RadianceCollector.cpp:17,88,112-118) */
typedef struct DsRadianceSettings {
    uint32_t max_thread_count;    /* MAX_THREAD_COUNT = 10 * 2048 */
    uint32_t launches_per_update; /* 100 */
    uint32_t max_updates;         /* safety cap on update() calls; 0 = unlimited */
    float relative_ci;            /* 2e-2 */
    float absolute_ci;            /* 1e-4 */
    uint32_t zero_radiance_min_experiments; /* 100000 */
} DsRadianceSettings;

/* ---------------------------------------------------------------- context */

/* optix::Context::create (DG/installers.cpp:111) */
int ds_context_create(int device, DsContext** out);
/* context->destroy (GuiExecutionLoop.cpp:93-97) */
int ds_context_destroy(DsContext* ctx);
/* message of the last failing call on this context (ctx may be NULL for create failures) */
const char* ds_last_error(DsContext* ctx);
/* run all work of this context on an externally owned cudaStream_t (e.g. torch's current stream) */
int ds_context_set_stream(DsContext* ctx, void* cuda_stream);
/* block until all work queued by this context is complete */
int ds_sync(DsContext* ctx);
/* named integer options: "precision" (DsPrecision), "variant", "block_threads", "blocks_per_sm",
 * "skip_empty", "primary_cache", "march_keep_quarters", "march_keep32", "march_max_iters", "march_unroll", "regen_min", "skip_min",
 * "skip_max_iters", "skip_open_dist", "zero_check_min", "smem_carveout", "staging_subframes", "spec_percent", "region_pixels" -- tuning knobs
 * of the estimator kernels; "fused_volume" (FAST estimator: march through one RG8 {density, sun transmittance} array, default 1),
 * "escape_octants" (FAST estimator: per-cell flags that end a path in empty space whose whole octant ahead is empty, default 1) -- both
 * leave every result bit unchanged;
 * "radiance_scheduler", "radiance_quota" -- the radiance collector (0 = the reference's host schedule, 1 = device-resident);
 * "stream_offset" -- added to the subframe id to form the RNG stream id;
 * neural renderer: "mlp_bf16" / "mlp_fp16" (FAST flavour of the model on bfloat16 / IEEE half instead of tf32 operands), "descriptor_hw" (-1 = the FAST network-input passes
 * sample a mip-mapped texture, the collectors never; 0 = never; 1 = also the float collector), "compact_reverse" (test hook: process the
 * scattering pixels in the opposite order);
 * "profile_events" (1: CUDA events around the trace / model launches, read back through ds_get_launch_stats and "mlp_last_us"; 2: also the
 * instrumented instantiation of the model kernel, ds_disney_model_profile); read-only: "mlp_last_us", "volume_generation" */
int ds_set_option(DsContext* ctx, const char* name, int value);
int ds_get_option(DsContext* ctx, const char* name, int* value);
int ds_get_counters(DsContext* ctx, DsCounters* out);
int ds_reset_counters(DsContext* ctx);
/* launch accounting since the last ds_reset_counters: kernels launched by the estimator entry points; with
 * option "profile_events" = 1 also the summed CUDA-event duration (ms) of the path-tracing kernel launches */
int ds_get_launch_stats(DsContext* ctx, uint64_t* kernel_launches, uint64_t* trace_launches_timed, double* trace_ms_total);
/* library / device description as a JSON string owned by the context */
const char* ds_describe(DsContext* ctx);

/* ---------------------------------------------------------------- volume (cloud importer back end) */

/* Resources::loadVolumeBuffer after the VDB parse (Resources.cpp:103-148): dense u8 level 0,
 * optional box-filter mip chain generateMipmaps (Resources.cpp:169-209) built on the device. */
int ds_volume_upload(DsContext* ctx, const uint8_t* level0, int nx, int ny, int nz, int build_mips);
/* Resources.cpp:127-141: value / max_density * 255 truncated to u8, then as ds_volume_upload */
int ds_volume_upload_float(DsContext* ctx, const float* dense, int nx, int ny, int nz, double max_density, int build_mips);
/* synthetic procedural grid (include/ds_synth.h) generated on the device: n^3, kind 0/1/2 */
int ds_volume_synth(DsContext* ctx, int n, int kind, uint32_t seed, int build_mips);
int ds_volume_level_count(DsContext* ctx, int* count);
int ds_volume_level_dims(DsContext* ctx, int level, int dims[3]);
int ds_volume_download_level(DsContext* ctx, int level, uint8_t* out);

/* ---------------------------------------------------------------- scene */

void ds_scene_params_default(DsSceneParams* p);
/* Sun::init (Sun.cpp:13-18) + VDBCloud::setupVariables (VDBCloud.cpp:88-117) + CloudMaterial (CloudMaterial.cpp:14,23) */
int ds_scene_set(DsContext* ctx, const DsSceneParams* p);
/* derived variables as the reference sets them: out[0..2] bboxSize, [3..5] textureScale, [6] densityMultiplier,
 * [7] voxelSizeInMeters, [8] voxelSizeInTermsOfFreePath, [9..11] normalised lightDirection */
int ds_scene_get_derived(DsContext* ctx, float out[12]);
/* VDBCloud::InitInScatter (VDBCloud.cpp:57-86) launching CU/inScatter.cu:40-66 */
int ds_bake_sun_transmittance(DsContext* ctx);
int ds_inscatter_download(DsContext* ctx, uint8_t* out);
int ds_inscatter_upload(DsContext* ctx, const uint8_t* in);

/* sutil::calculateCameraVariables with fov_is_vertical = false (DG/Util/sutil.cpp:501-524), host only */
void ds_camera_look_at(const float eye[3], const float lookat[3], const float up[3], float hfov_deg, float aspect, DsCamera* out);
/* Camera::init defaults: eye (2.5,-0.4,0), lookat 0, up +y, hfov 30 (Camera.cpp:37-39,102) */
void ds_camera_default(int width, int height, DsCamera* out);

/* ---------------------------------------------------------------- progressive renderer */

/* Camera::init buffers frameResult/progressive/variance/screen (Camera.cpp:45-48) */
int ds_frame_create(DsContext* ctx, int width, int height);
/* Camera::reset -> clearScreen (Camera.cpp:77-86, CU/progressive.cu:29-34) */
int ds_frame_clear(DsContext* ctx);
/* ARenderer::render (DG/Scene/Cameras/ARenderer.h:15; PathTracingRenderer.cpp:21-31): one new sample per
 * pixel for `subframe_id`; frame_result_out (float4 [H][W], may be NULL) receives frameResultBuffer. */
int ds_render_frame_result(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t subframe_id, float* frame_result_out);
/* Camera::render loop body (Camera.cpp:189-199) for subframes first..first+n-1: render + updateFrameResult
 * (CU/progressive.cu:17-27) into the device-resident progressive / variance buffers. */
int ds_render_subframes(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t first_subframe, uint32_t n);
/* as ds_render_subframes, but with HOST accumulation buffers: uploads progressive/variance (float4 [H][W]),
 * renders, downloads them again -- the plugin-boundary call with all copies inside. */
int ds_render_subframes_host(DsContext* ctx, const DsCamera* cam, DsMode mode, uint32_t first_subframe, uint32_t n,
                             float* progressive_inout, float* variance_inout);
int ds_frame_download(DsContext* ctx, float* progressive_out, float* variance_out);
int ds_frame_upload(DsContext* ctx, const float* progressive, const float* variance);
/* raw device pointers of the float4 buffers (for torch.distributed / NCCL plumbing) */
int ds_frame_device_ptrs(DsContext* ctx, void** progressive, void** variance);
/* reinhard firstPass/secondPass/applyReinhard (CU/reinhard.cu:26-83; Camera.cpp:202-210): uchar4 [H][W] */
int ds_tonemap(DsContext* ctx, float exposure, uint8_t* screen_out, float* average_luminance_out);
/* Camera::isConverged (Camera.cpp:232-268): number of pixels failing the 95 % CI test at `subframe_id` */
int ds_frame_unconverged(DsContext* ctx, uint32_t subframe_id, uint32_t* unconverged_out);
/* Multi-GPU merge support.  Exports per pixel 8 doubles {n*mean_rgba, M2_rgba + n*mean_rgba^2} (device
 * pointer, 8*W*H doubles) for `n` subframes; sums over ranks can then be re-imported with the total n. */
int ds_frame_export_moments_device(DsContext* ctx, uint32_t n, double* moments_device);
int ds_frame_import_moments_device(DsContext* ctx, uint32_t n_total, const double* moments_device);

/* ---------------------------------------------------------------- multi-GPU accumulation-buffer reduce (NCCL)
 * The reference is single-GPU.  One context per GPU renders its own subframe ids; the per-GPU accumulation
 * buffers are combined with ONE ncclReduce (sum, float64) of the mergeable moments over NVLink.  NCCL is loaded at run time
 * (dlopen "libnccl.so.2"): single-GPU users need no NCCL, and inside a torch process the copy torch already loaded is shared. */
#define DS_COMM_ID_BYTES 128
/* ncclGetUniqueId on the calling process (rank 0); ship the bytes to the other ranks by any means */
int ds_comm_unique_id(uint8_t id_out[DS_COMM_ID_BYTES]);
/* ncclCommInitRank on the context's device and stream; collective over all ranks */
int ds_comm_init(DsContext* ctx, int n_ranks, int rank, const uint8_t id[DS_COMM_ID_BYTES]);
int ds_comm_destroy(DsContext* ctx);
/* Collective.  The frame of this context holds `n_local` accumulated subframes (0 allowed); after the call the frame of
 * `root` holds the statistics of all `n_total` = sum n_local subframes (mean and M2 as if one GPU had rendered them all);
 * root < 0: every rank does (ncclAllReduce).  Runs export -> ncclReduce -> import on the context's stream, no host sync. */
int ds_frame_reduce(DsContext* ctx, uint32_t n_local, uint32_t n_total, int root);

/* ---------------------------------------------------------------- generic path tracing (tests, tools) */

/* rtTrace of one radiance ray per (origin, direction) in world space; seed = tea<4>(seed_val0[i], stream[i])
 * (CU/cloudRadianceMaterials.cu:21 with clock() replaced by stream).  radiance_out: float3 per path. */
int ds_trace_paths(DsContext* ctx, DsMode mode, uint32_t n, const float* origins, const float* directions,
                   const uint32_t* seed_val0, const uint32_t* stream, float* radiance_out);

/* ---------------------------------------------------------------- dataset generation */

/* ScatterSampleCollector::collect launch (ScatterSampleCollector.cpp:35-40; CU/pointGeneratorCamera.cu:20-42,
 * CU/cloudFirstScatterMaterial.cu:8-28) for launch ids first_index..first_index+n-1 */
int ds_generate_points(DsContext* ctx, uint32_t first_index, uint32_t n, uint32_t stream, float* positions_out, float* directions_out);
/* DisneyDescriptorCollector::collect (DisneyDescriptorCollector.cpp:61-67; CU/disneyDescriptorCollector.cu:22-29,
 * CU/DisneyDescriptor.cuh:72-112): n descriptors of 10*225 u8 */
int ds_collect_descriptors(DsContext* ctx, const float* positions, const float* directions, uint32_t n, uint8_t* descriptors_out);
/* float variant (TElement = float, the density members of DisneyNetworkInput) and optional floor-corner voxel
 * addresses {x, y, z, level} per tap for index parity */
int ds_collect_descriptors_float(DsContext* ctx, const float* positions, const float* directions, uint32_t n, float* out,
                                 int32_t* tap_index_out);
/* IntersectionInfo (CU/rayData.cuh:28-33): float3 radiance, float transmittance, bool hasScattered (4-byte slot), 20 bytes */
typedef struct DsIntersectionInfo {
    float radiance[3];
    float transmittance;
    uint32_t has_scattered;
} DsIntersectionInfo;
/* First launch of the neural renderers' renderRect (DG/Scene/Cameras/DisneyRenderer.cpp:84-88): CU/disneyCamera.cu:20-36
 * (pinholeCamera for pixel launchID + rectOrigin) with CU/disneyDescriptorMaterial.cu:14-46 (sampleDisneyDescriptor) as
 * closest hit.  Per pixel of the rectangle: the transmittance of the whole ray, a collision forced inside the cloud
 * (xi = 1 - rnd * (1 - T)), the direct sun radiance there with the full Mie phase, and DisneyNetworkInput = 10 layers of
 * 225 float densities + the light / view angle (CU/DisneyDescriptor.h) -- the tensor the reference hands to its TorchScript
 * model (DisneyRenderer.cpp:30-36).  network_input_out: [rect_h][rect_w][10][226] floats (densities zero where nothing
 * scattered); info_out: [rect_h][rect_w].  The RNG seed is tea<4>(launchID.x * 4096 + launchID.y, stream) with the
 * rectangle-local launch index, as in the reference (clock() -> stream).  The model: ds_disney_model_forward below, or the caller's own. */
int ds_render_network_input(DsContext* ctx, const DsCamera* cam, uint32_t frame_width, uint32_t frame_height, uint32_t rect_x, uint32_t rect_y,
                            uint32_t rect_w, uint32_t rect_h, uint32_t stream, float* network_input_out, DsIntersectionInfo* info_out);
/* copyToFrameResult (CU/disneyCamera.cu:38-46), host side: frameResult[pixel] = (predicted + radiance) * (1 - transmittance) for
 * the pixels that scattered; frame_result_inout is float4 [frame_height][frame_width] */
int ds_blit_predicted(uint32_t frame_width, uint32_t frame_height, uint32_t rect_x, uint32_t rect_y, uint32_t rect_w, uint32_t rect_h,
                      const float* predicted, const DsIntersectionInfo* info, float* frame_result_inout);
/* ---- the radiance-predicting network of the neural renderer ----
 * DisneyRenderer::init loads a TorchScript export of DeepestScatter_Train/Disney/DisneyModel.py (DisneyRenderer.cpp:19-22) and
 * renderRect evaluates it on the rectangle's network inputs (:104).  Here the model is ONE flat float32 array: the tensors of
 * DisneyModel().state_dict() in their own order, row-major as torch stores them --
 *   for i in 0..9: blocks.i.f1z.weight [200][226], .f1z.bias [200], .f1o.weight [200][200], .f1o.bias [200], .f2.weight [200][200],
 *   .f2.bias [200] (DisneyBlock.py:13-15); fullyConnected.0.weight [200][200], .0.bias [200], .2.weight [200][200], .2.bias [200],
 *   .4.weight [1][200], .4.bias [1] (DisneyModel.py:52-59)
 * = ds_disney_model_weight_count() = 1 338 601 floats (deepestscatter_b200/disney_model.py: flatten_state_dict). */
size_t ds_disney_model_weight_count(void);
int ds_disney_model_load(DsContext* ctx, const float* weights, size_t count);
/* Introspection, host only (no device needed): the program the tensor-core kernel runs for these weights -- the weight stream in UMMA
 * canonical K-major no-swizzle layout (8-row x 16-byte core matrices; row groups 128 B apart, 16-byte K groups 26 * 128 B apart; 208 rows;
 * bf16 = 0: tf32-rounded floats, 4 per K group; bf16 = 1: bfloat16, bf16 = 2: IEEE half, 8 per K group) and the chunk table (20-byte records: u32 stream offset,
 * u32 bytes, u16 MMA steps (two K groups each), u16 first K group (activations) or first k (descriptor layer), u8 source, u8 layer,
 * u8 accumulator, u8 flags 1 = overwrite / 2 = last of its GEMM / 4 = first chunk after an epilogue, u8 epilogue 1 = relu -> activations /
 * 2 = same + residual kept in tensor memory / 3 = output, u8 GEMM index, 2 pad).  Either output pointer may be NULL to query the sizes. */
int ds_disney_model_pack(const float* weights, size_t count, int bf16, void* stream_out, size_t stream_capacity, void* chunks_out,
                         size_t chunks_capacity, size_t* stream_bytes, size_t* chunk_count);
/* module->forward (DisneyRenderer.cpp:104; DisneyModel.forward, DisneyModel.py:16-29): network_input [n][10][226] floats ->
 * predicted_out [n] (radiance for a sun of 1e6).  DS_PRECISION_EXACT: fp32 FMA kernel; DS_PRECISION_FAST: tcgen05 tensor-core kernel, inputs
 * and activations rounded to tf32 (option "mlp_bf16" = 1: to bfloat16, "mlp_fp16" = 1: to IEEE half), fp32 accumulation and residual path */
int ds_disney_model_forward(DsContext* ctx, const float* network_input, uint32_t n, float* predicted_out);
/* Introspection: cycle accounting of block 0 of the last tensor-core model launch made with option "profile_events" = 2 (an instrumented
 * build of the kernel) -- SM clock cycles; [0] MMA-issuing thread: total, [1] weight producer: waiting for a free weight stage, [2] issuer
 * waiting for the previous epilogue, [3] for a staged descriptor
 * chunk, [4] for a weight chunk to land; [8] worker thread 0: total, [9] waiting for a free descriptor stage, [10] for a GEMM to finish,
 * [11] inside the epilogues, [12] inside the descriptor staging; the rest 0. */
int ds_disney_model_profile(DsContext* ctx, uint64_t* cycles16);
/* Introspection of the FAST estimator's direction sampling (getNewDirection, CU/cloud.cuh:160-188): for n values of the uniform variate in
 * [0, 1), cos_theta_out = the closed-form inverse of the piecewise-linear chopped-Mie CDF that cloud.cuh:167-178 bisects (two-level guide +
 * four fixed probes, exactly the code k_trace_fast runs), and phase_out = the kernel's half-precision copy of the chopped phase sampler
 * (Mie.cpp:8206-8282) read at u = value with tex1D semantics.  Host buffers. */
int ds_invert_phase_cdf(DsContext* ctx, const float* values, uint32_t n, float* cos_theta_out, float* phase_out);
/* DisneyRenderer::render (DisneyRenderer.cpp:58-110): every 128 x 128 rectangle of the frame (x outer, y inner; clipped at the frame
 * edge) -> network-input launch, the model on the pixels that scattered, copyToFrameResult.  Rectangle k uses RNG stream
 * `stream + k` (clock() in the reference).  frame_result_out: float4 [frame_height][frame_width], zero where nothing scattered;
 * what ARenderer::render leaves in frameResultBuffer, ready for the progressive accumulation (ds_frame_* / Camera.cpp:195-199). */
int ds_render_disney(DsContext* ctx, const DsCamera* cam, uint32_t frame_width, uint32_t frame_height, uint32_t stream, float* frame_result_out);
/* Camera::render with DisneyRenderer registered as the ARenderer (Tasks.cpp:87, Camera.cpp:189-199): n subframes of the neural renderer,
 * each followed by updateFrameResult into the context's progressive / variance buffers (ds_frame_create); nothing leaves the device.
 * Subframe s draws its forced collisions from stream s * 4096 (+ rectangle ordinal). */
int ds_render_disney_subframes(DsContext* ctx, const DsCamera* cam, uint32_t first_subframe, uint32_t n);
void ds_radiance_settings_default(DsRadianceSettings* s);
/* RadianceCollector::init/update loop until all samples converge (RadianceCollector.cpp:19-54,73-141,176-192;
 * CU/pointEmissionCamera.cu:20-33; CU/PointRadianceTask.h).  tasks_out[i] is the merged representative of
 * sample i, converged_out[i] its flag; returns the number of update() rounds in *updates_out. */
int ds_point_radiance_run(DsContext* ctx, const float* positions, const float* directions, uint32_t n,
                          const DsRadianceSettings* settings, DsPointRadianceTask* tasks_out, uint8_t* converged_out,
                          uint32_t* updates_out);

/* ---------------------------------------------------------------- records (protobuf wire format, host only) */

/* Persistance::ScatterSample (DeepestScatter_Train/Protocols/ScatterSample.proto; ScatterSampleCollector.cpp:48-56).
 * Returns the encoded length (<= 64) or a negative error. */
int ds_record_scatter_sample(const float point[3], const float view_direction[3], uint8_t* out, size_t cap);
/* Persistance::DisneyDescriptor{bytes grid} (DisneyDescriptorCollector.cpp:76-96) */
int ds_record_disney_descriptor(const uint8_t* grid, size_t grid_len, uint8_t* out, size_t cap);
/* Persistance::Result{light_intensity, is_converged} (RadianceCollector.cpp:160-164) */
int ds_record_result(float light_intensity, int is_converged, uint8_t* out, size_t cap);
/* Persistance::SceneSetup (DG/ExecutionLoop/Tasks.cpp:77-85) */
int ds_record_scene_setup(const char* cloud_path, float cloud_size_m, const float light_direction[3], uint8_t* out, size_t cap);

/* ---------------------------------------------------------------- cloud importer front end (host side) */

/* Resources::loadVolumeBuffer (DG/Util/Resources.cpp:68-155) for what can be read without OpenVDB: `path` is a dense NumPy
 * array file (<file>.npy, C order, shape (nz, ny, nx), float32 / float64 / uint8) or "synth:<n>[:<kind>[:<seed>]]".  The
 * grid is cropped to the bounding box of its non-zero ("active") voxels expanded by one voxel (Resources.cpp:97-101),
 * quantised uint8(v / max * 255) on the device (Resources.cpp:137) and mip-mapped (Resources.cpp:169-209).  The last cloud
 * of a context is kept (Resources::volumeCache, Resources.cpp:22,73-78): loading the same path again is free.
 * size_out (may be NULL) receives the grid size in voxels.  Errors: DS_ERR_IO with ds_cloud_last_error(). */
int ds_cloud_load(DsContext* ctx, const char* path, int build_mips, int size_out[3]);
/* drop the importer's cache entry of a context (ds_context_destroy does it as well) */
void ds_cloud_forget(DsContext* ctx);
const char* ds_cloud_last_error(void);
/* the crop step alone, on the host: active bounding box + one voxel of zero padding.  dims_out = cropped size; when `out`
 * is not NULL it receives the cropped grid (out_capacity in floats). */
int ds_cloud_crop_active(const float* dense, int nx, int ny, int nz, float* out, size_t out_capacity, int dims_out[3], double* max_density_out);
/* The .vdb front end alone, on the host (Resources.cpp:80-141 up to the quantisation): first grid of an OpenVDB file as a FloatGrid
 * (host/VdbReader.hpp), maximum over its active values, active bounding box expanded by one voxel, accessor value of every voxel of
 * that box.  Same calling convention as ds_cloud_crop_active: dims_out always, `out` ([nz][ny][nx] floats) when not NULL. */
int ds_cloud_read_vdb(const char* path, float* out, size_t out_capacity, int dims_out[3], double* max_density_out);

/* Camera::saveToDisk (DG/Scene/Cameras/Camera.cpp:149-175), host only: the float4 [height][width] progressive buffer as a
 * single-part scanline OpenEXR file with FLOAT channels R, G, B, lineOrder DECREASING_Y, scanline y = buffer row y
 * (uncompressed; written without the OpenEXR library, deepestscatter_b200/host/ExrWriter.hpp).  Errors: DS_ERR_IO with
 * ds_cloud_last_error(). */
int ds_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgba);

/* ---------------------------------------------------------------- dataset store (LMDB data file, host only) */

/* DeepestScatter::Dataset (DG/Util/Dataset/Dataset.h:87-232, Dataset.cpp:8-18): one LMDB environment opened
 * MDB_NOSUBDIR, one MDB_INTEGERKEY sub-database per record type named after the protobuf message, key = int32 record
 * id (native little endian), value = proto3 bytes.  The file is written in LMDB 0.9's on-disk format by this library
 * itself (no liblmdb dependency; deepestscatter_b200/host/LmdbFile.hpp) and is what
 * DeepestScatter_Train/LmdbDataset.py opens.  An existing file is loaded and appended to (CollectMode::Continue,
 * DG/ExecutionLoop/Tasks.h:65-68). */
typedef struct DsDataset DsDataset;
int ds_dataset_open(const char* path, DsDataset** out);
/* commits and closes (Dataset::~Dataset, Dataset.cpp:20-36) */
int ds_dataset_close(DsDataset* ds);
/* message of the last failing call (ds may be NULL for open / close failures) */
const char* ds_dataset_last_error(DsDataset* ds);
/* mdb_txn_commit: write the B+tree pages and flip the meta page; records appended since the previous commit become
 * durable and visible to readers.  The reference commits once per batch (Dataset.h:203-232). */
int ds_dataset_commit(DsDataset* ds);
/* mdb_put of already encoded record bytes (Dataset::tryAppend, Dataset.h:170-200) */
int ds_dataset_put(DsDataset* ds, const char* table, int32_t id, const uint8_t* data, size_t n);
/* mdb_get (Dataset::getRecord, Dataset.h:96-118): returns the record length (copied to out when cap suffices) or a negative error */
long long ds_dataset_get(DsDataset* ds, const char* table, int32_t id, uint8_t* out, size_t cap);
/* mdb_stat ms_entries (Dataset::getRecordsCount, Dataset.h:80-93) */
long long ds_dataset_count(DsDataset* ds, const char* table);
/* mdb_drop(dbi, 0) (Dataset::dropTable, Dataset.h:120-152, without the console confirmation) */
int ds_dataset_drop(DsDataset* ds, const char* table);
/* copy every record of another dataset file into this one (merging per-GPU shards) */
int ds_dataset_merge(DsDataset* ds, const char* other_path);
/* SceneSetup record of scene `scene_id` (DG/ExecutionLoop/Tasks.cpp:77-85) */
int ds_dataset_append_scene_setup(DsDataset* ds, int32_t scene_id, const char* cloud_path, float cloud_size_m, const float light_direction[3]);
/* ScatterSampleCollector::collect record loop (ScatterSampleCollector.cpp:42-61): ids start_id .. start_id + n - 1 */
int ds_dataset_append_scatter_samples(DsDataset* ds, int32_t start_id, uint32_t n, const float* positions, const float* directions);
/* DisneyDescriptorCollector::recordToDataset (DisneyDescriptorCollector.cpp:76-103): n descriptors of descriptor_bytes (2250) each */
int ds_dataset_append_descriptors(DsDataset* ds, int32_t start_id, uint32_t n, const uint8_t* descriptors, size_t descriptor_bytes);
/* RadianceCollector::recordToDataset (RadianceCollector.cpp:148-169) */
int ds_dataset_append_results(DsDataset* ds, int32_t start_id, uint32_t n, const float* light_intensity, const uint8_t* is_converged);

#ifdef __cplusplus
}
#endif

#endif /* DS_ABI_H */
