/*
 * ref_api.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
 *
 * C entry points over the REFERENCE'S OWN classes, compiled unmodified from /root/reference by oracle/ref_shim/Makefile:
 *   Scene, Sun, VDBCloud, CloudMaterial, Camera, PathTracingRenderer, ScatterSampleCollector, DisneyDescriptorCollector,
 *   RadianceCollector, Resources, Mie (host side) driving the CUDA/ *.cu programs (device side) through the OptiX emulation of
 *   dsref_runtime.h.  Scene composition follows DG/installers.cpp:28-41,65-105 and DG/ExecutionLoop/Tasks.cpp:86-153
 *   (item order Sun, VDBCloud, CloudMaterial, Camera, collector; light normalised by installSceneSetup AND by
 *   DirectionalLight's constructor).
 * tests/test_oracle_vs_ref.py holds oracle/ds_oracle.cpp to the outputs of this library bit for bit.
 *
 * `#define private public` below only opens the reference's classes for READING results (buffers, task lists) and for
 * seeding Resources::volumeCache with an in-memory u8 grid; it changes no layout and no behaviour.
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include "dsref_runtime.h"
#include "dsref_gsl.h"

#define private public
#include "Mie.h"
#include "Scene/CloudMaterial.h"
#include "Scene/Cameras/Camera.h"
#include "Scene/Cameras/EmptyRenderer.h"
#include "Scene/Cameras/PathTracingRenderer.h"
#include "Scene/DisneyDescriptorCollector.h"
#include "Scene/RadianceCollector.h"
#include "Scene/ScatterSampleCollector.h"
#include "Scene/Scene.h"
#include "Scene/Sun.h"
#include "Scene/VDBCloud.h"
#include "Util/BufferBind.h"
#include "Util/Resources.h"
#undef private

#include "CUDA/rayData.cuh"
#include "ScatterSample.pb.h"
#include "Result.pb.h"
#include "DisneyDescriptor.pb.h"
#include "include/OpenEXR/ImfOutputFile.h"

using namespace DeepestScatter;

namespace {

/* the reference prints progress on std::cout; keep the test logs quiet */
struct QuietCout {
    std::streambuf* old;
    std::ostringstream sink;
    QuietCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~QuietCout() { std::cout.rdbuf(old); }
};

struct RefScene {
    std::shared_ptr<optix::Context> context;
    std::shared_ptr<Resources> resources;
    std::string volumePath;
    bool mipmaps = true;
    int width = 0, height = 0;

    std::shared_ptr<SceneDescription> description;
    std::shared_ptr<Sun> sun;
    std::shared_ptr<VDBCloud> cloud;
    std::shared_ptr<CloudMaterial> material;
    std::shared_ptr<ARenderer> renderer;
    std::shared_ptr<Camera> camera;
    std::shared_ptr<Dataset> dataset;
    std::shared_ptr<BatchSettings> batch;
    std::shared_ptr<ScatterSampleCollector> sampleCollector;
    std::shared_ptr<DisneyDescriptorCollector> descriptorCollector;
    std::shared_ptr<RadianceCollector> radianceCollector;
    std::shared_ptr<Scene> scene;

    void dropScene()
    {
        scene.reset();
        radianceCollector.reset();
        descriptorCollector.reset();
        sampleCollector.reset();
        camera.reset();
        renderer.reset();
        material.reset();
        cloud.reset();
        sun.reset();
        description.reset();
        if (context) {
            (*context)->destroy();
            context.reset();
        }
        resources.reset();
    }
    ~RefScene() { dropScene(); }
};

int gNextVolume = 0;
bool gSkipBake = false; /* ref_set_skip_bake: VDBCloud::disableRendering() before init, the caller supplies the baked volume */

void newContext(RefScene& s)
{
    s.dropScene();
    /* installFramework (installers.cpp:107-120) */
    s.context = std::make_shared<optix::Context>(optix::Context::create());
    s.resources = std::make_shared<Resources>(s.context);
}

enum Collector { NONE = 0, SAMPLES = 1, DESCRIPTORS = 2, RADIANCE = 3 };

/* installSceneSetup + installApp + (Tasks::collect | renderCloudSingleTask), then Scene::init */
void buildScene(RefScene& s, float cloudSizeM, float sampleStep, const float* light, int mode, int width, int height, int pathTracer,
                Collector collector, int batchStart, int batchSize)
{
    QuietCout quiet;
    newContext(s);
    /* installers.cpp:73-77: the light direction is normalised here, and again by DirectionalLight's constructor */
    const optix::float3 lightDirection = optix::normalize(optix::make_float3(light[0], light[1], light[2]));
    const Cloud::Rendering::Mode renderingMode = mode == 0   ? Cloud::Rendering::Mode::SunAndSkyAllScatter
                                                 : mode == 1 ? Cloud::Rendering::Mode::SunMultipleScatter
                                                             : Cloud::Rendering::Mode::SunSingleScatter;
    SceneDescription description{
        Cloud{Cloud::Rendering{Cloud::Rendering::SampleStep{sampleStep}, renderingMode},
              Cloud::Model{s.volumePath, s.mipmaps ? Cloud::Model::Mipmaps::On : Cloud::Model::Mipmaps::Off, Cloud::Model::Size{Meter{cloudSizeM}}}},
        DirectionalLight{lightDirection, Color{optix::make_float3(1, 1, 1)}, 1e6}};
    s.description = std::make_shared<SceneDescription>(description);
    s.width = width;
    s.height = height;

    s.sun = std::make_shared<Sun>(std::make_shared<DirectionalLight>(description.light), s.context);
    s.cloud = std::make_shared<VDBCloud>(std::make_shared<Cloud::Model>(description.cloud.model), s.context, s.resources);
    s.material = std::make_shared<CloudMaterial>(std::make_shared<Cloud::Rendering>(description.cloud.rendering), s.context, s.resources);
    if (gSkipBake) s.cloud->disableRendering(); /* what the two sample collectors do (ScatterSampleCollector.h:32): no inScatter launch */
    if (pathTracer)
        s.renderer = std::make_shared<PathTracingRenderer>(s.context, s.resources);
    else
        s.renderer = std::make_shared<EmptyRenderer>();
    s.camera = std::make_shared<Camera>(std::make_shared<Camera::Settings>((uint32_t)width, (uint32_t)height, std::filesystem::path("ref.exr")), s.context,
                                        s.resources, s.renderer);
    std::vector<std::shared_ptr<SceneItem>> items{s.sun, s.cloud, s.material, s.camera};
    if (collector != NONE) {
        if (!s.dataset) s.dataset = std::make_shared<Dataset>(std::make_shared<Dataset::Settings>("memory"));
        s.batch = std::make_shared<BatchSettings>(batchStart, batchSize);
    }
    if (collector == SAMPLES) {
        s.sampleCollector = std::make_shared<ScatterSampleCollector>(s.context, s.resources, s.dataset, s.batch, s.description, s.cloud);
        items.push_back(s.sampleCollector);
    } else if (collector == DESCRIPTORS) {
        s.descriptorCollector = std::make_shared<DisneyDescriptorCollector>(s.context, s.resources, s.dataset, s.batch, s.cloud);
        items.push_back(s.descriptorCollector);
    } else if (collector == RADIANCE) {
        s.radianceCollector = std::make_shared<RadianceCollector>(s.context, s.resources, s.dataset, s.batch);
        items.push_back(s.radianceCollector);
    }
    s.scene = std::make_shared<Scene>(items, s.context);
    s.scene->init();
}

void readBuffer(optix::Buffer b, unsigned level, void* out)
{
    RTsize w, h, d;
    b->getMipLevelSize(level, w, h, d);
    memcpy(out, b->map(level), w * h * d * b->getElementSize());
    b->unmap(level);
}

template <class T> void getContextFloats(RefScene& s, const char* name, T* out, size_t n)
{
    const dsref::VariableObj* v = (*s.context)->find(name);
    for (size_t i = 0; i < n; i++) out[i] = 0;
    if (v && v->kind == dsref::VariableObj::BYTES) memcpy(out, v->bytes.data(), std::min(n * sizeof(T), v->bytes.size()));
}

/* run fn(part, parts) in `procs` forked children; each child writes its share of a MAP_SHARED output */
void forkParts(int procs, const std::function<void(int, int)>& fn)
{
    if (procs <= 1) {
        fn(0, 1);
        return;
    }
    std::vector<pid_t> kids;
    for (int p = 0; p < procs; p++) {
        const pid_t pid = fork();
        if (pid == 0) {
            fn(p, procs);
            _exit(0);
        }
        if (pid > 0) kids.push_back(pid);
    }
    for (pid_t k : kids) {
        int status = 0;
        waitpid(k, &status, 0);
    }
}

} // namespace

extern "C" {

void* ref_create() { return new RefScene(); }
void ref_destroy(void* h) { delete (RefScene*)h; }

/* An in-memory u8 grid enters as Resources::volumeCache (Resources.cpp:73-78, 211-233): the mip chain is built by the
 * reference's Resources::generateMipmaps (:169-209). */
int ref_volume_set_u8(void* h, const uint8_t* data, int nx, int ny, int nz, int buildMipmaps)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    try {
        newContext(s);
        s.mipmaps = buildMipmaps != 0;
        s.volumePath = "memory://volume" + std::to_string(gNextVolume++);
        optix::Buffer buffer = (*s.context)->createBuffer(RT_BUFFER_INPUT, RT_FORMAT_UNSIGNED_BYTE);
        int levelCount = 1;
        if (buildMipmaps) {
            /* Resources.cpp:107-117 */
            size_t maxSize = std::max({(size_t)nx, (size_t)ny, (size_t)nz});
            while (maxSize /= 2) levelCount++;
        }
        buffer->setMipLevelCount(levelCount);
        buffer->setSize(nx, ny, nz);
        memcpy(buffer->map(0), data, (size_t)nx * ny * nz);
        buffer->unmap(0);
        if (buildMipmaps) s.resources->generateMipmaps(buffer);
        Resources::volumeCache = std::make_unique<Resources::VolumeCache>(s.volumePath, buffer, optix::make_float3((float)nx, (float)ny, (float)nz));
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_volume_set_u8: %s\n", e.what());
        return 1;
    }
}

/* A dense float grid goes through Resources::loadVolumeBuffer itself (Resources.cpp:68-155): active bounding box + 1, /max*255,
 * mip chain.  `path` names a DSDENSE1 container (include/openvdb/openvdb.h). */
int ref_volume_load(void* h, const char* path, int buildMipmaps)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    try {
        newContext(s);
        s.mipmaps = buildMipmaps != 0;
        s.volumePath = path;
        Resources::volumeCache.reset();
        s.resources->loadVolumeBuffer(path, buildMipmaps != 0); /* fills the process-wide cache the scene then reads */
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_volume_load: %s\n", e.what());
        return 1;
    }
}

int ref_volume_level_count() { return Resources::volumeCache ? (int)Resources::volumeCache->cache.size() : 0; }
void ref_volume_level_dims(int level, int* dims)
{
    const optix::size_t3 n = Resources::volumeCache->size;
    dims[0] = (int)std::max<size_t>(1, n.x >> level);
    dims[1] = (int)std::max<size_t>(1, n.y >> level);
    dims[2] = (int)std::max<size_t>(1, n.z >> level);
}
void ref_volume_level_get(int level, uint8_t* out)
{
    const std::vector<uint8_t>& v = Resources::volumeCache->cache.at((size_t)level);
    memcpy(out, v.data(), v.size());
}
void ref_volume_float_size(float* out)
{
    out[0] = Resources::volumeCache->floatSize.x;
    out[1] = Resources::volumeCache->floatSize.y;
    out[2] = Resources::volumeCache->floatSize.z;
}

/* collector: 0 none, 1 ScatterSampleCollector, 2 DisneyDescriptorCollector, 3 RadianceCollector (the latter two read their
 * ScatterSamples from the in-memory dataset: ref_dataset_put_samples first) */
int ref_scene_init(void* h, float cloudSizeM, float sampleStep, const float* lightDir, int mode, int width, int height, int pathTracer,
                   int collector, int batchStart, int batchSize)
{
    RefScene& s = *(RefScene*)h;
    try {
        buildScene(s, cloudSizeM, sampleStep, lightDir, mode, width, height, pathTracer, (Collector)collector, batchStart, batchSize);
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_scene_init: %s\n", e.what());
        return 1;
    }
}

/* bbox[3], texScale[3], mult, voxelM, voxelFP, light[3] -- the context variables VDBCloud.cpp:91-110 and Sun.cpp:15 set */
void ref_scene_get_derived(void* h, float* out)
{
    RefScene& s = *(RefScene*)h;
    getContextFloats(s, "bboxSize", out + 0, 3);
    getContextFloats(s, "textureScale", out + 3, 3);
    getContextFloats(s, "densityMultiplier", out + 6, 1);
    getContextFloats(s, "voxelSizeInMeters", out + 7, 1);
    getContextFloats(s, "voxelSizeInTermsOfFreePath", out + 8, 1);
    getContextFloats(s, "lightDirection", out + 9, 3);
}

/* the three 4096-entry sampler buffers Scene::init builds through Mie.cpp:8206-8297 */
void ref_get_mie(void* h, float* mie, float* chopped, float* integral)
{
    RefScene& s = *(RefScene*)h;
    const char* names[3] = {"mie", "choppedMie", "choppedMieIntegral"};
    float* outs[3] = {mie, chopped, integral};
    for (int i = 0; i < 3; i++) {
        const dsref::VariableObj* v = (*s.context)->find(names[i]);
        readBuffer(optix::Buffer(v->sampler->buffer), 0, outs[i]);
    }
}

/* Benchmarks only: skip the (single-threaded, minutes at 512^3) inScatter launch of the next ref_scene_init calls and install a
 * volume baked elsewhere.  tests/test_oracle_vs_ref.py shows the oracle's bake equal to the reference's byte for byte, which is
 * what makes the oracle's OpenMP bake a legitimate way to PREPARE this input outside any timed region. */
void ref_set_skip_bake(int on) { gSkipBake = on != 0; }

int ref_inscatter_set(void* h, const uint8_t* in)
{
    RefScene& s = *(RefScene*)h;
    try {
        RTsize nx, ny, nz;
        s.cloud->densityBuffer->getSize(nx, ny, nz);
        s.cloud->inScatterBuffer->setSize(nx, ny, nz);
        memcpy(s.cloud->inScatterBuffer->map(0), in, nx * ny * nz);
        s.cloud->inScatterBuffer->unmap(0);
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_inscatter_set: %s\n", e.what());
        return 1;
    }
}

void ref_inscatter_get(void* h, uint8_t* out)
{
    RefScene& s = *(RefScene*)h;
    readBuffer(s.cloud->inScatterBuffer, 0, out);
}

/* rtTrace of explicit rays against the scene (ray type 0 -> the closest-hit program CloudMaterial chose by mode), launch index
 * (seedVal0 >> 12, seedVal0 & 4095) so that launchID.x * 4096 + launchID.y == seedVal0 (cloudRadianceMaterials.cu:21) */
void ref_trace_paths(void* h, int n, const float* origins, const float* dirs, const uint32_t* seedVal0, const uint32_t* stream, float* radiance,
                     int procs)
{
    RefScene& s = *(RefScene*)h;
    float* shared = (float*)mmap(nullptr, (size_t)n * 3 * sizeof(float) + 64, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    (*s.context)->bindScene();
    forkParts(procs, [&](int part, int parts) {
        for (int i = part; i < n; i += parts) {
            dsref::setStreamOverride(true, stream[i], false);
            RadianceRayData prd;
            prd.result = optix::make_float3(0);
            prd.importance = 1;
            const optix::Ray ray(optix::make_float3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                                 optix::make_float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), RadianceRayData::rayId, 0.0f);
            (*s.context)->traceFrom(seedVal0[i] >> 12, seedVal0[i] & 4095u, ray, &prd, sizeof(prd));
            shared[3 * i + 0] = prd.result.x;
            shared[3 * i + 1] = prd.result.y;
            shared[3 * i + 2] = prd.result.z;
        }
    });
    dsref::setStreamOverride(false, 0, false);
    memcpy(radiance, shared, (size_t)n * 3 * sizeof(float));
    munmap(shared, (size_t)n * 3 * sizeof(float) + 64);
}

/* camera: Camera::init places the eye at (2.5, -0.4, 0) looking at the origin (Camera.cpp:37-42); tests may move it */
void ref_camera_set(void* h, const float* eye, const float* lookat, const float* up)
{
    RefScene& s = *(RefScene*)h;
    s.camera->cameraEye = optix::make_float3(eye[0], eye[1], eye[2]);
    s.camera->cameraLookat = optix::make_float3(lookat[0], lookat[1], lookat[2]);
    s.camera->cameraUp = optix::make_float3(up[0], up[1], up[2]);
    s.camera->updatePosition();
    QuietCout quiet;
    s.camera->reset();
}

/* eye, U, V, W as Camera::updatePosition (Camera.cpp:100-134) set them on the renderer's camera program */
void ref_camera_get(void* h, float* cam)
{
    RefScene& s = *(RefScene*)h;
    optix::Program p = s.renderer->getCamera();
    const char* names[4] = {"eye", "U", "V", "W"};
    for (int i = 0; i < 4; i++) {
        const dsref::VariableObj* v = p->find(names[i]);
        memcpy(cam + 3 * i, v->bytes.data(), 12);
    }
}

/* ARenderer::render (PathTracingRenderer.cpp:21-31) for one subframe: frameResultBuffer float4[h][w].  With procs > 1 the rows
 * are split over forked children (the launch is embarrassingly parallel; statics make the emulator single-threaded). */
void ref_render_frame_result(void* h, uint32_t subframeId, float* frameResult, int procs)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    (*s.context)["subframeId"]->setUint(subframeId);
    const size_t bytes = (size_t)s.width * s.height * 16;
    if (procs <= 1) {
        s.renderer->render(s.camera->frameResultBuffer);
        readBuffer(s.camera->frameResultBuffer, 0, frameResult);
        return;
    }
    float* shared = (float*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    optix::Program camera = s.renderer->getCamera();
    camera["frameResultBuffer"]->setBuffer(s.camera->frameResultBuffer);
    (*s.context)->setRayGenerationProgram(0, camera);
    forkParts(procs, [&](int part, int parts) {
        /* interleaved rows balance the cloud-covered part of the frame */
        for (int y = part; y < s.height; y += parts) {
            (*s.context)->launchRect(s.width, s.height, 0, s.width, y, y + 1);
            const float* row = (const float*)s.camera->frameResultBuffer->map(0) + (size_t)y * s.width * 4;
            memcpy(shared + (size_t)y * s.width * 4, row, (size_t)s.width * 16);
        }
    });
    memcpy(frameResult, shared, bytes);
    memcpy(s.camera->frameResultBuffer->map(0), shared, bytes);
    munmap(shared, bytes);
}

/* Camera::update (Camera.cpp:68-75, 177-229) `updates` times: 10 subframes each, Welford accumulation, Reinhard passes */
int ref_camera_update(void* h, int updates)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    try {
        s.camera->completed = false; /* renderCloudSingleTask, Tasks.cpp:98-99 */
        for (int i = 0; i < updates; i++) s.camera->update();
        return (int)s.camera->subframeId;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_camera_update: %s\n", e.what());
        return -1;
    }
}

void ref_frame_get(void* h, float* progressive, float* variance, uint8_t* screen, float* frameResult)
{
    RefScene& s = *(RefScene*)h;
    if (progressive) readBuffer(s.camera->progressiveBuffer, 0, progressive);
    if (variance) readBuffer(s.camera->varianceBuffer, 0, variance);
    if (screen) readBuffer(s.camera->screenBuffer, 0, screen);
    if (frameResult) readBuffer(s.camera->frameResultBuffer, 0, frameResult);
}

float ref_average_luminance(void* h)
{
    RefScene& s = *(RefScene*)h;
    float v = 0;
    readBuffer(s.camera->reinhardAverageLuminance, 0, &v);
    return v;
}

int ref_camera_is_converged(void* h)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    return s.camera->isConverged() ? 1 : 0;
}

/* progressive.cu:17-27 and reinhard.cu:26-83 on caller-supplied buffers (so the oracle's and the product's inputs can be fed) */
void ref_update_frame_result(void* h, const float* frameResult, float* progressive, float* variance, uint32_t subframeId)
{
    RefScene& s = *(RefScene*)h;
    const size_t bytes = (size_t)s.width * s.height * 16;
    memcpy(s.camera->frameResultBuffer->map(0), frameResult, bytes);
    memcpy(s.camera->progressiveBuffer->map(0), progressive, bytes);
    memcpy(s.camera->varianceBuffer->map(0), variance, bytes);
    (*s.context)["subframeId"]->setUint(subframeId);
    (*s.context)->setRayGenerationProgram(0, s.camera->updateFrameResult);
    (*s.context)->launch(0, s.width, s.height);
    readBuffer(s.camera->progressiveBuffer, 0, progressive);
    readBuffer(s.camera->varianceBuffer, 0, variance);
}

float ref_tonemap(void* h, const float* progressive, float exposure, uint8_t* screen)
{
    RefScene& s = *(RefScene*)h;
    memcpy(s.camera->progressiveBuffer->map(0), progressive, (size_t)s.width * s.height * 16);
    /* Camera.cpp:202-210 */
    (*s.context)->setRayGenerationProgram(0, s.camera->reinhardFirstPass);
    (*s.context)->launch(0, s.width, 1);
    (*s.context)->setRayGenerationProgram(0, s.camera->reinhardSecondPass);
    (*s.context)->launch(0, 1, 1);
    (*s.context)["exposure"]->setFloat(exposure);
    (*s.context)->setRayGenerationProgram(0, s.camera->reinhardLastPass);
    (*s.context)->launch(0, s.width, s.height);
    readBuffer(s.camera->screenBuffer, 0, screen);
    return ref_average_luminance(h);
}

/* what Camera::saveToDisk handed to OpenEXR (Camera.cpp:149-175): rows in file order */
int ref_last_exr(float* rgb, int* width, int* height, int* decreasingY)
{
    const dsref::ExrImage& img = dsref::lastExr();
    *width = img.width;
    *height = img.height;
    *decreasingY = img.decreasingY ? 1 : 0;
    if (rgb && !img.rgb.empty()) memcpy(rgb, img.rgb.data(), img.rgb.size() * sizeof(float));
    return img.rgb.empty() ? 0 : 1;
}

/* ScatterSampleCollector::update (ScatterSampleCollector.cpp:23-62) with launch indices 0..batchSize-1.  clock() is `stream`
 * in the ray-generation program and the 1-based attempt number in the closest-hit program (oracle/ds_oracle.cpp,
 * orc_generate_points). */
int ref_generate_points(void* h, uint32_t stream, float* positions, float* directions)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    try {
        dsref::setStreamOverride(true, stream, true);
        s.sampleCollector->update();
        dsref::setStreamOverride(false, 0, false);
        const int n = s.batch->batchSize;
        for (int i = 0; i < n; i++) {
            const Persistance::ScatterSample r = s.dataset->getRecord<Persistance::ScatterSample>(s.batch->batchStartId + i);
            positions[3 * i + 0] = r.point().x();
            positions[3 * i + 1] = r.point().y();
            positions[3 * i + 2] = r.point().z();
            directions[3 * i + 0] = r.view_direction().x();
            directions[3 * i + 1] = r.view_direction().y();
            directions[3 * i + 2] = r.view_direction().z();
        }
        return 0;
    } catch (const std::exception& e) {
        dsref::setStreamOverride(false, 0, false);
        fprintf(stderr, "ref_generate_points: %s\n", e.what());
        return 1;
    }
}

void ref_dataset_put_samples(void* h, int startId, int n, const float* positions, const float* directions)
{
    RefScene& s = *(RefScene*)h;
    if (!s.dataset) s.dataset = std::make_shared<Dataset>(std::make_shared<Dataset::Settings>("memory"));
    std::vector<Persistance::ScatterSample> samples((size_t)n);
    for (int i = 0; i < n; i++) {
        samples[i].mutable_point()->set_x(positions[3 * i]);
        samples[i].mutable_point()->set_y(positions[3 * i + 1]);
        samples[i].mutable_point()->set_z(positions[3 * i + 2]);
        samples[i].mutable_view_direction()->set_x(directions[3 * i]);
        samples[i].mutable_view_direction()->set_y(directions[3 * i + 1]);
        samples[i].mutable_view_direction()->set_z(directions[3 * i + 2]);
    }
    s.dataset->batchAppend(gsl::make_span(samples), startId);
}

/* DisneyDescriptor records written by DisneyDescriptorCollector::init (DisneyDescriptorCollector.cpp:13-53,76-103): 2250 B each */
int ref_dataset_get_descriptors(void* h, int startId, int n, uint8_t* out)
{
    RefScene& s = *(RefScene*)h;
    try {
        for (int i = 0; i < n; i++) {
            const Persistance::DisneyDescriptor r = s.dataset->getRecord<Persistance::DisneyDescriptor>(startId + i);
            if (r.grid().size() != 2250) return 2;
            memcpy(out + (size_t)i * 2250, r.grid().data(), 2250);
        }
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_dataset_get_descriptors: %s\n", e.what());
        return 1;
    }
}

/* RadianceCollector::update (RadianceCollector.cpp:73-141) `updates` times (100 launches of <= 20480 threads each).
 * Returns the converged count.  tasksOut[id] = the last merged representative of each sample (the converged record once it
 * has converged), threadsOut (optional, 20480 x 40 B) = the task buffer after the last update's reschedule. */
int ref_radiance_update(void* h, int updates, void* tasksOut, uint8_t* converged, void* threadsOut, int* recorded)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    try {
        RadianceCollector& rc = *s.radianceCollector;
        const int n = s.batch->batchSize;
        for (int u = 0; u < updates && !rc.isCompleted(); u++) rc.update();
        Gpu::PointRadianceTask* out = (Gpu::PointRadianceTask*)tasksOut;
        for (int i = 0; i < n; i++) converged[i] = 0;
        for (const Gpu::PointRadianceTask& t : rc.convergedTasks) {
            out[t.id] = t;
            converged[t.id] = 1;
        }
        if (!rc.isCompleted()) {
            /* unconverged representatives sit at i * taskRepeatCount after scheduleTasks (:176-192) */
            BufferBind<Gpu::PointRadianceTask> bind(rc.tasksBuffer);
            const uint32_t remaining = (uint32_t)rc.getRemainingCount();
            for (uint32_t i = 0; i < remaining; i++) {
                const Gpu::PointRadianceTask& t = bind[i * rc.taskRepeatCount];
                out[t.id] = t;
            }
        }
        if (threadsOut) readBuffer(rc.tasksBuffer, 0, threadsOut);
        if (recorded) *recorded = (int)s.dataset->getRecordsCount<Persistance::Result>();
        return rc.getConvergedCount();
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_radiance_update: %s\n", e.what());
        return -1;
    }
}

int ref_dataset_get_results(void* h, int startId, int n, float* lightIntensity, uint8_t* isConverged)
{
    RefScene& s = *(RefScene*)h;
    try {
        for (int i = 0; i < n; i++) {
            const Persistance::Result r = s.dataset->getRecord<Persistance::Result>(startId + i);
            lightIntensity[i] = r.light_intensity();
            isConverged[i] = r.is_converged() ? 1 : 0;
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

/* Launch 0 of DisneyRenderer::renderRect (DisneyRenderer.cpp:84-88): disneyCamera.cu pinholeCamera over a rectangle with
 * sampleDisneyDescriptor as closest hit.  input [rectH][rectW][10][226], info [rectH][rectW][5] (radiance rgb,
 * transmittance, hasScattered).  The buffers are cleared first to the values the oracle reports for untouched pixels. */
int ref_network_input(void* h, int rectX, int rectY, int rectW, int rectH, uint32_t stream, float* input, float* info)
{
    RefScene& s = *(RefScene*)h;
    QuietCout quiet;
    try {
        optix::Context ctx = *s.context;
        optix::Program camera = s.resources->loadProgram("disneyCamera.cu", "pinholeCamera");
        optix::Buffer networkInput = ctx->createBuffer(RT_BUFFER_INPUT_OUTPUT, RT_FORMAT_USER, rectW, rectH);
        networkInput->setElementSize(sizeof(Gpu::DisneyNetworkInput));
        optix::Buffer direct = ctx->createBuffer(RT_BUFFER_INPUT_OUTPUT, RT_FORMAT_USER, rectW, rectH);
        direct->setElementSize(sizeof(IntersectionInfo));
        IntersectionInfo* di = (IntersectionInfo*)direct->map();
        for (int i = 0; i < rectW * rectH; i++) {
            di[i].radiance = optix::make_float3(0);
            di[i].transmittance = 1.0f;
            di[i].hasScattered = false;
        }
        camera["networkInputBuffer"]->setBuffer(networkInput);
        camera["directRadianceBuffer"]->setBuffer(direct);
        camera["frameResultBuffer"]->setBuffer(s.camera->frameResultBuffer);
        camera["rectOrigin"]->setUint((unsigned)rectX, (unsigned)rectY);
        /* eye, U, V, W: Camera::updatePosition would set them on this program; copy them from the path tracer's */
        optix::Program pt = s.renderer->getCamera();
        const char* names[4] = {"eye", "U", "V", "W"};
        for (int i = 0; i < 4; i++) {
            float v[3];
            memcpy(v, pt->find(names[i])->bytes.data(), 12);
            camera[names[i]]->setFloat(v[0], v[1], v[2]);
        }
        ctx["subframeId"]->setUint(stream);
        ctx->setRayGenerationProgram(0, camera);
        ctx->launch(0, rectW, rectH);
        const Gpu::DisneyNetworkInput* ni = (const Gpu::DisneyNetworkInput*)networkInput->map();
        static_assert(sizeof(Gpu::DisneyNetworkInput) == 2260 * 4, "DisneyNetworkInput layout");
        memcpy(input, ni, (size_t)rectW * rectH * 2260 * 4);
        for (int i = 0; i < rectW * rectH; i++) {
            info[5 * i + 0] = di[i].radiance.x;
            info[5 * i + 1] = di[i].radiance.y;
            info[5 * i + 2] = di[i].radiance.z;
            info[5 * i + 3] = di[i].transmittance;
            info[5 * i + 4] = di[i].hasScattered ? 1.0f : 0.0f;
        }
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_network_input: %s\n", e.what());
        return 1;
    }
}

/* work counters of the emulator: rtTrace calls, rtTex3D-by-id fetches (march steps), bound tex3D fetches (scatter events) */
void ref_counters_get(unsigned long long* out)
{
    out[0] = dsref::counters().traces;
    out[1] = dsref::counters().boundTaps;
    out[2] = dsref::counters().bindlessTaps;
}
void ref_counters_reset() { dsref::counters() = dsref::Counters(); }

} /* extern "C" */
