/*
 * DataGen.hpp -- headless host side of the DataGen subsystem (C++17) on top of the C ABI (include/ds_abi.h).
 *
 * Keeps the reference's class shapes and call order -- SceneItem (DG/Scene/SceneItem.h:5-17), Scene (Scene.cpp:36-62),
 * Sun / VDBCloud / CloudMaterial / Camera as scene items (installers.cpp:27-38), ARenderer + PathTracingRenderer
 * (ARenderer.h:8-16, PathTracingRenderer.cpp:21-31), the three dataset collectors (ScatterSampleCollector.cpp,
 * DisneyDescriptorCollector.cpp, RadianceCollector.cpp), Tasks (ExecutionLoop/Tasks.cpp) and an execution loop without
 * GLUT (GuiExecutionLoop.cpp:84-125) -- but every optix::Context variable, buffer and launch is one call into
 * libdeepestscatter_b200.so.  The Hypodermic container is replaced by a plain struct of shared pointers.
 */
#pragma once

#include <cmath>
#include <cstdio>
#include <csignal>
#include <functional>
#include <fstream>
#include <iostream>
#include <memory>
#include <queue>
#include <string>
#include <vector>

#include "../../include/ds_abi.h"
#include "CloudImporter.hpp"
#include "Dataset.hpp"
#include "ExrWriter.hpp"

namespace DeepestScatter {

inline void dsCheck(DsContext* ctx, int rc)
{
    if (rc != DS_OK) throw std::runtime_error(ds_last_error(ctx));
}

/* ---- DG/Scene/SceneDescription.h ---- */
using Meter = float;
struct Color {
    float r, g, b;
};
struct DirectionalLight {
    float direction[3];
    Color color;
    float intensity;
};
struct Cloud {
    struct Rendering {
        enum class Mode { SunAndSkyAllScatter = 0, SunMultipleScatter = 1, SunSingleScatter = 2 };
        float sampleStep;
        Mode mode;
    };
    struct Model {
        enum class Mipmaps : bool { Off = false, On = true };
        std::string vdbPath;
        Mipmaps mipmapsOn;
        Meter size;
        Meter meanFreePath = 10.0f; /* SceneDescription.h:80 */
    };
    Rendering rendering;
    Model model;
};
struct SceneDescription {
    Cloud cloud;
    DirectionalLight light;
};

/* one context per process; the reference creates one optix::Context per task (GuiExecutionLoop.cpp:93-97) -- here the
 * context outlives tasks so that the imported cloud stays on the device (Resources::volumeCache) */
struct Device {
    explicit Device(int index)
    {
        const int rc = ds_context_create(index, &ctx);
        if (rc != DS_OK) throw std::runtime_error(ds_last_error(nullptr));
        importer.reset(new CloudImporter(ctx));
    }
    ~Device() { ds_context_destroy(ctx); }
    Device(const Device&) = delete;
    DsContext* ctx = nullptr;
    std::unique_ptr<CloudImporter> importer;
};

/* ---- DG/Scene/SceneItem.h ---- */
class SceneItem {
public:
    virtual ~SceneItem() = default;
    virtual void init() = 0;
    virtual void reset() = 0;
    virtual void update() = 0;
    virtual bool isCompleted() { return true; }
};

/* Sun (Sun.cpp:13-18), VDBCloud (VDBCloud.cpp:16-21, 88-117) and CloudMaterial (CloudMaterial.cpp:14, 23) set OptiX
 * variables; here they fill the one parameter block and the cloud item imports, sets and bakes. */
class VDBCloud : public SceneItem {
public:
    VDBCloud(std::shared_ptr<Device> device, const SceneDescription& scene) : device(std::move(device)), scene(scene) {}
    void init() override
    {
        int size[3];
        std::cout << "Loading cloud... " << scene.cloud.model.vdbPath << std::endl;
        device->importer->load(scene.cloud.model.vdbPath, scene.cloud.model.mipmapsOn == Cloud::Model::Mipmaps::On, size); /* InitVolume */
        DsSceneParams p;
        ds_scene_params_default(&p);
        p.cloud_size_m = scene.cloud.model.size;
        p.mean_free_path_m = scene.cloud.model.meanFreePath;
        p.sample_step = scene.cloud.rendering.sampleStep;
        for (int i = 0; i < 3; i++) p.light_direction[i] = scene.light.direction[i];
        p.light_color[0] = scene.light.color.r;
        p.light_color[1] = scene.light.color.g;
        p.light_color[2] = scene.light.color.b;
        p.light_intensity = scene.light.intensity;
        dsCheck(device->ctx, ds_scene_set(device->ctx, &p));                 /* setupVariables */
        dsCheck(device->ctx, ds_bake_sun_transmittance(device->ctx));       /* InitInScatter */
    }
    void reset() override {}
    void update() override {}

private:
    std::shared_ptr<Device> device;
    SceneDescription scene;
};

/* ---- DG/Scene/Cameras/ARenderer.h ---- */
class ARenderer {
public:
    virtual ~ARenderer() = default;
    virtual void init() = 0;
    /* leaves `count` new samples per pixel, subframes first .. first + count - 1, accumulated into the progressive
     * and variance buffers (render + updateFrameResult of Camera::render, Camera.cpp:189-199) */
    virtual void render(const DsCamera& camera, uint32_t firstSubframe, uint32_t count) = 0;
};

/* PathTracingRenderer.cpp:21-31: pinhole camera + rtTrace of radiance rays */
class PathTracingRenderer : public ARenderer {
public:
    PathTracingRenderer(std::shared_ptr<Device> device, Cloud::Rendering::Mode mode) : device(std::move(device)), mode(mode) {}
    void init() override {}
    void render(const DsCamera& camera, uint32_t firstSubframe, uint32_t count) override
    {
        dsCheck(device->ctx, ds_render_subframes(device->ctx, &camera, (DsMode)mode, firstSubframe, count));
    }

private:
    std::shared_ptr<Device> device;
    Cloud::Rendering::Mode mode;
};

/* DisneyRenderer.cpp:17-110: the neural renderer.  init() loads the model -- here the flat float32 export of DisneyModel.state_dict()
 * (deepestscatter_b200/disney_model.py) instead of the TorchScript file of DisneyRenderer.cpp:19-22 -- and render() runs the network-input
 * launch, the model and copyToFrameResult for the whole frame, then the accumulation, on the device */
class DisneyRenderer : public ARenderer {
public:
    static constexpr const char* NAME = "Disney";
    DisneyRenderer(std::shared_ptr<Device> device, std::string modelPath) : device(std::move(device)), modelPath(std::move(modelPath)) {}
    void init() override
    {
        std::vector<float> w(ds_disney_model_weight_count());
        std::ifstream f(modelPath, std::ios::binary);
        if (!f.read(reinterpret_cast<char*>(w.data()), (std::streamsize)(w.size() * sizeof(float))))
            throw std::runtime_error("cannot read " + std::to_string(w.size()) + " float32 weights from " + modelPath);
        dsCheck(device->ctx, ds_disney_model_load(device->ctx, w.data(), w.size()));
    }
    void render(const DsCamera& camera, uint32_t firstSubframe, uint32_t count) override
    {
        dsCheck(device->ctx, ds_render_disney_subframes(device->ctx, &camera, firstSubframe, count));
    }

private:
    std::shared_ptr<Device> device;
    std::string modelPath;
};

/* EmptyRenderer (dataset tasks register it: Tasks.cpp:139): renders nothing */
class EmptyRenderer : public ARenderer {
public:
    void init() override {}
    void render(const DsCamera&, uint32_t, uint32_t) override {}
};

/* ---- DG/Scene/Cameras/Camera.cpp ---- */
class Camera : public SceneItem {
public:
    struct Settings {
        uint32_t width, height;
        std::string outputFile;     /* linear image: scanline EXR, R G B float, DECREASING_Y (Camera.cpp:149-175) */
        uint32_t maxSubframes = 0;  /* 0: until converged (Camera.cpp:232-268) */
        uint32_t subframesPerUpdate = 10; /* Camera.cpp:186 */
    };
    Camera(std::shared_ptr<Device> device, std::shared_ptr<ARenderer> renderer, const Settings& settings)
        : device(std::move(device)), renderer(std::move(renderer)), settings(settings)
    {
    }

    void init() override
    {
        renderer->init();
        dsCheck(device->ctx, ds_frame_create(device->ctx, (int)settings.width, (int)settings.height));
        ds_camera_default((int)settings.width, (int)settings.height, &camera); /* Camera.cpp:37-39, 102 */
        reset();
    }
    void reset() override
    {
        subframeId = 0;
        dsCheck(device->ctx, ds_frame_clear(device->ctx));
    }
    void update() override
    {
        if (!isCompleted()) render();
    }
    bool isCompleted() override { return completed; }
    void lookAt(const float eye[3], const float lookat[3], const float up[3])
    {
        ds_camera_look_at(eye, lookat, up, 30.0f, (float)settings.width / (float)settings.height, &camera);
        reset();
    }
    uint32_t subframes() const { return subframeId; }
    float exposure = 0.4f; /* Camera.h:90 */
    bool completed = false;

    void saveToDisk() const
    {
        if (settings.outputFile.empty()) return;
        const size_t pixels = (size_t)settings.width * settings.height;
        std::vector<float> rgba(pixels * 4);
        dsCheck(device->ctx, ds_frame_download(device->ctx, rgba.data(), nullptr));
        std::cout << rgba[4 * (pixels / 2 + settings.width / 2)] << std::endl; /* Camera.cpp:161 */
        writeExrRGB(settings.outputFile, settings.width, settings.height, rgba.data(), /*decreasingY=*/true); /* Camera.cpp:154-174 */
        FILE* f;
        /* the display image (reinhard.cu) next to it */
        std::vector<uint8_t> screen(pixels * 4);
        dsCheck(device->ctx, ds_tonemap(device->ctx, exposure, screen.data(), nullptr));
        const std::string ppm = settings.outputFile + ".ppm";
        f = fopen(ppm.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write " + ppm);
        fprintf(f, "P6\n%u %u\n255\n", settings.width, settings.height);
        for (uint32_t y = settings.height; y-- > 0;) /* PPM rows run top to bottom */
            for (uint32_t x = 0; x < settings.width; x++) fwrite(&screen[4 * ((size_t)y * settings.width + x)], 1, 3, f);
        fclose(f);
    }

private:
    void render()
    {
        if (!isConverged()) {
            uint32_t n = settings.subframesPerUpdate;
            if (settings.maxSubframes) n = std::min(n, settings.maxSubframes - subframeId);
            renderer->render(camera, subframeId + 1, n);
            subframeId += n;
            if (subframeId % 40 == 0) saveToDisk(); /* Camera.cpp:211 */
        } else {
            completed = true;
            saveToDisk();
            std::cout << "rendering subframe " << subframeId << std::endl;
        }
    }
    bool isConverged()
    {
        if (settings.maxSubframes) return subframeId >= settings.maxSubframes;
        if (subframeId < 100) return false;
        uint32_t left = 0;
        dsCheck(device->ctx, ds_frame_unconverged(device->ctx, subframeId, &left));
        std::cout << "Converged: " << (size_t)settings.width * settings.height - left << "/" << (size_t)settings.width * settings.height << " --- " << left
                  << "left" << std::endl;
        return left < 500; /* Camera.cpp:267 */
    }

    std::shared_ptr<Device> device;
    std::shared_ptr<ARenderer> renderer;
    Settings settings;
    DsCamera camera{};
    uint32_t subframeId = 0;
};

/* ---- dataset collectors ---- */

inline void readBatchSamples(Dataset& dataset, const BatchSettings& settings, std::vector<float>& positions, std::vector<float>& directions)
{
    positions.resize(3 * (size_t)settings.batchSize);
    directions.resize(3 * (size_t)settings.batchSize);
    for (int32_t i = 0; i < settings.batchSize; i++) {
        const auto sample = dataset.getRecord<Persistance::ScatterSample>(settings.batchStartId + i);
        positions[3 * i] = sample.point.x;
        positions[3 * i + 1] = sample.point.y;
        positions[3 * i + 2] = sample.point.z;
        directions[3 * i] = sample.view_direction.x;
        directions[3 * i + 1] = sample.view_direction.y;
        directions[3 * i + 2] = sample.view_direction.z;
    }
}

/* ScatterSampleCollector.cpp:24-62 */
class ScatterSampleCollector : public SceneItem {
public:
    ScatterSampleCollector(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset, BatchSettings settings)
        : device(std::move(device)), dataset(std::move(dataset)), settings(settings)
    {
    }
    void init() override {}
    void reset() override {}
    void update() override
    {
        if (done) return;
        std::cout << "Generating samples..." << std::endl;
        std::vector<float> positions(3 * (size_t)settings.batchSize), directions(3 * (size_t)settings.batchSize);
        /* RNG stream = scene id: the reference seeds with clock() (random.cuh:38) */
        dsCheck(device->ctx, ds_generate_points(device->ctx, 0, (uint32_t)settings.batchSize, (uint32_t)(settings.batchStartId / std::max(1, settings.batchSize)),
                                                positions.data(), directions.data()));
        std::vector<Persistance::ScatterSample> samples(settings.batchSize);
        for (int32_t i = 0; i < settings.batchSize; i++) {
            samples[i].point = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
            samples[i].view_direction = {directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]};
        }
        std::cout << "Writing samples..." << std::endl;
        dataset->batchAppend(samples, settings.batchStartId);
        done = true;
    }
    bool isCompleted() override { return done; }

private:
    std::shared_ptr<Device> device;
    std::shared_ptr<Dataset> dataset;
    BatchSettings settings;
    bool done = false;
};

/* DisneyDescriptorCollector.cpp:13-103 */
class DisneyDescriptorCollector : public SceneItem {
public:
    static constexpr size_t DESCRIPTOR_BYTES = 10 * 9 * 5 * 5; /* CU/DisneyDescriptor.h:8-55 */
    DisneyDescriptorCollector(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset, BatchSettings settings)
        : device(std::move(device)), dataset(std::move(dataset)), settings(settings)
    {
    }
    void init() override
    {
        std::vector<float> positions, directions;
        readBatchSamples(*dataset, settings, positions, directions);
        std::vector<uint8_t> descriptors(DESCRIPTOR_BYTES * (size_t)settings.batchSize);
        dsCheck(device->ctx, ds_collect_descriptors(device->ctx, positions.data(), directions.data(), (uint32_t)settings.batchSize, descriptors.data()));
        std::vector<Persistance::DisneyDescriptor> serialized(settings.batchSize);
        for (int32_t i = 0; i < settings.batchSize; i++)
            serialized[i].grid.assign(descriptors.begin() + DESCRIPTOR_BYTES * (size_t)i, descriptors.begin() + DESCRIPTOR_BYTES * (size_t)(i + 1));
        std::cout << "Writing descriptors..." << std::endl;
        dataset->batchAppend(serialized, settings.batchStartId);
    }
    void reset() override {}
    void update() override {}

private:
    std::shared_ptr<Device> device;
    std::shared_ptr<Dataset> dataset;
    BatchSettings settings;
};

/* RadianceCollector.cpp:19-193: the update loop (100 launches, merge, convergence test, reschedule) runs inside
 * ds_point_radiance_run */
class RadianceCollector : public SceneItem {
public:
    RadianceCollector(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset, BatchSettings settings, DsRadianceSettings radiance)
        : device(std::move(device)), dataset(std::move(dataset)), settings(settings), radiance(radiance)
    {
    }
    void init() override { readBatchSamples(*dataset, settings, positions, directions); }
    void reset() override {}
    void update() override
    {
        if (allPixelsConverged) return;
        std::vector<DsPointRadianceTask> tasks(settings.batchSize);
        std::vector<uint8_t> converged(settings.batchSize);
        uint32_t updates = 0;
        dsCheck(device->ctx, ds_point_radiance_run(device->ctx, positions.data(), directions.data(), (uint32_t)settings.batchSize, &radiance, tasks.data(),
                                                   converged.data(), &updates));
        std::vector<Persistance::Result> results(settings.batchSize);
        std::cout << "Serializing emissions..." << std::endl;
        /* The reference records only once EVERY task converged and then writes is_converged = true (RadianceCollector.cpp:130-136,
         * 163).  With the max_updates safety cap a sample may stop unconverged: it is recorded as such, never as converged. */
        size_t unconverged = 0;
        for (int32_t i = 0; i < settings.batchSize; i++) {
            results[i].light_intensity = tasks[i].radiance;
            results[i].is_converged = converged[i] != 0;
            unconverged += converged[i] == 0;
        }
        if (unconverged) std::cout << "WARNING: " << unconverged << " samples hit the update cap before converging" << std::endl;
        std::cout << "Writing emissions... (" << updates << " updates)" << std::endl;
        dataset->batchAppend(results, settings.batchStartId);
        allPixelsConverged = true; /* the batch is finished (all converged, or the cap ended it) */
    }
    bool isCompleted() override { return allPixelsConverged; }

private:
    std::shared_ptr<Device> device;
    std::shared_ptr<Dataset> dataset;
    BatchSettings settings;
    DsRadianceSettings radiance;
    std::vector<float> positions, directions;
    bool allPixelsConverged = false;
};

/* ---- DG/Scene/Scene.cpp ---- */
class Scene {
public:
    explicit Scene(std::vector<std::shared_ptr<SceneItem>> sceneItems) : sceneItems(std::move(sceneItems)) {}
    void init()
    {
        for (const auto& item : sceneItems) item->init();
    }
    void update()
    {
        for (const auto& item : sceneItems) item->update();
    }
    bool isCompleted()
    {
        for (const auto& item : sceneItems)
            if (!item->isCompleted()) return false;
        return true;
    }

private:
    std::vector<std::shared_ptr<SceneItem>> sceneItems;
};

/* ---- execution loop without GLUT (GuiExecutionLoop.cpp:84-125) ---- */
class ExecutionLoop {
public:
    using LazyTask = std::function<std::shared_ptr<Scene>()>;
    /* afterTask runs when a task (one scene = one batch of records) is complete: the driver commits the dataset there, which is
     * the reference's one-LMDB-transaction-per-batch (Dataset.h:203-232) and what makes CollectMode::Continue resume after a crash.
     * stopRequested() (set by SIGINT / SIGTERM in datagen.cpp) ends the loop between tasks. */
    void run(std::queue<LazyTask> tasks, const std::function<void()>& afterTask = nullptr)
    {
        while (!tasks.empty() && !stopRequested()) {
            std::shared_ptr<Scene> scene = tasks.front()();
            tasks.pop();
            scene->init();
            do {
                scene->update();
            } while (!scene->isCompleted());
            if (afterTask) afterTask();
        }
    }
    static volatile std::sig_atomic_t& stopFlag()
    {
        static volatile std::sig_atomic_t flag = 0;
        return flag;
    }
    static bool stopRequested() { return stopFlag() != 0; }
};

/* ---- DG/ExecutionLoop/Tasks.cpp ---- */
enum class LightDirection { Front, Back, Side };

inline const char* toString(LightDirection d) { return d == LightDirection::Front ? "Front" : d == LightDirection::Back ? "Back" : "Side"; }

inline void getLightDirection(LightDirection direction, float out[3])
{
    /* Tasks.cpp:52-66 */
    static const float dirs[3][3] = {{-0.586f, -0.766f, -0.271f}, {0.586f, -0.766f, -0.271f}, {-0.03f, -0.25f, 0.8f}};
    const float* d = dirs[direction == LightDirection::Front ? 0 : direction == LightDirection::Back ? 1 : 2];
    out[0] = d[0];
    out[1] = d[1];
    out[2] = d[2];
}

/* installSceneSetup (installers.cpp:66-105) */
inline SceneDescription makeSceneDescription(const Persistance::SceneSetup& setup, const std::string& cloudsRoot, Cloud::Rendering::Mode mode,
                                             Cloud::Model::Mipmaps mipmaps)
{
    std::string cloudPath = setup.cloud_path;
    const bool isSpec = cloudPath.rfind("synth:", 0) == 0;
    if (!isSpec && !cloudsRoot.empty() && cloudsRoot != "." && !cloudPath.empty() && cloudPath[0] != '/') cloudPath = cloudsRoot + "/" + cloudPath;
    const float lx = setup.light_direction.x, ly = setup.light_direction.y, lz = setup.light_direction.z;
    const float inv = 1.0f / std::sqrt(lx * lx + ly * ly + lz * lz); /* optix::normalize */
    SceneDescription s{Cloud{Cloud::Rendering{1.0f / 512.f, mode}, Cloud::Model{cloudPath, mipmaps, setup.cloud_size_m}},
                       DirectionalLight{{lx * inv, ly * inv, lz * inv}, Color{1, 1, 1}, 1e6f}};
    return s;
}

class Tasks {
public:
    enum class CollectMode { Reset, Continue };

    struct RenderSettings {
        uint32_t width = 512u, height = 256u; /* Tasks.cpp:49-50 */
        uint32_t maxSubframes = 0;
        Cloud::Rendering::Mode mode = Cloud::Rendering::Mode::SunAndSkyAllScatter;
        std::string outputDir = ".";
        std::string disneyModel; /* non-empty: `using TRenderer = DisneyRenderer` (Tasks.cpp:86) with this weight file */
    };

    /* renderCloudSingleTask (Tasks.cpp:68-106) with the path-tracing renderer */
    static ExecutionLoop::LazyTask renderCloudSingleTask(std::shared_ptr<Device> device, const std::string& cloudPath, float sizeM, LightDirection light,
                                                         const RenderSettings& rs)
    {
        return [=]() {
            Persistance::SceneSetup setup;
            setup.cloud_path = cloudPath;
            setup.cloud_size_m = sizeM;
            float d[3];
            getLightDirection(light, d);
            setup.light_direction = {d[0], d[1], d[2]};
            const SceneDescription scene = makeSceneDescription(setup, ".", rs.mode, Cloud::Model::Mipmaps::On);
            std::string base = cloudPath;
            const size_t slash = base.find_last_of('/');
            if (slash != std::string::npos) base = base.substr(slash + 1);
            for (char& c : base)
                if (c == ':') c = '_';
            const size_t dot = base.find_last_of('.');
            if (dot != std::string::npos) base = base.substr(0, dot);
            const bool neural = !rs.disneyModel.empty();
            Camera::Settings cs{rs.width, rs.height, rs.outputDir + "/" + base + "." + toString(light) + "." + (neural ? DisneyRenderer::NAME : "PathTracing") + ".exr",
                                rs.maxSubframes};
            std::shared_ptr<ARenderer> renderer;
            if (neural)
                renderer = std::make_shared<DisneyRenderer>(device, rs.disneyModel);
            else
                renderer = std::make_shared<PathTracingRenderer>(device, rs.mode);
            std::vector<std::shared_ptr<SceneItem>> items{std::make_shared<VDBCloud>(device, scene), std::make_shared<Camera>(device, renderer, cs)};
            return std::make_shared<Scene>(items);
        };
    }

    /* Tasks::renderCloud (Tasks.cpp:108-116): Side, then Back */
    static std::queue<ExecutionLoop::LazyTask> renderCloud(std::shared_ptr<Device> device, const std::string& cloudPath, float sizeM, const RenderSettings& rs)
    {
        std::queue<ExecutionLoop::LazyTask> tasks;
        tasks.push(renderCloudSingleTask(device, cloudPath, sizeM, LightDirection::Side, rs));
        tasks.push(renderCloudSingleTask(device, cloudPath, sizeM, LightDirection::Back, rs));
        return tasks;
    }

    struct CollectSettings {
        int32_t batchSize = 2048;    /* Tasks.cpp:137 */
        int shard = 0, shards = 1;   /* scenes with sceneId % shards == shard */
        DsRadianceSettings radiance; /* RadianceCollector constants */
        CollectSettings() { ds_radiance_settings_default(&radiance); }
    };

    /* Tasks::collect<T> (Tasks.h:43-71, Tasks.cpp:118-153); T is the record type being collected */
    template <class T>
    static std::queue<ExecutionLoop::LazyTask> collect(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset, const std::string& cloudRoot,
                                                       CollectMode mode, const CollectSettings& cs)
    {
        /* Reset drops the table; Continue resumes after the scenes already collected.  The reference resumes at
         * count / 2048 (Tasks.h:62-68), which assumes one contiguous writer; with scene sharding the test is per scene:
         * a scene is done when the last record of its batch exists (same answer in the contiguous case). */
        if (mode == CollectMode::Reset) dataset->dropTable<T>();
        std::queue<ExecutionLoop::LazyTask> tasks;
        const size_t sceneCount = dataset->getRecordsCount<Persistance::SceneSetup>();
        std::vector<uint8_t> probe;
        for (int32_t i = 0; i < (int32_t)sceneCount; i++) {
            if (cs.shards > 1 && i % cs.shards != cs.shard) continue;
            if (mode == CollectMode::Continue && dataset->lmdb().get(T::name(), (uint32_t)((i + 1) * cs.batchSize - 1), probe)) continue;
            tasks.push([=]() {
                const auto setup = dataset->getRecord<Persistance::SceneSetup>(i);
                const SceneDescription scene = makeSceneDescription(setup, cloudRoot, Cloud::Rendering::Mode::SunMultipleScatter, Cloud::Model::Mipmaps::On);
                const BatchSettings batch(i * cs.batchSize, cs.batchSize);
                std::vector<std::shared_ptr<SceneItem>> items{std::make_shared<VDBCloud>(device, scene), makeCollector<T>(device, dataset, batch, cs)};
                return std::make_shared<Scene>(items);
            });
        }
        return tasks;
    }

private:
    template <class T>
    static std::shared_ptr<SceneItem> makeCollector(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset, BatchSettings batch, const CollectSettings& cs);
};

template <>
inline std::shared_ptr<SceneItem> Tasks::makeCollector<Persistance::ScatterSample>(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset,
                                                                                  BatchSettings batch, const CollectSettings&)
{
    return std::make_shared<ScatterSampleCollector>(device, dataset, batch);
}
template <>
inline std::shared_ptr<SceneItem> Tasks::makeCollector<Persistance::DisneyDescriptor>(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset,
                                                                                     BatchSettings batch, const CollectSettings&)
{
    return std::make_shared<DisneyDescriptorCollector>(device, dataset, batch);
}
template <>
inline std::shared_ptr<SceneItem> Tasks::makeCollector<Persistance::Result>(std::shared_ptr<Device> device, std::shared_ptr<Dataset> dataset, BatchSettings batch,
                                                                           const CollectSettings& cs)
{
    return std::make_shared<RadianceCollector>(device, dataset, batch, cs.radiance);
}

} // namespace DeepestScatter
