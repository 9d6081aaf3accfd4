/*
 * datagen -- headless DataGen driver (replaces DG/main.cpp + GuiExecutionLoop): renders clouds and collects the Deep
 * Scattering dataset through libdeepestscatter_b200.so.  See host/DataGen.hpp for the class layer it drives.
 *
 *   datagen render <cloud> [--size M] [--width W --height H] [--spp N] [--mode all|multi|single] [--out DIR]
 *                  [--renderer pathtracing|disney --model WEIGHTS.f32]   (the reference's `using TRenderer = ...`, Tasks.cpp:86)
 *       Tasks::renderCloud (Tasks.cpp:108-116): sun "Side", then "Back"; linear image as PFM + tone-mapped PPM
 *   datagen render <cloud> --gpus N --spp S [--light Side|Back|Front] [--chunk 64] ...
 *       one frame of S subframes split over N GPUs (one host thread + context each), per-GPU accumulation buffers combined by one
 *       ncclReduce inside the library (ds_frame_reduce); prints render / reduce seconds of the slowest rank
 *   datagen scenes <db> --clouds a.npy,b.npy,... [--scenes-per-cloud 30] [--seed 566]
 *       DeepestScatter_Train/Utils/GenerateSceneSetups.py: SceneSetup records (size log-uniform 1..12 km, sun uniform on the sphere)
 *   datagen collect <db> --what samples|descriptors|results|all [--cloud-root DIR] [--mode continue|reset]
 *                   [--batch-size 2048] [--shard r/R] [--device D] [--max-threads N] [--launches N] [--opt name=value,...]
 *       Tasks::collect<T> (Tasks.h:43-71): one task per scene; --shard writes only scenes with id % R == r
 *   datagen merge <out-db> <shard-db>...      copy the records of per-GPU shards into one dataset
 *   datagen stat <db>                         record counts per table
 * <cloud> is a dense .npy grid or synth:<n>[:<kind>[:<seed>]] (host/CloudImporter.hpp).
 */
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <algorithm>
#include <chrono>
#include <sstream>
#include <thread>

#include "DataGen.hpp"

using namespace DeepestScatter;

namespace {

[[noreturn]] void usage(const char* argv0)
{
    std::cerr << "Usage: " << argv0 << " render|scenes|collect|merge|stat ... (see the header of host/datagen.cpp)\n";
    exit(1);
}

struct Args {
    std::vector<std::string> positional;
    std::map<std::string, std::string> opt;
    std::string get(const std::string& k, const std::string& d) const
    {
        const auto it = opt.find(k);
        return it == opt.end() ? d : it->second;
    }
    long num(const std::string& k, long d) const
    {
        const auto it = opt.find(k);
        return it == opt.end() ? d : atol(it->second.c_str());
    }
};

Args parse(int argc, char** argv, int first)
{
    Args a;
    for (int i = first; i < argc; i++) {
        if (strncmp(argv[i], "--", 2) == 0) {
            if (i + 1 >= argc) throw std::runtime_error(std::string("option ") + argv[i] + " needs a value");
            a.opt[argv[i] + 2] = argv[i + 1];
            i++;
        } else {
            a.positional.push_back(argv[i]);
        }
    }
    return a;
}

std::vector<std::string> split(const std::string& s, char sep)
{
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, sep))
        if (!item.empty()) out.push_back(item);
    return out;
}

/*
 * render --gpus N: one host thread and one context per GPU.  Every GPU loads the cloud, bakes its own sun-transmittance volume
 * (deterministic, so the replicas are identical) and renders a contiguous share of the frame's subframe ids
 * [offset + 1, offset + count] (option stream_offset keeps the per-(pixel, subframe) RNG streams those of a single-GPU run);
 * the per-GPU accumulation buffers are then combined by ONE ncclReduce inside the library (ds_frame_reduce) and GPU 0 writes
 * the EXR.  Prints the wall time of the slowest rank for the render and for the reduce.
 */
int cmdRenderMultiGpu(const Args& a, int gpus)
{
    const uint32_t total = (uint32_t)a.num("spp", 1024);
    const uint32_t width = (uint32_t)a.num("width", 512), height = (uint32_t)a.num("height", 256); /* Tasks.cpp:49-50 */
    const float sizeM = (float)atof(a.get("size", "7000").c_str());
    const std::string modeName = a.get("mode", "all"), lightName = a.get("light", "Side"), cloudPath = a.positional[0];
    const Cloud::Rendering::Mode mode = modeName == "single" ? Cloud::Rendering::Mode::SunSingleScatter
                                        : modeName == "multi" ? Cloud::Rendering::Mode::SunMultipleScatter
                                                              : Cloud::Rendering::Mode::SunAndSkyAllScatter;
    const LightDirection light = lightName == "Front" ? LightDirection::Front : lightName == "Back" ? LightDirection::Back : LightDirection::Side;
    const uint32_t chunk = (uint32_t)a.num("chunk", 64);
    uint8_t id[DS_COMM_ID_BYTES];
    if (ds_comm_unique_id(id) != DS_OK) throw std::runtime_error(ds_last_error(nullptr));
    std::vector<std::string> errors((size_t)gpus);
    std::vector<double> renderSeconds((size_t)gpus, 0.0), reduceSeconds((size_t)gpus, 0.0);
    std::vector<std::thread> threads;
    for (int rank = 0; rank < gpus; rank++)
        threads.emplace_back([&, rank] {
            try {
                auto device = std::make_shared<Device>(rank);
                for (const std::string& kv : split(a.get("opt", ""), ',')) {
                    const size_t eq = kv.find('=');
                    if (eq == std::string::npos) throw std::runtime_error("bad --opt " + kv);
                    dsCheck(device->ctx, ds_set_option(device->ctx, kv.substr(0, eq).c_str(), atoi(kv.c_str() + eq + 1)));
                }
                Persistance::SceneSetup setup;
                setup.cloud_path = cloudPath;
                setup.cloud_size_m = sizeM;
                float d[3];
                getLightDirection(light, d);
                setup.light_direction = {d[0], d[1], d[2]};
                VDBCloud cloud(device, makeSceneDescription(setup, ".", mode, Cloud::Model::Mipmaps::On));
                cloud.init();
                dsCheck(device->ctx, ds_frame_create(device->ctx, (int)width, (int)height));
                dsCheck(device->ctx, ds_comm_init(device->ctx, gpus, rank, id));
                DsCamera camera;
                ds_camera_default((int)width, (int)height, &camera);
                /* contiguous split of the subframe ids 1..total */
                const uint32_t base = total / (uint32_t)gpus, rem = total % (uint32_t)gpus;
                const uint32_t count = base + ((uint32_t)rank < rem ? 1u : 0u);
                const uint32_t offset = (uint32_t)rank * base + std::min((uint32_t)rank, rem);
                dsCheck(device->ctx, ds_set_option(device->ctx, "stream_offset", (int)offset));
                dsCheck(device->ctx, ds_set_option(device->ctx, "staging_subframes", (int)std::max(1u, std::min(chunk, count))));
                dsCheck(device->ctx, ds_sync(device->ctx));
                const auto t0 = std::chrono::steady_clock::now();
                for (uint32_t done = 0; done < count;) {
                    const uint32_t n = std::min(chunk, count - done);
                    dsCheck(device->ctx, ds_render_subframes(device->ctx, &camera, (DsMode)mode, done + 1, n));
                    done += n;
                }
                dsCheck(device->ctx, ds_sync(device->ctx));
                const auto t1 = std::chrono::steady_clock::now();
                dsCheck(device->ctx, ds_frame_reduce(device->ctx, count, total, 0));
                dsCheck(device->ctx, ds_sync(device->ctx));
                const auto t2 = std::chrono::steady_clock::now();
                renderSeconds[(size_t)rank] = std::chrono::duration<double>(t1 - t0).count();
                reduceSeconds[(size_t)rank] = std::chrono::duration<double>(t2 - t1).count();
                if (rank == 0) {
                    std::vector<float> rgba((size_t)width * height * 4);
                    dsCheck(device->ctx, ds_frame_download(device->ctx, rgba.data(), nullptr));
                    const std::string out = a.get("out", ".") + "/multigpu." + toString(light) + ".PathTracing.exr";
                    if (ds_write_exr(out.c_str(), width, height, rgba.data()) != DS_OK) throw std::runtime_error("cannot write " + out);
                    double mean = 0;
                    for (size_t i = 0; i < rgba.size(); i += 4) mean += rgba[i];
                    std::cout << "wrote " << out << " mean radiance " << mean / ((double)width * height) << std::endl;
                }
                dsCheck(device->ctx, ds_comm_destroy(device->ctx));
            } catch (const std::exception& e) {
                errors[(size_t)rank] = e.what();
            }
        });
    for (std::thread& t : threads) t.join();
    for (int rank = 0; rank < gpus; rank++)
        if (!errors[(size_t)rank].empty()) throw std::runtime_error("GPU " + std::to_string(rank) + ": " + errors[(size_t)rank]);
    const double render = *std::max_element(renderSeconds.begin(), renderSeconds.end());
    const double reduce = *std::max_element(reduceSeconds.begin(), reduceSeconds.end());
    std::cout << "{\"gpus\": " << gpus << ", \"width\": " << width << ", \"height\": " << height << ", \"spp\": " << total << ", \"render_s\": " << render
              << ", \"reduce_s\": " << reduce << ", \"mpaths_s\": " << (double)width * height * total / (render + reduce) / 1e6 << "}" << std::endl;
    return 0;
}

int cmdRender(const Args& a)
{
    if (a.positional.empty()) throw std::runtime_error("render needs a cloud");
    const int gpus = (int)a.num("gpus", 1);
    if (gpus > 1) return cmdRenderMultiGpu(a, gpus);
    auto device = std::make_shared<Device>((int)a.num("device", 0));
    Tasks::RenderSettings rs;
    rs.width = (uint32_t)a.num("width", rs.width);
    rs.height = (uint32_t)a.num("height", rs.height);
    rs.maxSubframes = (uint32_t)a.num("spp", 0);
    rs.outputDir = a.get("out", ".");
    if (a.get("renderer", "pathtracing") == "disney") {
        rs.disneyModel = a.get("model", "");
        if (rs.disneyModel.empty()) throw std::runtime_error("--renderer disney needs --model <flat float32 state_dict>");
    }
    const std::string mode = a.get("mode", "all");
    rs.mode = mode == "single" ? Cloud::Rendering::Mode::SunSingleScatter : mode == "multi" ? Cloud::Rendering::Mode::SunMultipleScatter
                                                                                           : Cloud::Rendering::Mode::SunAndSkyAllScatter;
    ExecutionLoop loop;
    loop.run(Tasks::renderCloud(device, a.positional[0], (float)atof(a.get("size", "7000").c_str()), rs)); /* main.cpp:63 */
    return 0;
}

int cmdScenes(const Args& a)
{
    if (a.positional.empty()) throw std::runtime_error("scenes needs a dataset path");
    const std::vector<std::string> clouds = split(a.get("clouds", ""), ',');
    if (clouds.empty()) throw std::runtime_error("--clouds a,b,... is required");
    const long perCloud = a.num("scenes-per-cloud", 30);
    std::mt19937_64 rng((uint64_t)a.num("seed", 566));
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    Dataset dataset{Dataset::Settings(a.positional[0])};
    for (const std::string& cloud : clouds)
        for (long i = 0; i < perCloud; i++) {
            Persistance::SceneSetup s;
            s.cloud_path = cloud;
            s.cloud_size_m = (float)std::exp(std::log(1000.0) + uni(rng) * (std::log(12000.0) - std::log(1000.0)));
            const double cosTheta = -1.0 + 2.0 * uni(rng), phi = uni(rng) * 2.0 * 3.14159265358979323846;
            const double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
            s.light_direction = {(float)(std::cos(phi) * sinTheta), (float)(std::sin(phi) * sinTheta), (float)cosTheta};
            dataset.append(s);
        }
    dataset.commit();
    std::cout << dataset.getRecordsCount<Persistance::SceneSetup>() << " scene setups in " << a.positional[0] << std::endl;
    return 0;
}

int cmdCollect(const Args& a)
{
    if (a.positional.empty()) throw std::runtime_error("collect needs a dataset path");
    auto device = std::make_shared<Device>((int)a.num("device", 0));
    auto dataset = std::make_shared<Dataset>(Dataset::Settings(a.positional[0]));
    Tasks::CollectSettings cs;
    cs.batchSize = (int32_t)a.num("batch-size", cs.batchSize);
    const std::string shard = a.get("shard", "0/1");
    if (sscanf(shard.c_str(), "%d/%d", &cs.shard, &cs.shards) != 2 || cs.shards < 1 || cs.shard < 0 || cs.shard >= cs.shards) throw std::runtime_error("bad --shard r/R");
    cs.radiance.max_thread_count = (uint32_t)a.num("max-threads", cs.radiance.max_thread_count);
    cs.radiance.launches_per_update = (uint32_t)a.num("launches", cs.radiance.launches_per_update);
    for (const std::string& kv : split(a.get("opt", ""), ',')) { /* library tuning options, name=value[,name=value...] */
        const size_t eq = kv.find('=');
        if (eq == std::string::npos) throw std::runtime_error("bad --opt " + kv);
        dsCheck(device->ctx, ds_set_option(device->ctx, kv.substr(0, eq).c_str(), atoi(kv.c_str() + eq + 1)));
    }
    const Tasks::CollectMode mode = a.get("mode", "continue") == "reset" ? Tasks::CollectMode::Reset : Tasks::CollectMode::Continue;
    const std::string root = a.get("cloud-root", ".");
    const std::string what = a.get("what", "all");
    ExecutionLoop loop;
    /* One commit per finished batch, like the reference's transaction per batchAppend: a kill, OOM or power cut loses at most
     * the batch in flight, and `--mode continue` resumes at count / batchSize (Tasks.h:65-68).  LmdbFile commits incrementally
     * (only the tail of an appended table is rewritten), so this costs megabytes per batch.  SIGINT / SIGTERM finish the
     * current batch, commit and exit. */
    std::signal(SIGINT, [](int) { ExecutionLoop::stopFlag() = 1; });
    std::signal(SIGTERM, [](int) { ExecutionLoop::stopFlag() = 1; });
    const auto commitBatch = [&] { dataset->commit(); };
    /* the reference runs one collector type per program run, in this order (main.cpp:61, Tasks.cpp:155-178) */
    if (what == "samples" || what == "all") loop.run(Tasks::collect<Persistance::ScatterSample>(device, dataset, root, mode, cs), commitBatch);
    if (what == "descriptors" || what == "all") loop.run(Tasks::collect<Persistance::DisneyDescriptor>(device, dataset, root, mode, cs), commitBatch);
    if (what == "results" || what == "all") loop.run(Tasks::collect<Persistance::Result>(device, dataset, root, mode, cs), commitBatch);
    dataset->commit();
    if (ExecutionLoop::stopRequested()) {
        std::cerr << "interrupted: every finished batch is committed; rerun with --mode continue" << std::endl;
        return 130;
    }
    return 0;
}

int cmdMerge(const Args& a)
{
    if (a.positional.size() < 2) throw std::runtime_error("merge needs an output and at least one shard");
    Dataset out{Dataset::Settings(a.positional[0])};
    for (size_t i = 1; i < a.positional.size(); i++) {
        Dataset in{Dataset::Settings(a.positional[i], /*create=*/false, /*readonly=*/true)}; /* a mistyped shard path is an error */
        out.mergeFrom(in);
    }
    out.commit();
    return 0;
}

int cmdStat(const Args& a)
{
    if (a.positional.empty()) throw std::runtime_error("stat needs a dataset path");
    dslmdb::LmdbFile f(a.positional[0], /*create=*/false, /*readonly=*/true);
    std::cout << "txn " << f.txnid() << ", " << f.lastPage() + 1 << " pages of " << f.pageSize() << " bytes\n";
    for (const auto& t : f.tables()) std::cout << t.first << " " << t.second.size() << "\n";
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 2) usage(argv[0]);
    try {
        const std::string cmd = argv[1];
        const Args a = parse(argc, argv, 2);
        if (cmd == "render") return cmdRender(a);
        if (cmd == "scenes") return cmdScenes(a);
        if (cmd == "collect") return cmdCollect(a);
        if (cmd == "merge") return cmdMerge(a);
        if (cmd == "stat") return cmdStat(a);
        usage(argv[0]);
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl; /* main.cpp:67-71 */
        return 1;
    }
}
