/*
 * datagen -- headless DataGen driver (replaces DG/main.cpp + GuiExecutionLoop): renders clouds and collects the Deep
 * Scattering dataset through libdeepestscatter_b200.so.  See host/DataGen.hpp for the class layer it drives.
 *
 *   datagen render <cloud> [--size M] [--width W --height H] [--spp N] [--mode all|multi|single] [--out DIR]
 *                  [--renderer pathtracing|disney --model WEIGHTS.f32]   (the reference's `using TRenderer = ...`, Tasks.cpp:86)
 *       Tasks::renderCloud (Tasks.cpp:108-116): sun "Side", then "Back"; linear image as PFM + tone-mapped PPM
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
#include <sstream>

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

int cmdRender(const Args& a)
{
    if (a.positional.empty()) throw std::runtime_error("render needs a cloud");
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
