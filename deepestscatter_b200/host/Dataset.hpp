/*
 * Dataset.hpp -- record store of the dataset generator (host side, C++17, no third-party dependency).
 *
 * Mirrors the interface of the reference's DeepestScatter::Dataset (DG/Util/Dataset/Dataset.h:87-232,
 * Dataset.cpp): tables named after the protobuf message (`T::descriptor()->name()`), keys are int32 record ids,
 * values are the proto3 wire bytes.  The reference keeps them in one LMDB environment; liblmdb does not exist in
 * this environment, so the store is an append-only record log (`*.dsrec`) with the same table / key / value
 * contract, resumable (`CollectMode::Continue`, Tasks.h:65-68) and mergeable by concatenation (one shard per GPU).
 *
 * Record messages (DeepestScatter_Train/Protocols/*.proto) are plain structs whose serialize() goes through the
 * C-ABI encoders (ds_record_*), i.e. the exact bytes the reference's generated protobuf code writes.
 */
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ds_abi.h"

namespace Persistance {

struct Vector3 {
    float x = 0, y = 0, z = 0;
};

/* minimal proto3 reader for the four messages */
struct WireReader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    uint64_t varint()
    {
        uint64_t v = 0;
        int shift = 0;
        while (p < end) {
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false;
        return 0;
    }
    float fixed32()
    {
        float f = 0;
        if (end - p < 4) {
            ok = false;
            return 0;
        }
        memcpy(&f, p, 4);
        p += 4;
        return f;
    }
    void skip(int wireType)
    {
        if (wireType == 0)
            varint();
        else if (wireType == 5)
            p += 4;
        else if (wireType == 1)
            p += 8;
        else if (wireType == 2)
            p += varint();
        else
            ok = false;
        if (p > end) ok = false;
    }
};

inline Vector3 parseVector3(const uint8_t* data, size_t n)
{
    Vector3 v;
    WireReader r{data, data + n};
    while (r.ok && r.p < r.end) {
        const uint64_t tag = r.varint();
        const int field = (int)(tag >> 3), wt = (int)(tag & 7);
        if (wt == 5 && field >= 1 && field <= 3) {
            const float f = r.fixed32();
            (field == 1 ? v.x : field == 2 ? v.y : v.z) = f;
        } else {
            r.skip(wt);
        }
    }
    if (!r.ok) throw std::runtime_error("malformed Vector3 record");
    return v;
}

struct SceneSetup {
    static const char* name() { return "SceneSetup"; }
    std::string cloud_path;
    float cloud_size_m = 0;
    Vector3 light_direction;
    std::vector<uint8_t> serialize() const
    {
        std::vector<uint8_t> out(cloud_path.size() + 64);
        const float l[3] = {light_direction.x, light_direction.y, light_direction.z};
        const int n = ds_record_scene_setup(cloud_path.c_str(), cloud_size_m, l, out.data(), out.size());
        if (n < 0) throw std::runtime_error("ds_record_scene_setup failed");
        out.resize(n);
        return out;
    }
    static SceneSetup parse(const uint8_t* data, size_t n)
    {
        SceneSetup s;
        WireReader r{data, data + n};
        while (r.ok && r.p < r.end) {
            const uint64_t tag = r.varint();
            const int field = (int)(tag >> 3), wt = (int)(tag & 7);
            if (field == 1 && wt == 2) {
                const uint64_t len = r.varint();
                if ((uint64_t)(r.end - r.p) < len) throw std::runtime_error("malformed SceneSetup record");
                s.cloud_path.assign((const char*)r.p, len);
                r.p += len;
            } else if (field == 2 && wt == 5) {
                s.cloud_size_m = r.fixed32();
            } else if (field == 3 && wt == 2) {
                const uint64_t len = r.varint();
                if ((uint64_t)(r.end - r.p) < len) throw std::runtime_error("malformed SceneSetup record");
                s.light_direction = parseVector3(r.p, len);
                r.p += len;
            } else {
                r.skip(wt);
            }
        }
        if (!r.ok) throw std::runtime_error("malformed SceneSetup record");
        return s;
    }
};

struct ScatterSample {
    static const char* name() { return "ScatterSample"; }
    Vector3 point, view_direction;
    std::vector<uint8_t> serialize() const
    {
        std::vector<uint8_t> out(64);
        const float p[3] = {point.x, point.y, point.z}, d[3] = {view_direction.x, view_direction.y, view_direction.z};
        const int n = ds_record_scatter_sample(p, d, out.data(), out.size());
        if (n < 0) throw std::runtime_error("ds_record_scatter_sample failed");
        out.resize(n);
        return out;
    }
    static ScatterSample parse(const uint8_t* data, size_t n)
    {
        ScatterSample s;
        WireReader r{data, data + n};
        while (r.ok && r.p < r.end) {
            const uint64_t tag = r.varint();
            const int field = (int)(tag >> 3), wt = (int)(tag & 7);
            if ((field == 2 || field == 3) && wt == 2) {
                const uint64_t len = r.varint();
                if ((uint64_t)(r.end - r.p) < len) throw std::runtime_error("malformed ScatterSample record");
                (field == 2 ? s.point : s.view_direction) = parseVector3(r.p, len);
                r.p += len;
            } else {
                r.skip(wt);
            }
        }
        if (!r.ok) throw std::runtime_error("malformed ScatterSample record");
        return s;
    }
};

struct DisneyDescriptor {
    static const char* name() { return "DisneyDescriptor"; }
    std::vector<uint8_t> grid; /* 10 * 9 * 5 * 5 bytes */
    std::vector<uint8_t> serialize() const
    {
        std::vector<uint8_t> out(grid.size() + 16);
        const int n = ds_record_disney_descriptor(grid.data(), grid.size(), out.data(), out.size());
        if (n < 0) throw std::runtime_error("ds_record_disney_descriptor failed");
        out.resize(n);
        return out;
    }
    static DisneyDescriptor parse(const uint8_t* data, size_t n)
    {
        DisneyDescriptor d;
        WireReader r{data, data + n};
        while (r.ok && r.p < r.end) {
            const uint64_t tag = r.varint();
            const int field = (int)(tag >> 3), wt = (int)(tag & 7);
            if (field == 1 && wt == 2) {
                const uint64_t len = r.varint();
                if ((uint64_t)(r.end - r.p) < len) throw std::runtime_error("malformed DisneyDescriptor record");
                d.grid.assign(r.p, r.p + len);
                r.p += len;
            } else {
                r.skip(wt);
            }
        }
        if (!r.ok) throw std::runtime_error("malformed DisneyDescriptor record");
        return d;
    }
};

struct Result {
    static const char* name() { return "Result"; }
    float light_intensity = 0;
    bool is_converged = false;
    std::vector<uint8_t> serialize() const
    {
        std::vector<uint8_t> out(16);
        const int n = ds_record_result(light_intensity, is_converged ? 1 : 0, out.data(), out.size());
        if (n < 0) throw std::runtime_error("ds_record_result failed");
        out.resize(n);
        return out;
    }
    static Result parse(const uint8_t* data, size_t n)
    {
        Result res;
        WireReader r{data, data + n};
        while (r.ok && r.p < r.end) {
            const uint64_t tag = r.varint();
            const int field = (int)(tag >> 3), wt = (int)(tag & 7);
            if (field == 1 && wt == 5)
                res.light_intensity = r.fixed32();
            else if (field == 2 && wt == 0)
                res.is_converged = r.varint() != 0;
            else
                r.skip(wt);
        }
        if (!r.ok) throw std::runtime_error("malformed Result record");
        return res;
    }
};

} // namespace Persistance

namespace DeepestScatter {

/* Dataset (DG/Util/Dataset/Dataset.h): tables by message name, int32 keys, proto3 values. */
class Dataset {
public:
    struct Settings {
        explicit Settings(std::string path) : path(std::move(path)) {}
        std::string path;
    };
    using TableName = std::string;

    explicit Dataset(const Settings& settings) : path(settings.path)
    {
        /* "Opening Dataset..." (Dataset.cpp:10): load the index of an existing log */
        FILE* f = fopen(path.c_str(), "rb");
        if (f) {
            load(f);
            fclose(f);
        }
        file = fopen(path.c_str(), "ab");
        if (!file) throw std::runtime_error("cannot open dataset " + path);
    }
    ~Dataset()
    {
        if (file) fclose(file);
    }
    Dataset(const Dataset&) = delete;
    Dataset& operator=(const Dataset&) = delete;

    template <class T>
    size_t getRecordsCount()
    {
        return tables[T::name()].size();
    }

    template <class T>
    T getRecord(int32_t recordId)
    {
        const auto& table = tables[T::name()];
        const auto it = table.find(recordId);
        if (it == table.end()) throw std::runtime_error(std::string("MDB_NOTFOUND: no record ") + std::to_string(recordId) + " in table " + T::name());
        return T::parse(it->second.data(), it->second.size());
    }

    /* raw bytes, for tests and the LMDB exporter */
    const std::map<TableName, std::map<int32_t, std::vector<uint8_t>>>& allTables() const { return tables; }

    template <class T>
    void dropTable()
    {
        put(T::name(), DROP_KEY, nullptr, 0);
        tables[T::name()].clear();
        nextIds[T::name()] = 0;
        fflush(file);
    }

    template <class T>
    void append(const T& example)
    {
        const int32_t id = nextIds[T::name()];
        const std::vector<uint8_t> bytes = example.serialize();
        put(T::name(), id, bytes.data(), bytes.size());
        tables[T::name()][id] = bytes;
        nextIds[T::name()] = id + 1;
        fflush(file);
    }

    /* one transaction per batch (Dataset.h:203-232) */
    template <class T>
    void batchAppend(const std::vector<T>& examples, int32_t startId)
    {
        int32_t id = startId;
        for (const T& e : examples) {
            const std::vector<uint8_t> bytes = e.serialize();
            put(T::name(), id, bytes.data(), bytes.size());
            tables[T::name()][id] = bytes;
            id++;
        }
        nextIds[T::name()] = startId + (int32_t)examples.size();
        fflush(file);
    }

    /* append every record of another store (shard merge) */
    void mergeFrom(const Dataset& other)
    {
        for (const auto& t : other.tables)
            for (const auto& kv : t.second) {
                put(t.first, kv.first, kv.second.data(), kv.second.size());
                tables[t.first][kv.first] = kv.second;
            }
        fflush(file);
    }

private:
    static constexpr int32_t DROP_KEY = INT32_MIN; /* log entry that empties a table */
    static constexpr uint32_t MAGIC = 0x43525344u; /* "DSRC" */

    void put(const TableName& table, int32_t key, const uint8_t* data, size_t n)
    {
        const uint32_t magic = MAGIC, len = (uint32_t)n;
        const uint8_t nameLen = (uint8_t)table.size();
        if (fwrite(&magic, 4, 1, file) != 1 || fwrite(&nameLen, 1, 1, file) != 1 || fwrite(table.data(), 1, nameLen, file) != nameLen ||
            fwrite(&key, 4, 1, file) != 1 || fwrite(&len, 4, 1, file) != 1 || (n && fwrite(data, 1, n, file) != n))
            throw std::runtime_error("dataset write failed: " + path);
    }

    void load(FILE* f)
    {
        for (;;) {
            uint32_t magic, len;
            uint8_t nameLen;
            int32_t key;
            char name[256];
            if (fread(&magic, 4, 1, f) != 1) break;
            if (magic != MAGIC || fread(&nameLen, 1, 1, f) != 1 || fread(name, 1, nameLen, f) != nameLen || fread(&key, 4, 1, f) != 1 ||
                fread(&len, 4, 1, f) != 1)
                break; /* truncated tail of an interrupted batch: ignore */
            std::vector<uint8_t> bytes(len);
            if (len && fread(bytes.data(), 1, len, f) != len) break;
            const std::string table(name, nameLen);
            if (key == DROP_KEY) {
                tables[table].clear();
                nextIds[table] = 0;
            } else {
                tables[table][key] = std::move(bytes);
                nextIds[table] = std::max(nextIds[table], key + 1);
            }
        }
    }

    std::string path;
    FILE* file = nullptr;
    std::map<TableName, std::map<int32_t, std::vector<uint8_t>>> tables;
    std::map<TableName, int32_t> nextIds;
};

/* DG/Util/Dataset/BatchSettings.h */
struct BatchSettings {
    explicit BatchSettings(int32_t batchStartId, int32_t batchSize)
        : batchStartId(batchStartId), batchSize(batchSize), batchEndId(batchStartId + batchSize)
    {
    }
    const int32_t batchStartId;
    const int32_t batchSize;
    const int32_t batchEndId;
};

} // namespace DeepestScatter
