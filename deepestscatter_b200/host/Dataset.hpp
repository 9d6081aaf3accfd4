/*
 * Dataset.hpp -- record store of the dataset generator (host side, C++17, no third-party dependency).
 *
 * Mirrors the interface of the reference's DeepestScatter::Dataset (DG/Util/Dataset/Dataset.h:87-232,
 * Dataset.cpp): tables named after the protobuf message (`T::descriptor()->name()`), keys are int32 record ids,
 * values are the proto3 wire bytes, all in one LMDB data file that DeepestScatter_Train/LmdbDataset.py opens
 * unchanged.  liblmdb does not exist in this environment; the file format is written by host/LmdbFile.hpp.
 * Resumable (`CollectMode::Continue`, Tasks.h:65-68: an existing file is loaded and appended to) and mergeable
 * (one shard per GPU, mergeFrom).
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
#include "LmdbFile.hpp"

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
        while (p < end && shift < 64) { /* a varint is at most 10 bytes: never shift by 64 or more */
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
        /* every advance is checked against the bytes that are left BEFORE the pointer moves: a corrupt length must not wrap it */
        uint64_t n = 0;
        if (wireType == 0) {
            varint();
            return;
        } else if (wireType == 5)
            n = 4;
        else if (wireType == 1)
            n = 8;
        else if (wireType == 2)
            n = varint();
        else
            ok = false;
        if (!ok || n > (uint64_t)(end - p)) {
            ok = false;
            p = end;
            return;
        }
        p += n;
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

/*
 * Dataset (DG/Util/Dataset/Dataset.h:87-232, Dataset.cpp): tables by message name, int32 keys, proto3 values, kept in
 * one LMDB data file (MDB_NOSUBDIR, sub-databases MDB_INTEGERKEY | MDB_CREATE) written by host/LmdbFile.hpp.
 * A batchAppend is the reference's one-transaction-per-batch; the B+tree pages are written by commit() (destructor,
 * or explicitly after every N batches for crash safety).
 */
class Dataset {
public:
    struct Settings {
        explicit Settings(std::string path, bool create = true, bool readonly = false) : path(std::move(path)), create(create), readonly(readonly) {}
        std::string path;
        bool create;   /* false: a missing file is an error (merge sources, readers) instead of a fresh empty dataset */
        bool readonly; /* opened O_RDONLY, never committed */
    };
    using TableName = std::string;

    explicit Dataset(const Settings& settings) : file(settings.path, settings.create && !settings.readonly, settings.readonly), readonly(settings.readonly)
    {
        /* "Opening Dataset..." (Dataset.cpp:10) */
        for (const auto& t : file.tables()) nextIds[t.first] = t.second.empty() ? 0 : (int32_t)t.second.rbegin()->first + 1;
    }
    Dataset(const Dataset&) = delete;
    Dataset& operator=(const Dataset&) = delete;

    template <class T>
    size_t getRecordsCount()
    {
        if (!readonly) file.createTable(T::name()); /* getTable opens with MDB_CREATE (Dataset.cpp:78-90) */
        return file.count(T::name());
    }

    template <class T>
    T getRecord(int32_t recordId)
    {
        std::vector<uint8_t> bytes;
        if (!file.get(T::name(), (uint32_t)recordId, bytes))
            throw std::runtime_error(std::string("MDB_NOTFOUND: no record ") + std::to_string(recordId) + " in table " + T::name());
        return T::parse(bytes.data(), bytes.size());
    }

    template <class T>
    void dropTable()
    {
        file.drop(T::name());
        nextIds[T::name()] = 0;
    }

    template <class T>
    void append(const T& example)
    {
        const int32_t id = nextIds[T::name()]; /* zero if not initialised (Dataset.h:150-151) */
        const std::vector<uint8_t> bytes = example.serialize();
        file.put(T::name(), (uint32_t)id, bytes.data(), bytes.size());
        nextIds[T::name()] = id + 1;
    }

    /* one transaction per batch (Dataset.h:203-232) */
    template <class T>
    void batchAppend(const std::vector<T>& examples, int32_t startId)
    {
        int32_t id = startId;
        for (const T& e : examples) {
            const std::vector<uint8_t> bytes = e.serialize();
            file.put(T::name(), (uint32_t)id, bytes.data(), bytes.size());
            id++;
        }
        nextIds[T::name()] = startId + (int32_t)examples.size();
    }

    /* raw bytes (already encoded records) */
    void putRaw(const TableName& table, int32_t id, const uint8_t* data, size_t n)
    {
        file.put(table, (uint32_t)id, data, n);
        nextIds[table] = std::max(nextIds[table], id + 1);
    }

    /* copy every record of another dataset (merging the shards written by different GPUs) */
    void mergeFrom(Dataset& other)
    {
        std::vector<uint8_t> bytes;
        for (const auto& t : other.file.tables()) {
            file.createTable(t.first);
            for (const auto& kv : t.second) {
                other.file.get(t.first, kv.first, bytes);
                putRaw(t.first, (int32_t)kv.first, bytes.data(), bytes.size());
            }
        }
    }

    /* mdb_txn_commit: make everything appended so far durable and visible to readers.
     * DeepestScatter_Train/LmdbDataset.py:36-40 opens five tables the moment it opens a file, with create = False for readers, so a file the
     * training side can open holds all of them -- also BakedInterpolationSet, whose collector is outside this library's scope (in the reference
     * a table appears once anything asks for it: getTable opens with MDB_CREATE, Dataset.cpp:78-90).  They are added by the first commit that
     * writes anything; a dataset that was only opened stays untouched. */
    void commit()
    {
        if (readonly) return;
        if (file.dirty())
            for (const char* name : {"SceneSetup", "ScatterSample", "DisneyDescriptor", "BakedInterpolationSet", "Result"}) file.createTable(name);
        file.commit();
    }
    ~Dataset()
    {
        try {
            commit();
        } catch (...) {
        }
    }

    dslmdb::LmdbFile& lmdb() { return file; }

private:
    dslmdb::LmdbFile file;
    bool readonly = false;
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
