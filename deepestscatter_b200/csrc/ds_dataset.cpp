/*
 * ds_dataset.cpp -- C ABI of the dataset store (host only): the LMDB data file the reference's collectors append to
 * (DG/Util/Dataset/Dataset.h:87-232, Dataset.cpp) and DeepestScatter_Train/LmdbDataset.py reads.  Thin wrapper over
 * host/Dataset.hpp / host/LmdbFile.hpp; C++ exceptions stop here.
 */
#include <cstring>
#include <memory>
#include <string>

#include "../../include/ds_abi.h"
#include "../host/Dataset.hpp"

struct DsDataset {
    std::unique_ptr<DeepestScatter::Dataset> ds;
    std::string err;
};

namespace {
thread_local std::string g_openError;

template <class F>
int guarded(DsDataset* d, F&& f)
{
    if (!d || !d->ds) return DS_ERR_INVALID;
    try {
        return f();
    } catch (const dslmdb::Error& e) {
        d->err = e.what();
        return DS_ERR_IO;
    } catch (const std::exception& e) {
        d->err = e.what();
        return DS_ERR_INVALID;
    }
}
} // namespace

extern "C" {

int ds_dataset_open(const char* path, DsDataset** out)
{
    if (!path || !out) return DS_ERR_INVALID;
    *out = nullptr;
    try {
        std::unique_ptr<DsDataset> d(new DsDataset);
        d->ds.reset(new DeepestScatter::Dataset(DeepestScatter::Dataset::Settings(path)));
        *out = d.release();
        return DS_OK;
    } catch (const std::exception& e) {
        g_openError = e.what();
        return DS_ERR_IO;
    }
}

int ds_dataset_close(DsDataset* d)
{
    if (!d) return DS_ERR_INVALID;
    int rc = DS_OK;
    try {
        if (d->ds) d->ds->commit();
    } catch (const std::exception& e) {
        g_openError = e.what();
        rc = DS_ERR_IO;
    }
    delete d;
    return rc;
}

const char* ds_dataset_last_error(DsDataset* d) { return d ? d->err.c_str() : g_openError.c_str(); }

int ds_dataset_commit(DsDataset* d)
{
    return guarded(d, [&] {
        d->ds->commit();
        return DS_OK;
    });
}

int ds_dataset_put(DsDataset* d, const char* table, int32_t id, const uint8_t* data, size_t n)
{
    if (!table || (!data && n)) return DS_ERR_INVALID;
    return guarded(d, [&] {
        d->ds->putRaw(table, id, data, n);
        return DS_OK;
    });
}

long long ds_dataset_get(DsDataset* d, const char* table, int32_t id, uint8_t* out, size_t cap)
{
    if (!table) return DS_ERR_INVALID;
    long long len = 0;
    const int rc = guarded(d, [&] {
        std::vector<uint8_t> bytes;
        if (!d->ds->lmdb().get(table, (uint32_t)id, bytes)) {
            d->err = std::string("MDB_NOTFOUND: no record ") + std::to_string(id) + " in table " + table;
            return DS_ERR_STATE;
        }
        len = (long long)bytes.size();
        if (out && cap >= bytes.size() && !bytes.empty()) memcpy(out, bytes.data(), bytes.size());
        return DS_OK;
    });
    return rc ? rc : len;
}

long long ds_dataset_count(DsDataset* d, const char* table)
{
    if (!table) return DS_ERR_INVALID;
    long long n = 0;
    const int rc = guarded(d, [&] {
        n = (long long)d->ds->lmdb().count(table);
        return DS_OK;
    });
    return rc ? rc : n;
}

int ds_dataset_drop(DsDataset* d, const char* table)
{
    if (!table) return DS_ERR_INVALID;
    return guarded(d, [&] {
        d->ds->lmdb().drop(table);
        return DS_OK;
    });
}

int ds_dataset_merge(DsDataset* d, const char* other_path)
{
    if (!other_path) return DS_ERR_INVALID;
    return guarded(d, [&] {
        /* read-only, never created: a mistyped shard path must fail, not merge an empty dataset (and a read-only shard must merge) */
        DeepestScatter::Dataset other{DeepestScatter::Dataset::Settings(other_path, /*create=*/false, /*readonly=*/true)};
        d->ds->mergeFrom(other);
        return DS_OK;
    });
}

int ds_dataset_append_scene_setup(DsDataset* d, int32_t scene_id, const char* cloud_path, float cloud_size_m, const float light_direction[3])
{
    if (!cloud_path || !light_direction) return DS_ERR_INVALID;
    return guarded(d, [&] {
        Persistance::SceneSetup s;
        s.cloud_path = cloud_path;
        s.cloud_size_m = cloud_size_m;
        s.light_direction = {light_direction[0], light_direction[1], light_direction[2]};
        const std::vector<uint8_t> b = s.serialize();
        d->ds->putRaw(Persistance::SceneSetup::name(), scene_id, b.data(), b.size());
        return DS_OK;
    });
}

int ds_dataset_append_scatter_samples(DsDataset* d, int32_t start_id, uint32_t n, const float* positions, const float* directions)
{
    if (n && (!positions || !directions)) return DS_ERR_INVALID;
    return guarded(d, [&] {
        uint8_t buf[64];
        for (uint32_t i = 0; i < n; i++) {
            const int len = ds_record_scatter_sample(positions + 3 * (size_t)i, directions + 3 * (size_t)i, buf, sizeof(buf));
            if (len < 0) return len;
            d->ds->putRaw(Persistance::ScatterSample::name(), start_id + (int32_t)i, buf, (size_t)len);
        }
        return DS_OK;
    });
}

int ds_dataset_append_descriptors(DsDataset* d, int32_t start_id, uint32_t n, const uint8_t* descriptors, size_t descriptor_bytes)
{
    if (n && !descriptors) return DS_ERR_INVALID;
    return guarded(d, [&] {
        std::vector<uint8_t> buf(descriptor_bytes + 16);
        for (uint32_t i = 0; i < n; i++) {
            const int len = ds_record_disney_descriptor(descriptors + (size_t)i * descriptor_bytes, descriptor_bytes, buf.data(), buf.size());
            if (len < 0) return len;
            d->ds->putRaw(Persistance::DisneyDescriptor::name(), start_id + (int32_t)i, buf.data(), (size_t)len);
        }
        return DS_OK;
    });
}

int ds_dataset_append_results(DsDataset* d, int32_t start_id, uint32_t n, const float* light_intensity, const uint8_t* is_converged)
{
    if (n && (!light_intensity || !is_converged)) return DS_ERR_INVALID;
    return guarded(d, [&] {
        uint8_t buf[16];
        for (uint32_t i = 0; i < n; i++) {
            const int len = ds_record_result(light_intensity[i], is_converged[i] ? 1 : 0, buf, sizeof(buf));
            if (len < 0) return len;
            d->ds->putRaw(Persistance::Result::name(), start_id + (int32_t)i, buf, (size_t)len);
        }
        return DS_OK;
    });
}

} /* extern "C" */
