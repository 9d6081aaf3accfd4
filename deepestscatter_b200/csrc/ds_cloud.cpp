/*
 * ds_cloud.cpp -- C ABI of the cloud importer front end (host/CloudImporter.hpp): Resources::loadVolumeBuffer
 * (DG/Util/Resources.cpp:68-155) for dense .npy grids and synth:<n> specs.  Host code only; the device work goes through
 * ds_volume_upload / ds_volume_upload_float / ds_volume_synth.
 */
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/ds_abi.h"
#include "../host/CloudImporter.hpp"
#include "../host/ExrWriter.hpp"

namespace {
std::mutex g_mutex;
std::map<DsContext*, std::unique_ptr<DeepestScatter::CloudImporter>> g_importers; /* one volume cache per context */
thread_local std::string g_error;
} // namespace

extern "C" {

int ds_cloud_load(DsContext* ctx, const char* path, int build_mips, int size_out[3])
{
    if (!ctx || !path) return DS_ERR_INVALID;
    try {
        DeepestScatter::CloudImporter* imp;
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            auto& slot = g_importers[ctx];
            if (!slot) slot.reset(new DeepestScatter::CloudImporter(ctx));
            imp = slot.get();
        }
        int size[3];
        imp->load(path, build_mips != 0, size);
        if (size_out) memcpy(size_out, size, sizeof(size));
        return DS_OK;
    } catch (const std::exception& e) {
        g_error = e.what();
        return DS_ERR_IO;
    }
}

void ds_cloud_forget(DsContext* ctx)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_importers.erase(ctx);
}

const char* ds_cloud_last_error(void) { return g_error.c_str(); }

int ds_cloud_crop_active(const float* dense, int nx, int ny, int nz, float* out, size_t out_capacity, int dims_out[3], double* max_density_out)
{
    if (!dense || !dims_out || nx <= 0 || ny <= 0 || nz <= 0) return DS_ERR_INVALID;
    try {
        DeepestScatter::DenseGrid g;
        g.nx = nx;
        g.ny = ny;
        g.nz = nz;
        g.f32.assign(dense, dense + (size_t)nx * ny * nz);
        const DeepestScatter::DenseGrid c = DeepestScatter::cropToActive(g);
        dims_out[0] = c.nx;
        dims_out[1] = c.ny;
        dims_out[2] = c.nz;
        if (max_density_out) *max_density_out = c.maxDensity;
        if (out) {
            if (out_capacity < c.f32.size()) return DS_ERR_INVALID;
            memcpy(out, c.f32.data(), c.f32.size() * sizeof(float));
        }
        return DS_OK;
    } catch (const std::exception& e) {
        g_error = e.what();
        return DS_ERR_INVALID;
    }
}

int ds_cloud_read_vdb(const char* path, float* out, size_t out_capacity, int dims_out[3], double* max_density_out)
{
    if (!path || !dims_out) return DS_ERR_INVALID;
    try {
        std::vector<float> dense;
        double mx = 0.0;
        dsvdb::toDense(dsvdb::readFirstFloatGrid(dsvdb::readFile(path)), dense, dims_out, mx);
        if (max_density_out) *max_density_out = mx;
        if (out) {
            if (out_capacity < dense.size()) return DS_ERR_INVALID;
            memcpy(out, dense.data(), dense.size() * sizeof(float));
        }
        return DS_OK;
    } catch (const std::exception& e) {
        g_error = e.what();
        return DS_ERR_IO;
    }
}

int ds_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgba)
{
    if (!path || !rgba) return DS_ERR_INVALID;
    try {
        DeepestScatter::writeExrRGB(path, width, height, rgba, true);
        return DS_OK;
    } catch (const std::exception& e) {
        g_error = e.what();
        return DS_ERR_IO;
    }
}

} /* extern "C" */
