/*
 * ds_records.cpp -- protobuf wire encoding of the dataset records (host only, no protobuf dependency).
 *
 * The reference serialises with protobuf 3.6.1 generated C++ (DeepestScatter_Train/CppProtocols/*.pb.cc).
 * proto3 rules reproduced here, as that generated code applies them:
 *   - scalar fields are written only when `value != 0` (Vector.pb.cc:286-296, Result.pb.cc:266-271), so
 *     +0.0f and -0.0f are both omitted and NaN is written;
 *   - sub-messages that were touched with mutable_*() are always written, even when empty
 *     (ScatterSample.pb.cc:329-340; the collectors always call mutable_point()/mutable_view_direction(),
 *     DG/Scene/ScatterSampleCollector.cpp:48-56);
 *   - ScatterSample.scene_setup_id is never set by the reference and therefore never written;
 *   - bytes/string fields are written when non-empty.
 * Golden vectors: tests/golden/records.json (generated from the reference's PythonProtocols modules).
 */
#include <cstring>

#include "../../include/ds_abi.h"

namespace {

struct Writer {
    uint8_t* out;
    size_t cap;
    size_t len = 0;
    bool overflow = false;
    void byte(uint8_t b)
    {
        if (len < cap)
            out[len] = b;
        else
            overflow = true;
        len++;
    }
    void varint(uint64_t v)
    {
        while (v >= 0x80) {
            byte((uint8_t)(v | 0x80));
            v >>= 7;
        }
        byte((uint8_t)v);
    }
    void bytes(const void* p, size_t n)
    {
        if (len + n <= cap)
            memcpy(out + len, p, n);
        else
            overflow = true;
        len += n;
    }
    void fixed32(int field, float f)
    {
        byte((uint8_t)((field << 3) | 5));
        bytes(&f, 4); /* little-endian host */
    }
    int done() const { return overflow ? DS_ERR_INVALID : (int)len; }
};

size_t vector3Size(const float v[3])
{
    size_t n = 0;
    for (int i = 0; i < 3; i++)
        if (v[i] != 0) n += 5;
    return n;
}

void writeVector3(Writer& w, int field, const float v[3])
{
    w.byte((uint8_t)((field << 3) | 2));
    w.varint(vector3Size(v));
    for (int i = 0; i < 3; i++)
        if (v[i] != 0) w.fixed32(i + 1, v[i]);
}

} // namespace

extern "C" {

int ds_record_scatter_sample(const float point[3], const float view_direction[3], uint8_t* out, size_t cap)
{
    if (!point || !view_direction || !out) return DS_ERR_INVALID;
    Writer w{out, cap};
    writeVector3(w, 2, point);
    writeVector3(w, 3, view_direction);
    return w.done();
}

int ds_record_disney_descriptor(const uint8_t* grid, size_t grid_len, uint8_t* out, size_t cap)
{
    if ((!grid && grid_len) || !out) return DS_ERR_INVALID;
    Writer w{out, cap};
    if (grid_len > 0) {
        w.byte(0x0a);
        w.varint(grid_len);
        w.bytes(grid, grid_len);
    }
    return w.done();
}

int ds_record_result(float light_intensity, int is_converged, uint8_t* out, size_t cap)
{
    if (!out) return DS_ERR_INVALID;
    Writer w{out, cap};
    if (light_intensity != 0) w.fixed32(1, light_intensity);
    if (is_converged) {
        w.byte(0x10);
        w.byte(0x01);
    }
    return w.done();
}

int ds_record_scene_setup(const char* cloud_path, float cloud_size_m, const float light_direction[3], uint8_t* out, size_t cap)
{
    if (!cloud_path || !light_direction || !out) return DS_ERR_INVALID;
    Writer w{out, cap};
    const size_t n = strlen(cloud_path);
    if (n > 0) {
        w.byte(0x0a);
        w.varint(n);
        w.bytes(cloud_path, n);
    }
    if (cloud_size_m != 0) w.fixed32(2, cloud_size_m);
    writeVector3(w, 3, light_direction);
    return w.done();
}

} /* extern "C" */
