/* TEST INFRASTRUCTURE (oracle/_ref build only).  protobuf 3.6.1 is not installed, so the generated classes of
 * DeepestScatter_Train/CppProtocols/*.pb.h cannot be compiled; these plain structs give the collectors
 * (ScatterSampleCollector.cpp:45-59, DisneyDescriptorCollector.cpp:24-39,76-103, RadianceCollector.cpp:27-46,157-167) the
 * accessors they call.  The wire bytes are pinned elsewhere (tests/golden/records.json, from the reference's *_pb2.py). */
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
namespace Persistance {
class Vector3 {
public:
    float x() const { return x_; }
    float y() const { return y_; }
    float z() const { return z_; }
    void set_x(float v) { x_ = v; }
    void set_y(float v) { y_ = v; }
    void set_z(float v) { z_ = v; }

private:
    float x_ = 0, y_ = 0, z_ = 0;
};
class ScatterSample {
public:
    const Vector3& point() const { return point_; }
    const Vector3& view_direction() const { return view_; }
    Vector3* mutable_point() { return &point_; }
    Vector3* mutable_view_direction() { return &view_; }
    int32_t scene_setup_id() const { return scene_; }
    void set_scene_setup_id(int32_t v) { scene_ = v; }

private:
    int32_t scene_ = 0;
    Vector3 point_, view_;
};
class Result {
public:
    float light_intensity() const { return intensity_; }
    bool is_converged() const { return converged_; }
    void set_light_intensity(float v) { intensity_ = v; }
    void set_is_converged(bool v) { converged_ = v; }

private:
    float intensity_ = 0;
    bool converged_ = false;
};
class DisneyDescriptor {
public:
    const std::string& grid() const { return grid_; }
    void set_grid(const void* p, size_t n) { grid_.assign((const char*)p, n); }

private:
    std::string grid_;
};
class SceneSetup {
public:
    const std::string& cloud_path() const { return path_; }
    float cloud_size_m() const { return size_; }
    const Vector3& light_direction() const { return light_; }
    void set_cloud_path(const std::string& s) { path_ = s; }
    void set_cloud_size_m(float v) { size_ = v; }
    Vector3* mutable_light_direction() { return &light_; }

private:
    std::string path_;
    float size_ = 0;
    Vector3 light_;
};
} // namespace Persistance
