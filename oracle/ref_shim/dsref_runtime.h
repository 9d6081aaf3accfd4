/*
 * dsref_runtime.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * A small single-threaded emulation of the OptiX 5.1 object model, enough for the reference's DataGen sources to be
 * compiled UNMODIFIED by g++ and run on the host:
 *   - host side (optixu/optixpp_namespace.h): Context, Program, Buffer, TextureSampler, Variable, Geometry, Material,
 *     GeometryInstance, GeometryGroup, Acceleration as Handle<...Obj>, with the calls the reference makes;
 *   - device side (optix.h / optix_device.h): rtDeclareVariable, rtBuffer, rtTextureSampler, tex1D/tex3D/rtTex3D/
 *     rtTex3DLod, rtTrace, rtPotentialIntersection/rtReportIntersection, RT_PROGRAM.
 * A "PTX module" is one reference .cu compiled as its own translation unit (oracle/ref_shim/modules/): its
 * rtDeclareVariable's become file-static objects registered by name; Context::launch copies the values the host set
 * (program scope, then geometry instance / material / geometry, then context -- the OptiX lookup order) into them, then runs
 * the ray-generation function once per launch index.  The scene is the single box of cloudBBox.cu; rtTrace runs the bound
 * intersection program and then the closest-hit (or miss) program of the ray type.  No BVH, no any-hit, no recursion limit.
 *
 * Texture fetches (the SDK/hardware part that cannot be compiled) are DEFINED in dsref_runtime.cpp exactly as the header of
 * oracle/ds_oracle.cpp documents them: unnormalised coordinate x = u*N - 0.5, exact fp32 weights, clamp-to-edge,
 * normalized-float read for u8, trilinear with fmaf in x, y, z order; LOD fetches lerp two levels.
 */
#ifndef DSREF_RUNTIME_H
#define DSREF_RUNTIME_H

#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "dsref_vec.h"

/* what cuda's host_defines.h gives a host compiler (the reference's shared headers carry these qualifiers) */
#ifndef __device__
#define __device__
#endif
#ifndef __host__
#define __host__
#endif

typedef size_t RTsize;
enum RTbuffertype { RT_BUFFER_INPUT = 1, RT_BUFFER_OUTPUT = 2, RT_BUFFER_INPUT_OUTPUT = 3 };
enum RTformat {
    RT_FORMAT_UNKNOWN = 0,
    RT_FORMAT_FLOAT,
    RT_FORMAT_FLOAT2,
    RT_FORMAT_FLOAT3,
    RT_FORMAT_FLOAT4,
    RT_FORMAT_UNSIGNED_BYTE,
    RT_FORMAT_UNSIGNED_BYTE4,
    RT_FORMAT_USER
};
enum RTwrapmode { RT_WRAP_REPEAT, RT_WRAP_CLAMP_TO_EDGE };
enum RTfiltermode { RT_FILTER_NEAREST, RT_FILTER_LINEAR, RT_FILTER_NONE };
enum RTtextureindexmode { RT_TEXTURE_INDEX_NORMALIZED_COORDINATES, RT_TEXTURE_INDEX_ARRAY_INDEX };
enum RTobjecttype { RT_OBJECTTYPE_UNKNOWN = 0, RT_OBJECTTYPE_FLOAT, RT_OBJECTTYPE_UNSIGNED_INT, RT_OBJECTTYPE_INT, RT_OBJECTTYPE_OBJECT };

struct rtObject { void* p; };

namespace dsref {

/* ------------------------------------------------------------------ device-side registry */
struct LevelView {
    const void* data = nullptr;
    size_t nx = 1, ny = 1, nz = 1;
};
struct SamplerState {
    RTformat format = RT_FORMAT_UNKNOWN;
    int dims = 1;
    std::vector<LevelView> levels;
};
float fetch1D(const SamplerState* s, float u);
float fetch3D(const SamplerState* s, float u, float v, float w);
float fetch3DLod(const SamplerState* s, float u, float v, float w, float lod);
const SamplerState* samplerById(int id);

struct DevBufferBase {
    void* data = nullptr;
    size_t dim[3] = {0, 0, 0};
};
struct DevTexBase {
    const SamplerState* state = nullptr;
};
struct VarDesc {
    std::string name, semantic;
    void* addr;
    size_t size;
};
struct BufDesc {
    std::string name;
    DevBufferBase* buf;
};
struct TexDesc {
    std::string name;
    DevTexBase* tex;
};
struct ProgDesc {
    std::string name;
    void (*fn)() = nullptr;
    void (*fnIntersect)(int) = nullptr;
    void (*fnBounds)(int, float*) = nullptr;
};
struct Module {
    std::string name;
    std::vector<VarDesc> vars;
    std::vector<BufDesc> bufs;
    std::vector<TexDesc> texs;
    std::vector<ProgDesc> progs;
};
Module* registerModule(const char* name);
int addVar(Module* m, const char* name, const char* semantic, void* addr, size_t size);
int addBuf(Module* m, const char* name, DevBufferBase* b);
int addTex(Module* m, const char* name, DevTexBase* t);
int addProg(Module* m, const char* name, void (*fn)());
int addProg(Module* m, const char* name, void (*fn)(int));
int addProg(Module* m, const char* name, void (*fn)(int, float*));

/* rtTrace / intersection reporting, called from the device-side shim */
void traceRay(const optix::Ray& ray, void* payload, size_t payloadSize);
bool potentialIntersection(float t);
bool reportIntersection(unsigned int material);

/* value of clock() inside tea<N> (CUDA/random.cuh:38): an explicit stream id, see oracle/ds_oracle.cpp header */
uint32_t streamId();

/* ------------------------------------------------------------------ host-side objects */
struct BufferObj;
struct TextureSamplerObj;
struct GeometryGroupObj;
struct ContextObj;

template <class T>
class Handle {
public:
    Handle() : ptr(nullptr) {}
    Handle(T* p) : ptr(p) {}
    T* operator->() const { return ptr; }
    T* get() const { return ptr; }
    operator bool() const { return ptr != nullptr; }
    Handle<struct VariableObj> operator[](const std::string& name) const { return ptr->queryVariable(name); }
    Handle<struct VariableObj> operator[](const char* name) const { return ptr->queryVariable(name); }
    static Handle<T> create() { return Handle<T>(T::createNew()); }

private:
    T* ptr;
};

struct VariableObj {
    enum Kind { UNSET, BYTES, BUFFER, SAMPLER, GROUP };
    Kind kind = UNSET;
    RTobjecttype type = RT_OBJECTTYPE_UNKNOWN;
    std::vector<uint8_t> bytes;
    BufferObj* buffer = nullptr;
    TextureSamplerObj* sampler = nullptr;
    GeometryGroupObj* group = nullptr;

    void setBytes(const void* p, size_t n, RTobjecttype t)
    {
        kind = BYTES;
        type = t;
        bytes.assign((const uint8_t*)p, (const uint8_t*)p + n);
    }
    void setFloat(float a) { setBytes(&a, 4, RT_OBJECTTYPE_FLOAT); }
    void setFloat(float a, float b, float c)
    {
        const float v[3] = {a, b, c};
        setBytes(v, 12, RT_OBJECTTYPE_FLOAT);
    }
    void setFloat(const float3& v) { setBytes(&v, 12, RT_OBJECTTYPE_FLOAT); }
    void setUint(unsigned int a) { setBytes(&a, 4, RT_OBJECTTYPE_UNSIGNED_INT); }
    void setUint(unsigned int a, unsigned int b)
    {
        const unsigned int v[2] = {a, b};
        setBytes(v, 8, RT_OBJECTTYPE_UNSIGNED_INT);
    }
    void setInt(int a) { setBytes(&a, 4, RT_OBJECTTYPE_INT); }
    void setBuffer(Handle<BufferObj> b)
    {
        kind = BUFFER;
        type = RT_OBJECTTYPE_OBJECT;
        buffer = b.get();
    }
    void setTextureSampler(Handle<TextureSamplerObj> s)
    {
        kind = SAMPLER;
        type = RT_OBJECTTYPE_OBJECT;
        sampler = s.get();
    }
    void set(Handle<GeometryGroupObj> g)
    {
        kind = GROUP;
        type = RT_OBJECTTYPE_OBJECT;
        group = g.get();
    }
    RTobjecttype getType() const { return type; }
    void getUint(unsigned int& v) const
    {
        v = 0;
        if (kind == BYTES && bytes.size() >= 4) memcpy(&v, bytes.data(), 4);
    }
    unsigned int getUint() const
    {
        unsigned int v;
        getUint(v);
        return v;
    }
};

struct ScopedObj {
    std::map<std::string, std::unique_ptr<VariableObj>> vars;
    Handle<VariableObj> queryVariable(const std::string& name)
    {
        auto& slot = vars[name];
        if (!slot) slot.reset(new VariableObj());
        return Handle<VariableObj>(slot.get());
    }
    const VariableObj* find(const std::string& name) const
    {
        auto it = vars.find(name);
        return it == vars.end() || it->second->kind == VariableObj::UNSET ? nullptr : it->second.get();
    }
    virtual ~ScopedObj() = default;
};

struct BufferObj {
    ContextObj* ctx = nullptr;
    RTformat format = RT_FORMAT_UNKNOWN;
    size_t elementSize = 0;
    unsigned dimensionality = 1;
    size_t size[3] = {1, 1, 1};
    unsigned levelCount = 1;
    std::vector<std::vector<uint8_t>> levels;

    void allocate();
    void setFormat(RTformat f);
    void setElementSize(size_t s)
    {
        elementSize = s;
        allocate();
    }
    size_t getElementSize() const { return elementSize; }
    void setSize(RTsize w)
    {
        dimensionality = 1;
        size[0] = w; size[1] = 1; size[2] = 1;
        allocate();
    }
    void setSize(RTsize w, RTsize h)
    {
        dimensionality = 2;
        size[0] = w; size[1] = h; size[2] = 1;
        allocate();
    }
    void setSize(RTsize w, RTsize h, RTsize d)
    {
        dimensionality = 3;
        size[0] = w; size[1] = h; size[2] = d;
        allocate();
    }
    void getSize(RTsize& w) const { w = size[0]; }
    void getSize(RTsize& w, RTsize& h) const { w = size[0]; h = size[1]; }
    void getSize(RTsize& w, RTsize& h, RTsize& d) const { w = size[0]; h = size[1]; d = size[2]; }
    /* OptiX: each mip level halves every dimension, never below 1 */
    void getMipLevelSize(unsigned level, RTsize& w, RTsize& h, RTsize& d) const
    {
        w = std::max<size_t>(1, size[0] >> level);
        h = std::max<size_t>(1, size[1] >> level);
        d = std::max<size_t>(1, size[2] >> level);
    }
    void setMipLevelCount(unsigned n)
    {
        levelCount = n;
        allocate();
    }
    unsigned getMipLevelCount() const { return levelCount; }
    unsigned getDimensionality() const { return dimensionality; }
    void* map(unsigned level = 0) { return levels.at(level).data(); }
    void unmap(unsigned = 0) {}
    void destroy() {}
};

struct TextureSamplerObj {
    ContextObj* ctx = nullptr;
    int id = 0;
    BufferObj* buffer = nullptr;
    SamplerState state; /* refreshed from the buffer at launch time */
    void setWrapMode(unsigned, RTwrapmode m)
    {
        if (m != RT_WRAP_CLAMP_TO_EDGE) throw std::runtime_error("dsref: only RT_WRAP_CLAMP_TO_EDGE is emulated");
    }
    void setFilteringModes(RTfiltermode minF, RTfiltermode magF, RTfiltermode)
    {
        if (minF != RT_FILTER_LINEAR || magF != RT_FILTER_LINEAR) throw std::runtime_error("dsref: only linear filtering is emulated");
    }
    void setIndexingMode(RTtextureindexmode m)
    {
        if (m != RT_TEXTURE_INDEX_NORMALIZED_COORDINATES) throw std::runtime_error("dsref: only normalized coordinates are emulated");
    }
    void setBuffer(Handle<BufferObj> b) { buffer = b.get(); }
    int getId() const { return id; }
    void refresh();
    void destroy() {}
};

struct ProgramObj : ScopedObj {
    ContextObj* ctx = nullptr;
    Module* module = nullptr;
    ProgDesc* prog = nullptr;
    bool destroyed = false;
    void destroy() { destroyed = true; }
};

struct GeometryObj : ScopedObj {
    ProgramObj* boundsProgram = nullptr;
    ProgramObj* intersectProgram = nullptr;
    unsigned primitiveCount = 0;
    void setBoundingBoxProgram(Handle<ProgramObj> p) { boundsProgram = p.get(); }
    void setIntersectionProgram(Handle<ProgramObj> p) { intersectProgram = p.get(); }
    void setPrimitiveCount(unsigned n) { primitiveCount = n; }
};

struct MaterialObj : ScopedObj {
    std::map<unsigned, ProgramObj*> closestHit;
    void setClosestHitProgram(unsigned rayType, Handle<ProgramObj> p) { closestHit[rayType] = p.get(); }
};

struct GeometryInstanceObj : ScopedObj {
    GeometryObj* geometry = nullptr;
    std::vector<MaterialObj*> materials;
};

struct AccelerationObj {
};

struct GeometryGroupObj {
    std::vector<GeometryInstanceObj*> children;
    AccelerationObj* acceleration = nullptr;
    void addChild(Handle<GeometryInstanceObj> gi) { children.push_back(gi.get()); }
    void setAcceleration(Handle<AccelerationObj> a) { acceleration = a.get(); }
};

struct ContextObj : ScopedObj {
    std::vector<std::unique_ptr<BufferObj>> buffers;
    std::vector<std::unique_ptr<TextureSamplerObj>> samplers;
    std::vector<std::unique_ptr<ProgramObj>> programs;
    std::vector<std::unique_ptr<GeometryObj>> geometries;
    std::vector<std::unique_ptr<MaterialObj>> materials;
    std::vector<std::unique_ptr<GeometryInstanceObj>> instances;
    std::vector<std::unique_ptr<GeometryGroupObj>> groups;
    std::vector<std::unique_ptr<AccelerationObj>> accelerations;
    ProgramObj* rayGen = nullptr;
    ProgramObj* exceptionProgram = nullptr;
    std::map<unsigned, ProgramObj*> missPrograms;
    unsigned rayTypeCount = 0, entryPointCount = 0;
    unsigned long long launchCount = 0;

    static ContextObj* createNew();
    void destroy();
    ~ContextObj() override;

    Handle<BufferObj> createBuffer(unsigned type);
    Handle<BufferObj> createBuffer(unsigned type, RTformat format);
    Handle<BufferObj> createBuffer(unsigned type, RTformat format, RTsize w);
    Handle<BufferObj> createBuffer(unsigned type, RTformat format, RTsize w, RTsize h);
    Handle<BufferObj> createBuffer(unsigned type, RTformat format, RTsize w, RTsize h, RTsize d);
    Handle<TextureSamplerObj> createTextureSampler();
    Handle<ProgramObj> createProgramFromPTXFile(const std::string& path, const std::string& name);
    Handle<GeometryObj> createGeometry();
    Handle<MaterialObj> createMaterial();
    template <class It>
    Handle<GeometryInstanceObj> createGeometryInstance(Handle<GeometryObj> g, It matBegin, It matEnd)
    {
        instances.emplace_back(new GeometryInstanceObj());
        GeometryInstanceObj* gi = instances.back().get();
        gi->geometry = g.get();
        for (It it = matBegin; it != matEnd; ++it) gi->materials.push_back(it->get());
        return Handle<GeometryInstanceObj>(gi);
    }
    Handle<GeometryGroupObj> createGeometryGroup();
    Handle<AccelerationObj> createAcceleration(const std::string&, const std::string&);

    void setRayTypeCount(unsigned n) { rayTypeCount = n; }
    void setEntryPointCount(unsigned n) { entryPointCount = n; }
    void setRayGenerationProgram(unsigned, Handle<ProgramObj> p) { rayGen = p.get(); }
    Handle<ProgramObj> getRayGenerationProgram(unsigned) const { return Handle<ProgramObj>(rayGen); }
    void setExceptionProgram(unsigned, Handle<ProgramObj> p) { exceptionProgram = p.get(); }
    void setMissProgram(unsigned rayType, Handle<ProgramObj> p) { missPrograms[rayType] = p.get(); }
    void validate() {}
    void launch(unsigned entry, RTsize w);
    void launch(unsigned entry, RTsize w, RTsize h);
    void launch(unsigned entry, RTsize w, RTsize h, RTsize d);

    /* test hooks (not part of the OptiX API): run rays against the bound scene from a caller-chosen launch index */
    void bindScene();
    void traceFrom(unsigned lx, unsigned ly, const optix::Ray& ray, void* payload, size_t payloadSize);
    /* launch only the sub-rectangle [x0, x1) x [y0, y1) of a (w, h) launch */
    void launchRect(RTsize w, RTsize h, RTsize x0, RTsize x1, RTsize y0, RTsize y1);
};

/* stream policy for clock(): by default the context variable "subframeId" at launch time (what Camera.cpp:192 and
 * RadianceCollector.cpp:91 set before every launch); the API layer may override it. */
void setStreamOverride(bool enabled, uint32_t raygenStream, bool closestHitCountsAttempts);

/* work counters: rtTrace calls, rtTex3D fetches by id (the density taps of cloud.cuh:58-62 = march steps) and tex3D fetches
 * through a bound sampler (the inScatter taps of cloud.cuh:64-68 = scatter events; the density taps of the bake) */
struct Counters {
    unsigned long long traces = 0, bindlessTaps = 0, boundTaps = 0;
};
Counters& counters();

} // namespace dsref

namespace optix {
typedef dsref::Handle<dsref::ContextObj> Context;
typedef dsref::Handle<dsref::ProgramObj> Program;
typedef dsref::Handle<dsref::BufferObj> Buffer;
typedef dsref::Handle<dsref::TextureSamplerObj> TextureSampler;
typedef dsref::Handle<dsref::VariableObj> Variable;
typedef dsref::Handle<dsref::GeometryObj> Geometry;
typedef dsref::Handle<dsref::MaterialObj> Material;
typedef dsref::Handle<dsref::GeometryInstanceObj> GeometryInstance;
typedef dsref::Handle<dsref::GeometryGroupObj> GeometryGroup;
typedef dsref::Handle<dsref::AccelerationObj> Acceleration;
using dsref::Handle;
typedef dsref::ContextObj ContextObj;
typedef dsref::ProgramObj ProgramObj;
} // namespace optix

#endif /* DSREF_RUNTIME_H */
