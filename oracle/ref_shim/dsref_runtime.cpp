/*
 * dsref_runtime.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only).  See dsref_runtime.h.
 */
#include "dsref_runtime.h"

#include <cmath>
#include <cstdio>

namespace dsref {

/* ------------------------------------------------------------------ registry */
static std::vector<std::unique_ptr<Module>>& modules()
{
    static std::vector<std::unique_ptr<Module>> m;
    return m;
}

Module* registerModule(const char* name)
{
    modules().emplace_back(new Module());
    modules().back()->name = name;
    return modules().back().get();
}
int addVar(Module* m, const char* name, const char* semantic, void* addr, size_t size)
{
    m->vars.push_back(VarDesc{name, semantic, addr, size});
    return 0;
}
int addBuf(Module* m, const char* name, DevBufferBase* b)
{
    m->bufs.push_back(BufDesc{name, b});
    return 0;
}
int addTex(Module* m, const char* name, DevTexBase* t)
{
    m->texs.push_back(TexDesc{name, t});
    return 0;
}
int addProg(Module* m, const char* name, void (*fn)())
{
    ProgDesc p;
    p.name = name;
    p.fn = fn;
    m->progs.push_back(p);
    return 0;
}
int addProg(Module* m, const char* name, void (*fn)(int))
{
    ProgDesc p;
    p.name = name;
    p.fnIntersect = fn;
    m->progs.push_back(p);
    return 0;
}
int addProg(Module* m, const char* name, void (*fn)(int, float*))
{
    ProgDesc p;
    p.name = name;
    p.fnBounds = fn;
    m->progs.push_back(p);
    return 0;
}

static Module* findModule(const std::string& name)
{
    for (auto& m : modules())
        if (m->name == name) return m.get();
    return nullptr;
}

/* ------------------------------------------------------------------ texture fetch definitions */
static Counters gCounters;
Counters& counters() { return gCounters; }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* One axis of a linear, clamp-to-edge, normalized-coordinate fetch: texel-centre convention x = u*N - 0.5 */
struct Axis {
    int i0, i1;
    float t;
};
static inline Axis axis(float u, size_t n)
{
    const float x = u * (float)n - 0.5f;
    const float f0 = floorf(x);
    Axis a;
    a.t = x - f0;
    /* clamp in float first so huge coordinates cannot overflow the int cast */
    const float c = fminf(fmaxf(f0, -2.0f), (float)n + 1.0f);
    const int i = (int)c;
    a.i0 = clampi(i, 0, (int)n - 1);
    a.i1 = clampi(i + 1, 0, (int)n - 1);
    return a;
}

float fetch1D(const SamplerState* s, float u)
{
    const LevelView& L = s->levels[0];
    const Axis a = axis(u, L.nx);
    if (s->format == RT_FORMAT_FLOAT) {
        const float* d = (const float*)L.data;
        return fmaf(a.t, d[a.i1] - d[a.i0], d[a.i0]);
    }
    const uint8_t* d = (const uint8_t*)L.data;
    return fmaf(a.t, (float)d[a.i1] - (float)d[a.i0], (float)d[a.i0]) * (1.0f / 255.0f);
}

static float fetchLevel3D(const SamplerState* s, const LevelView& L, float u, float v, float w)
{
    if (s->format != RT_FORMAT_UNSIGNED_BYTE) throw std::runtime_error("dsref: 3-D fetches are emulated for u8 volumes only");
    const Axis ax = axis(u, L.nx), ay = axis(v, L.ny), az = axis(w, L.nz);
    const uint8_t* d = (const uint8_t*)L.data;
    const size_t sy = L.nx, sz = L.nx * L.ny;
    auto at = [&](int x, int y, int z) { return (float)d[(size_t)x + (size_t)y * sy + (size_t)z * sz]; };
    const float c00 = fmaf(ax.t, at(ax.i1, ay.i0, az.i0) - at(ax.i0, ay.i0, az.i0), at(ax.i0, ay.i0, az.i0));
    const float c10 = fmaf(ax.t, at(ax.i1, ay.i1, az.i0) - at(ax.i0, ay.i1, az.i0), at(ax.i0, ay.i1, az.i0));
    const float c01 = fmaf(ax.t, at(ax.i1, ay.i0, az.i1) - at(ax.i0, ay.i0, az.i1), at(ax.i0, ay.i0, az.i1));
    const float c11 = fmaf(ax.t, at(ax.i1, ay.i1, az.i1) - at(ax.i0, ay.i1, az.i1), at(ax.i0, ay.i1, az.i1));
    const float c0 = fmaf(ay.t, c10 - c00, c00);
    const float c1 = fmaf(ay.t, c11 - c01, c01);
    return fmaf(az.t, c1 - c0, c0) * (1.0f / 255.0f);
}

float fetch3D(const SamplerState* s, float u, float v, float w)
{
    return fetchLevel3D(s, s->levels[0], u, v, w);
}

float fetch3DLod(const SamplerState* s, float u, float v, float w, float lod)
{
    const int last = (int)s->levels.size() - 1;
    const float l = fminf(fmaxf(lod, 0.0f), (float)last);
    const float lf = floorf(l);
    const int l0 = (int)lf;
    const float t = l - lf;
    const float a = fetchLevel3D(s, s->levels[l0], u, v, w);
    if (l0 >= last || t == 0.0f) return a;
    const float b = fetchLevel3D(s, s->levels[l0 + 1], u, v, w);
    return fmaf(t, b - a, a);
}

static std::vector<TextureSamplerObj*>& samplerTable()
{
    static std::vector<TextureSamplerObj*> t(1, nullptr); /* id 0 is RT_TEXTURE_ID_NULL */
    return t;
}
const SamplerState* samplerById(int id) { return &samplerTable().at((size_t)id)->state; }

/* ------------------------------------------------------------------ host objects */
static size_t formatSize(RTformat f)
{
    switch (f) {
    case RT_FORMAT_FLOAT: return 4;
    case RT_FORMAT_FLOAT2: return 8;
    case RT_FORMAT_FLOAT3: return 12;
    case RT_FORMAT_FLOAT4: return 16;
    case RT_FORMAT_UNSIGNED_BYTE: return 1;
    case RT_FORMAT_UNSIGNED_BYTE4: return 4;
    default: return 0;
    }
}

void BufferObj::setFormat(RTformat f)
{
    format = f;
    if (f != RT_FORMAT_USER) elementSize = formatSize(f);
    allocate();
}

void BufferObj::allocate()
{
    levels.resize(levelCount);
    for (unsigned l = 0; l < levelCount; l++) {
        RTsize w, h, d;
        getMipLevelSize(l, w, h, d);
        levels[l].resize(w * h * d * elementSize); /* keeps what is there; new bytes are zero */
    }
}

void TextureSamplerObj::refresh()
{
    state.levels.clear();
    if (!buffer) return;
    state.format = buffer->format;
    state.dims = (int)buffer->dimensionality;
    for (unsigned l = 0; l < buffer->levelCount; l++) {
        LevelView v;
        buffer->getMipLevelSize(l, v.nx, v.ny, v.nz);
        v.data = buffer->levels[l].data();
        state.levels.push_back(v);
    }
}

ContextObj* ContextObj::createNew() { return new ContextObj(); }
void ContextObj::destroy() { delete this; }
ContextObj::~ContextObj()
{
    for (auto& s : samplers) samplerTable()[(size_t)s->id] = nullptr;
}

Handle<BufferObj> ContextObj::createBuffer(unsigned)
{
    buffers.emplace_back(new BufferObj());
    buffers.back()->ctx = this;
    return Handle<BufferObj>(buffers.back().get());
}
Handle<BufferObj> ContextObj::createBuffer(unsigned type, RTformat format)
{
    Handle<BufferObj> b = createBuffer(type);
    b->setFormat(format);
    return b;
}
Handle<BufferObj> ContextObj::createBuffer(unsigned type, RTformat format, RTsize w)
{
    Handle<BufferObj> b = createBuffer(type, format);
    b->setSize(w);
    return b;
}
Handle<BufferObj> ContextObj::createBuffer(unsigned type, RTformat format, RTsize w, RTsize h)
{
    Handle<BufferObj> b = createBuffer(type, format);
    b->setSize(w, h);
    return b;
}
Handle<BufferObj> ContextObj::createBuffer(unsigned type, RTformat format, RTsize w, RTsize h, RTsize d)
{
    Handle<BufferObj> b = createBuffer(type, format);
    b->setSize(w, h, d);
    return b;
}
Handle<TextureSamplerObj> ContextObj::createTextureSampler()
{
    samplers.emplace_back(new TextureSamplerObj());
    TextureSamplerObj* s = samplers.back().get();
    s->ctx = this;
    s->id = (int)samplerTable().size();
    samplerTable().push_back(s);
    return Handle<TextureSamplerObj>(s);
}
Handle<ProgramObj> ContextObj::createProgramFromPTXFile(const std::string& path, const std::string& name)
{
    /* "./CUDA/<file>.cu.ptx" (Resources.cpp:159-166) -> module "<file>.cu" */
    std::string file = path;
    const size_t slash = file.find_last_of('/');
    if (slash != std::string::npos) file = file.substr(slash + 1);
    const std::string ext = ".ptx";
    if (file.size() > ext.size() && file.compare(file.size() - ext.size(), ext.size(), ext) == 0) file.resize(file.size() - ext.size());
    Module* m = findModule(file);
    if (!m) {
        /* a module outside the path (the light-probe research variant CloudMaterial.cpp:37-39 installs for ray type 3) is
         * not compiled; the handle exists so the host code runs, and reaching it is an error */
        programs.emplace_back(new ProgramObj());
        programs.back()->ctx = this;
        return Handle<ProgramObj>(programs.back().get());
    }
    for (auto& p : m->progs)
        if (p.name == name) {
            programs.emplace_back(new ProgramObj());
            ProgramObj* po = programs.back().get();
            po->ctx = this;
            po->module = m;
            po->prog = &p;
            return Handle<ProgramObj>(po);
        }
    throw std::runtime_error("dsref: program " + name + " not found in " + file);
}
Handle<GeometryObj> ContextObj::createGeometry()
{
    geometries.emplace_back(new GeometryObj());
    return Handle<GeometryObj>(geometries.back().get());
}
Handle<MaterialObj> ContextObj::createMaterial()
{
    materials.emplace_back(new MaterialObj());
    return Handle<MaterialObj>(materials.back().get());
}
Handle<GeometryGroupObj> ContextObj::createGeometryGroup()
{
    groups.emplace_back(new GeometryGroupObj());
    return Handle<GeometryGroupObj>(groups.back().get());
}
Handle<AccelerationObj> ContextObj::createAcceleration(const std::string&, const std::string&)
{
    accelerations.emplace_back(new AccelerationObj());
    return Handle<AccelerationObj>(accelerations.back().get());
}

/* ------------------------------------------------------------------ binding and launch */
struct SemanticSlots {
    std::vector<std::pair<void*, size_t>> launchIndex, currentRay, payload, tHit;
};

struct BoundProgram {
    ProgramObj* program = nullptr;
    SemanticSlots sem;
};

static void bindProgram(BoundProgram& bp, ProgramObj* p, const std::vector<const ScopedObj*>& outerScopes)
{
    bp.program = p;
    bp.sem = SemanticSlots();
    if (!p) return;
    if (!p->prog) {
        bp.program = nullptr; /* unresolved module: never bound, so a ray of its type finds no program */
        return;
    }
    if (p->destroyed) throw std::runtime_error("dsref: launch of a destroyed program " + p->prog->name);
    std::vector<const ScopedObj*> scopes;
    scopes.push_back(p);
    for (const ScopedObj* s : outerScopes) scopes.push_back(s);
    auto lookup = [&](const std::string& name) -> const VariableObj* {
        for (const ScopedObj* s : scopes)
            if (const VariableObj* v = s->find(name)) return v;
        return nullptr;
    };
    Module* m = p->module;
    for (VarDesc& v : m->vars) {
        if (v.semantic == "rtLaunchIndex") bp.sem.launchIndex.emplace_back(v.addr, v.size);
        else if (v.semantic == "rtCurrentRay") bp.sem.currentRay.emplace_back(v.addr, v.size);
        else if (v.semantic == "rtPayload") bp.sem.payload.emplace_back(v.addr, v.size);
        else if (v.semantic == "rtIntersectionDistance") bp.sem.tHit.emplace_back(v.addr, v.size);
        else if (const VariableObj* h = lookup(v.name)) {
            if (h->kind == VariableObj::BYTES) memcpy(v.addr, h->bytes.data(), std::min(v.size, h->bytes.size()));
        }
    }
    for (BufDesc& b : m->bufs) {
        const VariableObj* h = lookup(b.name);
        if (h && h->kind == VariableObj::BUFFER && h->buffer) {
            b.buf->data = h->buffer->levels[0].data();
            for (int i = 0; i < 3; i++) b.buf->dim[i] = h->buffer->size[i];
        }
    }
    for (TexDesc& t : m->texs) {
        const VariableObj* h = lookup(t.name);
        if (h && h->kind == VariableObj::SAMPLER && h->sampler) t.tex->state = &h->sampler->state;
    }
}

struct LaunchState {
    ContextObj* ctx = nullptr;
    BoundProgram rayGen, intersect;
    std::map<unsigned, BoundProgram> closestHit, miss;
    unsigned index[3] = {0, 0, 0};
    /* current rtTrace */
    float tmin = 0, tmax = 0, tPotential = 0;
    bool hit = false;
    unsigned tracesThisIndex = 0;
    bool inClosestHit = false;
};
static LaunchState gLaunch;

static bool gStreamOverride = false;
static uint32_t gRaygenStream = 0;
static bool gClosestHitCountsAttempts = false;
static uint32_t gContextStream = 0;

void setStreamOverride(bool enabled, uint32_t raygenStream, bool closestHitCountsAttempts)
{
    gStreamOverride = enabled;
    gRaygenStream = raygenStream;
    gClosestHitCountsAttempts = closestHitCountsAttempts;
}

uint32_t streamId()
{
    if (gStreamOverride) {
        if (gLaunch.inClosestHit && gClosestHitCountsAttempts) return gLaunch.tracesThisIndex; /* 1-based attempt number */
        return gRaygenStream;
    }
    return gContextStream;
}

static void setSlots(const std::vector<std::pair<void*, size_t>>& slots, const void* src, size_t srcSize)
{
    for (auto& s : slots) memcpy(s.first, src, std::min(s.second, srcSize));
}

static void setLaunchIndex(unsigned x, unsigned y, unsigned z)
{
    gLaunch.index[0] = x;
    gLaunch.index[1] = y;
    gLaunch.index[2] = z;
    gLaunch.tracesThisIndex = 0;
    setSlots(gLaunch.rayGen.sem.launchIndex, gLaunch.index, 12);
    setSlots(gLaunch.intersect.sem.launchIndex, gLaunch.index, 12);
    for (auto& kv : gLaunch.closestHit) setSlots(kv.second.sem.launchIndex, gLaunch.index, 12);
    for (auto& kv : gLaunch.miss) setSlots(kv.second.sem.launchIndex, gLaunch.index, 12);
}

void ContextObj::bindScene()
{
    for (auto& s : samplers) s->refresh();
    gLaunch.ctx = this;
    gLaunch.closestHit.clear();
    gLaunch.miss.clear();
    gLaunch.intersect = BoundProgram();
    const VariableObj* sid = find("subframeId");
    gContextStream = sid ? sid->getUint() : 0;

    const VariableObj* root = find("objectRoot");
    if (root && root->kind == VariableObj::GROUP && root->group && !root->group->children.empty()) {
        GeometryInstanceObj* gi = root->group->children[0];
        if (gi->geometry && gi->geometry->intersectProgram) bindProgram(gLaunch.intersect, gi->geometry->intersectProgram, {gi, gi->geometry, this});
        if (!gi->materials.empty()) {
            MaterialObj* mat = gi->materials[0];
            for (auto& kv : mat->closestHit) {
                /* a closest-hit module that was not compiled (the light-probe research variant) can never be reached by
                 * the ray types this library traces */
                if (kv.second) bindProgram(gLaunch.closestHit[kv.first], kv.second, {gi, mat, this});
            }
        }
    }
    for (auto& kv : missPrograms) bindProgram(gLaunch.miss[kv.first], kv.second, {this});
}

void traceRay(const optix::Ray& ray, void* payload, size_t payloadSize)
{
    LaunchState& L = gLaunch;
    gCounters.traces++;
    L.tracesThisIndex++;
    L.tmin = ray.tmin;
    L.tmax = ray.tmax;
    L.hit = false;
    if (L.intersect.program) {
        setSlots(L.intersect.sem.currentRay, &ray, sizeof(ray));
        L.intersect.program->prog->fnIntersect(0);
    }
    BoundProgram* target = nullptr;
    if (L.hit) {
        auto it = L.closestHit.find(ray.ray_type);
        if (it != L.closestHit.end()) target = &it->second;
    } else {
        auto it = L.miss.find(ray.ray_type);
        if (it != L.miss.end()) target = &it->second;
    }
    if (!target || !target->program) return;
    setSlots(target->sem.currentRay, &ray, sizeof(ray));
    const float tHit = L.tmax;
    setSlots(target->sem.tHit, &tHit, 4);
    setSlots(target->sem.payload, payload, payloadSize);
    L.inClosestHit = L.hit;
    target->program->prog->fn();
    L.inClosestHit = false;
    for (auto& s : target->sem.payload) memcpy(payload, s.first, std::min(s.second, payloadSize));
}

bool potentialIntersection(float t)
{
    if (t > gLaunch.tmin && t < gLaunch.tmax) {
        gLaunch.tPotential = t;
        return true;
    }
    return false;
}

bool reportIntersection(unsigned int)
{
    gLaunch.tmax = gLaunch.tPotential;
    gLaunch.hit = true;
    return true;
}

void ContextObj::launch(unsigned entry, RTsize w) { launch(entry, w, 1, 1); }
void ContextObj::launch(unsigned entry, RTsize w, RTsize h) { launch(entry, w, h, 1); }
void ContextObj::launch(unsigned, RTsize w, RTsize h, RTsize d)
{
    if (!rayGen || !rayGen->prog) throw std::runtime_error("dsref: launch without a (compiled) ray generation program");
    launchCount++;
    bindScene();
    bindProgram(gLaunch.rayGen, rayGen, {this});
    void (*fn)() = rayGen->prog->fn;
    for (RTsize z = 0; z < d; z++)
        for (RTsize y = 0; y < h; y++)
            for (RTsize x = 0; x < w; x++) {
                setLaunchIndex((unsigned)x, (unsigned)y, (unsigned)z);
                fn();
            }
}

void ContextObj::launchRect(RTsize, RTsize, RTsize x0, RTsize x1, RTsize y0, RTsize y1)
{
    if (!rayGen || !rayGen->prog) throw std::runtime_error("dsref: launch without a (compiled) ray generation program");
    launchCount++;
    bindScene();
    bindProgram(gLaunch.rayGen, rayGen, {this});
    void (*fn)() = rayGen->prog->fn;
    for (RTsize y = y0; y < y1; y++)
        for (RTsize x = x0; x < x1; x++) {
            setLaunchIndex((unsigned)x, (unsigned)y, 0);
            fn();
        }
}

void ContextObj::traceFrom(unsigned lx, unsigned ly, const optix::Ray& ray, void* payload, size_t payloadSize)
{
    gLaunch.rayGen = BoundProgram();
    setLaunchIndex(lx, ly, 0);
    traceRay(ray, payload, payloadSize);
}

} // namespace dsref
