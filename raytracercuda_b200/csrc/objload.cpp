// objload.cpp -- Wavefront OBJ import for the host side.  The reference imports through Assimp
// (TestProgram/Model.cpp:26-54: Triangulate | JoinIdenticalVertices | CalcTangentSpace | SortByPType, one IMesh per aiMesh, i.e. per
// material), which exists in its tree only as a Windows binary.  This loader keeps what parity depends on: vertex positions exactly as
// parsed, in file order; faces in file order, polygons fan-triangulated (v0, v[i], v[i+1]); optional split into one mesh per `usemtl`
// group in order of first use; `vn` normals de-indexed to per-position normals (first use wins) so that the reference's shading
// path (VERTEX_DATA_NORMAL indexed by the position index, BuildTree.cu:486-496) has data.  It is the C++ twin of
// raytracercuda_b200/meshio.py (tests/test_objload.py checks they agree on the reference's bundled files).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <string>
#include <vector>
#include "../../include/beam_host_c.h"

namespace
{
struct ObjGroup { std::string name; std::vector<uint32_t> idx; std::string texture; };
struct ObjFile
{
    std::vector<float> pos, nrm, uv;      // nrm: per position (3 floats each) or empty; uv: per position (2 floats each) or empty
    std::vector<ObjGroup> groups;
};

std::string dir_of(const char* path)
{
    std::string p(path);
    size_t k = p.find_last_of("/\\");
    return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

// .mtl: `newmtl NAME` ... `map_Kd FILE` -> texture path per material (Content/f16.mtl)
void parse_mtl(const std::string& path, const std::string& dir, std::map<std::string, std::string>& texOf)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return;
    char line[4096], name[1024];
    std::string cur;
    while (fgets(line, sizeof line, f))
    {
        char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (strncmp(p, "newmtl", 6) == 0 && sscanf(p + 6, " %1023s", name) == 1) cur = name;
        else if (strncmp(p, "map_Kd", 6) == 0 && !cur.empty() && sscanf(p + 6, " %1023s", name) == 1) texOf[cur] = dir + name;
    }
    fclose(f);
}

// 1-based or negative (relative) OBJ index -> 0-based; anything outside [0, count) comes back as -1
int resolve(long v, size_t count) { long r = v > 0 ? v - 1 : (long)count + v; return (v != 0 && r >= 0 && (size_t)r < count) ? (int)r : -1; }

bool parse_obj(const char* path, bool splitByMaterial, ObjFile& out, std::string& err)
{
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<float> vn, vt;
    std::vector<int> vnOfV, vtOfV;        // normal / texture-coordinate index per position, -1 = unset
    std::map<std::string, std::string> texOf;
    const std::string dir = dir_of(path);
    std::map<std::string, size_t> groupOf;
    std::string cur = "__default__";
    char line[4096];
    std::vector<int> vs, ns, ts;
    while (fgets(line, sizeof line, f))
    {
        char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t'))
        {
            char* q = p + 1;
            for (int k = 0; k < 3; k++) out.pos.push_back(strtof(q, &q));
        }
        else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t'))
        {
            char* q = p + 2;
            for (int k = 0; k < 3; k++) vn.push_back(strtof(q, &q));
        }
        else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t'))
        {
            char* q = p + 2;
            for (int k = 0; k < 2; k++) vt.push_back(strtof(q, &q));
        }
        else if (strncmp(p, "mtllib", 6) == 0 && (p[6] == ' ' || p[6] == '\t'))
        {
            char name[1024] = { 0 };
            if (sscanf(p + 6, " %1023s", name) == 1) parse_mtl(dir + name, dir, texOf);
        }
        else if (strncmp(p, "usemtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t'))
        {
            if (splitByMaterial)
            {
                char name[1024] = { 0 };
                if (sscanf(p + 6, " %1023s", name) == 1) cur = name;
            }
        }
        else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'))
        {
            vs.clear(); ns.clear(); ts.clear();
            char* q = p + 1;
            for (;;)
            {
                while (*q == ' ' || *q == '\t') ++q;
                if (*q == 0 || *q == '\n' || *q == '\r') break;
                char* e;
                long v = strtol(q, &e, 10);
                if (e == q) break;
                q = e;
                long n = 0, t = 0; bool hasN = false, hasT = false;
                if (*q == '/')
                {
                    ++q;
                    if (*q != '/') { char* e1; t = strtol(q, &e1, 10); hasT = e1 != q; q = e1; }
                    if (*q == '/') { ++q; char* e2; n = strtol(q, &e2, 10); hasN = e2 != q; q = e2; }
                }
                // indices are validated the moment they are resolved (a negative or huge one must never size or index a table)
                int rv = resolve(v, out.pos.size() / 3), rn = hasN ? resolve(n, vn.size() / 3) : -1, rt = hasT ? resolve(t, vt.size() / 2) : -1;
                if (rv < 0 || (size_t)rv >= out.pos.size() / 3) { fclose(f); err = "face index out of range"; return false; }
                // a normal / texture-coordinate reference outside the tables read so far is ignored (files in the wild name `vt` indices
                // without any `vt` line); resolve() has already turned it into -1, so it can never size or index a table
                vs.push_back(rv); ns.push_back(rn); ts.push_back(rt);
            }
            if (vs.size() < 3) continue;
            auto it = groupOf.find(cur);
            if (it == groupOf.end()) { groupOf[cur] = out.groups.size(); out.groups.push_back({ cur, {} }); it = groupOf.find(cur); }
            std::vector<uint32_t>& g = out.groups[it->second].idx;
            for (size_t i = 1; i + 1 < vs.size(); i++) { g.push_back((uint32_t)vs[0]); g.push_back((uint32_t)vs[i]); g.push_back((uint32_t)vs[i + 1]); }
            for (size_t i = 0; i < vs.size(); i++)
            {
                if (ns[i] >= 0)
                {
                    if ((size_t)vs[i] >= vnOfV.size()) vnOfV.resize(vs[i] + 1, -1);
                    if (vnOfV[vs[i]] < 0) vnOfV[vs[i]] = ns[i];
                }
                if (ts[i] >= 0 && vs[i] >= 0)
                {
                    if ((size_t)vs[i] >= vtOfV.size()) vtOfV.resize(vs[i] + 1, -1);
                    if (vtOfV[vs[i]] < 0) vtOfV[vs[i]] = ts[i];
                }
            }
        }
    }
    fclose(f);
    size_t V = out.pos.size() / 3;
    for (auto& g : out.groups) for (uint32_t i : g.idx) if (i >= V) { err = "face index out of range"; return false; }
    bool any = false;
    for (int n : vnOfV) if (n >= 0) { any = true; break; }
    if (!vn.empty() && any)
    {
        out.nrm.assign(V * 3, 0.f);
        for (size_t v = 0; v < V; v++) out.nrm[3 * v + 2] = 1.f;
        for (size_t v = 0; v < vnOfV.size() && v < V; v++)
            if (vnOfV[v] >= 0 && (size_t)vnOfV[v] < vn.size() / 3) { out.nrm[3 * v] = vn[3 * vnOfV[v]]; out.nrm[3 * v + 1] = vn[3 * vnOfV[v] + 1]; out.nrm[3 * v + 2] = vn[3 * vnOfV[v] + 2]; }
    }
    any = false;
    for (int t : vtOfV) if (t >= 0) { any = true; break; }
    if (!vt.empty() && any)
    {
        out.uv.assign(V * 2, 0.f);
        for (size_t v = 0; v < vtOfV.size() && v < V; v++)
            if (vtOfV[v] >= 0 && (size_t)vtOfV[v] < vt.size() / 2) { out.uv[2 * v] = vt[2 * vtOfV[v]]; out.uv[2 * v + 1] = vt[2 * vtOfV[v] + 1]; }
    }
    for (auto& g : out.groups) { auto it = texOf.find(g.name); if (it != texOf.end()) g.texture = it->second; }
    return true;
}

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
}

struct beamObj { ObjFile file; std::string error; };

extern "C"
{
void* beamObjLoad(const char* path, int splitByMaterial)
{
    beamObj* o = new beamObj();
    try
    {
        if (!path || !parse_obj(path, splitByMaterial != 0, o->file, o->error)) { if (!path) o->error = "null path"; o->file = ObjFile(); }
    }
    catch (const std::exception& ex) { o->error = std::string("OBJ loader: ") + ex.what(); o->file = ObjFile(); }     // no exception crosses the C ABI
    catch (...) { o->error = "OBJ loader: unknown failure"; o->file = ObjFile(); }
    return o;
}
const char* beamObjError(void* h) { beamObj* o = (beamObj*)h; return o->error.empty() ? nullptr : o->error.c_str(); }
void beamObjFree(void* h) { delete (beamObj*)h; }
uint32_t beamObjNumVertices(void* h) { return (uint32_t)(((beamObj*)h)->file.pos.size() / 3); }
uint32_t beamObjNumMeshes(void* h) { return (uint32_t)((beamObj*)h)->file.groups.size(); }
int beamObjHasNormals(void* h) { return !((beamObj*)h)->file.nrm.empty(); }
const float* beamObjPositions(void* h) { return ((beamObj*)h)->file.pos.data(); }
const float* beamObjNormals(void* h) { beamObj* o = (beamObj*)h; return o->file.nrm.empty() ? nullptr : o->file.nrm.data(); }
uint32_t beamObjMeshNumIndices(void* h, uint32_t mesh) { beamObj* o = (beamObj*)h; return mesh < o->file.groups.size() ? (uint32_t)o->file.groups[mesh].idx.size() : 0; }
const uint32_t* beamObjMeshIndices(void* h, uint32_t mesh) { beamObj* o = (beamObj*)h; return mesh < o->file.groups.size() ? o->file.groups[mesh].idx.data() : nullptr; }
const char* beamObjMeshName(void* h, uint32_t mesh) { beamObj* o = (beamObj*)h; return mesh < o->file.groups.size() ? o->file.groups[mesh].name.c_str() : nullptr; }

int beamObjHasUVs(void* h) { return !((beamObj*)h)->file.uv.empty(); }
const float* beamObjUVs(void* h) { beamObj* o = (beamObj*)h; return o->file.uv.empty() ? nullptr : o->file.uv.data(); }
const char* beamObjMeshTexture(void* h, uint32_t mesh)
{
    beamObj* o = (beamObj*)h;
    return (mesh < o->file.groups.size() && !o->file.groups[mesh].texture.empty()) ? o->file.groups[mesh].texture.c_str() : nullptr;
}

// Windows BMP, BI_RGB 24 or 32 bits per pixel (Content/F16s.bmp, F16t.bmp are 24-bit), bottom-up or top-down rows
uint32_t* beamBmpLoad(const char* path, uint32_t* width, uint32_t* height)
{
    if (!path || !width || !height) return nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return nullptr;
    unsigned char hdr[54];
    uint32_t* out = nullptr;
    if (fread(hdr, 1, 54, f) == 54 && hdr[0] == 'B' && hdr[1] == 'M')
    {
        uint32_t off = rd32(hdr + 10), comp = rd32(hdr + 30);
        int32_t w = (int32_t)rd32(hdr + 18), h = (int32_t)rd32(hdr + 22);
        uint32_t bpp = (uint32_t)hdr[28] | ((uint32_t)hdr[29] << 8);
        bool topDown = h < 0;
        if (topDown) h = -h;
        if (w > 0 && h > 0 && w <= 32768 && h <= 32768 && comp == 0 && (bpp == 24 || bpp == 32) && fseek(f, (long)off, SEEK_SET) == 0)
        {
            size_t stride = (((size_t)w * bpp / 8) + 3) & ~(size_t)3;
            std::vector<unsigned char> row(stride);
            out = (uint32_t*)malloc((size_t)w * h * sizeof(uint32_t));
            for (int32_t y = 0; out && y < h; y++)
            {
                if (fread(row.data(), 1, stride, f) != stride) { free(out); out = nullptr; break; }
                uint32_t* dst = out + (size_t)(topDown ? y : h - 1 - y) * w;
                for (int32_t x = 0; x < w; x++)
                {
                    const unsigned char* px = row.data() + (size_t)x * (bpp / 8);          // B, G, R [, A]
                    dst[x] = ((uint32_t)px[2] << 16) | ((uint32_t)px[1] << 8) | (uint32_t)px[0];
                }
            }
            if (out) { *width = (uint32_t)w; *height = (uint32_t)h; }
        }
    }
    fclose(f);
    return out;
}
void beamImageFree(uint32_t* texels) { free(texels); }

// Model::load (TestProgram/Model.cpp:26-126) in one call: every mesh of the file becomes an IMesh added to `scene` (positions slot 0,
// normals slot 1 when present).  Returns the number of meshes added, or -1.
int beamSceneLoadObj(void* scene, const char* path, int splitByMaterial)
{
    void* h = beamObjLoad(path, splitByMaterial);
    if (beamObjError(h)) { beamObjFree(h); return -1; }
    uint32_t V = beamObjNumVertices(h), M = beamObjNumMeshes(h);
    int added = 0;
    for (uint32_t m = 0; m < M; m++)
    {
        void* mesh = beamMeshCreate();
        uint32_t e = beamMeshSetIndices(mesh, beamObjMeshIndices(h, m), beamObjMeshNumIndices(h, m), 0);
        if (!e) e = beamMeshSetVertexData(mesh, beamObjPositions(h), V, 3, 0, 0);
        if (!e && beamObjHasNormals(h)) e = beamMeshSetVertexData(mesh, beamObjNormals(h), V, 3, 1, 0);
        if (!e && beamObjHasUVs(h)) e = beamMeshSetVertexData(mesh, beamObjUVs(h), V, 2, 2, 0);
        if (!e && beamObjMeshTexture(h, m))
        {
            uint32_t tw = 0, th = 0;
            uint32_t* tex = beamBmpLoad(beamObjMeshTexture(h, m), &tw, &th);     // a missing or unsupported image leaves the mesh untextured
            if (tex) { e = beamMeshSetTexture(mesh, tex, tw, th); beamImageFree(tex); }
        }
        if (!e) { beamSceneAddMesh(scene, mesh); added++; }
        beamMeshDestroy(mesh);               // the scene holds its own reference
        if (e) { beamObjFree(h); return -1; }
    }
    beamObjFree(h);
    return added;
}

// Headless stand-in for the viewer's present path (GLinterop.h:24-36 uploads the u32 buffer as BGRA8, i.e. bits 16-23 = red):
// writes the pixels as a binary PPM.  hostPixels = width*height u32 in row-major order (IRenderTarget::readPixels).
int beamWritePPM(const char* path, const uint32_t* hostPixels, uint32_t width, uint32_t height)
{
    if (!path || !hostPixels || !width || !height) return 2;
    FILE* f = fopen(path, "wb");
    if (!f) return 4;
    fprintf(f, "P6\n%u %u\n255\n", width, height);
    std::vector<unsigned char> row(3 * (size_t)width);
    for (uint32_t y = 0; y < height; y++)
    {
        for (uint32_t x = 0; x < width; x++)
        {
            uint32_t p = hostPixels[(size_t)y * width + x];
            row[3 * x] = (unsigned char)((p >> 16) & 0xFF); row[3 * x + 1] = (unsigned char)((p >> 8) & 0xFF); row[3 * x + 2] = (unsigned char)(p & 0xFF);
        }
        fwrite(row.data(), 1, row.size(), f);
    }
    fclose(f);
    return 0;
}
}
