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
#include <map>
#include <string>
#include <vector>
#include "../../include/beam_host_c.h"

namespace
{
struct ObjGroup { std::string name; std::vector<uint32_t> idx; };
struct ObjFile
{
    std::vector<float> pos, nrm;          // nrm: per position (3 floats each) or empty
    std::vector<ObjGroup> groups;
};

int resolve(long v, size_t count) { return v > 0 ? (int)(v - 1) : (int)((long)count + v); }

bool parse_obj(const char* path, bool splitByMaterial, ObjFile& out, std::string& err)
{
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<float> vn;
    std::vector<int> vnOfV;               // normal index per position, -1 = unset
    std::map<std::string, size_t> groupOf;
    std::string cur = "__default__";
    char line[4096];
    std::vector<int> vs, ns;
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
            vs.clear(); ns.clear();
            char* q = p + 1;
            for (;;)
            {
                while (*q == ' ' || *q == '\t') ++q;
                if (*q == 0 || *q == '\n' || *q == '\r') break;
                char* e;
                long v = strtol(q, &e, 10);
                if (e == q) break;
                q = e;
                long n = 0; bool hasN = false;
                if (*q == '/')
                {
                    ++q;
                    if (*q != '/') strtol(q, &q, 10);          // texture index, unused
                    if (*q == '/') { ++q; char* e2; n = strtol(q, &e2, 10); hasN = e2 != q; q = e2; }
                }
                vs.push_back(resolve(v, out.pos.size() / 3));
                ns.push_back(hasN ? resolve(n, vn.size() / 3) : -1);
            }
            if (vs.size() < 3) continue;
            auto it = groupOf.find(cur);
            if (it == groupOf.end()) { groupOf[cur] = out.groups.size(); out.groups.push_back({ cur, {} }); it = groupOf.find(cur); }
            std::vector<uint32_t>& g = out.groups[it->second].idx;
            for (size_t i = 1; i + 1 < vs.size(); i++) { g.push_back((uint32_t)vs[0]); g.push_back((uint32_t)vs[i]); g.push_back((uint32_t)vs[i + 1]); }
            for (size_t i = 0; i < vs.size(); i++)
                if (ns[i] >= 0)
                {
                    if ((size_t)vs[i] >= vnOfV.size()) vnOfV.resize(vs[i] + 1, -1);
                    if (vnOfV[vs[i]] < 0) vnOfV[vs[i]] = ns[i];
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
    return true;
}
}

struct beamObj { ObjFile file; std::string error; };

extern "C"
{
void* beamObjLoad(const char* path, int splitByMaterial)
{
    beamObj* o = new beamObj();
    if (!path || !parse_obj(path, splitByMaterial != 0, o->file, o->error)) { if (!path) o->error = "null path"; o->file = ObjFile(); }
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
