// oracle/ref_shim/ref_obj.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The text side of the reference's Wavefront OBJ loader executed as written: oracle/Makefile cuts these pieces out of /root/reference at
// build time (oracle/_ref/ref_obj_*.inc, ref_cam_*.inc) and this file pastes them into a class that only supplies what they touch:
//   src/shapes/obj.cpp :155-163 OBJTriangle, :165-187 fetch_line (trailing blanks, backslash continuation), :245-328 the line loop of
//                      WavefrontOBJ(props) with the closing createMesh, :371-390 parse (v, v/vt, v//vn, v/vt/vn),
//                      :577-715 Vertex, vertex_key_order, createMesh (negative indices, bounds errors, toWorld, merge of equal vertices)
//   src/libcore/util.cpp :83-104 tokenize, trim
//   include/mitsuba/core/transform.h :108-125 operator()(Point), :203-211 operator()(Normal) (inverse transpose);
//   include/mitsuba/core/matrix.inl :138-193 Matrix::invert -- Transform(const Matrix4x4 &), what a <matrix> element of a scene file runs
// Point and Normal are distinct types here.  `collapse` is true (one mesh per file, as the product loads it).  TriMesh::computeNormals,
// which Shape::configure() runs on the result, is pinned separately (ref_trimesh.cpp).  Part of oracle/_ref/libref_geom.so.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace refobj {
typedef float Float;
enum ELogLevel { EDebug, EInfo, EWarn, EError };
#define SIZE_T_FMT "%zu"
#define MTS_EXPORT_CORE
#define BOOST_STATIC_ASSERT(x) static_assert(x, "")
inline std::string formatString(const char *fmt, ...) { char buf[1024]; va_list a; va_start(a, fmt); vsnprintf(buf, sizeof(buf), fmt, a); va_end(a); return buf; }
#undef Log
#undef SLog
#define Log(level, ...) do { if (level >= EError) throw std::runtime_error(formatString(__VA_ARGS__)); } while (0)   // logger.cpp:100,147: EError throws
#define SLog(level, ...) Log(level, __VA_ARGS__)

struct Normal { Float x, y, z; Normal() : x(0), y(0), z(0) {} Normal(Float x, Float y, Float z) : x(x), y(y), z(z) {} explicit Normal(Float v) : x(v), y(v), z(v) {}
    bool isZero() const { return x == 0 && y == 0 && z == 0; }
    Float length() const { return (Float) std::sqrt((Float) (x * x + y * y + z * z)); }
    Normal operator/(Float f) const { Float recip = (Float) 1 / f; return Normal(x * recip, y * recip, z * recip); } };        // vector.h
inline Normal normalize(const Normal &n) { return n / n.length(); }
struct Point { Float x, y, z; Point() : x(0), y(0), z(0) {} Point(Float x, Float y, Float z) : x(x), y(y), z(z) {}
    Point operator/(Float f) const { Float recip = (Float) 1 / f; return Point(x * recip, y * recip, z * recip); } };
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x, Float y) : x(x), y(y) {} explicit Point2(Float v) : x(v), y(v) {} };
struct AABB { void expandBy(const Point &) {} };
struct Triangle { uint32_t idx[3]; };                                                      // include/mitsuba/core/triangle.h:34-36
struct Timer { int getMilliseconds() const { return 0; } };
template <typename T> struct ref { T *p; ref(T *p = nullptr) : p(p) {} T *operator->() const { return p; } operator T *() const { return p; } };

template <int M, int N, typename T> struct Matrix { T m[M][N]; bool invert(Matrix &target) const; };
#include "ref_cam_invert.inc"
struct Matrix4x4 : public Matrix<4, 4, Float> { std::string toString() const { return "matrix"; } };
struct Transform {
    Matrix4x4 m_transform, m_invTransform;
    explicit Transform(const Matrix4x4 &trafo) : m_transform(trafo) {                                          // transform.h:50-55
        bool success = m_transform.invert(m_invTransform);
        if (!success) SLog(EError, "Unable to invert singular matrix %s", trafo.toString().c_str()); }
#include "ref_obj_apply.inc"
};

#include "ref_obj_util.inc"

struct TriMesh {
    std::string name; std::vector<Triangle> tris; std::vector<Point> pos; std::vector<Normal> nrm; std::vector<Point2> uv; AABB aabb;
    bool hasNormals, hasTexcoords;
    TriMesh(const std::string &name, size_t triangleCount, size_t vertexCount, bool hasNormals, bool hasTexcoords, bool, bool, bool)
        : name(name), tris(triangleCount), pos(vertexCount), nrm(hasNormals ? vertexCount : 0), uv(hasTexcoords ? vertexCount : 0), hasNormals(hasNormals), hasTexcoords(hasTexcoords) {}
    Triangle *getTriangles() { return tris.data(); } Point *getVertexPositions() { return pos.data(); }
    Normal *getVertexNormals() { return nrm.data(); } Point2 *getVertexTexcoords() { return uv.data(); } AABB &getAABB() { return aabb; }
    void incRef() {}
};
namespace fs { struct path { std::string p; path() {} path(const std::string &s) : p(s) {} bool empty() const { return p.empty(); } }; }
struct FileResolver { fs::path resolve(const std::string &s) const { return fs::path(s); } };

struct WavefrontOBJ {
    std::string m_name = "obj"; bool m_collapse = true, m_flipNormals = false, m_faceNormals = false;
    std::vector<TriMesh *> m_meshes; std::vector<std::string> m_materialAssignment;
    ~WavefrontOBJ() { for (TriMesh *m : m_meshes) delete m; }
#include "ref_obj_members.inc"
    void load(std::istream &is, const Transform &objectToWorld, bool flipTexCoords) {
        FileResolver resolver, *fileResolver = &resolver;
        const int shapeIndex = -1;
        // locals of WavefrontOBJ(props), obj.cpp:232-243
        std::string buf;
        std::vector<Point> vertices;
        std::vector<Normal> normals;
        std::vector<Point2> texcoords;
        std::vector<OBJTriangle> triangles;
        std::string name = m_name, line;
        std::set<std::string> geomNames;
        std::vector<Vertex> vertexBuffer;
        fs::path materialLibrary;
        int geomIndex = 0;
        bool nameBeforeGeometry = false;
        std::string materialName;
#include "ref_obj_loop.inc"
        (void) nameBeforeGeometry; (void) fileResolver;
    }
};
}

// Returns 0 and the mesh counts (then ref_obj_copy + ref_obj_free), or -1 with the message of the reference's Log(EError) in err.
extern "C" int ref_obj_load(const char *filename, const float toWorld[16], int flipTexCoords, void **handle, size_t *nVerts, size_t *nTris, int *hasNormals, int *hasTexcoords,
                            char *err, size_t errCap) {
    using namespace refobj;
    WavefrontOBJ *o = new WavefrontOBJ();
    try {
        std::ifstream is(filename);
        if (is.bad() || is.fail()) Log(EError, "Wavefront OBJ file '%s' not found!", filename);
        Matrix4x4 m; std::memcpy(m.m, toWorld, sizeof(m.m));
        o->load(is, Transform(m), flipTexCoords != 0);
        if (o->m_meshes.size() != 1) Log(EError, "expected one collapsed mesh, got %zu", o->m_meshes.size());
    } catch (const std::exception &e) {
        if (err && errCap) { std::strncpy(err, e.what(), errCap - 1); err[errCap - 1] = 0; }
        delete o; *handle = nullptr; return -1;
    }
    TriMesh *t = o->m_meshes[0];
    *handle = o; *nVerts = t->pos.size(); *nTris = t->tris.size(); *hasNormals = t->hasNormals; *hasTexcoords = t->hasTexcoords;
    return 0;
}
extern "C" void ref_obj_copy(void *handle, float *positions, float *normals, float *texcoords, uint32_t *triangles) {
    refobj::TriMesh *t = static_cast<refobj::WavefrontOBJ *>(handle)->m_meshes[0];
    if (positions) std::memcpy(positions, t->pos.data(), t->pos.size() * 12);
    if (normals && t->hasNormals) std::memcpy(normals, t->nrm.data(), t->nrm.size() * 12);
    if (texcoords && t->hasTexcoords) std::memcpy(texcoords, t->uv.data(), t->uv.size() * 8);
    if (triangles) std::memcpy(triangles, t->tris.data(), t->tris.size() * 12);
}
extern "C" void ref_obj_free(void *handle) { delete static_cast<refobj::WavefrontOBJ *>(handle); }
