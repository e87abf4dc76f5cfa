// oracle/ref_shim/ref_trimesh.cpp -- TEST INFRASTRUCTURE ONLY.
// TriMesh::computeNormals of the reference executed as written (src/librender/trimesh.cpp:606-676, angle-weighted vertex normals after
// Thuermer & Wuethrich) with unitAngle (include/mitsuba/core/util.h:309-314), both cut out at build time (oracle/_ref/ref_trimesh_*.inc)
// into a class that supplies the members they touch.  The `obj` loader runs it on every mesh without `vn` records (obj.cpp:380-383).
// Part of oracle/_ref/libref_geom.so.
#include <cstring>
#include <string>
#include "mitsuba_shim.h"

namespace mitsuba {
#include "ref_trimesh_unitangle.inc"
struct Triangle { uint32_t idx[3]; };                  // include/mitsuba/core/triangle.h:34-36
struct TriMesh {
    std::string m_name = "mesh";
    Point *m_positions = nullptr; Normal *m_normals = nullptr; Triangle *m_triangles = nullptr;
    size_t m_triangleCount = 0, m_vertexCount = 0;
    bool m_faceNormals = false, m_flipNormals = false;
    void computeNormals(bool force = false);
};
#include "ref_trimesh_normals.inc"
}
// positions: 3 floats per vertex; triangles: 3 indices each (swapped in place when face normals + flip); normals_out: 3 floats per vertex
// (NULL when the mesh ends up without vertex normals).  Returns 1 when vertex normals exist afterwards.
extern "C" int ref_compute_normals(const float *positions, size_t nVerts, uint32_t *triangles, size_t nTris, int faceNormals, int flipNormals, float *normals_out) {
    using namespace mitsuba;
    TriMesh m;
    m.m_positions = new Point[nVerts];
    for (size_t i = 0; i < nVerts; ++i) m.m_positions[i] = Point(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
    m.m_triangles = reinterpret_cast<Triangle *>(triangles);
    m.m_triangleCount = nTris; m.m_vertexCount = nVerts; m.m_faceNormals = faceNormals != 0; m.m_flipNormals = flipNormals != 0;
    m.computeNormals();
    const int has = m.m_normals != nullptr;
    if (has && normals_out) for (size_t i = 0; i < nVerts; ++i) { normals_out[3 * i] = m.m_normals[i].x; normals_out[3 * i + 1] = m.m_normals[i].y; normals_out[3 * i + 2] = m.m_normals[i].z; }
    delete[] m.m_positions; delete[] m.m_normals;
    return has;
}
