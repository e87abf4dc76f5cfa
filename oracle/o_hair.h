// oracle/o_hair.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Restates the HairShape geometry path of the reference: file loader + vertex merge
// (src/shapes/hair.cpp:609-785), segment/miter helpers (:551-596), segment bounds (:246-286,
// :368-397), the FP64 mitred-cylinder test (:485-542), the two-level closest/any query
// (src/librender/skdtree.cpp:112-142,207-226; hair.cpp:200-237) and the intersection frame
// (hair.cpp:825-862; include/mitsuba/render/skdtree.h:426-427).
// Triangle meshes in the same scene (T1): TriAccel::load/rayIntersect (include/mitsuba/render/triaccel.h:61-158),
// the triangle branch of ShapeKDTree::intersect (skdtree.h:293-304,330-338) and of fillIntersectionRecord
// (skdtree.h:346-427), Triangle::getAABB (include/mitsuba/core/triangle.h:40-45).
// The reference's SAH kd-trees are replaced by (a) a brute-force loop in segment order and
// (b) a plain binned-SAH BVH used only to make oracle renders finish; both apply the
// reference's interval logic per primitive test, so results agree except for equal-t ties.
#pragma once
#include "o_math.h"
#include <cstdio>
#include <fstream>
#include <sstream>
#include <thread>
#include <atomic>
#include <mutex>

namespace orc {

struct AABB {
    V3 mn, mx;
    AABB() : mn(kInf), mx(-kInf) {}
    void expand(const V3 &p) {
        mn = V3(std::min(mn.x, p.x), std::min(mn.y, p.y), std::min(mn.z, p.z));
        mx = V3(std::max(mx.x, p.x), std::max(mx.y, p.y), std::max(mx.z, p.z));
    }
    void expand(const AABB &b) { expand(b.mn); expand(b.mx); }
    V3 center() const { return (mx + mn) * 0.5f; } // aabb.h getCenter
    // include/mitsuba/core/aabb.h:308-338
    bool rayIntersect(const V3 &o, const V3 &d, const V3 &dRcp, float &nearT, float &farT) const {
        nearT = -kInf; farT = kInf;
        for (int i = 0; i < 3; i++) {
            const float origin = o[i], minVal = mn[i], maxVal = mx[i];
            if (d[i] == 0) {
                if (origin < minVal || origin > maxVal) return false;
            } else {
                float t1 = (minVal - origin) * dRcp[i];
                float t2 = (maxVal - origin) * dRcp[i];
                if (t1 > t2) std::swap(t1, t2);
                nearT = std::max(t1, nearT);
                farT = std::min(t2, farT);
                if (!(nearT <= farT)) return false;
            }
        }
        return true;
    }
};

// KDTreeBase::buildInternal, include/mitsuba/render/gkdtree.h:1213-1220: after the build the tree's bounding box is enlarged by
// MTS_KD_AABB_EPSILON = 1e-3 (gkdtree.h:50), relative + absolute ("necessary e.g. when the scene is planar"); the second line
// already sees the lowered minimum.  getAABB() returns this enlarged box -- it is what HairKDTree::rayIntersect clips rays against
// (hair.cpp:205), what HairShape::getAABB() hands to the scene-level tree (hair.cpp:944-946), and, enlarged once more after the
// scene-level build, what ShapeKDTree::rayIntersect clips against (skdtree.cpp:124).
static inline void enlargeKDTreeBounds(AABB &a) {
    const float eps = 1e-3f;
    a.mn = a.mn - ((a.mx - a.mn) * eps + V3(eps));
    a.mx = a.mx + ((a.mx - a.mn) * eps + V3(eps));
}

struct Ray {
    V3 o, d, dRcp;
    float mint, maxt;
    Ray() {}
    Ray(V3 o_, V3 d_, float mint_ = kEpsilon, float maxt_ = kInf) : o(o_), d(d_), mint(mint_), maxt(maxt_) {
        dRcp = V3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // include/mitsuba/core/ray.h:72-93
    }
};

// include/mitsuba/render/triaccel.h:29-158 (Wald's projection test, 48 bytes per triangle)
struct TriAccel {
    uint32_t k = 3;
    float n_u = 0, n_v = 0, n_d = 0, a_u = 0, a_v = 0, b_nu = 0, b_nv = 0, c_nu = 0, c_nv = 0;
    uint32_t shapeIndex = 0, primIndex = 0;

    int load(const V3 &A, const V3 &B, const V3 &C) { // :61-97
        static const int waldModulo[4] = {1, 2, 0, 1};
        V3 b = C - A, c = B - A, N = cross(c, b);
        k = 0;
        for (int j = 0; j < 3; j++) if (std::abs(N[j]) > std::abs(N[k])) k = (uint32_t) j;
        uint32_t u = (uint32_t) waldModulo[k], v = (uint32_t) waldModulo[k + 1];
        const float n_k = N[k], denom = b[u] * c[v] - b[v] * c[u];
        if (denom == 0) { k = 3; return 1; }
        n_u = N[u] / n_k; n_v = N[v] / n_k; n_d = dot(A, N) / n_k;
        b_nu = b[u] / denom; b_nv = -b[v] / denom;
        a_u = A[u]; a_v = A[v];
        c_nu = c[v] / denom; c_nv = -c[u] / denom;
        return 0;
    }
    bool rayIntersect(const V3 &ro, const V3 &rd, float mint, float maxt, float &u, float &v, float &t) const { // :99-158
        float o_u, o_v, o_k, d_u, d_v, d_k;
        switch (k) {
            case 0: o_u = ro.y; o_v = ro.z; o_k = ro.x; d_u = rd.y; d_v = rd.z; d_k = rd.x; break;
            case 1: o_u = ro.z; o_v = ro.x; o_k = ro.y; d_u = rd.z; d_v = rd.x; d_k = rd.y; break;
            case 2: o_u = ro.x; o_v = ro.y; o_k = ro.z; d_u = rd.x; d_v = rd.y; d_k = rd.z; break;
            default: return false;
        }
        t = (n_d - o_u * n_u - o_v * n_v - o_k) / (d_u * n_u + d_v * n_v + d_k);
        if (t < mint || t > maxt) return false;
        const float hu = o_u + t * d_u - a_u;
        const float hv = o_v + t * d_v - a_v;
        u = hv * b_nu + hu * b_nv;
        v = hu * c_nu + hv * c_nv;
        return u >= 0 && v >= 0 && u + v <= 1.0f;
    }
};

// A triangle mesh as ShapeKDTree sees it (TriMesh::getVertexPositions/Normals/Triangles after configure()).
// The `rectangle` shape (src/shapes/rectangle.cpp:79-183): the unit square [-1,1]^2 x {0} under toWorld, one primitive of the
// top-level tree, intersected analytically in object space (no triangles).
struct Rectangle {
    M44 objectToWorld = M44::identity(), worldToObject = M44::identity();
    V3 dpdu, dpdv; Frame frame;
    // ctor :81-86 + configure :100-112.  `toWorldInv` is the inverse Transform() carries (Gauss-Jordan for a <matrix>)
    void configure(const M44 &toWorld, const M44 &toWorldInv, bool flipNormals) {
        objectToWorld = toWorld; worldToObject = toWorldInv;
        if (flipNormals) {                                                   // m_objectToWorld * Transform::scale(Vector(1, 1, -1)), transform.cpp:28-31,66-80
            M44 sc = M44::identity(); sc.m[2][2] = -1.0f;
            M44 scInv = M44::identity(); scInv.m[2][2] = 1.0f / -1.0f;
            objectToWorld = mul(toWorld, sc); worldToObject = mul(scInv, toWorldInv);
        }
        dpdu = xfmVector(objectToWorld, V3(2, 0, 0)); dpdv = xfmVector(objectToWorld, V3(0, 2, 0));
        // Transform::operator()(Normal): the transpose of the inverse (transform.h:203-211)
        const M44 &iv = worldToObject; const V3 nl(0, 0, 1);
        const V3 nw(iv.m[0][0] * nl.x + iv.m[1][0] * nl.y + iv.m[2][0] * nl.z, iv.m[0][1] * nl.x + iv.m[1][1] * nl.y + iv.m[2][1] * nl.z,
                    iv.m[0][2] * nl.x + iv.m[1][2] * nl.y + iv.m[2][2] * nl.z);
        frame.s = normalize(dpdu); frame.t = normalize(dpdv); frame.n = normalize(nw);
        if (std::abs(dot(normalize(dpdu), normalize(dpdv))) > kEpsilon) throw std::runtime_error("Error: 'toWorld' transformation contains shear!");
    }
    AABB getAABB() const {                                                   // :114-121
        AABB a; a.expand(xfmPoint(objectToWorld, V3(-1, -1, 0))); a.expand(xfmPoint(objectToWorld, V3(1, -1, 0)));
        a.expand(xfmPoint(objectToWorld, V3(1, 1, 0))); a.expand(xfmPoint(objectToWorld, V3(-1, 1, 0))); return a;
    }
    // :127-151 after Transform::transformAffine(Ray) (transform.h:128-146,292-307)
    bool rayIntersect(const V3 &ro, const V3 &rd, float mint, float maxt, float &lx, float &ly, float &t) const {
        const M44 &w = worldToObject;
        const V3 o(w.m[0][0] * ro.x + w.m[0][1] * ro.y + w.m[0][2] * ro.z + w.m[0][3], w.m[1][0] * ro.x + w.m[1][1] * ro.y + w.m[1][2] * ro.z + w.m[1][3],
                   w.m[2][0] * ro.x + w.m[2][1] * ro.y + w.m[2][2] * ro.z + w.m[2][3]);
        const V3 d = xfmVector(w, rd);
        const float hit = -o.z / d.z;
        if (!(hit >= mint && hit <= maxt)) return false;
        const V3 local = o + d * hit;                                        // Ray::operator()(t) = o + t * d
        if (std::abs(local.x) <= 1 && std::abs(local.y) <= 1) { t = hit; lx = local.x; ly = local.y; return true; }
        return false;
    }
};

struct TriMesh {
    std::vector<V3> pos, nrm;       // nrm empty = face normals (skdtree.h:389-391)
    std::vector<float> uv;          // 2 per vertex; empty = no texture coordinates (its.uv = barycentrics, skdtree.h:399-406)
    std::vector<uint32_t> idx;      // 3 per triangle
    std::vector<TriAccel> accel;    // skdtree.cpp:78-110
    bool isRect = false; Rectangle rect;   // a `rectangle` shape rides in the mesh slot of the shape list as ONE analytic primitive
    size_t triCount() const { return isRect ? 1 : idx.size() / 3; }
    AABB triAABB(uint32_t j) const { // triangle.h:40-45
        if (isRect) return rect.getAABB();
        AABB r; r.expand(pos[idx[3 * j]]); r.expand(pos[idx[3 * j + 1]]); r.expand(pos[idx[3 * j + 2]]); return r;
    }
    // the primitive test of the top-level tree (skdtree.h:293-312): TriAccel for triangles, Shape::rayIntersect otherwise
    bool intersectPrim(uint32_t j, const V3 &ro, const V3 &rd, float mint, float maxt, float &u, float &v, float &t) const {
        if (isRect) return rect.rayIntersect(ro, rd, mint, maxt, u, v, t);
        return accel[j].rayIntersect(ro, rd, mint, maxt, u, v, t);
    }
};

struct HairShape {
    bool isMesh = false;              // a TriMesh entry of the shape list (then only `mesh`, `bsdf`, `aabb` are used)
    TriMesh mesh;
    std::vector<V3> verts;
    std::vector<uint8_t> startsFiber; // size verts.size()+1, sentinel = 1 (hair.cpp:782)
    std::vector<uint32_t> segIndex;   // iv of each segment (hair.cpp:117-124)
    float radius = 0;
    int bsdf = 0;
    AABB aabb;                        // union of segment bounds (gkdtree.h:997-1002)

    // --- hair.cpp:551-596 ---
    V3 firstVertex(uint32_t iv) const { return verts[iv]; }
    V3 secondVertex(uint32_t iv) const { return verts[iv + 1]; }
    bool prevSegmentExists(uint32_t iv) const { return !startsFiber[iv]; }
    bool nextSegmentExists(uint32_t iv) const { return !startsFiber[iv + 2]; }
    V3 tangent(uint32_t iv) const { return normalize(verts[iv + 1] - verts[iv]); }
    D3 tangentD(uint32_t iv) const { return normalize(D3(verts[iv + 1]) - D3(verts[iv])); }
    V3 prevTangent(uint32_t iv) const { return normalize(verts[iv] - verts[iv - 1]); }
    D3 prevTangentD(uint32_t iv) const { return normalize(D3(verts[iv]) - D3(verts[iv - 1])); }
    V3 nextTangent(uint32_t iv) const { return normalize(verts[iv + 2] - verts[iv + 1]); }
    D3 nextTangentD(uint32_t iv) const { return normalize(D3(verts[iv + 2]) - D3(verts[iv + 1])); }
    V3 firstMiterNormal(uint32_t iv) const {
        return prevSegmentExists(iv) ? normalize(prevTangent(iv) + tangent(iv)) : tangent(iv);
    }
    V3 secondMiterNormal(uint32_t iv) const {
        return nextSegmentExists(iv) ? normalize(tangent(iv) + nextTangent(iv)) : tangent(iv);
    }
    D3 firstMiterNormalD(uint32_t iv) const {
        return prevSegmentExists(iv) ? normalize(prevTangentD(iv) + tangentD(iv)) : tangentD(iv);
    }
    D3 secondMiterNormalD(uint32_t iv) const {
        return nextSegmentExists(iv) ? normalize(tangentD(iv) + nextTangentD(iv)) : tangentD(iv);
    }

    // hair.cpp:246-286 (intersectCylPlane)
    static bool intersectCylPlane(V3 planePt, V3 planeNrml, V3 cylPt, V3 cylD, float radius,
                                  V3 &center, V3 *axes, float *lengths) {
        if (std::abs(dot(planeNrml, cylD)) < kEpsilon) return false;
        V3 B, A = cylD - dot(cylD, planeNrml) * planeNrml;
        float len = length(A);
        bool same = planeNrml.x == cylD.x && planeNrml.y == cylD.y && planeNrml.z == cylD.z;
        if (len > kEpsilon && !same) {
            A = A / len;
            B = cross(planeNrml, A);
        } else {
            coordinateSystem(planeNrml, A, B);
        }
        V3 delta = planePt - cylPt, deltaProj = delta - cylD * dot(delta, cylD);
        float aDotD = dot(A, cylD), bDotD = dot(B, cylD);
        float c0 = 1 - aDotD * aDotD, c1 = 1 - bDotD * bDotD;
        float c2 = 2 * dot(A, deltaProj), c3 = 2 * dot(B, deltaProj);
        float c4 = dot(delta, deltaProj) - radius * radius;
        float lambda = (c2 * c2 / (4 * c0) + c3 * c3 / (4 * c1) - c4) / (c0 * c1);
        float alpha0 = -c2 / (2 * c0), beta0 = -c3 / (2 * c1);
        lengths[0] = std::sqrt(c1 * lambda);
        lengths[1] = std::sqrt(c0 * lambda);
        center = planePt + alpha0 * A + beta0 * B;
        axes[0] = A; axes[1] = B;
        return true;
    }

    // hair.cpp:368-397 (tight bound of the two miter-cut end ellipses, radius*(1-Epsilon))
    AABB segmentAABB(uint32_t iv) const {
        V3 center, axes[2]; float lengths[2];
        AABB result;
        for (int end = 0; end < 2; ++end) {
            V3 pt = end == 0 ? firstVertex(iv) : secondVertex(iv);
            V3 nrm = end == 0 ? firstMiterNormal(iv) : secondMiterNormal(iv);
            bool ok = intersectCylPlane(pt, nrm, pt, tangent(iv), radius * (1 - kEpsilon), center, axes, lengths);
            (void) ok;
            axes[0] *= lengths[0]; axes[1] *= lengths[1];
            for (int i = 0; i < 3; ++i) {
                float range = std::sqrt(axes[0][i] * axes[0][i] + axes[1][i] * axes[1][i]);
                result.mn[i] = std::min(result.mn[i], center[i] - range);
                result.mx[i] = std::max(result.mx[i], center[i] + range);
            }
        }
        return result;
    }

    // skdtree.cpp:78-110: TriAccel per triangle; bounds = union of the triangle boxes (what the kd-tree build sees)
    void finalizeMesh(uint32_t shapeIndex) {
        isMesh = true;
        aabb = AABB();
        if (mesh.isRect) { aabb = mesh.rect.getAABB(); return; }
        mesh.accel.resize(mesh.triCount());
        for (uint32_t j = 0; j < mesh.triCount(); ++j) {
            mesh.accel[j].load(mesh.pos[mesh.idx[3 * j]], mesh.pos[mesh.idx[3 * j + 1]], mesh.pos[mesh.idx[3 * j + 2]]);
            mesh.accel[j].shapeIndex = shapeIndex; mesh.accel[j].primIndex = j;
            aabb.expand(mesh.triAABB(j));
        }
    }

    void finalize() {
        // hair.cpp:117-124
        segIndex.clear();
        for (size_t i = 0; i + 1 < verts.size(); i++)
            if (!startsFiber[i + 1]) segIndex.push_back((uint32_t) i);
        aabb = AABB();
        for (uint32_t iv : segIndex) aabb.expand(segmentAABB(iv));
        if (!segIndex.empty()) enlargeKDTreeBounds(aabb);       // HairKDTree is a KDTreeBase: its getAABB() is the enlarged box
    }

    // hair.cpp:485-542.  Returns true on hit; t is the fp32-rounded root, p the fp32 hit point.
    bool intersect(const Ray &ray, uint32_t iv, float mint, float maxt, float &t, V3 &p) const {
        D3 axis = tangentD(iv);
        D3 rayO(ray.o), rayD(ray.d), v1(verts[iv]);
        D3 relOrigin = rayO - v1;
        D3 projOrigin = relOrigin - dot(axis, relOrigin) * axis;
        D3 projDirection = rayD - dot(axis, rayD) * axis;
        const double A = dot(projDirection, projDirection);
        const double B = 2 * dot(projOrigin, projDirection);
        const double C = dot(projOrigin, projOrigin) - radius * radius; // m_radius*m_radius is a float product
        double nearT, farT;
        if (!solveQuadraticDouble(A, B, C, nearT, farT)) return false;
        if (!(nearT <= maxt && farT >= mint)) return false;
        D3 pointNear = rayO + rayD * nearT, pointFar = rayO + rayD * farT;
        D3 n1 = firstMiterNormalD(iv), n2 = secondMiterNormalD(iv), v2(verts[iv + 1]);
        if (dot(pointNear - v1, n1) >= 0 && dot(pointNear - v2, n2) <= 0 && nearT >= mint) {
            D3 q = rayO + rayD * nearT;
            p = V3((float) q.x, (float) q.y, (float) q.z);
            t = (float) nearT;
        } else if (dot(pointFar - v1, n1) >= 0 && dot(pointFar - v2, n2) <= 0) {
            if (farT > maxt) return false;
            D3 q = rayO + rayD * farT;
            p = V3((float) q.x, (float) q.y, (float) q.z);
            t = (float) farT;
        } else {
            return false;
        }
        return true;
    }
};

// hair.cpp:609-785.  `reduction` > 0 skips whole fibers with one draw of Random() per fiber marker (:629,672-673,769-770) and thickens the rest.
static inline void loadHairFile(const std::string &path, float radius, float angleThresholdDeg,
                                const M44 &toWorld, HairShape &out, float reduction = 0.0f) {
    float angleThreshold = angleThresholdDeg * (kPi / 180.0f);
    float dpThresh = std::cos(angleThreshold);
    if (reduction < 0 || reduction >= 1) throw std::runtime_error("The 'reduction' parameter must have a value in [0, 1)!");
    else if (reduction > 0) { float correction = 1.0f / (1 - reduction); radius *= correction; }
    MitsubaRandom random;
    bool ignore = false;
    radius *= length(xfmVector(toWorld, V3(0, 0, 1)));
    std::ifstream bs(path, std::ios::binary);
    if (!bs) throw std::runtime_error("oracle: cannot open hair file " + path);
    char temp[11] = {0};
    bs.read(temp, 11);
    if (bs.gcount() != 11) throw std::runtime_error("Read less data than expected (11 bytes required) from \"" + path + "\"");   // FileStream::read, fstream.cpp:317
    bool binaryFormat = std::memcmp(temp, "BINARY_HAIR", 11) == 0;

    std::vector<V3> &vertices = out.verts;
    std::vector<uint8_t> &vsf = out.startsFiber;
    vertices.clear(); vsf.clear();
    V3 tangent(0.0f), p, lastP(0.0f);
    bool newFiber = true;

    auto consume = [&](V3 pt) {
        p = xfmPoint(toWorld, pt);
        if (newFiber) {
            vertices.push_back(p); vsf.push_back(1);
            lastP = p; tangent = V3(0.0f);
        } else if (!(p.x == lastP.x && p.y == lastP.y && p.z == lastP.z)) {
            if (isZero(tangent)) {
                vertices.push_back(p); vsf.push_back(0);
                tangent = normalize(p - lastP);
                lastP = p;
            } else {
                V3 nextTangent = normalize(p - lastP);
                if (dot(nextTangent, tangent) > dpThresh) {
                    tangent = normalize(p - vertices[vertices.size() - 2]);
                    vertices[vertices.size() - 1] = p;
                } else {
                    vertices.push_back(p); vsf.push_back(0);
                    tangent = nextTangent;
                }
                lastP = p;
            }
        }
        newFiber = false;
    };

    if (binaryFormat) {
        uint32_t vertexCount = 0;
        bs.read((char *) &vertexCount, 4);
        size_t verticesRead = 0;
        auto readSingle = [&]() { float f = 0; bs.read((char *) &f, 4); if (!bs) throw std::runtime_error("oracle: truncated hair file"); return f; };
        while (verticesRead != vertexCount) {
            float value = readSingle();
            V3 q;
            if (std::isinf(value)) {
                q.x = readSingle(); q.y = readSingle(); q.z = readSingle();
                newFiber = true;
                if (reduction > 0) ignore = random.nextFloat() < reduction;
            } else {
                q.x = value; q.y = readSingle(); q.z = readSingle();
            }
            verticesRead++;
            if (ignore) newFiber = false; else consume(q);
        }
    } else {
        std::ifstream is(path);
        std::string line;
        while (is.good()) {
            std::getline(is, line);
            if (line.length() > 0 && line[0] == '#') { newFiber = true; continue; }
            std::istringstream iss(line);
            V3 q;
            iss >> q.x >> q.y >> q.z;
            if (!iss.fail()) { if (ignore) newFiber = false; else consume(q); }
            else { newFiber = true; if (reduction > 0) ignore = random.nextFloat() < reduction; }
        }
    }
    vsf.push_back(1);
    out.radius = radius;
    out.finalize();
}

struct Hit {
    float t = kInf;
    int shape = -1;
    uint32_t iv = 0;
    V3 p;      // fp32 hit point stored by HairKDTree::intersect
    float u = 0, v = 0; // barycentrics stored by ShapeKDTree::intersect for triangles (skdtree.h:296-300); iv = triangle index
};

// Result of fillIntersectionRecord (hair.cpp:825-862 + skdtree.h:426-427)
struct Intersection {
    bool valid = false;
    float t = kInf;
    int shape = -1;
    uint32_t iv = 0;
    V3 p;
    Frame geoFrame, shFrame;
    V3 wi;
    float u = 0, v = 0;   // its.uv
};

struct BVHNode { AABB box; uint32_t left, right; uint32_t first, count; }; // leaf if count>0

struct Geometry {
    std::vector<HairShape> shapes;
    AABB aabb; // scene kd-tree bounds (union of shape bounds; skdtree.h:213-221)
    // oracle-only acceleration structure
    struct PrimRef { uint32_t shape, iv; };
    std::vector<PrimRef> prims;
    std::vector<BVHNode> nodes;

    void finalize() {
        aabb = AABB();
        for (auto &s : shapes) aabb.expand(s.aabb);   // hair: the shape's (enlarged) tree box; mesh: union of its triangle boxes (skdtree.h:213-221)
        if (!shapes.empty()) enlargeKDTreeBounds(aabb);
        buildBVH();
    }

    // skdtree.cpp:112-142 (closest) / :207-226 (shadow): scene-level interval set-up.
    // Returns false if the ray misses the scene bounds or the interval is empty.
    bool sceneInterval(const Ray &ray, bool shadow, float &mint, float &maxt) const {
        if (!aabb.rayIntersect(ray.o, ray.d, ray.dRcp, mint, maxt)) return false;
        float rayMinT = ray.mint;
        if (rayMinT == kEpsilon) {
            float m = std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
            if (!shadow) m = std::max(m, kEpsilon);
            rayMinT *= m;
        }
        if (rayMinT > mint) mint = rayMinT;
        if (ray.maxt < maxt) maxt = ray.maxt;
        return maxt > mint;
    }
    // hair.cpp:200-217: per-shape interval clip
    bool shapeInterval(const HairShape &s, const Ray &ray, float _mint, float _maxt, float &mint, float &maxt) const {
        if (!s.aabb.rayIntersect(ray.o, ray.d, ray.dRcp, mint, maxt)) return false;
        if (_mint > mint) mint = _mint;
        if (_maxt < maxt) maxt = _maxt;
        return maxt > mint;
    }

    // Brute force in (shape, segment) order -- the reference semantics with a trivial visiting order.
    bool intersectBrute(const Ray &ray, bool shadow, Hit &hit) const {
        float mint, maxt;
        hit = Hit();
        if (!sceneInterval(ray, shadow, mint, maxt)) return false;
        bool found = false;
        for (size_t si = 0; si < shapes.size(); ++si) {
            const HairShape &s = shapes[si];
            if (s.isMesh) { // triangles live in the top-level tree: they see [mint, maxt] directly (skdtree.h:293-304)
                for (uint32_t j = 0; j < s.mesh.triCount(); ++j) {
                    float t, u, v;
                    if (s.mesh.intersectPrim(j, ray.o, ray.d, mint, maxt, u, v, t)) {
                        if (shadow) { hit.t = t; hit.shape = (int) si; hit.iv = j; return true; }
                        maxt = t; hit.t = t; hit.shape = (int) si; hit.iv = j; hit.u = u; hit.v = v; found = true;
                    }
                }
                continue;
            }
            float smin, smax;
            if (!shapeInterval(s, ray, mint, maxt, smin, smax)) continue;
            for (uint32_t iv : s.segIndex) {
                float t; V3 p;
                if (s.intersect(ray, iv, smin, smax, t, p)) {
                    if (shadow) { hit.t = t; hit.shape = (int) si; hit.iv = iv; return true; }
                    smax = t; maxt = t; // sahkdtree3.h:287-291 (maxt = t at both levels)
                    hit.t = t; hit.shape = (int) si; hit.iv = iv; hit.p = p;
                    found = true;
                }
            }
        }
        return found;
    }

#ifdef ORC_FAST
    // ------------------------------------------------------------------------------------------------------------------------
    // CPU-BASELINE ACCELERATOR (timing build liboracle_fast.so only; the checker liboracle.so keeps the plain BVH above).
    // What `mitsuba -p N` has and a plain BVH over whole segments lacks is (a) spatial subdivision of long thin segments -- its SAH
    // kd-tree clips cylinders against the split planes (src/shapes/hair.cpp:246-444, include/mitsuba/render/gkdtree.h:958-2400) --
    // and (b) a cheap rejection in front of the FP64 cylinder test (the kd leaf boxes are tight).  This build gets both the same way
    // the CUDA path does: every segment is referenced by up to 8 boxes cut along its axis, a binned-SAH binary BVH is built over
    // the references (in parallel), a conservative fp32 distance test runs before the exact test, and an 8-entry mailbox
    // (sahkdtree3.h:138-152) skips segments already tested for this ray.  The exact test and the interval logic are the
    // unchanged oracle functions, so results are identical to the checker build except for equal-t ties
    // (tests/test_oracle_cpu.py::test_fast_accelerator_matches_checker).
    struct FastRef { uint32_t shapeFlags, iv; };     // shapeFlags: bits 3.. shape, bit 0 mesh triangle, bits 1 / 2 mild miter joint at the first / second vertex (8 bytes: the vertices are read from the shape)
    struct FastNode { float lo[3], hi[3]; uint32_t a, b; };                         // inner: a = right child (left = self + 1), b = 0; leaf: a = first reference, b = count
    std::vector<FastRef> frefs;
    FastNode *fnodes = nullptr; size_t fnodeCap = 0;                                 // binary build product, collapsed into `wide` and released
    struct Wide { float lo[3][8], hi[3][8]; uint32_t child[8]; };                    // 8-wide node, SoA boxes: child >= 0x80000000: reference index | 0x80000000; 0xffffffff: empty
    std::vector<Wide> wide;
    int fastMaxSplit = 8, fastLeaf = 1;
    ~Geometry() { free(fnodes); }
    Geometry() = default;
    Geometry(const Geometry &) = delete; Geometry &operator=(const Geometry &) = delete;

    static const char *accelDescription() { return "8-wide BVH (top-down spatial-median splits on the SAH-cheapest axis, collapsed from a binary tree with one reference per leaf; SoA boxes tested 8 at a time) over segment references pre-split up to 8x along the fiber axis, fp32 pre-test before the FP64 cylinder test, 8-entry mailbox (parallel build)"; }

    void buildBVH() {
        frefs.clear(); free(fnodes); fnodes = nullptr; fnodeCap = 0;
        if (const char *e = getenv("ORC_FAST_MAX_SPLIT")) fastMaxSplit = std::max(1, atoi(e));
        if (const char *e = getenv("ORC_FAST_LEAF")) fastLeaf = std::max(1, atoi(e));
        std::vector<AABB> boxes;
        for (size_t si = 0; si < shapes.size(); ++si) {
            const HairShape &s = shapes[si];
            if (s.isMesh) {
                for (uint32_t j = 0; j < s.mesh.triCount(); ++j) {
                    AABB b = s.mesh.triAABB(j);
                    V3 pad = (b.mx - b.mn) * 1e-5f + V3(1e-6f) + V3(std::abs(b.mx.x) + std::abs(b.mn.x), std::abs(b.mx.y) + std::abs(b.mn.y), std::abs(b.mx.z) + std::abs(b.mn.z)) * 1e-6f;
                    b.mn = b.mn - pad; b.mx = b.mx + pad;
                    FastRef r{}; r.shapeFlags = ((uint32_t) si << 3) | 1u; r.iv = j;
                    frefs.push_back(r); boxes.push_back(b);
                }
                continue;
            }
            auto mild = [&](uint32_t v) {          // the miter joint at vertex v bends by at most ~117 degrees (or v ends a fiber): the cut overshoots the end by < 2 r
                const bool hasPrev = !s.startsFiber[v] && v > 0, hasNext = v + 1 < s.verts.size() && !s.startsFiber[v + 1];
                if (!(hasPrev && hasNext)) return true;
                return dot(normalize(s.verts[v] - s.verts[v - 1]), normalize(s.verts[v + 1] - s.verts[v])) >= -0.45f;
            };
            for (uint32_t iv : s.segIndex) {
                const AABB whole = s.segmentAABB(iv);
                const V3 p1 = s.verts[iv], a = s.verts[iv + 1] - p1;
                const float ext = std::max(std::max(std::abs(a.x), std::abs(a.y)), std::abs(a.z));
                const int k = std::max(1, std::min((int) (ext / (8.0f * s.radius)), fastMaxSplit));
                const float invLen = 1.0f / length(a);
                const V3 e(s.radius * std::sqrt(std::max(0.0f, 1.0f - (a.x * invLen) * (a.x * invLen))), s.radius * std::sqrt(std::max(0.0f, 1.0f - (a.y * invLen) * (a.y * invLen))),
                           s.radius * std::sqrt(std::max(0.0f, 1.0f - (a.z * invLen) * (a.z * invLen))));
                for (int j = 0; j < k; ++j) {
                    AABB b;
                    if (k == 1) b = whole;
                    else {
                        // a piece's box: its stretch of the axis widened by the cross-section; the two end pieces also cover the miter-cut end
                        // ellipses by taking the whole segment's bound on the far side of their cut
                        const V3 q0 = p1 + a * ((float) j / k), q1 = p1 + a * ((float) (j + 1) / k);
                        for (int c = 0; c < 3; ++c) {
                            float lo = std::min(q0[c], q1[c]) - e[c] - s.radius * 1e-3f, hi = std::max(q0[c], q1[c]) + e[c] + s.radius * 1e-3f;
                            const bool towardsMin = a[c] >= 0;       // piece 0 touches the segment's minimum along c when the axis increases
                            if ((j == 0 && towardsMin) || (j == k - 1 && !towardsMin)) lo = std::min(lo, whole.mn[c]);
                            if ((j == k - 1 && towardsMin) || (j == 0 && !towardsMin)) hi = std::max(hi, whole.mx[c]);
                            b.mn[c] = std::max(lo, whole.mn[c]); b.mx[c] = std::min(hi, whole.mx[c]);
                        }
                    }
                    for (int c = 0; c < 3; ++c) {   // every point the FP64 test can accept lies strictly inside (the reference's bound uses radius (1 - Epsilon))
                        const float pad = s.radius * 4e-4f + 4e-7f * std::max(std::abs(b.mn[c]), std::abs(b.mx[c]));
                        b.mn[c] -= pad; b.mx[c] += pad;
                    }
                    FastRef r{};
                    r.shapeFlags = ((uint32_t) si << 3) | (mild(iv) ? 2u : 0u) | (mild(iv + 1) ? 4u : 0u); r.iv = iv;
                    frefs.push_back(r); boxes.push_back(b);
                }
            }
        }
        if (frefs.empty()) return;
        const uint32_t n = (uint32_t) frefs.size();
        std::vector<uint32_t> order(n);
        for (uint32_t i = 0; i < n; ++i) order[i] = i;
        fnodeCap = 2 * (size_t) n;
        fnodes = (FastNode *) calloc(fnodeCap, sizeof(FastNode));      // lazily committed: a subtree over m references owns 2m - 1 consecutive slots
        if (!fnodes) throw std::runtime_error("oracle: out of memory for the baseline accelerator");
        fastBuildRec(0, order, boxes, 0, n, 0);
        std::vector<FastRef> sorted(n);
        for (uint32_t i = 0; i < n; ++i) sorted[i] = frefs[order[i]];
        frefs.swap(sorted);
        // collapse into 8-wide nodes: the children of a wide node are found by opening, largest box first, the inner nodes below it
        wide.clear(); wide.reserve(n / 4 + 16);
        wide.push_back(Wide());
        if (fnodes[0].b) {       // a single leaf
            Wide &w = wide[0];
            for (int k = 0; k < 8; ++k) { for (int c = 0; c < 3; ++c) { w.lo[c][k] = kInf; w.hi[c][k] = -kInf; } w.child[k] = 0xffffffffu; }
            for (uint32_t k = 0; k < fnodes[0].b && k < 8; ++k) { for (int c = 0; c < 3; ++c) { w.lo[c][k] = fnodes[0].lo[c]; w.hi[c][k] = fnodes[0].hi[c]; } w.child[k] = 0x80000000u | (fnodes[0].a + k); }
        } else {
            struct Work { uint32_t wi, bi, cnt; };
            std::vector<Work> work;       // (wide index, binary node, references below it)
            work.push_back({0u, 0u, n});
            while (!work.empty()) {
                const Work wk = work.back(); work.pop_back();
                const uint32_t wi = wk.wi, bi = wk.bi;
                uint32_t cand[8], ccnt[8]; int nc = 0;
                ccnt[nc] = (fnodes[bi].a - bi) / 2; cand[nc++] = bi + 1; ccnt[nc] = wk.cnt - ccnt[0]; cand[nc++] = fnodes[bi].a;
                auto areaOf = [&](uint32_t i) { const FastNode &f = fnodes[i]; const float ex = f.hi[0] - f.lo[0], ey = f.hi[1] - f.lo[1], ez = f.hi[2] - f.lo[2]; return ex * ey + ey * ez + ez * ex; };
                while (nc < 8) {
                    int bestK = -1; float bestA = -1;
                    for (int k = 0; k < nc; ++k) if (!fnodes[cand[k]].b) { const float a = areaOf(cand[k]); if (a > bestA) { bestA = a; bestK = k; } }
                    if (bestK < 0) break;
                    const uint32_t open = cand[bestK], oc = ccnt[bestK];
                    cand[bestK] = open + 1; ccnt[bestK] = (fnodes[open].a - open) / 2; ccnt[nc] = oc - ccnt[bestK]; cand[nc++] = fnodes[open].a;
                }
                Wide w;
                for (int k = 0; k < 8; ++k) {
                    if (k >= nc) { for (int c = 0; c < 3; ++c) { w.lo[c][k] = kInf; w.hi[c][k] = -kInf; } w.child[k] = 0xffffffffu; continue; }
                    const FastNode &f = fnodes[cand[k]];
                    for (int c = 0; c < 3; ++c) { w.lo[c][k] = f.lo[c]; w.hi[c][k] = f.hi[c]; }
                    if (f.b) w.child[k] = 0x80000000u | f.a;      // binary leaves hold exactly one reference
                    else { w.child[k] = (uint32_t) wide.size(); wide.push_back(Wide()); work.push_back({w.child[k], cand[k], ccnt[k]}); }
                }
                wide[wi] = w;
            }
        }
        free(fnodes); fnodes = nullptr; fnodeCap = 0;
        nodes.clear(); nodes.push_back(BVHNode());     // "built" marker for callers that look at nodes.empty()
    }

    void fastBuildRec(size_t ni, std::vector<uint32_t> &order, const std::vector<AABB> &boxes, uint32_t lo, uint32_t hi, int depth) {
        AABB box, cbox;
        for (uint32_t i = lo; i < hi; ++i) { box.expand(boxes[order[i]]); cbox.expand(boxes[order[i]].center()); }
        FastNode &nd = fnodes[ni];
        for (int c = 0; c < 3; ++c) { nd.lo[c] = box.mn[c]; nd.hi[c] = box.mx[c]; }
        const uint32_t n = hi - lo;
        if (n <= (uint32_t) fastLeaf) { nd.a = lo; nd.b = n; return; }
        V3 ext = cbox.mx - cbox.mn;
        uint32_t mid = (lo + hi) / 2;
        auto area = [](const AABB &b) { V3 e = b.mx - b.mn; return 2 * (e.x * e.y + e.y * e.z + e.z * e.x); };
        const int NB = 16;
        float best = kInf; int bestAxis = -1, bestSplit = -1;
        // Split rule.  Default: the spatial median of the centroid box, on the axis whose two halves have the lowest SAH cost.  On hair (long
        // thin references, mostly empty boxes) this beats the 16-bin SAH sweep that round 1 / early round 2 used by 3x to 9x in node visits per
        // ray (hair-curl, 32 M references: 16.7 against 149 wide nodes per ray, 2.2 against 0.44 Mrays/s on 8 threads); the sweep stays
        // selectable with ORC_FAST_SPLITMODE=0 for the record.
        static const int splitMode = getenv("ORC_FAST_SPLITMODE") ? atoi(getenv("ORC_FAST_SPLITMODE")) : 2;
#ifdef ORC_FAST_STATS
        if (splitMode == 5 || splitMode == 6) {   // experiment: what a Morton-code LBVH does -- global grid, axes x, y, z in turn (5: every axis normalised by its own extent, 6: by the largest)
            static AABB g; static std::once_flag once; static V3 gext;
            std::call_once(once, [&]() { g = cbox; V3 e = cbox.mx - cbox.mn; float m = std::max(e.x, std::max(e.y, e.z)); gext = splitMode == 6 ? V3(m, m, m) : e; });
            // find the first grid bit (from the top) at which the centroids of [lo, hi) differ
            auto cell = [&](uint32_t id, int axis) { float f = (boxes[id].center()[axis] - g.mn[axis]) / gext[axis]; return (uint32_t) std::min(std::max(f * 2097152.0f, 0.0f), 2097151.0f); };
            bool done = false;
            for (int bit = 20; bit >= 0 && !done; --bit) for (int axis = 0; axis < 3 && !done; ++axis) {
                const uint32_t first = cell(order[lo], axis) >> bit; bool differ = false;
                for (uint32_t i = lo + 1; i < hi; ++i) if ((cell(order[i], axis) >> bit) != first) { differ = true; break; }
                if (!differ) continue;
                auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t id) { return ((cell(id, axis) >> bit) & 1u) == 0u; });
                mid = (uint32_t) (it - order.begin()); done = true;
            }
        } else
#endif
        if (splitMode == 2) {
            float bestC = kInf; int bestA = -1;
            for (int axis = 0; axis < 3; ++axis) {
                if (!(ext[axis] > 0)) continue;
                const float midp = 0.5f * (cbox.mn[axis] + cbox.mx[axis]);
                AABB l, r; uint32_t nl = 0, nr = 0;
                for (uint32_t i = lo; i < hi; ++i) { const AABB &b = boxes[order[i]]; if (b.center()[axis] < midp) { l.expand(b); nl++; } else { r.expand(b); nr++; } }
                if (!nl || !nr) continue;
                const float c = area(l) * nl + area(r) * nr;
                if (c < bestC) { bestC = c; bestA = axis; }
            }
            if (bestA >= 0) {
                const float midp = 0.5f * (cbox.mn[bestA] + cbox.mx[bestA]);
                auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t id) { return boxes[id].center()[bestA] < midp; });
                mid = (uint32_t) (it - order.begin());
            }
        } else
        for (int axis = 0; axis < 3; ++axis) {
            if (!(ext[axis] > 0)) continue;
            AABB bb[NB]; uint32_t bc[NB] = {0};
            const float k = NB * (1 - 1e-6f) / ext[axis];
            for (uint32_t i = lo; i < hi; ++i) { const uint32_t id = order[i]; const int b = clampi((int) ((boxes[id].center()[axis] - cbox.mn[axis]) * k), 0, NB - 1); bb[b].expand(boxes[id]); bc[b]++; }
            float rightArea[NB]; uint32_t rightCount[NB];
            AABB acc; uint32_t cnt = 0;
            for (int b = NB - 1; b > 0; --b) { acc.expand(bb[b]); cnt += bc[b]; rightArea[b] = cnt ? area(acc) : 0; rightCount[b] = cnt; }
            acc = AABB(); cnt = 0;
            for (int b = 0; b < NB - 1; ++b) {
                acc.expand(bb[b]); cnt += bc[b];
                if (cnt == 0 || rightCount[b + 1] == 0) continue;
                const float cost = area(acc) * cnt + rightArea[b + 1] * rightCount[b + 1];
                if (cost < best) { best = cost; bestAxis = axis; bestSplit = b; }
            }
        }
        if (bestAxis >= 0) {
            const float k = NB * (1 - 1e-6f) / ext[bestAxis];
            auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t id) { return clampi((int) ((boxes[id].center()[bestAxis] - cbox.mn[bestAxis]) * k), 0, NB - 1) <= bestSplit; });
            mid = (uint32_t) (it - order.begin());
        }
        if (mid == lo || mid == hi) {
            const int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
            mid = (lo + hi) / 2;
            std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](uint32_t a, uint32_t b) { return boxes[a].center()[axis] < boxes[b].center()[axis]; });
        }
        const size_t left = ni + 1, right = ni + 1 + (2 * (size_t) (mid - lo) - 1);
        nd.a = (uint32_t) right; nd.b = 0;
        if (depth < 4 && n > 200000) {          // the top of the tree forks: up to 16 subtrees are built concurrently
            std::thread t([&]() { fastBuildRec(left, order, boxes, lo, mid, depth + 1); });
            fastBuildRec(right, order, boxes, mid, hi, depth + 1);
            t.join();
        } else {
            fastBuildRec(left, order, boxes, lo, mid, depth + 1);
            fastBuildRec(right, order, boxes, mid, hi, depth + 1);
        }
    }

    static inline bool slab(const FastNode &nd, const Ray &ray, float mint, float maxt, float &tnear) {
        float t0 = mint, t1 = maxt;
        for (int c = 0; c < 3; ++c) {
            if (ray.d[c] == 0) { if (ray.o[c] < nd.lo[c] || ray.o[c] > nd.hi[c]) return false; continue; }
            float a = (nd.lo[c] - ray.o[c]) * ray.dRcp[c], b = (nd.hi[c] - ray.o[c]) * ray.dRcp[c];
            if (a > b) std::swap(a, b);
            t0 = std::max(t0, a); t1 = std::min(t1, b);
        }
        tnear = t0;
        return t0 <= t1 * 1.0000004f;
    }

#ifdef ORC_FAST_STATS
    static inline std::atomic<uint64_t> stNodes{0}, stPre{0}, stExact{0}, stRays{0}, stLeafPop{0};
#define ORC_ST(...) __VA_ARGS__
#else
#define ORC_ST(...)
#endif
    bool intersectBVH(const Ray &ray, bool shadow, Hit &hit) const {
        float mint, maxt;
        hit = Hit();
        ORC_ST(uint64_t cN = 0, cP = 0, cE = 0, cL = 0; struct Fl { uint64_t &a, &b, &c, &d; ~Fl() { stNodes += a; stPre += b; stExact += c; stLeafPop += d; stRays += 1; } } fl{cN, cP, cE, cL};)
        if (wide.empty() || !sceneInterval(ray, shadow, mint, maxt)) return false;
        float pmin[16], pmax[16]; bool pok[16];
        if (shapes.size() > 16) throw std::runtime_error("oracle: >16 shapes unsupported in BVH path");
        for (size_t si = 0; si < shapes.size(); ++si) pok[si] = shapes[si].isMesh ? true : shapeInterval(shapes[si], ray, mint, maxt, pmin[si], pmax[si]);
        uint64_t mailbox[8]; for (int i = 0; i < 8; ++i) mailbox[i] = ~0ull;
        const V3 o = ray.o, d = ray.d;
        const float dd = dot(d, d);
        // slab arithmetic without branches: a zero direction component keeps the exact containment test of aabb.h:315-318 by
        // turning the two plane distances into -inf / +inf (inside) or +inf / -inf (outside)
        float ro[3], rinv[3]; bool flat[3];
        for (int c = 0; c < 3; ++c) { ro[c] = o[c]; flat[c] = d[c] == 0; rinv[c] = flat[c] ? 0.0f : ray.dRcp[c]; }
        bool found = false;
        struct Entry { uint32_t node; float tnear; };
        Entry stack[256]; int sp = 0;
        stack[sp++] = {0u, mint};
        while (sp) {
            const Entry e = stack[--sp];
            if (!shadow && e.tnear > maxt) continue;
            if (e.node & 0x80000000u) {
                const FastRef &r = frefs[e.node & 0x7fffffffu];
                const uint32_t rshape = r.shapeFlags >> 3;
                if (r.shapeFlags & 1u) {
                    float t, u, v;
                    if (shapes[rshape].mesh.intersectPrim(r.iv, ray.o, ray.d, mint, maxt, u, v, t)) {
                        if (shadow) { hit.t = t; hit.shape = (int) rshape; hit.iv = r.iv; return true; }
                        maxt = t; hit.t = t; hit.shape = (int) rshape; hit.iv = r.iv; hit.u = u; hit.v = v; found = true;
                    }
                    continue;
                }
                if (!pok[rshape]) continue;
                ORC_ST(cP++;)
                const HairShape &sh = shapes[rshape];
                const float radius = sh.radius;
                // conservative fp32 rejection (never rejects a hit the FP64 test would accept; same bound as csrc/cp_traverse.cuh)
                const V3 p1 = sh.verts[r.iv], a = sh.verts[r.iv + 1] - p1, w = p1 - o, nrm = cross(d, a);
                const float nn = dot(nrm, nrm), aa = dot(a, a), wn = dot(w, nrm);
                const float sin2 = nn / (aa * dd);
                if (sin2 > 4e-4f) {
                    const float wmax = std::max(std::max(std::abs(w.x), std::abs(w.y)), std::abs(w.z));
                    const float rsin = 1.0f / std::sqrt(sin2);
                    const float R = radius * 1.02f + wmax * (2e-6f + 2e-6f * rsin);
                    if (wn * wn > R * R * nn) continue;
                    const float inn = 1.0f / nn;
                    const float tcl = dot(cross(w, a), nrm) * inn;
                    const float slack = 1.01f * R * rsin / std::sqrt(dd) + 1e-5f * std::abs(tcl);
                    if (tcl + slack < mint || tcl - slack > maxt) continue;
                    if ((r.shapeFlags & 6u) == 6u) {
                        const float scl = dot(cross(w, d), nrm) * inn;
                        const float sslack = 1.01f * (R * rsin + 2.0f * radius) / std::sqrt(aa) + 1e-5f * (1.0f + std::abs(scl));
                        if (scl + sslack < 0.0f || scl - sslack > 1.0f) continue;
                    }
                }
                const uint64_t key = ((uint64_t) rshape << 32) | r.iv;
                uint64_t &slot = mailbox[(r.iv * 2654435761u) >> 29];
                if (slot == key) continue;
                slot = key;
                const float hi = std::min(pmax[rshape], maxt);
                if (!(hi > pmin[rshape])) continue;
                float t; V3 p;
                ORC_ST(cE++;)
                if (sh.intersect(ray, r.iv, pmin[rshape], hi, t, p)) {
                    if (shadow) { hit.t = t; hit.shape = (int) rshape; hit.iv = r.iv; return true; }
                    maxt = t; hit.t = t; hit.shape = (int) rshape; hit.iv = r.iv; hit.p = p; found = true;
                }
                continue;
            }
            ORC_ST(cN++;)
            const Wide &nd = wide[e.node];
            ORC_ST({ bool anyRef = false; for (int k = 0; k < 8; ++k) if (nd.child[k] != 0xffffffffu && (nd.child[k] & 0x80000000u)) anyRef = true; if (anyRef) cL++; })
            float tn[8], tf[8];
            for (int k = 0; k < 8; ++k) { tn[k] = mint; tf[k] = maxt; }
            for (int c = 0; c < 3; ++c) {
                if (flat[c]) {
                    for (int k = 0; k < 8; ++k) if (ro[c] < nd.lo[c][k] || ro[c] > nd.hi[c][k]) tf[k] = -kInf;
                } else {
                    for (int k = 0; k < 8; ++k) {
                        const float a = (nd.lo[c][k] - ro[c]) * rinv[c], b = (nd.hi[c][k] - ro[c]) * rinv[c];
                        tn[k] = std::max(tn[k], std::min(a, b)); tf[k] = std::min(tf[k], std::max(a, b));
                    }
                }
            }
            // children that are entered, nearest last (so that it is popped first)
            int first = sp;
            for (int k = 0; k < 8; ++k) {
                if (nd.child[k] == 0xffffffffu || !(tn[k] <= tf[k] * 1.0000004f)) continue;
                int pos = sp++;
                while (pos > first && stack[pos - 1].tnear < tn[k]) { stack[pos] = stack[pos - 1]; --pos; }
                stack[pos] = {nd.child[k], tn[k]};
                if (nd.child[k] & 0x80000000u) __builtin_prefetch(&frefs[nd.child[k] & 0x7fffffffu]);
                else { const char *pf = (const char *) &wide[nd.child[k]]; __builtin_prefetch(pf); __builtin_prefetch(pf + 64); __builtin_prefetch(pf + 128); __builtin_prefetch(pf + 192); }
            }
        }
        return found;
    }
#else
    static const char *accelDescription() { return "plain binned-SAH BVH over whole segments (checker build)"; }
    void buildBVH() {
        prims.clear(); nodes.clear();
        std::vector<AABB> boxes;
        for (size_t si = 0; si < shapes.size(); ++si) {
            if (shapes[si].isMesh) {
                for (uint32_t j = 0; j < shapes[si].mesh.triCount(); ++j) {
                    prims.push_back({(uint32_t) si, j});
                    AABB b = shapes[si].mesh.triAABB(j);   // padded: a box only decides which tests run, and flat boxes must not lose edge hits to rounding
                    V3 pad = (b.mx - b.mn) * 1e-5f + V3(1e-6f) + V3(std::abs(b.mx.x) + std::abs(b.mn.x), std::abs(b.mx.y) + std::abs(b.mn.y), std::abs(b.mx.z) + std::abs(b.mn.z)) * 1e-6f;
                    b.mn = b.mn - pad; b.mx = b.mx + pad;
                    boxes.push_back(b);
                }
                continue;
            }
            for (uint32_t iv : shapes[si].segIndex) {
                prims.push_back({(uint32_t) si, iv});
                boxes.push_back(shapes[si].segmentAABB(iv));
            }
        }
        if (prims.empty()) return;
        std::vector<uint32_t> order(prims.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t) i;
        nodes.reserve(prims.size() * 2);
        nodes.push_back(BVHNode());
        buildRec(0, order, boxes, 0, (uint32_t) order.size());
        std::vector<PrimRef> sorted(prims.size());
        for (size_t i = 0; i < order.size(); ++i) sorted[i] = prims[order[i]];
        prims.swap(sorted);
    }

    void buildRec(uint32_t ni, std::vector<uint32_t> &order, const std::vector<AABB> &boxes, uint32_t lo, uint32_t hi) {
        AABB box, cbox;
        for (uint32_t i = lo; i < hi; ++i) { box.expand(boxes[order[i]]); cbox.expand(boxes[order[i]].center()); }
        nodes[ni].box = box;
        uint32_t n = hi - lo;
        if (n <= 4) { nodes[ni].first = lo; nodes[ni].count = n; nodes[ni].left = nodes[ni].right = 0; return; }
        V3 ext = cbox.mx - cbox.mn;
        int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
        uint32_t mid = (lo + hi) / 2;
        if (ext[axis] > 0) {
            const int NB = 16;
            AABB bb[NB]; uint32_t bc[NB] = {0};
            float k = NB * (1 - 1e-6f) / ext[axis];
            auto binOf = [&](uint32_t id) { return clampi((int) ((boxes[id].center()[axis] - cbox.mn[axis]) * k), 0, NB - 1); };
            for (uint32_t i = lo; i < hi; ++i) { int b = binOf(order[i]); bb[b].expand(boxes[order[i]]); bc[b]++; }
            auto area = [](const AABB &b) { V3 e = b.mx - b.mn; return 2 * (e.x * e.y + e.y * e.z + e.z * e.x); };
            float best = kInf; int bestSplit = -1;
            float rightArea[NB]; uint32_t rightCount[NB];
            AABB acc; uint32_t cnt = 0;
            for (int b = NB - 1; b > 0; --b) { acc.expand(bb[b]); cnt += bc[b]; rightArea[b] = cnt ? area(acc) : 0; rightCount[b] = cnt; }
            acc = AABB(); cnt = 0;
            for (int b = 0; b < NB - 1; ++b) {
                acc.expand(bb[b]); cnt += bc[b];
                if (cnt == 0 || rightCount[b + 1] == 0) continue;
                float cost = area(acc) * cnt + rightArea[b + 1] * rightCount[b + 1];
                if (cost < best) { best = cost; bestSplit = b; }
            }
            if (bestSplit >= 0) {
                auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t id) { return binOf(id) <= bestSplit; });
                mid = (uint32_t) (it - order.begin());
            }
            if (mid == lo || mid == hi) {
                mid = (lo + hi) / 2;
                std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                                 [&](uint32_t a, uint32_t b) { return boxes[a].center()[axis] < boxes[b].center()[axis]; });
            }
        }
        uint32_t l = (uint32_t) nodes.size(); nodes.push_back(BVHNode());
        uint32_t r = (uint32_t) nodes.size(); nodes.push_back(BVHNode());
        nodes[ni].left = l; nodes[ni].right = r; nodes[ni].count = 0; nodes[ni].first = 0;
        buildRec(l, order, boxes, lo, mid);
        buildRec(r, order, boxes, mid, hi);
    }

    // BVH query with the reference's per-primitive interval logic.
    bool intersectBVH(const Ray &ray, bool shadow, Hit &hit) const {
        float mint, maxt;
        hit = Hit();
        if (nodes.empty() || !sceneInterval(ray, shadow, mint, maxt)) return false;
        // per-shape clipped intervals (hair.cpp:205-209), computed once per ray
        float smin[16], smaxv[16]; bool sok[16];
        size_t ns = std::min<size_t>(shapes.size(), 16);
        std::vector<float> vmin, vmax; std::vector<char> vok;
        float *pmin = smin, *pmax = smaxv; bool *pok = sok;
        if (shapes.size() > 16) throw std::runtime_error("oracle: >16 shapes unsupported in BVH path");
        for (size_t si = 0; si < ns; ++si) pok[si] = shapes[si].isMesh ? true : shapeInterval(shapes[si], ray, mint, maxt, pmin[si], pmax[si]);
        bool found = false;
        uint32_t stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BVHNode &nd = nodes[stack[--sp]];
            float n0, f0;
            if (!nd.box.rayIntersect(ray.o, ray.d, ray.dRcp, n0, f0)) continue;
            if (n0 > maxt || f0 < mint) continue;
            if (nd.count) {
                for (uint32_t i = nd.first; i < nd.first + nd.count; ++i) {
                    const PrimRef &pr = prims[i];
                    if (shapes[pr.shape].isMesh) {
                        float t, u, v;
                        if (shapes[pr.shape].mesh.intersectPrim(pr.iv, ray.o, ray.d, mint, maxt, u, v, t)) {
                            if (shadow) { hit.t = t; hit.shape = (int) pr.shape; hit.iv = pr.iv; return true; }
                            maxt = t; hit.t = t; hit.shape = (int) pr.shape; hit.iv = pr.iv; hit.u = u; hit.v = v; found = true;
                        }
                        continue;
                    }
                    if (!pok[pr.shape]) continue;
                    float hi = std::min(pmax[pr.shape], maxt);
                    if (!(hi > pmin[pr.shape])) continue; // hair.cpp:209 `maxt > mint`
                    float t; V3 p;
                    if (shapes[pr.shape].intersect(ray, pr.iv, pmin[pr.shape], hi, t, p)) {
                        if (shadow) { hit.t = t; hit.shape = (int) pr.shape; hit.iv = pr.iv; return true; }
                        maxt = t; hit.t = t; hit.shape = (int) pr.shape; hit.iv = pr.iv; hit.p = p; found = true;
                    }
                }
            } else {
                float nl, fl, nr, fr;
                bool hl = nodes[nd.left].box.rayIntersect(ray.o, ray.d, ray.dRcp, nl, fl) && nl <= maxt && fl >= mint;
                bool hr = nodes[nd.right].box.rayIntersect(ray.o, ray.d, ray.dRcp, nr, fr) && nr <= maxt && fr >= mint;
                if (hl && hr) {
                    if (nl < nr) { stack[sp++] = nd.right; stack[sp++] = nd.left; }
                    else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
                } else if (hl) stack[sp++] = nd.left;
                else if (hr) stack[sp++] = nd.right;
            }
        }
        return found;
    }

#endif // ORC_FAST

    // hair.cpp:825-862 followed by skdtree.h:426-427
    void fillIntersection(const Ray &ray, const Hit &hit, Intersection &its) const {
        const HairShape &s = shapes[hit.shape];
        its.valid = true; its.t = hit.t; its.shape = hit.shape; its.iv = hit.iv;
        if (s.isMesh && s.mesh.isRect) { // rectangle.cpp:158-171, then skdtree.h:426-427
            const Rectangle &r = s.mesh.rect;
            its.geoFrame = r.frame;
            its.shFrame.n = r.frame.n;
            its.u = 0.5f * (hit.u + 1); its.v = 0.5f * (hit.v + 1);
            its.p = ray.o + ray.d * hit.t;
            computeShadingFrame(its.shFrame.n, r.dpdu, its.shFrame);
            its.wi = its.shFrame.toLocal(-ray.d);
            return;
        }
        if (s.isMesh) { // skdtree.h:346-427 with BarycentricPos = true (skdtree.cpp:136)
            const TriMesh &m = s.mesh;
            const V3 b(1 - hit.u - hit.v, hit.u, hit.v);
            const uint32_t idx0 = m.idx[3 * hit.iv], idx1 = m.idx[3 * hit.iv + 1], idx2 = m.idx[3 * hit.iv + 2];
            const V3 &p0 = m.pos[idx0], &p1 = m.pos[idx1], &p2 = m.pos[idx2];
            its.p = p0 * b.x + p1 * b.y + p2 * b.z;
            V3 side1(p1 - p0), side2(p2 - p0);
            V3 faceNormal = cross(side1, side2);
            float length = std::sqrt(dot(faceNormal, faceNormal));
            if (!(faceNormal.x == 0 && faceNormal.y == 0 && faceNormal.z == 0)) faceNormal = faceNormal / length;
            V3 dpdu = side1;
            if (!m.nrm.empty()) {
                its.shFrame.n = normalize(m.nrm[idx0] * b.x + m.nrm[idx1] * b.y + m.nrm[idx2] * b.z);
                if (dot(faceNormal, its.shFrame.n) < 0) faceNormal = -faceNormal;
            } else its.shFrame.n = faceNormal;
            its.geoFrame = Frame(faceNormal);
            if (!m.uv.empty()) {                                              // skdtree.h:399-406
                its.u = m.uv[2 * idx0] * b.x + m.uv[2 * idx1] * b.y + m.uv[2 * idx2] * b.z;
                its.v = m.uv[2 * idx0 + 1] * b.x + m.uv[2 * idx1 + 1] * b.y + m.uv[2 * idx2 + 1] * b.z;
            } else { its.u = b.y; its.v = b.z; }
            computeShadingFrame(its.shFrame.n, dpdu, its.shFrame);
            its.wi = its.shFrame.toLocal(-ray.d);
            return;
        }
        its.p = hit.p;
        const V3 axis = s.tangent(hit.iv);
        its.geoFrame.s = axis;
        const V3 relHitPoint = its.p - s.firstVertex(hit.iv);
        its.geoFrame.n = normalize(relHitPoint - dot(axis, relHitPoint) * axis);
        its.geoFrame.t = cross(its.geoFrame.n, its.geoFrame.s);
        const V3 local = its.geoFrame.toLocal(relHitPoint);
        its.p += its.geoFrame.n * (s.radius - std::sqrt(local.y * local.y + local.z * local.z));
        its.shFrame = its.geoFrame;
        V3 dpdu = its.geoFrame.s;
        computeShadingFrame(its.shFrame.n, dpdu, its.shFrame);
        its.wi = its.shFrame.toLocal(-ray.d);
    }

    bool rayIntersect(const Ray &ray, Intersection &its) const {
        Hit h;
        its = Intersection();
        if (!intersectBVH(ray, false, h)) return false;
        fillIntersection(ray, h, its);
        return true;
    }
    bool rayOccluded(const Ray &ray) const { Hit h; return intersectBVH(ray, true, h); }
};

} // namespace orc
