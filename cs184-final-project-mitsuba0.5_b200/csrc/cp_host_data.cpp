// cp_host_data.cpp -- host-side set-up that stays on the CPU: file parsing and tiny tables.
//
// Replaces (reference file:line):
//   HairShape::HairShape(props) file loaders + vertex merge   src/shapes/hair.cpp:609-785
//   RoughTransmittance load / setEta / setAlpha / evalDiffuse  src/bsdfs/rtrans.h:81-149, 249-290, 292-384
//     (cubic tensor-spline lookups: src/libcore/spline.cpp:23-60, 236-304, evalCubicInterp3D)
//   GaussLegendre<140> nodes and weights                       src/bsdfs/gausssexylingerie.hpp:14-68
#include "cp_host.h"
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <array>

namespace cp {

// ------------------------------------------------------------------------------------------ Gauss-Legendre
namespace {
double legendreP(double x, int n) {
    if (n == 0) return 1.0;
    if (n == 1) return x;
    double pPrev = 1.0, pCur = x;
    for (int i = 2; i <= n; ++i) {
        double pNext = ((2.0 * i - 1.0) * x * pCur - (i - 1.0) * pPrev) / i;
        pPrev = pCur; pCur = pNext;
    }
    return pCur;
}
double legendreDP(double x, int n) { return n / (x * x - 1.0) * (x * legendreP(x, n) - legendreP(x, n - 1)); }
}
void gauss_legendre_140(float *points, float *weights) {
    const int N = 140;
    for (int k = 1; k <= N; ++k) {
        // Tricomi's initial guess (with the reference's fp32 pi), then Newton until |P_N| < 1e-6
        double x = std::cos(3.14159265358979323846f * (4.0 * k - 1.0) / (4.0 * N + 2.0)) * (1.0 - 1.0 / (8.0 * N * N) + 1.0 / (8.0 * N * N * N));
        for (int it = 0; it < 100; ++it) {
            double f = legendreP(x, N);
            x -= f / legendreDP(x, N);
            if (std::fabs(f) < 1e-6) break;
        }
        const float xf = float(x);
        points[k - 1] = xf;
        // the weight is evaluated at the fp32-rounded node, with fp32 (1 - x*x) promoted to double
        const double dp = legendreDP(xf, N);
        weights[k - 1] = float(2.0 / ((1.0 - xf * xf) * dp * dp));
    }
}

// ------------------------------------------------------------------------------------------ rough transmittance
namespace {
struct Axis { size_t knot; float w[4]; };
// Catmull-Rom knot weights on a uniform [0,1] grid with one-sided differences at the ends
bool axisWeights(float p, size_t size, Axis &a) {
    if (!(p >= 0.0f && p <= 1.0f)) return false;
    float t = (p * (size - 1)) / 1.0f;
    a.knot = std::min((size_t) t, size - 2);
    t -= (float) a.knot;
    const float t2 = t * t, t3 = t2 * t;
    a.w[0] = 0.0f; a.w[1] = 2 * t3 - 3 * t2 + 1; a.w[2] = -2 * t3 + 3 * t2; a.w[3] = 0.0f;
    const float d0 = t3 - 2 * t2 + t, d1 = t3 - t2;
    if (a.knot > 0) { a.w[2] += 0.5f * d0; a.w[0] -= 0.5f * d0; } else { a.w[2] += d0; a.w[1] -= d0; }
    if (a.knot + 2 < size) { a.w[3] += 0.5f * d1; a.w[1] -= 0.5f * d1; } else { a.w[2] += d1; a.w[1] -= d1; }
    return true;
}
float spline1(float x, const float *v, size_t n) {       // evalCubicInterp1D with [min,max] = [0,1]
    if (!(x >= 0.0f && x <= 1.0f)) return 0.0f;
    float t = (x * (n - 1)) / 1.0f;
    size_t k = std::min((size_t) t, n - 2);
    const float f0 = v[k], f1 = v[k + 1];
    const float d0 = k > 0 ? 0.5f * (v[k + 1] - v[k - 1]) : v[k + 1] - v[k];
    const float d1 = k + 2 < n ? 0.5f * (v[k + 2] - v[k]) : v[k + 1] - v[k];
    t -= (float) k;
    const float t2 = t * t, t3 = t2 * t;
    return (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
}
float spline2(float px, float py, const float *v, size_t nx, size_t ny) {
    Axis ax, ay;
    if (!axisWeights(px, nx, ax) || !axisWeights(py, ny, ay)) return 0.0f;
    float r = 0.0f;
    for (int j = -1; j <= 2; ++j) for (int i = -1; i <= 2; ++i) {
        const float w = ax.w[i + 1] * ay.w[j + 1];
        if (w == 0) continue;
        r += v[(ay.knot + j) * nx + ax.knot + i] * w;
    }
    return r;
}
float spline3(float px, float py, float pz, const float *v, size_t nx, size_t ny, size_t nz) {
    Axis ax, ay, az;
    if (!axisWeights(px, nx, ax) || !axisWeights(py, ny, ay) || !axisWeights(pz, nz, az)) return 0.0f;
    float r = 0.0f;
    for (int k = -1; k <= 2; ++k) for (int j = -1; j <= 2; ++j) {
        const float wyz = ay.w[j + 1] * az.w[k + 1];
        for (int i = -1; i <= 2; ++i) {
            const float w = ax.w[i + 1] * wyz;
            if (w == 0) continue;
            r += v[((az.knot + k) * ny + (ay.knot + j)) * nx + ax.knot + i] * w;
        }
    }
    return r;
}
}

// The slice only depends on (file, eta, alpha): results are remembered per process, so the BSDFs of a scene that share their parameters (the four
// hair colours of models/hair-curl) read and reduce the 2 MB table once.
static bool rough_transmittance_slice_uncached(const std::string &path, float eta, float alpha, std::vector<float> &outT, float &outFdr, std::string &err);
bool rough_transmittance_slice(const std::string &dataDir, int distribution, float eta, float alpha,
                               std::vector<float> &outT, float &outFdr, std::string &err) {
    static const char *names[3] = {"beckmann", "ggx", "phong"};
    if (distribution < 0 || distribution > 2) { err = "RoughTransmittance: unsupported distribution type!"; return false; }
    const std::string path = dataDir + "/microfacet/" + names[distribution] + ".dat";
    struct Entry { std::vector<float> T; float Fdr; };
    static std::mutex m; static std::map<std::tuple<std::string, uint32_t, uint32_t>, Entry> memo;
    uint32_t eb, ab; std::memcpy(&eb, &eta, 4); std::memcpy(&ab, &alpha, 4);
    const auto key = std::make_tuple(path, eb, ab);
    {
        std::lock_guard<std::mutex> g(m);
        auto it = memo.find(key);
        if (it != memo.end()) { outT = it->second.T; outFdr = it->second.Fdr; return true; }
    }
    if (!rough_transmittance_slice_uncached(path, eta, alpha, outT, outFdr, err)) return false;
    std::lock_guard<std::mutex> g(m);
    if (memo.size() < 256) memo[key] = Entry{outT, outFdr};
    return true;
}
static bool rough_transmittance_slice_uncached(const std::string &path, float eta, float alpha, std::vector<float> &outT, float &outFdr, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open \"" + path + "\" (set the data directory to a Mitsuba data/ tree)"; return false; }
    char hdr[17]; uint64_t dims[3]; float rng[4];
    bool ok = std::fread(hdr, 1, 17, f) == 17 && std::memcmp(hdr, "MTS_TRANSMITTANCE", 17) == 0 &&
              std::fread(dims, 8, 3, f) == 3 && std::fread(rng, 4, 4, f) == 4;
    if (!ok) { std::fclose(f); err = "Encountered an invalid transmittance data file!"; return false; }
    const size_t nEta = dims[0], nAlpha = dims[1], nTheta = dims[2];
    const float etaMin = rng[0], etaMax = rng[1], alphaMin = rng[2], alphaMax = rng[3];
    std::vector<float> raw((size_t) 2 * nEta * nAlpha * (nTheta + 1));
    ok = std::fread(raw.data(), 4, raw.size(), f) == raw.size();
    std::fclose(f);
    if (!ok) { err = "truncated transmittance data file"; return false; }
    // de-interleave: per (eta block, alpha) row = nTheta transmittance samples followed by one diffuse value
    std::vector<float> trans((size_t) 2 * nEta * nAlpha * nTheta), diff((size_t) 2 * nEta * nAlpha);
    for (size_t r = 0; r < 2 * nEta * nAlpha; ++r) {
        std::memcpy(&trans[r * nTheta], &raw[r * (nTheta + 1)], nTheta * 4);
        diff[r] = raw[r * (nTheta + 1) + nTheta];
    }
    if (alpha < alphaMin || alpha > alphaMax) { err = "the requested roughness value is outside of the supported range"; return false; }
    { float e = eta < 1 ? 1 / eta : eta; if (e < etaMin || e > etaMax) { err = "the requested relative index of refraction is outside of the supported range"; return false; } }
    const float warpedAlpha = std::pow((alpha - alphaMin) / (alphaMax - alphaMin), 0.25f);
    auto warpEta = [&](float e) { if (e < etaMin) e = etaMin; return std::pow((e - etaMin) / (etaMax - etaMin), 0.25f); };

    // external transmittance: setEta(eta) then setAlpha(alpha) -> nTheta samples
    {
        const float *block = trans.data(); float e = eta;
        if (e < 1) { block += nEta * nAlpha * nTheta; e = 1.0f / e; }
        const float we = warpEta(e);
        const float dAlpha = 1.0f / (nAlpha - 1), dTheta = 1.0f / (nTheta - 1);
        std::vector<float> slice2(nAlpha * nTheta);
        for (size_t i = 0; i < nAlpha; ++i) for (size_t j = 0; j < nTheta; ++j)
            slice2[i * nTheta + j] = spline3(j * dTheta, i * dAlpha, we, block, nTheta, nAlpha, nEta);
        outT.resize(nTheta);
        for (size_t i = 0; i < nTheta; ++i) outT[i] = spline2(i * dTheta, warpedAlpha, slice2.data(), nTheta, nAlpha);
    }
    // internal diffuse transmittance: clone -> setEta(1/eta) -> evalDiffuse(alpha); Fdr = 1 - that
    {
        const float *block = diff.data(); float e = 1 / eta;
        if (e < 1) { block += nEta * nAlpha; e = 1.0f / e; }
        const float we = warpEta(e);
        const float dAlpha = 1.0f / (nAlpha - 1);
        std::vector<float> slice1(nAlpha);
        for (size_t i = 0; i < nAlpha; ++i) slice1[i] = spline2(i * dAlpha, we, block, nAlpha, nEta);
        float v = spline1(warpedAlpha, slice1.data(), nAlpha);
        v = std::min(1.0f, std::max(0.0f, v));
        outFdr = 1 - v;
    }
    return true;
}

// ------------------------------------------------------------------------------------------ hair files
namespace {
struct Vec { float x, y, z; };
inline Vec sub(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dotv(Vec a, Vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec unit(Vec a) { float r = 1.0f / std::sqrt(dotv(a, a)); return {a.x * r, a.y * r, a.z * r}; }
inline Vec xfmP(const float *m, Vec p) {
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11], w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return {x, y, z};
    float r = 1.0f / w; return {x * r, y * r, z * r};
}
// Incremental fiber builder: drops duplicate points and merges a vertex into its predecessor while the
// tangent turns by less than the angle threshold (hair.cpp:686-715)
struct FiberBuilder {
    std::vector<float> &xyz; std::vector<uint8_t> &starts; float dpThresh;
    Vec tangent{0, 0, 0}, last{0, 0, 0}; bool tangentSet = false;
    void push(Vec p, bool s) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); starts.push_back(s ? 1 : 0); }
    Vec at(size_t i) const { return {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}; }
    void add(Vec p, bool newFiber) {
        if (newFiber) { push(p, true); last = p; tangentSet = false; return; }
        if (p.x == last.x && p.y == last.y && p.z == last.z) return;   // degenerate segment
        if (!tangentSet) { push(p, false); tangent = unit(sub(p, last)); tangentSet = !(tangent.x == 0 && tangent.y == 0 && tangent.z == 0); last = p; return; }
        Vec next = unit(sub(p, last));
        if (dotv(next, tangent) > dpThresh) {
            const size_t n = starts.size();
            tangent = unit(sub(p, at(n - 2)));
            xyz[3 * (n - 1)] = p.x; xyz[3 * (n - 1) + 1] = p.y; xyz[3 * (n - 1) + 2] = p.z;
        } else { push(p, false); tangent = next; }
        last = p;
    }
};
}

// Mitsuba's `Random` as the hair loader uses it (`reduction`, hair.cpp:629,672-673,769-770): SFMT-19937 seeded with the default 5489
// (include/mitsuba/core/random.h:113; src/libcore/random.cpp:72-96 parameters, :129-220 recursion, :325-347 period certification, :350-390
// regeneration, :396-405 init_gen_rand, :551-553 nextULong, :630-639 nextFloat).  The state is kept as 32-bit words; the reference's 64- and
// 128-bit views are the little-endian overlays of the same words.
namespace {
struct MitsubaRandom {
    static constexpr int N = 19937 / 128 + 1, N32 = N * 4, N64 = N * 2, POS1 = 122;
    uint32_t w[N32]; int idx;
    explicit MitsubaRandom(uint64_t seed = 5489ULL) {
        uint64_t prev = seed;
        w[0] = (uint32_t) prev; w[1] = (uint32_t) (prev >> 32);
        for (int i = 1; i < N64; ++i) { prev = 6364136223846793005ULL * (prev ^ (prev >> 62)) + (uint64_t) i; w[2 * i] = (uint32_t) prev; w[2 * i + 1] = (uint32_t) (prev >> 32); }
        idx = N32;
        static const uint32_t parity[4] = {0x00000001U, 0x00000000U, 0x00000000U, 0x13c9e684U};
        uint32_t inner = 0;
        for (int i = 0; i < 4; ++i) inner ^= w[i] & parity[i];
        for (int i = 16; i > 0; i >>= 1) inner ^= inner >> i;
        if ((inner & 1u) == 0u) {
            for (int i = 0; i < 4; ++i) { uint32_t work = 1; bool done = false;
                for (int j = 0; j < 32; ++j) { if (work & parity[i]) { w[i] ^= work; done = true; break; } work <<= 1; }
                if (done) break; }
        }
    }
    void recursion(uint32_t *r, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d) const {
        static const uint32_t msk[4] = {0xdfffffefU, 0xddfecb7fU, 0xbffaffffU, 0xbffffff6U};
        // x = a << 8 and y = c >> 8 as 128-bit integers (SL2 = SR2 = 1 byte)
        const uint64_t al = a[0] | ((uint64_t) a[1] << 32), ah = a[2] | ((uint64_t) a[3] << 32), cl = c[0] | ((uint64_t) c[1] << 32), ch = c[2] | ((uint64_t) c[3] << 32);
        const uint64_t xh = (ah << 8) | (al >> 56), xl = al << 8, yh = ch >> 8, yl = (cl >> 8) | (ch << 56);
        const uint32_t x[4] = {(uint32_t) xl, (uint32_t) (xl >> 32), (uint32_t) xh, (uint32_t) (xh >> 32)}, y[4] = {(uint32_t) yl, (uint32_t) (yl >> 32), (uint32_t) yh, (uint32_t) (yh >> 32)};
        for (int k = 0; k < 4; ++k) r[k] = a[k] ^ x[k] ^ ((b[k] >> 11) & msk[k]) ^ y[k] ^ (d[k] << 18);
    }
    void regenerate() {
        const uint32_t *r1 = w + 4 * (N - 2), *r2 = w + 4 * (N - 1);
        int i = 0;
        for (; i < N - POS1; ++i) { recursion(w + 4 * i, w + 4 * i, w + 4 * (i + POS1), r1, r2); r1 = r2; r2 = w + 4 * i; }
        for (; i < N; ++i) { recursion(w + 4 * i, w + 4 * i, w + 4 * (i + POS1 - N), r1, r2); r1 = r2; r2 = w + 4 * i; }
    }
    uint64_t nextULong() { if (idx >= N32) { regenerate(); idx = 0; } const uint64_t r = w[idx] | ((uint64_t) w[idx + 1] << 32); idx += 2; return r; }
    float nextFloat() { uint32_t u = (uint32_t) ((nextULong() & 0xFFFFFFFFull) >> 9) | 0x3f800000u; float f; std::memcpy(&f, &u, 4); return f - 1.0f; }
};
}
void mitsuba_random_floats(uint64_t seed, size_t n, float *out) { MitsubaRandom r(seed); for (size_t i = 0; i < n; ++i) out[i] = r.nextFloat(); }

bool load_hair_file(const std::string &path, float radius, float angleThresholdDeg, float reduction, const float toWorld[16],
                    HairFileData &out, std::string &err) {
    if (reduction < 0 || reduction >= 1) { err = "The 'reduction' parameter must have a value in [0, 1)!"; return false; }
    if (reduction > 0) radius *= 1.0f / (1 - reduction);      // hair.cpp:623-628: the surviving fibers are thickened
    MitsubaRandom random;                                     // :629; one draw per fiber marker decides whether the fiber is skipped
    bool ignore = false;
    out.xyz.clear(); out.startsFiber.clear();
    // radius scales with the length of toWorld * (0,0,1) (hair.cpp:632-633)
    { float x = toWorld[2], y = toWorld[6], z = toWorld[10]; out.radius = radius * std::sqrt(x * x + y * y + z * z); }
    FiberBuilder fb{out.xyz, out.startsFiber, std::cos(angleThresholdDeg * (3.14159265358979323846f / 180.0f))};
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "Could not open \"" + path + "\"!"; return false; }
    char magic[11] = {0};
    in.read(magic, 11);
    // the reference probes the 11-byte magic through FileStream::read, which throws on a shorter file whatever its format (fstream.cpp:317)
    if (in.gcount() != 11) { err = "Read less data than expected (11 bytes required) from \"" + path + "\""; return false; }
    if (std::memcmp(magic, "BINARY_HAIR", 11) == 0) {
        uint32_t count = 0;
        in.read((char *) &count, 4);
        bool newFiber = true;
        auto rd = [&](float &v) { in.read((char *) &v, 4); return (bool) in; };
        for (uint32_t n = 0; n < count; ++n) {
            float a; Vec p;
            if (!rd(a)) { err = "unexpected end of hair file"; return false; }
            if (std::isinf(a)) {
                newFiber = true; if (!rd(p.x) || !rd(p.y) || !rd(p.z)) { err = "unexpected end of hair file"; return false; }
                if (reduction > 0) ignore = random.nextFloat() < reduction;          // :672-673
            } else { p.x = a; if (!rd(p.y) || !rd(p.z)) { err = "unexpected end of hair file"; return false; } }
            if (!ignore) fb.add(xfmP(toWorld, p), newFiber);
            newFiber = false;
        }
    } else {
        in.close();
        std::ifstream txt(path);
        std::string line; bool newFiber = true;
        while (txt.good()) {
            std::getline(txt, line);
            if (!line.empty() && line[0] == '#') { newFiber = true; continue; }
            std::istringstream iss(line);
            Vec p;
            iss >> p.x >> p.y >> p.z;
            if (iss.fail()) { newFiber = true; if (reduction > 0) ignore = random.nextFloat() < reduction; continue; }      // :767-770
            if (!ignore) fb.add(xfmP(toWorld, p), newFiber);
            newFiber = false;
        }
    }
    return true;
}

// Matrix<4,4,float>::invert (include/mitsuba/core/matrix.inl:138-193): Gauss-Jordan with full pivoting, in place, fp32, one rounding per
// operation -- what Transform(const Matrix4x4 &) runs on every <matrix> of a scene file.  Row-major 4x4; false when singular.
bool mat4_invert_f32(const float *a, float *out) {
    const int N = 4;
    int indxc[N], indxr[N], ipiv[N] = {0, 0, 0, 0};
    float (*m)[4] = reinterpret_cast<float (*)[4]>(out);
    std::memcpy(out, a, 64);
    for (int i = 0; i < N; i++) {
        int irow = -1, icol = -1;
        float big = 0;
        for (int j = 0; j < N; j++) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < N; k++) {
                    if (ipiv[k] == 0) {
                        if (std::fabs(m[j][k]) >= big) { big = std::fabs(m[j][k]); irow = j; icol = k; }
                    } else if (ipiv[k] > 1) return false;
                }
            }
        }
        ++ipiv[icol];
        if (irow != icol) for (int k = 0; k < N; ++k) std::swap(m[irow][k], m[icol][k]);
        indxr[i] = irow; indxc[i] = icol;
        if (m[icol][icol] == 0) return false;
        const volatile float pivinv = 1.f / m[icol][icol];
        m[icol][icol] = 1.f;
        for (int j = 0; j < N; j++) m[icol][j] *= pivinv;
        for (int j = 0; j < N; j++) {
            if (j != icol) {
                const float save = m[j][icol];
                m[j][icol] = 0;
                for (int k = 0; k < N; k++) { const volatile float prod = m[icol][k] * save; m[j][k] -= prod; }    // no fused multiply-add, like the x86 build of the reference
            }
        }
    }
    for (int j = N - 1; j >= 0; j--)
        if (indxr[j] != indxc[j]) for (int k = 0; k < N; k++) std::swap(m[k][indxr[j]], m[k][indxc[j]]);
    return true;
}

// ------------------------------------------------------------------------------------------ Wavefront OBJ meshes
// WavefrontOBJ(props) with collapse = true semantics: every face of the file lands in ONE mesh (src/shapes/obj.cpp:186-349):
// `v` / `vn` / `vt` / `f` lines (n-gons as a fan :316-323, negative indices :640-645), vertices transformed by toWorld and merged
// when position, normal and uv agree (createMesh :608-700, key order :584-606), then TriMesh::computeNormals
// (src/librender/trimesh.cpp:608-672): faceNormals drops the normals (flipNormals then swaps the winding), given normals are kept
// (negated by flipNormals), missing normals are generated angle-weighted (Thuermer & Wuethrich).  Materials (mtllib / usemtl)
// and groups are ignored: the mesh takes the bsdf of the <shape> element.
namespace {
inline Vec crossv(Vec a, Vec b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float lenv(Vec a) { return std::sqrt(dotv(a, a)); }
inline Vec divv(Vec a, float f) { const float r = 1.0f / f; return {a.x * r, a.y * r, a.z * r}; }   // vector.h: multiply by the reciprocal
// util.h:309-314
inline float unit_angle(Vec u, Vec v) {
    const float kPiF = 3.14159265358979323846f;
    if (dotv(u, v) < 0) return kPiF - 2 * std::asin(0.5f * lenv({v.x + u.x, v.y + u.y, v.z + u.z}));
    return 2 * std::asin(0.5f * lenv(sub(v, u)));
}
}

bool load_obj_file(const std::string &path, const float toWorld[16], bool faceNormals, bool flipNormals, bool flipTexCoords, MeshFileData &out, std::string &err) {
    out = MeshFileData();
    std::ifstream is(path);
    if (!is) { err = "Wavefront OBJ file '" + path + "' not found!"; return false; }
    // normals transform with the transpose of the inverse (transform.h:203-211); the inverse is the one Transform(const Matrix4x4 &) computes
    float inv[16];
    if (!mat4_invert_f32(toWorld, inv)) { err = "obj: singular toWorld transform"; return false; }
    auto xfmN = [&](Vec v) -> Vec {
        return {inv[0] * v.x + inv[4] * v.y + inv[8] * v.z, inv[1] * v.x + inv[5] * v.y + inv[9] * v.z, inv[2] * v.x + inv[6] * v.y + inv[10] * v.z};
    };
    std::vector<Vec> vertices, normals; std::vector<std::array<float, 2>> texcoords;
    struct Corner { int p = 0, n = 0, uv = 0; };
    std::vector<std::array<Corner, 3>> faces;
    auto parseCorner = [&](const std::string &str, Corner &c) -> bool {          // obj.cpp:371-390
        std::vector<std::string> tok; size_t b = 0;
        while (b <= str.size()) { size_t e = str.find('/', b); if (e == std::string::npos) e = str.size(); if (e > b) tok.push_back(str.substr(b, e - b)); b = e + 1; }
        c = Corner();
        if (tok.size() == 1) c.p = atoi(tok[0].c_str());
        else if (tok.size() == 2) { c.p = atoi(tok[0].c_str()); if (str.find("//") == std::string::npos) c.uv = atoi(tok[1].c_str()); else c.n = atoi(tok[1].c_str()); }
        else if (tok.size() == 3) { c.p = atoi(tok[0].c_str()); c.uv = atoi(tok[1].c_str()); c.n = atoi(tok[2].c_str()); }
        else return false;
        return true;
    };
    // fetch_line (obj.cpp:165-187): trailing blanks / CR are cut; a line ending in a backslash is glued to the next one (nothing in between)
    std::function<bool(std::string &)> fetchLine = [&](std::string &l) -> bool {
        if (!std::getline(is, l)) return false;
        size_t n = l.size();
        while (n > 0 && (l[n - 1] == '\r' || l[n - 1] == '\n' || l[n - 1] == '\t' || l[n - 1] == ' ')) --n;
        if (n > 0 && l[n - 1] == '\\') { std::string next; fetchLine(next); l = l.substr(0, n - 1) + next; }
        else l.resize(n);
        return true;
    };
    std::string line, buf;
    while (is.good() && !is.eof() && fetchLine(line)) {
        std::istringstream iss(line);
        if (!(iss >> buf)) continue;
        if (buf == "v") { Vec p{0, 0, 0}; iss >> p.x >> p.y >> p.z; vertices.push_back(p); }
        else if (buf == "vn") { Vec n{0, 0, 0}; iss >> n.x >> n.y >> n.z; normals.push_back(n); }
        else if (buf == "vt") { float u = 0, v = 0; iss >> u >> v; if (flipTexCoords) v = 1 - v; texcoords.push_back({u, v}); }
        else if (buf == "f") {
            std::string tmp; std::array<Corner, 3> t;
            // obj.cpp:311-314 does not check the extraction: a face with fewer than three corners repeats its last token (a degenerate triangle)
            for (int k = 0; k < 3; ++k) { iss >> tmp; if (!parseCorner(tmp, t[k])) { err = "Invalid OBJ face format!"; return false; } }
            faces.push_back(t);
            while (iss >> tmp) { t[1] = t[2]; if (!parseCorner(tmp, t[2])) { err = "Invalid OBJ face format!"; return false; } faces.push_back(t); }
        }
    }
    if (faces.empty()) { err = "obj: the file contains no faces"; return false; }
    struct Key { float v[8]; bool operator<(const Key &o) const { for (int i = 0; i < 8; ++i) { if (v[i] < o.v[i]) return true; if (v[i] > o.v[i]) return false; } return false; } };
    std::map<Key, uint32_t> vertexMap;
    bool hasNormals = false;
    std::vector<Vec> P, N; std::vector<std::array<float, 2>> UV; bool hasTexcoords = false;
    for (auto &f : faces) {
        for (int j = 0; j < 3; ++j) {
            int vid = f[j].p, nid = f[j].n, uid = f[j].uv;
            if (vid < 0) vid += (int) vertices.size() + 1;
            if (nid < 0) nid += (int) normals.size() + 1;
            if (uid < 0) uid += (int) texcoords.size() + 1;
            if (vid > (int) vertices.size() || vid <= 0) { err = "Out of bounds: tried to access vertex " + std::to_string(vid) + " (max: " + std::to_string(vertices.size()) + ")"; return false; }
            const Vec p = xfmP(toWorld, vertices[vid - 1]);
            Vec n{0, 0, 0};
            if (nid != 0) {
                if (nid > (int) normals.size() || nid < 0) { err = "Out of bounds: tried to access normal " + std::to_string(nid) + " (max: " + std::to_string(normals.size()) + ")"; return false; }
                n = xfmN(normals[nid - 1]);
                if (!(n.x == 0 && n.y == 0 && n.z == 0)) n = divv(n, lenv(n));
                hasNormals = true;
            }
            float uv[2] = {0, 0};
            if (uid != 0) {
                if (uid > (int) texcoords.size() || uid < 0) { err = "Out of bounds: tried to access uv " + std::to_string(uid) + " (max: " + std::to_string(texcoords.size()) + ")"; return false; }
                uv[0] = texcoords[uid - 1][0]; uv[1] = texcoords[uid - 1][1];
                hasTexcoords = true;                                              // obj.cpp:633-636
            }
            const Key key{{p.x, p.y, p.z, n.x, n.y, n.z, uv[0], uv[1]}};
            auto it = vertexMap.find(key);
            uint32_t id;
            if (it != vertexMap.end()) id = it->second;
            else { id = (uint32_t) P.size(); vertexMap[key] = id; P.push_back(p); N.push_back(n); UV.push_back({uv[0], uv[1]}); }
            out.indices.push_back(id);
        }
    }
    const size_t nTri = out.indices.size() / 3;
    // TriMesh::computeNormals (trimesh.cpp:608-672)
    if (faceNormals) {
        hasNormals = false;
        if (flipNormals) for (size_t i = 0; i < nTri; ++i) std::swap(out.indices[3 * i], out.indices[3 * i + 1]);
    } else if (hasNormals) {
        if (flipNormals) for (Vec &n : N) n = {n.x * -1, n.y * -1, n.z * -1};
    } else {
        std::fill(N.begin(), N.end(), Vec{0, 0, 0});
        for (size_t i = 0; i < nTri; ++i) {
            Vec n{0, 0, 0};
            for (int k = 0; k < 3; ++k) {
                const Vec v0 = P[out.indices[3 * i + k]], v1 = P[out.indices[3 * i + (k + 1) % 3]], v2 = P[out.indices[3 * i + (k + 2) % 3]];
                const Vec sideA = sub(v1, v0), sideB = sub(v2, v0);
                if (k == 0) {
                    n = crossv(sideA, sideB);
                    const float length = lenv(n);
                    if (length == 0) break;
                    n = divv(n, length);
                }
                const float angle = unit_angle(divv(sideA, lenv(sideA)), divv(sideB, lenv(sideB)));
                Vec &dst = N[out.indices[3 * i + k]];
                dst = {dst.x + n.x * angle, dst.y + n.y * angle, dst.z + n.z * angle};
            }
        }
        for (Vec &n : N) {
            float length = lenv(n);
            if (flipNormals) length *= -1;
            if (length != 0) n = divv(n, length); else n = {1, 0, 0};
        }
        hasNormals = true;
    }
    out.xyz.reserve(3 * P.size());
    for (const Vec &p : P) { out.xyz.push_back(p.x); out.xyz.push_back(p.y); out.xyz.push_back(p.z); }
    if (hasNormals) { out.normals.reserve(3 * N.size()); for (const Vec &n : N) { out.normals.push_back(n.x); out.normals.push_back(n.y); out.normals.push_back(n.z); } }
    if (hasTexcoords) { out.uvs.reserve(2 * UV.size()); for (const auto &t : UV) { out.uvs.push_back(t[0]); out.uvs.push_back(t[1]); } }   // obj.cpp:672-675
    return true;
}

// ------------------------------------------------------------------------------------------------------------------------------
// fresnelDiffuseReflectance(eta, false), src/libcore/util.cpp:807-862: GaussLobattoIntegrator(1024, 0, 1e-5f) (src/libcore/quad.cpp:287-420,
// useConvergenceEstimate = true) over xi -> fresnelDielectricExt(sqrt(xi), eta) on [0, 1].  fp32 throughout, like the reference; the six
// sub-intervals of a refinement step are summed left to right.  `plastic` stores the result for eta and 1/eta (plastic.cpp:194-195).
namespace {
struct LobattoRule {
    float eta, absTol = 0; size_t evals = 0; static constexpr size_t maxEvals = 1024;
    const float alpha = (float) std::sqrt(2.0 / 3.0), beta = (float) (1.0 / std::sqrt(5.0));
    float f(float xi) const { return fresnelDielectricExt(std::sqrt(xi), eta); }
    float tolerance(float a, float b) {                                           // calculateAbsTolerance, relError = 1e-5, absError = 0
        const float x1 = (float) 0.94288241569547971906, x2 = (float) 0.64185334234578130578, x3 = (float) 0.23638319966214988028;
        const float m = (a + b) / 2, h = (b - a) / 2;
        const float y1 = f(a), y3 = f(m - alpha * h), y5 = f(m - beta * h), y7 = f(m), y9 = f(m + beta * h), y11 = f(m + alpha * h), y13 = f(b);
        const float p1 = f(m - x1 * h) + f(m + x1 * h), p2 = f(m - x2 * h) + f(m + x2 * h), p3 = f(m - x3 * h) + f(m + x3 * h);
        const float acc = h * ((float) 0.0158271919734801831 * (y1 + y13) + (float) 0.0942738402188500455 * p1 + (float) 0.1550719873365853963 * (y3 + y11)
                             + (float) 0.1888215739601824544 * p2 + (float) 0.1997734052268585268 * (y5 + y9) + (float) 0.2249264653333395270 * p3
                             + (float) 0.2426110719014077338 * y7);
        evals += 13;
        const float integral2 = (h / 6) * (y1 + y13 + 5 * (y5 + y9));
        const float integral1 = (h / 1470) * (77 * (y1 + y13) + 432 * (y3 + y11) + 625 * (y5 + y9) + 672 * y7);
        float r = 1.0f;
        if (std::abs(integral2 - acc) != 0.0f) r = std::abs(integral1 - acc) / std::abs(integral2 - acc);
        if (r == 0.0f || r > 1.0f) r = 1.0f;
        const float eps = std::numeric_limits<float>::epsilon();
        float result = std::numeric_limits<float>::infinity();
        if (acc != 0) result = acc * std::max(1e-5f, eps) / (r * eps);
        return result;
    }
    float refine(float a, float b, float fa, float fb) {                           // adaptiveGaussLobattoStep
        const float h = (b - a) / 2, m = (a + b) / 2;
        const float mll = m - alpha * h, ml = m - beta * h, mr = m + beta * h, mrr = m + alpha * h;
        const float fmll = f(mll), fml = f(ml), fm = f(m), fmr = f(mr), fmrr = f(mrr);
        const float integral2 = (h / 6) * (fa + fb + 5 * (fml + fmr));
        const float integral1 = (h / 1470) * (77 * (fa + fb) + 432 * (fmll + fmrr) + 625 * (fml + fmr) + 672 * fm);
        evals += 5;
        if (evals >= maxEvals) return integral1;
        const float dist = absTol + (integral1 - integral2);
        if (dist == absTol || mll <= a || b <= mrr) return integral1;
        float sum = refine(a, mll, fa, fmll);
        sum = sum + refine(mll, ml, fmll, fml); sum = sum + refine(ml, m, fml, fm); sum = sum + refine(m, mr, fm, fmr);
        sum = sum + refine(mr, mrr, fmr, fmrr); sum = sum + refine(mrr, b, fmrr, fb);
        return sum;
    }
};
}
float fresnel_diffuse_reflectance(float eta) {
    LobattoRule q; q.eta = eta;
    q.absTol = q.tolerance(0.0f, 1.0f);
    q.evals += 2;
    return q.refine(0.0f, 1.0f, q.f(0.0f), q.f(1.0f));
}

} // namespace cp

// ------------------------------------------------------------------------------------------------------------------------------
// Radiance RGBE (.hdr) reader: Bitmap::readRGBE, src/libcore/bitmap.cpp:3590-3678 (+ RGBE_ToFloat :3522-3530, RGBE_ReadPixels :3579-3586,
// Stream::readLine src/libcore/stream.cpp:392-414).  Output: top-down lat-long RGB fp32, what the `envmap` emitter hands to its MIP map.
namespace cp {
namespace {
struct ByteReader {
    const std::vector<unsigned char> &d; size_t pos = 0;
    explicit ByteReader(const std::vector<unsigned char> &v) : d(v) {}
    void read(unsigned char *out, size_t n) { if (pos + n > d.size()) throw std::runtime_error("readRGBE(): unexpected end of file"); std::memcpy(out, d.data() + pos, n); pos += n; }
    std::string line() {               // drops CR, ends at LF; a last line without LF is returned as it is
        std::string r;
        for (;;) {
            if (pos >= d.size()) { if (!r.empty()) return r; throw std::runtime_error("readRGBE(): unexpected end of file"); }
            const char c = (char) d[pos++];
            if (c == 10) return r;
            if (c != 13) r += c;
        }
    }
};
inline void rgbe_to_float(const unsigned char q[4], float *out) {
    if (q[3]) { const float f = std::ldexp(1.0f, (int) q[3] - (128 + 8)); for (int i = 0; i < 3; ++i) out[i] = q[i] * f; }
    else out[0] = out[1] = out[2] = 0.0f;
}
}

bool load_rgbe_file(const std::string &path, std::vector<float> &rgb, int &w, int &h, std::string &err) {
    try {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("Environment map file \"" + path + "\" could not be found!");
        std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        ByteReader in(bytes);
        std::string line = in.line();
        if (line.length() < 2 || line[0] != '#' || line[1] != '?') throw std::runtime_error("readRGBE(): Invalid header!");
        bool recognised = false;
        w = h = 0;
        for (;;) {
            line = in.line();
            if (line.rfind("FORMAT=32-bit_rle_rgbe", 0) == 0) recognised = true;
            if (line.rfind("-Y ", 0) == 0) {
                if (std::sscanf(line.c_str(), "-Y %i +X %i", &h, &w) < 2) throw std::runtime_error("readRGBE(): parser error!");
                break;
            }
        }
        if (!recognised) throw std::runtime_error("readRGBE(): invalid format!");
        if (w <= 0 || h <= 0) throw std::runtime_error("readRGBE(): invalid image size!");
        rgb.assign((size_t) 3 * w * h, 0.0f);
        float *data = rgb.data();
        auto flat = [&](float *dst, size_t n) { unsigned char q[4]; while (n-- > 0) { in.read(q, 4); rgbe_to_float(q, dst); dst += 3; } };
        if (w < 8 || w > 0x7fff) { flat(data, (size_t) w * h); return true; }        // run-length encoding is not allowed for these widths
        std::vector<unsigned char> buffer((size_t) 4 * w);
        for (int y = 0; y < h; ++y) {
            unsigned char q[4];
            in.read(q, 4);
            if (q[0] != 2 || q[1] != 2 || (q[2] & 0x80)) {                           // not run-length encoded: the rest of the file is flat
                rgbe_to_float(q, data);
                flat(data + 3, (size_t) w * h - 1);                                    // (as in the reference, also when y > 0)
                return true;
            }
            if ((((int) q[2]) << 8 | q[3]) != w) throw std::runtime_error("readRGBE(): wrong scanline width!");
            unsigned char *ptr = buffer.data();
            for (int i = 0; i < 4; ++i) {
                unsigned char *end = buffer.data() + (size_t) (i + 1) * w;
                while (ptr < end) {
                    unsigned char b[2];
                    in.read(b, 2);
                    if (b[0] > 128) {
                        int count = b[0] - 128;
                        if (count == 0 || count > end - ptr) throw std::runtime_error("readRGBE(): bad scanline data!");
                        while (count-- > 0) *ptr++ = b[1];
                    } else {
                        int count = b[0];
                        if (count == 0 || count > end - ptr) throw std::runtime_error("readRGBE(): bad scanline data!");
                        *ptr++ = b[1];
                        if (--count > 0) in.read(ptr, (size_t) count);
                        ptr += count;
                    }
                }
            }
            for (int i = 0; i < w; ++i) {
                q[0] = buffer[i]; q[1] = buffer[(size_t) w + i]; q[2] = buffer[(size_t) 2 * w + i]; q[3] = buffer[(size_t) 3 * w + i];
                rgbe_to_float(q, data);
                data += 3;
            }
        }
        return true;
    } catch (const std::exception &e) { err = e.what(); return false; }
}
} // namespace cp

// ------------------------------------------------------------------------------------------------------------------------------
// OpenEXR writer for `hdrfilm` (src/films/hdrfilm.cpp:213-246 defaults: fileFormat openexr, pixelFormat rgb, componentFormat float16;
// HDRFilm::develop -> Bitmap::write(EOpenEXR), src/libcore/bitmap.cpp writeOpenEXR): a single-part scan-line file, channels B, G, R
// (the order the format prescribes), increasing-y line order, one scan line per chunk, no compression -- every OpenEXR reader accepts it.
// float -> half rounds to nearest even with overflow to infinity, like OpenEXR's half(float).
namespace cp {
namespace {
inline uint16_t exr_half(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t) (sign | 0x7c00u | ((x > 0x7f800000u) ? (0x200u | ((x >> 13) & 0x3ffu)) : 0u));   // inf / NaN
    if (x >= 0x477ff000u) return (uint16_t) (sign | 0x7c00u);                                                               // rounds to or beyond 65520: infinity
    if (x < 0x33000001u) return (uint16_t) sign;                                                                             // below half of the smallest subnormal
    if (x < 0x38800000u) {                                                                                                  // subnormal half
        const int shift = 126 - (int) (x >> 23);                          // 14 .. 24: value = m 2^(E-150), one half subnormal = 2^-24
        const uint32_t m = (x & 0x7fffffu) | 0x800000u;
        const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        return (uint16_t) (sign | (q + ((rem > halfway || (rem == halfway && (q & 1u))) ? 1u : 0u)));
    }
    uint32_t h = ((x - 0x38000000u) >> 13);
    const uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t) (sign | h);
}
struct ExrOut {
    std::vector<unsigned char> b;
    void raw(const void *p, size_t n) { const unsigned char *q = (const unsigned char *) p; b.insert(b.end(), q, q + n); }
    void str(const char *s) { raw(s, std::strlen(s) + 1); }
    void i32(int32_t v) { raw(&v, 4); }
    void f32(float v) { raw(&v, 4); }
    void attr(const char *name, const char *type, int32_t size) { str(name); str(type); i32(size); }
};
}
bool write_exr_file(const std::string &path, const float *rgb, int w, int h, bool half, std::string &err) {
    if (!rgb || w <= 0 || h <= 0) { err = "write_exr: empty image"; return false; }
    ExrOut o;
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    o.raw(magic, 8);
    const int32_t ptype = half ? 1 : 2;
    o.attr("channels", "chlist", 3 * (2 + 16) + 1);
    for (const char *c : {"B", "G", "R"}) { o.str(c); o.i32(ptype); const unsigned char lin[4] = {0, 0, 0, 0}; o.raw(lin, 4); o.i32(1); o.i32(1); }
    o.b.push_back(0);
    o.attr("compression", "compression", 1); o.b.push_back(0);
    o.attr("dataWindow", "box2i", 16); o.i32(0); o.i32(0); o.i32(w - 1); o.i32(h - 1);
    o.attr("displayWindow", "box2i", 16); o.i32(0); o.i32(0); o.i32(w - 1); o.i32(h - 1);
    o.attr("lineOrder", "lineOrder", 1); o.b.push_back(0);
    o.attr("pixelAspectRatio", "float", 4); o.f32(1.0f);
    o.attr("screenWindowCenter", "v2f", 8); o.f32(0.0f); o.f32(0.0f);
    o.attr("screenWindowWidth", "float", 4); o.f32(1.0f);
    o.b.push_back(0);
    const size_t bpp = half ? 2 : 4, lineBytes = 3 * bpp * (size_t) w, tableAt = o.b.size();
    o.b.resize(tableAt + 8 * (size_t) h);
    for (int y = 0; y < h; ++y) {
        const uint64_t off = o.b.size();
        std::memcpy(&o.b[tableAt + 8 * (size_t) y], &off, 8);
        o.i32(y); o.i32((int32_t) lineBytes);
        for (int c = 2; c >= 0; --c) {                                   // B, G, R
            if (half) for (int x = 0; x < w; ++x) { const uint16_t v = exr_half(rgb[((size_t) y * w + x) * 3 + c]); o.raw(&v, 2); }
            else for (int x = 0; x < w; ++x) o.f32(rgb[((size_t) y * w + x) * 3 + c]);
        }
    }
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write \"" + path + "\""; return false; }
    const bool ok = fwrite(o.b.data(), 1, o.b.size(), f) == o.b.size();
    if (fclose(f) != 0 || !ok) { err = "cannot write \"" + path + "\""; return false; }
    return true;
}
void float_to_half_array(const float *in, size_t n, uint16_t *out) { for (size_t i = 0; i < n; ++i) out[i] = exr_half(in[i]); }
} // namespace cp
