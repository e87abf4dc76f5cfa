// cp_host_data.cpp -- host-side set-up that stays on the CPU: file parsing and tiny tables.
//
// Replaces (reference file:line):
//   HairShape::HairShape(props) file loaders + vertex merge   src/shapes/hair.cpp:609-785
//   RoughTransmittance load / setEta / setAlpha / evalDiffuse  src/bsdfs/rtrans.h:81-149, 249-290, 292-384
//     (cubic tensor-spline lookups: src/libcore/spline.cpp:23-60, 236-304, evalCubicInterp3D)
//   GaussLegendre<140> nodes and weights                       src/bsdfs/gausssexylingerie.hpp:14-68
#include "cp_host.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <algorithm>

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

bool rough_transmittance_slice(const std::string &dataDir, int distribution, float eta, float alpha,
                               std::vector<float> &outT, float &outFdr, std::string &err) {
    static const char *names[3] = {"beckmann", "ggx", "phong"};
    if (distribution < 0 || distribution > 2) { err = "RoughTransmittance: unsupported distribution type!"; return false; }
    const std::string path = dataDir + "/microfacet/" + names[distribution] + ".dat";
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

bool load_hair_file(const std::string &path, float radius, float angleThresholdDeg, float reduction, const float toWorld[16],
                    HairFileData &out, std::string &err) {
    if (reduction < 0 || reduction >= 1) { err = "The 'reduction' parameter must have a value in [0, 1)!"; return false; }
    if (reduction > 0) { err = "reduction > 0 draws from the reference's Mersenne-Twister stream and is not supported"; return false; }
    out.xyz.clear(); out.startsFiber.clear();
    // radius scales with the length of toWorld * (0,0,1) (hair.cpp:632-633)
    { float x = toWorld[2], y = toWorld[6], z = toWorld[10]; out.radius = radius * std::sqrt(x * x + y * y + z * z); }
    FiberBuilder fb{out.xyz, out.startsFiber, std::cos(angleThresholdDeg * (3.14159265358979323846f / 180.0f))};
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "Could not open \"" + path + "\"!"; return false; }
    char magic[11] = {0};
    in.read(magic, 11);
    if (in.gcount() == 11 && std::memcmp(magic, "BINARY_HAIR", 11) == 0) {
        uint32_t count = 0;
        in.read((char *) &count, 4);
        bool newFiber = true;
        auto rd = [&](float &v) { in.read((char *) &v, 4); return (bool) in; };
        for (uint32_t n = 0; n < count; ++n) {
            float a; Vec p;
            if (!rd(a)) { err = "unexpected end of hair file"; return false; }
            if (std::isinf(a)) { newFiber = true; if (!rd(p.x) || !rd(p.y) || !rd(p.z)) { err = "unexpected end of hair file"; return false; } }
            else { p.x = a; if (!rd(p.y) || !rd(p.z)) { err = "unexpected end of hair file"; return false; } }
            fb.add(xfmP(toWorld, p), newFiber);
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
            if (iss.fail()) { newFiber = true; continue; }
            fb.add(xfmP(toWorld, p), newFiber);
            newFiber = false;
        }
    }
    return true;
}

} // namespace cp
