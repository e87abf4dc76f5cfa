// oracle/ref_shim/ref_pieces.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the three pieces of the reference that build standalone, straight from where they lie
// under /root/reference (nothing is copied into this repository):
//   src/bsdfs/gausssexylingerie.hpp, src/bsdfs/InterpolatedDistribution1D.hpp,
//   src/emitters/sunsky/skymodel.cpp (+ skymodel.h, skymodeldata.h)
// plus the CIE 1931 tables, which oracle/Makefile extracts from src/libcore/spectrum.cpp:743-1141
// into oracle/_ref/cie_tables.inc at build time.  Output: oracle/_ref/libref_pieces.so.
#include <cassert>
#include <cmath>
#include <vector>
#include <memory>
#include <algorithm>
#include <cstring>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define M_PI_FLT 3.14159265358979323846f
namespace mitsuba { namespace math {
    template <typename T> inline T clamp(T v, T lo, T hi) { return std::min(hi, std::max(lo, v)); }
} }

#include REF_GAUSS_LEGENDRE
#include REF_INTERP_DIST
#include REF_SKYMODEL_H

typedef float Float;
static const int CIE_samples = 471;
#include "cie_tables.inc"

extern "C" {

void ref_gauss_legendre_140(float *points, float *weights) {
    mitsuba::GaussLegendre<140> g;
    std::memcpy(points, g.points().data(), 140 * sizeof(float));
    std::memcpy(weights, g.weights().data(), 140 * sizeof(float));
}

void ref_interp_dist(const float *weights, int size, int num, int n, const float *distribution, const float *u,
                     float *outU, int *outX, float *outPdf, float *outSum) {
    mitsuba::InterpolatedDistribution1D d(std::vector<float>(weights, weights + size * num), size, num);
    for (int i = 0; i < n; ++i) {
        float uu = u[i]; int x;
        d.warp(distribution[i], uu, x);
        outU[i] = uu; outX[i] = x; outPdf[i] = d.pdf(distribution[i], x); outSum[i] = d.sum(distribution[i]);
    }
}

void *ref_sky_alloc(double turbidity, double albedo, double elevation) {
    return arhosek_rgb_skymodelstate_alloc_init(turbidity, albedo, elevation);
}
double ref_sky_radiance(void *state, double theta, double gamma, int channel) {
    return arhosek_tristim_skymodel_radiance((ArHosekTristimSkyModelState *) state, theta, gamma, channel);
}
void ref_sky_free(void *state) { arhosek_tristim_skymodelstate_free((ArHosekTristimSkyModelState *) state); }

int ref_cie_tables(const float **wl, const float **x, const float **y, const float **z) {
    *wl = CIE_wavelengths; *x = CIE_X_entries; *y = CIE_Y_entries; *z = CIE_Z_entries;
    return CIE_samples;
}

}
