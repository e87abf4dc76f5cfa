// oracle/ref_shim/ref_develop.cpp -- TEST INFRASTRUCTURE ONLY.
// Film develop of the reference executed as written: FormatConverterImpl::undoGamma / applyGamma / convertScalar
// (src/libcore/fmtconv.cpp:1093-1160) cut out at build time (oracle/_ref/ref_fmtconv_scalar.inc) into a struct that supplies
// format_traits (:86-99) and detail::safe_cast (:70-73); the pixel loop around them is the ESpectrumAlphaWeight -> ERGB case of
// FormatConverterImpl::convert (:984-995) with SPECTRUM_SAMPLES = 3 (Spectrum::toLinearRGB is the identity then).
// What LDRFilm::develop asks for: Bitmap::convert(ERGB, EUInt8, gamma, 2^exposure) (src/films/ldrfilm.cpp:300-321), and
// HDRFilm::develop: convert(ERGB, EFloat32, 1.0) (src/films/hdrfilm.cpp).  Part of oracle/_ref/libref_geom.so.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <type_traits>

namespace refdevelop {
typedef float Float;
namespace detail { template <typename T, typename S> inline T safe_cast(S a) { return static_cast<T>(a); } }
struct Color3 { Float v[3]; Float &operator[](int i) { return v[i]; } const Float &operator[](int i) const { return v[i]; } };
struct Converter {
    template <typename FormatType> struct format_traits {
        static const bool is_float = std::is_same<FormatType, double>::value || std::is_same<FormatType, float>::value;
        static const bool is_compact = std::is_same<FormatType, uint8_t>::value || std::is_same<FormatType, uint16_t>::value;
    };
#include "ref_fmtconv_scalar.inc"
    template <typename DestFormat> static void convertRGB(const float *source, size_t count, DestFormat *dest, Float sourceGamma, Float destGamma, Float multiplier) {
        const Float invDestGamma = 1.0f / destGamma;                                              // fmtconv.cpp:147 (-1, the sRGB curve, stays -1)
        for (size_t i = 0; i < count; ++i) {                                                      // fmtconv.cpp:984-995
            Float spec[3];
            for (int j = 0; j < 3; ++j) spec[j] = convertScalar<Float>(*source++, sourceGamma);
            source++;
            Float weight = convertScalar<Float>(*source++), invWeight = (weight != 0) ? 1 / weight : weight;
            Float r = spec[0] * invWeight, g = spec[1] * invWeight, b = spec[2] * invWeight;
            *dest++ = convertScalar<DestFormat>(r, 1.0f, (DestFormat *) NULL, multiplier, invDestGamma);
            *dest++ = convertScalar<DestFormat>(g, 1.0f, (DestFormat *) NULL, multiplier, invDestGamma);
            *dest++ = convertScalar<DestFormat>(b, 1.0f, (DestFormat *) NULL, multiplier, invDestGamma);
        }
    }
};
}
extern "C" void ref_develop_ldr(const float *film, size_t count, float gamma, float exposure, uint8_t *out) {
    refdevelop::Converter::convertRGB<uint8_t>(film, count, out, 1.0f, gamma, std::pow(2.0f, exposure));      // ldrfilm.cpp:305,318-320
}
extern "C" void ref_develop_hdr(const float *film, size_t count, float *out) {
    refdevelop::Converter::convertRGB<float>(film, count, out, 1.0f, 1.0f, 1.0f);
}
