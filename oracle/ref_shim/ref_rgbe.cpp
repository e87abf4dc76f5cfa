// oracle/ref_shim/ref_rgbe.cpp -- TEST INFRASTRUCTURE ONLY.
// The Radiance RGBE reader of the reference executed as written: Bitmap::readRGBE (src/libcore/bitmap.cpp:3590-3678) with its helpers
// RGBE_ToFloat (:3522-3530) and RGBE_ReadPixels (:3579-3586), cut out at build time (oracle/_ref/ref_rgbe_*.inc) and pasted into a class that
// supplies the members they touch; the stream follows Stream::readLine (src/libcore/stream.cpp:392-414).  Part of oracle/_ref/libref_geom.so.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { inline bool starts_with(const std::string &s, const char *prefix) { return s.compare(0, std::strlen(prefix), prefix) == 0; } }
namespace refrgbe {
enum ELogLevel { EDebug, EInfo, EWarn, EError };
#define Log(level, ...) do { if (level >= EError) throw std::runtime_error(firstArg(__VA_ARGS__)); } while (0)
inline const char *firstArg(const char *fmt, ...) { return fmt; }
struct Stream {
    std::vector<unsigned char> d; size_t pos = 0;
    void read(void *dst, size_t n) { if (pos + n > d.size()) throw std::runtime_error("Read less data than expected"); std::memcpy(dst, d.data() + pos, n); pos += n; }
    std::string readLine() {                                              // stream.cpp:392-414
        std::string retval; bool nl = false;
        do {
            if (pos >= d.size()) { if (retval.size() != 0) return retval; throw std::runtime_error("Read less data than expected"); }
            const char data = (char) d[pos++];
            if (data != 13 && data != 10) retval += data; else if (data == 10) nl = true;
        } while (!nl);
        return retval;
    }
};
struct Vector2i { int x = 0, y = 0; };
namespace detail {
#include "ref_rgbe_helpers.inc"
}
struct Bitmap {
    enum EPixelFormat { ERGB = 3 }; enum EComponentFormat { EFloat32 = 7 };
    Vector2i m_size; int m_pixelFormat = 0, m_componentFormat = 0, m_channelCount = 0; float m_gamma = 0; uint8_t *m_data = nullptr; bool m_ownsData = false;
    size_t getBufferSize() const { return (size_t) m_size.x * m_size.y * 3 * sizeof(float); }
    void readRGBE(Stream *stream);
};
inline void *allocAligned(size_t n) { return std::calloc(1, n); }
#include "ref_rgbe_read.inc"
}
extern "C" int ref_load_rgbe(const char *path, float *out, int *w, int *h, char *err) {
    try {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("could not be found");
        refrgbe::Stream s; s.d.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        refrgbe::Bitmap b; b.readRGBE(&s);
        *w = b.m_size.x; *h = b.m_size.y;
        if (out) std::memcpy(out, b.m_data, b.getBufferSize());
        std::free(b.m_data);
        return 0;
    } catch (const std::exception &e) { if (err) std::snprintf(err, 256, "%s", e.what()); return -1; }
}
