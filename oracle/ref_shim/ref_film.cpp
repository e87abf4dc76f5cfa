// oracle/ref_shim/ref_film.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Film splatting of the reference executed as written: ImageBlock::put (include/mitsuba/render/imageblock.h:124-186, both overloads) is cut
// out of the reference at build time (oracle/_ref/ref_imageblock_put.inc) and pasted into a class that supplies the members it touches;
// the reconstruction filters are the reference's own files compiled unmodified: src/libcore/rfilter.cpp (configure(): the 31-tap
// discretisation and its normalisation) and src/rfilters/{tent,box,gaussian}.cpp.  Output: part of oracle/_ref/libref_geom.so.
#include <mitsuba/core/rfilter.h>
#define SPECTRUM_SAMPLES 3
namespace mitsuba {
struct Point2i { int x, y; Point2i() : x(0), y(0) {} Point2i(int x, int y) : x(x), y(y) {} };
struct Vector2iF { int x, y; };
struct FilmBitmap { int w = 0, h = 0; std::vector<Float> data; Vector2iF size;
    int getChannelCount() const { return SPECTRUM_SAMPLES + 2; } const Vector2iF &getSize() const { return size; } Float *getFloatData() { return data.data(); } };
#define Vector2i Vector2iF
#define FINLINE inline
class ImageBlock {
public:
    FilmBitmap *m_bitmap = nullptr; const ReconstructionFilter *m_filter = nullptr; Point2i m_offset; int m_borderSize = 0; bool m_warn = true;
    Float m_weightsX[64], m_weightsY[64];
#include "ref_imageblock_put.inc"
};
#undef Vector2i
}
using namespace mitsuba;
extern "C" {
void *ref_create_TentFilter(const Properties *); void *ref_create_BoxFilter(const Properties *); void *ref_create_GaussianFilter(const Properties *);
// type: 0 tent, 1 box, 2 gaussian (stddev = param, 0 = the plugin default); out32 = m_values, outInfo = radius, scale factor, border size
void *ref_filter_create(int type, float param, float *out32, float *outInfo) {
    Properties props;
    if (type == 2 && param > 0) props.setFloat("stddev", param);
    ReconstructionFilter *f = (ReconstructionFilter *) (type == 0 ? ref_create_TentFilter(&props) : type == 1 ? ref_create_BoxFilter(&props) : ref_create_GaussianFilter(&props));
    f->configure();
    for (int i = 0; i <= MTS_FILTER_RESOLUTION; ++i) out32[i] = f->evalDiscretized((i + 0.5f) * f->getRadius() / MTS_FILTER_RESOLUTION * (i == MTS_FILTER_RESOLUTION ? 2.0f : 1.0f));
    outInfo[0] = f->getRadius(); outInfo[1] = (float) f->getBorderSize();
    return f;
}
// ImageBlock::put(pos, spec, alpha) for n samples into a w x h block with offset 0 and no border (the film itself): film = w*h*5 floats
void ref_film_put(void *filter, int w, int h, int n, const float *pos, const float *rgb, const float *alpha, float *film, int *outOk) {
    FilmBitmap bmp; bmp.w = w; bmp.h = h; bmp.size.x = w; bmp.size.y = h; bmp.data.assign((size_t) w * h * 5, 0.0f);
    ImageBlock block; block.m_bitmap = &bmp; block.m_filter = (const ReconstructionFilter *) filter; block.m_warn = true;
    for (int i = 0; i < n; ++i) {
        Spectrum s; s[0] = rgb[3 * i]; s[1] = rgb[3 * i + 1]; s[2] = rgb[3 * i + 2];
        bool ok = false;
        try { ok = block.put(Point2(pos[2 * i], pos[2 * i + 1]), s, alpha[i]); } catch (...) { ok = false; }
        outOk[i] = ok ? 1 : 0;
    }
    std::memcpy(film, bmp.data.data(), bmp.data.size() * sizeof(float));
}
}
