// oracle/ref_shim/ref_sobol.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's `sobol` sampler plugin: src/samplers/sobol.cpp and src/samplers/sobolseq.cpp (the direction numbers) are
// compiled UNMODIFIED from /root/reference against fake_sampler/sampler_shim.h (oracle/Makefile -> oracle/_ref/libref_sobol.so).  The sampler is
// driven the way SamplingIntegrator::render / renderBlock drive it (src/librender/integrator.cpp:40-41,167-185): setFilmResolution(crop size, true),
// generate(pixel), next2D / next1D ..., advance().
#include "sampler_shim.h"
using namespace mitsuba;
extern "C" {
void *ref_create_SobolSampler(const Properties *);
void *ref_sobol_create(unsigned sampleCount, unsigned long long scramble, int filmW, int filmH) {
    try {
        Properties props; props.sizes["sampleCount"] = sampleCount; props.sizes["scramble"] = (size_t) scramble;
        Sampler *s = (Sampler *) ref_create_SobolSampler(&props);
        s->setFilmResolution(Vector2i(filmW, filmH), true);
        return s;
    } catch (...) { return nullptr; }
}
// pattern: 1 = next1D, 2 = next2D per request; the same requests are repeated for nSamples consecutive sample indices starting at firstSample
int ref_sobol_sequence(void *h, int px, int py, unsigned firstSample, unsigned nSamples, int nReq, const int *pattern, float *out) {
    try {
        Sampler *s = (Sampler *) h;
        s->generate(Point2i(px, py));
        for (unsigned k = 0; k < firstSample; ++k) s->advance();
        for (unsigned j = 0; j < nSamples; ++j) {
            for (int r = 0; r < nReq; ++r) { if (pattern[r] == 1) *out++ = s->next1D(); else { const Point2 p = s->next2D(); *out++ = p.x; *out++ = p.y; } }
            s->advance();
        }
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "ref_sobol_sequence: %s\n", e.what()); return -1; }
}
}
