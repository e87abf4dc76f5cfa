// cp_cli.cpp -- `cudapath_render`: renders a scene file of the reference's XML format through libcudapath.so.
//
// Mirrors the part of the reference's command line that concerns this path (src/mitsuba/mitsuba.cpp:154-260):
//   cudapath_render [-o out.{png,exr,pfm,ppm}] [-D name=value]... [-p N] [--gpus N] [-r sec] [-q] [--gpu i] [--spp n] [--seed s] scene.xml
//     -o   output file (mitsuba.cpp:190); default: the scene's name with the extension the film asks for (.png for ldrfilm, .exr for hdrfilm)
//     -r   `mitsuba -r sec` (mitsuba.cpp:228-229): write the (partial) output image every `sec` seconds while rendering.  The job is then rendered
//          as a sequence of sample-index ranges whose films add up (the random numbers are keyed by the sample index, so the final image is the
//          one a single call produces, up to the order of fp32 additions); a partial image is the film developed so far -- fewer samples per
//          pixel, correctly normalised
//     --faithful-sampler   draw the random numbers of the file's <sampler type="sobol"> (src/samplers/sobol.cpp) instead of the Philox counters
//     -D   parameter substitution of $name in the file (mitsuba.cpp:168)
//     -p   `mitsuba -p N` asks for N local workers (mitsuba.cpp:218-222,280-282); the workers of this path are GPUs: the job is split
//          over min(N, visible GPUs) devices (cudapath_create_multi: sample-range sharding, one ncclReduce of the film)
//     --gpus N   the same, but N GPUs must exist; --gpu i selects the first device
//     -q   quiet
// Ctrl-C cancels the render (cudapath_cancel) and exits with 130 without writing an image; a progress line goes to a terminal's stderr.
// and prints the "Render time" line of RenderJob::run (src/librender/renderjob.cpp:108) plus Mpaths/s and Mrays/s.
// Uses nothing but the C ABI of include/cudapath.h (this file is also the smallest example of a host program on that boundary).
// Image writers: PNG (8 bit, zlib "stored/deflate" through libz), PPM, PFM (linear float, bottom-up as the format wants it), OpenEXR (cudapath_write_exr: half RGB).
#include "../../include/cudapath.h"
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>
#include <unistd.h>

static void put_be32(std::vector<unsigned char> &v, uint32_t x) { for (int k = 3; k >= 0; --k) v.push_back((unsigned char) (x >> (8 * k))); }
static void png_chunk(FILE *f, const char *type, const std::vector<unsigned char> &data) {
    std::vector<unsigned char> buf; put_be32(buf, (uint32_t) data.size());
    fwrite(buf.data(), 1, 4, f);
    std::vector<unsigned char> body(type, type + 4); body.insert(body.end(), data.begin(), data.end());
    fwrite(body.data(), 1, body.size(), f);
    buf.clear(); put_be32(buf, (uint32_t) crc32(0L, body.data(), (uInt) body.size()));
    fwrite(buf.data(), 1, 4, f);
}
static bool write_png(const char *path, const uint8_t *rgb, int w, int h) {
    FILE *f = fopen(path, "wb"); if (!f) return false;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr; put_be32(ihdr, (uint32_t) w); put_be32(ihdr, (uint32_t) h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);     // 8 bit, RGB
    png_chunk(f, "IHDR", ihdr);
    std::vector<unsigned char> raw((size_t) h * (3 * (size_t) w + 1));
    for (int y = 0; y < h; ++y) { raw[(size_t) y * (3 * w + 1)] = 0; memcpy(&raw[(size_t) y * (3 * w + 1) + 1], rgb + (size_t) y * 3 * w, (size_t) 3 * w); }
    uLongf clen = compressBound((uLong) raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong) raw.size(), 6) != Z_OK) { fclose(f); return false; }
    comp.resize(clen);
    png_chunk(f, "IDAT", comp);
    png_chunk(f, "IEND", std::vector<unsigned char>());
    return fclose(f) == 0;
}
static bool write_ppm(const char *path, const uint8_t *rgb, int w, int h) {
    FILE *f = fopen(path, "wb"); if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    fwrite(rgb, 1, (size_t) 3 * w * h, f);
    return fclose(f) == 0;
}
static bool write_pfm(const char *path, const float *rgb, int w, int h) {
    FILE *f = fopen(path, "wb"); if (!f) return false;
    fprintf(f, "PF\n%d %d\n-1.0\n", w, h);                                   // negative scale: little endian
    for (int y = h - 1; y >= 0; --y) fwrite(rgb + (size_t) y * 3 * w, 4, (size_t) 3 * w, f);
    return fclose(f) == 0;
}
static bool ends_with(const std::string &s, const char *suffix) { const size_t n = strlen(suffix); return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0; }

// Ctrl-C / SIGTERM cancel the running job the way RenderJob::cancel does (include/mitsuba/render/renderjob.h:81): cudapath_cancel is the one
// entry point that may run beside a blocking render (it only sets an atomic flag, so it is safe inside a signal handler)
static cudapath_ctx *volatile g_ctx = nullptr;
static void on_signal(int) { if (g_ctx) cudapath_cancel(g_ctx); }
static void on_progress(void *, uint64_t done, uint64_t total) { fprintf(stderr, "\rRendering: %3.0f %%", 100.0 * (double) done / (double) total); if (done == total) fprintf(stderr, "\n"); }
static int die(const char *what) { fprintf(stderr, "cudapath_render: %s: %s\n", what, cudapath_last_error()); return 1; }

int main(int argc, char **argv) {
    std::string out, defines, scene, dataDir;
    int gpu = 0, workers = 0, gpus = 0; long spp = 0; unsigned long long seed = 0; bool quiet = false, faithful = false; double flushSec = -1; long chunk = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char *opt) -> const char * { if (i + 1 >= argc) { fprintf(stderr, "cudapath_render: %s needs an argument\n", opt); exit(2); } return argv[++i]; };
        if (a == "-o") out = need("-o");
        else if (a == "-D") { if (!defines.empty()) defines += ";"; defines += need("-D"); }
        else if (a.rfind("-D", 0) == 0 && a.size() > 2) { if (!defines.empty()) defines += ";"; defines += a.substr(2); }
        else if (a == "-p") workers = atoi(need("-p"));
        else if (a == "--gpus") gpus = atoi(need("--gpus"));
        else if (a == "-b") need(a.c_str());
        else if (a == "-r") flushSec = atof(need("-r"));
        else if (a == "--chunk") chunk = atol(need("--chunk"));
        else if (a == "--faithful-sampler") faithful = true;
        else if (a == "-q") quiet = true;
        else if (a == "--gpu") gpu = atoi(need("--gpu"));
        else if (a == "--spp") spp = atol(need("--spp"));
        else if (a == "--seed") seed = strtoull(need("--seed"), nullptr, 10);
        else if (a == "--data-dir") dataDir = need("--data-dir");
        else if (a == "-h" || a == "--help") { printf("usage: cudapath_render [-o out.{png,exr,pfm,ppm}] [-D name=value]... [-p N] [--gpus N] [-r sec] [--chunk spp] [--faithful-sampler] [-q] [--gpu i] [--spp n] [--seed s] [--data-dir dir] scene.xml\n"); return 0; }
        else if (!a.empty() && a[0] == '-') { fprintf(stderr, "cudapath_render: unknown option %s\n", a.c_str()); return 2; }
        else scene = a;
    }
    if (scene.empty()) { fprintf(stderr, "cudapath_render: no scene file given (-h for help)\n"); return 2; }
    if (dataDir.empty()) dataDir = getenv("CUDAPATH_DATA_DIR") ? getenv("CUDAPATH_DATA_DIR") : "refdata";

    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    cudapath_ctx *ctx = nullptr;
    const int visible = cudapath_visible_devices();
    int nDev = gpus > 0 ? gpus : (workers > 1 ? (workers < visible - gpu ? workers : visible - gpu) : 1);
    if (nDev < 1) nDev = 1;
    std::vector<int> devices;
    for (int k = 0; k < nDev; ++k) devices.push_back(gpu + k);
    if (cudapath_create_multi(devices.data(), nDev, &ctx) != 0) return die("cannot create a context");
    if (cudapath_set_data_dir(ctx, dataDir.c_str()) != 0) return die("data directory");
    if (faithful && cudapath_set_sampler(ctx, 2, 0) != 0) return die("sampler");     // the <sampler type="sobol"> of the file, number for number
    const double t0 = now();
    uint32_t fileSpp = 0;
    if (cudapath_load_scene_xml(ctx, scene.c_str(), defines.empty() ? nullptr : defines.c_str(), &fileSpp) != 0) return die("cannot load the scene");
    const double t1 = now();
    {   // build effort from the size of the job (cudapath_set_job_size_hint)
        int fw = 0, fh = 0; cudapath_film_size(ctx, &fw, &fh);
        const uint64_t jobSpp = spp > 0 ? (uint64_t) spp : fileSpp;
        cudapath_set_job_size_hint(ctx, (uint64_t) fw * fh * jobSpp / (uint64_t) nDev);
    }
    if (cudapath_build(ctx) != 0) return die("cannot build the scene");
    const double t2 = now();
    const uint32_t n = spp > 0 ? (uint32_t) spp : fileSpp;
    int w = 0, h = 0, hdr = 0; float gamma = -1, exposure = 0;
    cudapath_film_size(ctx, &w, &h);
    cudapath_get_film_output(ctx, &hdr, &gamma, &exposure);
    std::vector<float> film((size_t) w * h * 5);
    g_ctx = ctx; signal(SIGINT, on_signal); signal(SIGTERM, on_signal);
    if (!quiet && isatty(2)) cudapath_set_progress_callback(ctx, on_progress, nullptr);     // the progress bar of the reference's console (ProgressReporter)
    if (out.empty()) { out = scene; const size_t dot = out.rfind('.'); if (dot != std::string::npos) out.resize(dot); out += hdr ? ".exr" : ".png"; }
    auto writeImage = [&](const std::vector<float> &f) -> int {
        bool ok;
        if (ends_with(out, ".pfm") || ends_with(out, ".exr")) {
            std::vector<float> rgb((size_t) w * h * 3);
            cudapath_develop(f.data(), w, h, rgb.data());
            ok = ends_with(out, ".pfm") ? write_pfm(out.c_str(), rgb.data(), w, h) : cudapath_write_exr(out.c_str(), rgb.data(), w, h, 1) == 0;
        } else {
            std::vector<uint8_t> rgb8((size_t) w * h * 3);
            if (cudapath_develop_ldr(f.data(), w, h, gamma, exposure, rgb8.data()) != 0) return die("develop");
            ok = ends_with(out, ".ppm") ? write_ppm(out.c_str(), rgb8.data(), w, h) : write_png(out.c_str(), rgb8.data(), w, h);
        }
        if (!ok) { fprintf(stderr, "cudapath_render: cannot write %s\n", out.c_str()); return 1; }
        return 0;
    };
    int rc = 0, flushes = 0;
    if (flushSec < 0 && chunk <= 0) rc = cudapath_render(ctx, n, seed, 0, n, film.data());
    else {
        // progressive: ranges of sample indices, the film of each added to the running sum; flush when `-r` seconds have passed
        const uint32_t step = chunk > 0 ? (uint32_t) chunk : (n >= 16 ? n / 16 : 1);
        std::vector<float> part(film.size());
        double lastFlush = now();
        for (uint32_t s0 = 0; s0 < n && rc == 0; s0 += step) {
            const uint32_t s1 = s0 + step < n ? s0 + step : n;
            rc = cudapath_render(ctx, n, seed, s0, s1, part.data());
            if (rc != 0) break;
            for (size_t i = 0; i < film.size(); ++i) film[i] += part[i];
            if (flushSec >= 0 && s1 < n && now() - lastFlush >= flushSec) {
                if (writeImage(film) != 0) return 1;
                ++flushes; lastFlush = now();
                if (!quiet) printf("Flushed a partial image (%u of %u samples per pixel) to \"%s\"\n", s1, n, out.c_str());
            }
        }
    }
    g_ctx = nullptr; signal(SIGINT, SIG_DFL); signal(SIGTERM, SIG_DFL);
    if (rc != 0) {
        if (std::string(cudapath_last_error()) == "render cancelled") { fprintf(stderr, "\ncudapath_render: render cancelled, no image written\n"); cudapath_destroy(ctx); return 130; }
        return die("render failed");
    }
    const double t3 = now();
    cudapath_stats st; cudapath_get_stats(ctx, &st);
    if (writeImage(film) != 0) return 1;
    if (!quiet) {
        const double paths = (double) w * h * n, rays = (double) st.rays + (double) st.shadow_rays;
        printf("Loaded \"%s\" in %.3f s; %llu segments, %llu triangles -> %llu BVH references, %llu nodes (built in %.3f s)\n", scene.c_str(), t1 - t0,
               (unsigned long long) st.segments, (unsigned long long) st.triangles, (unsigned long long) st.bvh_references, (unsigned long long) st.bvh_nodes, t2 - t1);
        printf("Render time: %.4f s  (%dx%d, %u spp on %d GPU%s: %.1f Mpaths/s, %.1f Mrays/s; device %.4f s, film reduce %.3f ms)\n", t3 - t2, w, h, n, nDev, nDev > 1 ? "s" : "",
               paths / (t3 - t2) / 1e6, rays / (t3 - t2) / 1e6, st.render_ms * 1e-3, cudapath_last_reduce_ms(ctx));
        printf("Writing image to \"%s\" ..\n", out.c_str());
    }
    cudapath_destroy(ctx);
    return 0;
}
