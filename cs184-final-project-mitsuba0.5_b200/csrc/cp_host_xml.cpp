// cp_host_xml.cpp -- loads the reference's scene XML files (unchanged) and flattens them through the C ABI.
//
// Replaces, for the tags the hair scenes use (reference file:line):
//   SceneHandler tag table, $-substitution, <ref>, <default>, version check   src/librender/scenehandler.cpp:70-106, 228-250, 300-700
//   Properties handed to plugin constructors                                  src/libcore/properties.cpp
//   mitsuba -D name=value                                                     src/mitsuba/mitsuba.cpp:168
//   Transform::lookAt / rotate / translate / scale                            src/libcore/transform.cpp
//   PerspectiveCamera fov handling                                            src/librender/sensor.cpp:225-300
// Plugins understood: integrator `path`; sensor `perspective` (film `ldrfilm`/`hdrfilm`, rfilter `tent`/`box`/`gaussian`,
// any sampler: only sampleCount is used); bsdf `kajiyakay`, `marschner`, `marschner_fixed`, `marschnerdielectric`, `thindielectric`, `roughplastic`, `plastic`, `diffuse`,
// `twosided` (around diffuse / plastic / roughplastic); texture `checkerboard` (as the reflectance of diffuse / plastic); shape `hair`, `obj`, `rectangle`;
// emitter `sunsky`, `envmap` (Radiance .hdr file).  models/teapot/scene.xml loads unchanged.
// Anything else raises an error naming the plugin (the reference would dlopen plugins/<type>.so, src/libcore/plugin.cpp:222-245).
#include "../../include/cudapath.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct Node {
    std::string tag;
    std::map<std::string, std::string> attr;
    std::vector<std::unique_ptr<Node>> children;
    const std::string &get(const std::string &k) const { static const std::string empty; auto it = attr.find(k); return it == attr.end() ? empty : it->second; }
    bool has(const std::string &k) const { return attr.count(k) != 0; }
};

struct Parser {
    const std::string &s; size_t p = 0;
    explicit Parser(const std::string &src) : s(src) {}
    [[noreturn]] void err(const std::string &m) { throw std::runtime_error("XML parse error near byte " + std::to_string(p) + ": " + m); }
    void skipWs() { while (p < s.size() && isspace((unsigned char) s[p])) ++p; }
    bool starts(const char *t) const { return s.compare(p, strlen(t), t) == 0; }
    void skipMisc() {
        for (;;) {
            skipWs();
            if (starts("<?")) { size_t e = s.find("?>", p); if (e == std::string::npos) err("unterminated declaration"); p = e + 2; }
            else if (starts("<!--")) { size_t e = s.find("-->", p); if (e == std::string::npos) err("unterminated comment"); p = e + 3; }
            else if (starts("<!")) { size_t e = s.find('>', p); if (e == std::string::npos) err("unterminated directive"); p = e + 1; }
            else break;
        }
    }
    std::string name() { size_t b = p; while (p < s.size() && (isalnum((unsigned char) s[p]) || s[p] == '_' || s[p] == '-' || s[p] == ':' || s[p] == '.')) ++p; if (b == p) err("expected a name"); return s.substr(b, p - b); }
    static std::string decode(const std::string &v) {
        std::string o; o.reserve(v.size());
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i] != '&') { o += v[i]; continue; }
            if (!v.compare(i, 4, "&lt;")) { o += '<'; i += 3; } else if (!v.compare(i, 4, "&gt;")) { o += '>'; i += 3; }
            else if (!v.compare(i, 5, "&amp;")) { o += '&'; i += 4; } else if (!v.compare(i, 6, "&quot;")) { o += '"'; i += 5; }
            else if (!v.compare(i, 6, "&apos;")) { o += '\''; i += 5; } else o += v[i];
        }
        return o;
    }
    std::unique_ptr<Node> element() {
        if (s[p] != '<') err("expected '<'");
        ++p;
        std::unique_ptr<Node> n(new Node());
        n->tag = name();
        for (;;) {
            skipWs();
            if (p >= s.size()) err("unexpected end of file");
            if (s[p] == '/') { if (!starts("/>")) err("expected '/>'"); p += 2; return n; }
            if (s[p] == '>') { ++p; break; }
            std::string k = name();
            skipWs(); if (s[p] != '=') err("expected '='"); ++p; skipWs();
            char q = s[p]; if (q != '"' && q != '\'') err("expected a quoted attribute value"); ++p;
            size_t e = s.find(q, p); if (e == std::string::npos) err("unterminated attribute value");
            n->attr[k] = decode(s.substr(p, e - p)); p = e + 1;
        }
        for (;;) {
            // text content is ignored
            size_t lt = s.find('<', p); if (lt == std::string::npos) err("unexpected end of file inside <" + n->tag + ">");
            p = lt;
            if (starts("<!--")) { size_t e = s.find("-->", p); if (e == std::string::npos) err("unterminated comment"); p = e + 3; continue; }
            if (starts("</")) { p += 2; std::string c = name(); if (c != n->tag) err("mismatched closing tag </" + c + ">"); skipWs(); if (s[p] != '>') err("expected '>'"); ++p; return n; }
            n->children.push_back(element());
        }
    }
    std::unique_ptr<Node> document() { skipMisc(); auto n = element(); return n; }
};

struct Mat4 { double m[16]; };
Mat4 ident() { Mat4 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0); return r; }
Mat4 mul(const Mat4 &a, const Mat4 &b) { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += a.m[i * 4 + k] * b.m[k * 4 + j]; r.m[i * 4 + j] = s; } return r; }

std::vector<double> numbers(const std::string &v) {
    std::vector<double> out; std::string t = v;
    for (char &c : t) if (c == ',') c = ' ';
    std::istringstream is(t); std::string tok;
    while (is >> tok) { char *e = nullptr; double d = strtod(tok.c_str(), &e); if (*e) throw std::runtime_error("could not parse number \"" + tok + "\""); out.push_back(d); }
    return out;
}
double num(const std::string &v) { auto n = numbers(v); if (n.size() != 1) throw std::runtime_error("expected a single number, got \"" + v + "\""); return n[0]; }
bool boolean(const std::string &v) { if (v == "true") return true; if (v == "false") return false; throw std::runtime_error("could not parse boolean \"" + v + "\""); }
std::string lower(std::string s) { for (char &c : s) c = (char) tolower((unsigned char) c); return s; }

struct Loader {
    cudapath_ctx *ctx; std::string baseDir; std::map<std::string, std::string> defines;
    // dry run (cudapath_validate_scene_xml): no context, no GPU -- every plugin and parameter is still parsed and checked, the calls
    // into the C ABI are replaced by entries of `report`
    bool dry = false; int fakeIds = 0; std::vector<std::string> report;
    int note(const std::string &what) { report.push_back(what); return fakeIds++; }
    std::map<std::string, int> bsdfIds;
    uint32_t spp = 4;
    bool haveIntegrator = false, haveSensor = false;

    void substitute(Node &n) {                                       // scenehandler.cpp: '$name' replacement in every attribute
        for (auto &kv : n.attr) {
            std::string &v = kv.second; size_t pos = 0;
            while ((pos = v.find('$', pos)) != std::string::npos) {
                bool done = false;
                // longest matching parameter name wins
                size_t bestLen = 0; std::string bestVal;
                for (auto &d : defines) if (v.compare(pos + 1, d.first.size(), d.first) == 0 && d.first.size() > bestLen) { bestLen = d.first.size(); bestVal = d.second; done = true; }
                if (!done) throw std::runtime_error("The variable \"" + v.substr(pos) + "\" is undefined (use -D name=value)");
                v.replace(pos, bestLen + 1, bestVal); pos += bestVal.size();
            }
        }
        for (auto &c : n.children) {
            if (c->tag == "default") { if (!defines.count(c->get("name"))) defines[c->get("name")] = c->get("value"); continue; }
            substitute(*c);
        }
    }
    const Node *child(const Node &n, const std::string &tag, const std::string &name) const {
        for (auto &c : n.children) if (c->tag == tag && c->get("name") == name) return c.get();
        return nullptr;
    }
    double getFloat(const Node &n, const std::string &name, double def) const { auto c = child(n, "float", name); return c ? num(c->get("value")) : def; }
    long getInt(const Node &n, const std::string &name, long def) const { auto c = child(n, "integer", name); return c ? (long) num(c->get("value")) : def; }
    bool getBool(const Node &n, const std::string &name, bool def) const { auto c = child(n, "boolean", name); return c ? boolean(c->get("value")) : def; }
    std::string getString(const Node &n, const std::string &name, const std::string &def) const { auto c = child(n, "string", name); return c ? c->get("value") : def; }
    void getColor(const Node &n, const std::string &name, float def, float out[3]) const {
        out[0] = out[1] = out[2] = def;
        for (const char *tag : {"rgb", "spectrum", "srgb"}) {
            auto c = child(n, tag, name);
            if (!c) continue;
            if (std::string(tag) == "srgb") throw std::runtime_error("<srgb> colour values are not supported");
            if (c->get("value").find(':') != std::string::npos) throw std::runtime_error("wavelength:value spectra are not supported in RGB mode");
            auto v = numbers(c->get("value"));
            if (v.size() == 1) out[0] = out[1] = out[2] = (float) v[0];
            else if (v.size() == 3) { out[0] = (float) v[0]; out[1] = (float) v[1]; out[2] = (float) v[2]; }
            else throw std::runtime_error("could not parse colour \"" + c->get("value") + "\"");
            return;
        }
    }
    Mat4 getTransform(const Node &n, const std::string &name) const {
        Mat4 t = ident();
        auto c = child(n, "transform", name);
        if (!c) return t;
        for (auto &op : c->children) {
            Mat4 m = ident();
            auto xyz = [&](double dflt, double v[3]) {
                if (op->has("value")) { auto q = numbers(op->get("value")); if (q.size() == 1) v[0] = v[1] = v[2] = q[0]; else if (q.size() == 3) { v[0] = q[0]; v[1] = q[1]; v[2] = q[2]; } else throw std::runtime_error("bad vector value"); return; }
                v[0] = op->has("x") ? num(op->get("x")) : dflt; v[1] = op->has("y") ? num(op->get("y")) : dflt; v[2] = op->has("z") ? num(op->get("z")) : dflt;
            };
            if (op->tag == "matrix") { auto q = numbers(op->get("value")); if (q.size() != 16) throw std::runtime_error("<matrix> needs 16 values"); for (int i = 0; i < 16; ++i) m.m[i] = (float) q[i]; }
            else if (op->tag == "translate") { double v[3]; xyz(0, v); m.m[3] = v[0]; m.m[7] = v[1]; m.m[11] = v[2]; }
            else if (op->tag == "scale") { double v[3]; xyz(1, v); m.m[0] = v[0]; m.m[5] = v[1]; m.m[10] = v[2]; }
            else if (op->tag == "rotate") {
                double a[3]; xyz(0, a); const double ang = num(op->get("angle")) * M_PI / 180.0;
                const double len = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); if (len == 0) throw std::runtime_error("<rotate> needs an axis");
                const double x = a[0] / len, y = a[1] / len, z = a[2] / len, s = std::sin(ang), c2 = std::cos(ang);
                m.m[0] = x * x + (1 - x * x) * c2; m.m[1] = x * y * (1 - c2) - z * s; m.m[2] = x * z * (1 - c2) + y * s;
                m.m[4] = x * y * (1 - c2) + z * s; m.m[5] = y * y + (1 - y * y) * c2; m.m[6] = y * z * (1 - c2) - x * s;
                m.m[8] = x * z * (1 - c2) - y * s; m.m[9] = y * z * (1 - c2) + x * s; m.m[10] = z * z + (1 - z * z) * c2;
            } else if (op->tag == "lookat" || op->tag == "lookAt") {
                // scenehandler.cpp:362-398 + Transform::lookAt (transform.cpp:191-214): a missing or zero `up` picks an arbitrary axis with
                // coordinateSystem(normalize(t - o)) (util.cpp:592-601); coinciding points and a parallel `up` are errors
                auto o = numbers(op->get("origin")), tg = numbers(op->get("target")); std::vector<double> up = op->has("up") ? numbers(op->get("up")) : std::vector<double>{0, 0, 0};
                if (o.size() != 3) throw std::runtime_error("<lookat>: invalid 'origin' argument");
                if (tg.size() != 3) throw std::runtime_error("<lookat>: invalid 'target' argument");
                if (up.size() != 3 && !up.empty()) throw std::runtime_error("<lookat>: invalid 'up' argument");
                if (up.empty()) up = {0, 0, 0};
                double d[3] = {tg[0] - o[0], tg[1] - o[1], tg[2] - o[2]}; const double dl = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                if (dl == 0) throw std::runtime_error("lookAt(): 'origin' and 'target' coincide!");
                for (double &q : d) q /= dl;
                if (up[0] * up[0] + up[1] * up[1] + up[2] * up[2] == 0) {
                    double c[3];
                    if (std::abs(d[0]) > std::abs(d[1])) { const double il = 1.0 / std::sqrt(d[0] * d[0] + d[2] * d[2]); c[0] = d[2] * il; c[1] = 0; c[2] = -d[0] * il; }
                    else { const double il = 1.0 / std::sqrt(d[1] * d[1] + d[2] * d[2]); c[0] = 0; c[1] = d[2] * il; c[2] = -d[1] * il; }
                    up = {c[1] * d[2] - c[2] * d[1], c[2] * d[0] - c[0] * d[2], c[0] * d[1] - c[1] * d[0]};      // b = cross(c, a)
                }
                double l[3] = {up[1] * d[2] - up[2] * d[1], up[2] * d[0] - up[0] * d[2], up[0] * d[1] - up[1] * d[0]}; const double ll = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
                if (ll == 0) throw std::runtime_error("lookAt(): the forward and upward direction must be linearly independent!");
                for (double &q : l) q /= ll;
                double u[3] = {d[1] * l[2] - d[2] * l[1], d[2] * l[0] - d[0] * l[2], d[0] * l[1] - d[1] * l[0]};
                for (int r = 0; r < 3; ++r) { m.m[r * 4] = l[r]; m.m[r * 4 + 1] = u[r]; m.m[r * 4 + 2] = d[r]; m.m[r * 4 + 3] = o[r]; }
            } else throw std::runtime_error("unsupported transform element <" + op->tag + ">");
            t = mul(m, t);
        }
        return t;
    }
    static void toFloat(const Mat4 &m, float out[16]) { for (int i = 0; i < 16; ++i) out[i] = (float) m.m[i]; }
    void check(int rc) const { if (rc < 0) throw std::runtime_error(cudapath_last_error()); }

    // <texture type="checkerboard"> (src/textures/checkerboard.cpp:49-52; Texture2D parameters src/librender/texture.cpp:81-95)
    void setTexture(int bsdf, const Node &t) {
        if (t.get("type") != "checkerboard") throw std::runtime_error("texture plugin \"" + t.get("type") + "\" is not supported (checkerboard)");
        if (getString(t, "coordinates", "uv") != "uv") throw std::runtime_error("Only UV coordinates are supported at the moment!");
        float c0[3], c1[3]; getColor(t, "color0", 0.4f, c0); getColor(t, "color1", 0.2f, c1);
        const double uvscale = getFloat(t, "uvscale", 1.0);
        if (dry) { note("texture checkerboard"); return; }
        check(cudapath_bsdf_set_checkerboard(ctx, bsdf, c0, c1, (float) getFloat(t, "uoffset", 0.0), (float) getFloat(t, "voffset", 0.0),
                                             (float) getFloat(t, "uscale", uvscale), (float) getFloat(t, "vscale", uvscale)));
    }
    int loadBsdf(const Node &n) {
        const std::string type = n.get("type");
        int id;
        if (type == "kajiyakay") {
            float d[3], s[3]; getColor(n, "diffuseReflectance", 0.5f, d); getColor(n, "specularReflectance", 0.2f, s);
            id = (dry ? note("bsdf kajiyakay") : cudapath_add_bsdf_kajiyakay(ctx, d, s, (float) getFloat(n, "exponent", 30.0)));
        } else if (type == "marschner" || type == "marschner_diffuse") {
            // `marschner_diffuse` (models/*/scene_marschner_diffuse.xml, hair_curl_diffuse.xml) is the same class: the fork builds
            // marschner_diffuse.cpp AS plugins/marschner.so (src/bsdfs/SConscript:30-31), no plugin of the longer name exists
            auto ior = [&](const char *name, double def) {           // ior.h:95-100: a float, or a named material (only the defaults are known here)
                if (child(n, "float", name)) return getFloat(n, name, def);
                auto c = child(n, "string", name);
                if (!c) return def;
                std::string v = lower(c->get("value"));
                if (v == "air") return 1.000277; if (v == "bk7") return 1.5046; if (v == "vacuum") return 1.0; if (v == "water") return 1.3330;
                throw std::runtime_error("Unable to find an IOR value for \"" + v + "\"");
            };
            float d[3], s[3]; getColor(n, "diffuseReflectance", 0.5f, d); getColor(n, "specularReflectance", 0.5f, s);
            std::string distr = lower(getString(n, "distribution", "beckmann"));
            int di = distr == "beckmann" ? 0 : distr == "ggx" ? 1 : (distr == "phong" || distr == "as") ? 2 : -1;
            if (di < 0) throw std::runtime_error("Specified an invalid distribution \"" + distr + "\", must be \"beckmann\", \"ggx\", or \"phong\"/\"as\"!");
            if (child(n, "float", "alphaU") || child(n, "float", "alphaV")) throw std::runtime_error("The 'marschner' plugin does not support anisotropic microfacet distributions!");
            id = (dry ? note("bsdf marschner") : cudapath_add_bsdf_marschner(ctx, (float) ior("intIOR", 1.5046), (float) ior("extIOR", 1.000277), d, s, (float) getFloat(n, "alpha", 0.1),
                                             di, getBool(n, "nonlinear", false) ? 1 : 0));
        } else if (type == "roughplastic") {
            auto ior = [&](const char *name, double def) {
                if (child(n, "float", name)) return getFloat(n, name, def);
                auto c = child(n, "string", name);
                if (!c) return def;
                std::string v = lower(c->get("value"));
                if (v == "polypropylene") return 1.49; if (v == "amber") return 1.55; if (v == "air") return 1.000277; if (v == "bk7") return 1.5046;
                if (v == "vacuum") return 1.0; if (v == "water") return 1.3330;
                throw std::runtime_error("Unable to find an IOR value for \"" + v + "\"");
            };
            float d[3], s[3]; getColor(n, "diffuseReflectance", 0.5f, d); getColor(n, "specularReflectance", 1.0f, s);
            std::string distr = lower(getString(n, "distribution", "beckmann"));
            int di = distr == "beckmann" ? 0 : distr == "ggx" ? 1 : (distr == "phong" || distr == "as") ? 2 : -1;
            if (di < 0) throw std::runtime_error("Specified an invalid distribution \"" + distr + "\", must be \"beckmann\", \"ggx\", or \"phong\"/\"as\"!");
            if (child(n, "float", "alphaU") || child(n, "float", "alphaV")) throw std::runtime_error("The 'roughplastic' plugin currently does not support anisotropic microfacet distributions!");
            if (child(n, "texture", "alpha") || child(n, "texture", "diffuseReflectance") || child(n, "texture", "specularReflectance")) throw std::runtime_error("roughplastic: textured parameters are not supported");
            id = (dry ? note("bsdf roughplastic") : cudapath_add_bsdf_roughplastic(ctx, (float) ior("intIOR", 1.49), (float) ior("extIOR", 1.000277), d, s, (float) getFloat(n, "alpha", 0.1), di,
                                                getBool(n, "sampleVisible", true) ? 1 : 0, getBool(n, "nonlinear", false) ? 1 : 0));
        } else if (type == "marschner_fixed" || type == "marschner_full") {      // the class of src/bsdfs/marschner.cpp, which the fork's build leaves out
            auto ior = [&](const char *name, double def) {
                if (child(n, "float", name)) return getFloat(n, name, def);
                auto c = child(n, "string", name);
                if (!c) return def;
                std::string v = lower(c->get("value"));
                if (v == "amber") return 1.55; if (v == "air") return 1.000277; if (v == "bk7") return 1.5046; if (v == "vacuum") return 1.0; if (v == "water") return 1.3330;
                throw std::runtime_error("Unable to find an IOR value for \"" + v + "\"");
            };
            if (type == "marschner_full") {          // all three lobes, constants from the scene: sigmaA, betaR (or roughness), scaleAngle (degrees) / scaleAngleRad
                float sa[3]; getColor(n, "sigmaA", 0.22f, sa);
                const double beta = getFloat(n, "betaR", getFloat(n, "roughness", 0.1));
                const double ang = child(n, "float", "scaleAngle") ? getFloat(n, "scaleAngle", 0.0) * M_PI / 180.0 : getFloat(n, "scaleAngleRad", -0.1);
                id = (dry ? note("bsdf marschner_full") : cudapath_add_bsdf_marschner_full(ctx, (float) ior("intIOR", 1.55), (float) ior("extIOR", 1.000277), sa, (float) beta, (float) ang,
                                                                                           (int) getInt(n, "lobes", 7)));
            } else
            id = (dry ? note("bsdf marschner_fixed") : cudapath_add_bsdf_marschner_fixed(ctx, (float) ior("intIOR", 1.55), (float) ior("extIOR", 1.000277)));
        } else if (type == "twosided") {
            // TwoSidedBRDF::addChild / configure (src/bsdfs/twosided.cpp:84-110, 163-177): one nested BRDF serves both sides
            const Node *nested = nullptr;
            for (auto &c : n.children) if (c->tag == "bsdf") {
                if (nested) throw std::runtime_error("twosided: two different nested BRDFs are not supported");
                nested = c.get();
            }
            if (!nested) throw std::runtime_error("twosided: A nested one-sided material is required!");
            const std::string nt = nested->get("type");
            // models/teapot/dielectric.xml nests a `dielectric`: TwoSidedBRDF::configure refuses that in the reference as well (twosided.cpp:106-108)
            if (nt == "dielectric" || nt == "thindielectric" || nt == "roughdielectric" || nt == "marschnerdielectric") throw std::runtime_error("Only materials without a transmission component can be nested!");
            if (nt != "diffuse" && nt != "plastic" && nt != "roughplastic" && nt != "mirror") throw std::runtime_error("twosided: only a nested `diffuse`, `plastic`, `roughplastic` or `mirror` is supported on this path");
            id = loadBsdf(*nested);
            if (!dry) check(cudapath_bsdf_set_twosided(ctx, id)); else note("twosided adapter around the nested bsdf");
        } else if (type == "diffuse") {
            // `diffuse` (diffuse.cpp:70-76: "reflectance" or "diffuseReflectance"), constant or a checkerboard
            float r[3]; getColor(n, "reflectance", 0.5f, r);
            if (child(n, "rgb", "diffuseReflectance") || child(n, "spectrum", "diffuseReflectance")) getColor(n, "diffuseReflectance", 0.5f, r);
            id = (dry ? note("bsdf diffuse") : cudapath_add_bsdf_diffuse(ctx, r, 0));
            check(id);
            const Node *tex = child(n, "texture", "reflectance"); if (!tex) tex = child(n, "texture", "diffuseReflectance");
            if (tex) setTexture(id, *tex);
        } else if (type == "mirror") {                 // the fork's src/bsdfs/mirror.cpp
            if (child(n, "texture", "specularReflectance")) throw std::runtime_error("mirror: a textured specularReflectance is not supported");
            float sp[3]; getColor(n, "specularReflectance", 1.0f, sp);
            id = (dry ? note("bsdf mirror") : cudapath_add_bsdf_mirror(ctx, sp));
        } else if (type == "plastic") {
            auto ior = [&](const char *name, double def) {
                if (child(n, "float", name)) return getFloat(n, name, def);
                auto c = child(n, "string", name);
                if (!c) return def;
                std::string v = lower(c->get("value"));
                if (v == "polypropylene") return 1.49; if (v == "amber") return 1.55; if (v == "air") return 1.000277; if (v == "bk7") return 1.5046;
                if (v == "vacuum") return 1.0; if (v == "water") return 1.3330;
                throw std::runtime_error("Unable to find an IOR value for \"" + v + "\"");
            };
            if (child(n, "texture", "specularReflectance")) throw std::runtime_error("plastic: a textured specularReflectance is not supported");
            float d[3], sp[3]; getColor(n, "diffuseReflectance", 0.5f, d); getColor(n, "specularReflectance", 1.0f, sp);
            id = (dry ? note("bsdf plastic") : cudapath_add_bsdf_plastic(ctx, (float) ior("intIOR", 1.49), (float) ior("extIOR", 1.000277), d, sp, getBool(n, "nonlinear", false) ? 1 : 0));
            check(id);
            if (const Node *tex = child(n, "texture", "diffuseReflectance")) setTexture(id, *tex);
        } else if (type == "thindielectric" || type == "marschnerdielectric") {
            // src/bsdfs/thindielectric.cpp:73-91 / src/bsdfs/marschnerdielectric.cpp:128-167 (defaults bk7 resp. benzene over air)
            auto ior = [&](const char *name, double def) {
                if (child(n, "float", name)) return getFloat(n, name, def);
                auto c = child(n, "string", name);
                if (!c) return def;
                std::string v = lower(c->get("value"));
                if (v == "benzene") return 1.501; if (v == "amber") return 1.55; if (v == "air") return 1.000277; if (v == "bk7") return 1.5046;
                if (v == "vacuum") return 1.0; if (v == "water") return 1.3330;
                throw std::runtime_error("Unable to find an IOR value for \"" + v + "\"");
            };
            if (child(n, "texture", "specularReflectance") || child(n, "texture", "specularTransmittance") || child(n, "texture", "diffuseReflectance"))
                throw std::runtime_error(type + ": textured parameters are not supported");
            float r[3], t[3], d[3];
            if (type == "thindielectric") {
                getColor(n, "specularReflectance", 1.0f, r); getColor(n, "specularTransmittance", 1.0f, t);
                id = (dry ? note("bsdf thindielectric") : cudapath_add_bsdf_thindielectric(ctx, (float) ior("intIOR", 1.5046), (float) ior("extIOR", 1.000277), r, t));
            } else {
                getColor(n, "specularReflectance", 0.1f, r); getColor(n, "specularTransmittance", 0.1f, t); getColor(n, "diffuseReflectance", 0.5f, d);
                id = (dry ? note("bsdf marschnerdielectric") : cudapath_add_bsdf_marschnerdielectric(ctx, (float) ior("intIOR", 1.501), (float) ior("extIOR", 1.000277), d, r, t,
                                                                                                      (float) getFloat(n, "exponent", 30.0)));
            }
        } else throw std::runtime_error("bsdf plugin \"" + type + "\" is outside the hair hot path (supported: kajiyakay, marschner, marschner_fixed, marschner_full, marschnerdielectric, thindielectric, roughplastic, plastic, mirror, diffuse, twosided)");
        check(id);
        if (n.has("id")) bsdfIds[n.get("id")] = id;
        return id;
    }
    void loadShape(const Node &n) {
        const bool isObj = n.get("type") == "obj", isRect = n.get("type") == "rectangle";
        if (n.get("type") != "hair" && !isObj && !isRect) throw std::runtime_error("shape plugin \"" + n.get("type") + "\" is outside the hair hot path (supported: hair, obj, rectangle)");
        int bsdf = -1;
        for (auto &c : n.children) {
            if (c->tag == "bsdf") bsdf = loadBsdf(*c);
            else if (c->tag == "ref") { auto it = bsdfIds.find(c->get("id")); if (it == bsdfIds.end()) throw std::runtime_error("unknown reference id \"" + c->get("id") + "\""); bsdf = it->second; }
        }
        if (bsdf < 0) {          // Shape::configure falls back to a default `diffuse` (src/librender/shape.cpp)
            const float half[3] = {0.5f, 0.5f, 0.5f};
            check(bsdf = (dry ? note("bsdf diffuse (default)") : cudapath_add_bsdf_diffuse(ctx, half, 0)));
        }
        if (isRect) {                                  // src/shapes/rectangle.cpp:81-86
            float tw[16]; toFloat(getTransform(n, "toWorld"), tw);
            check(dry ? note("shape rectangle") : cudapath_add_rectangle(ctx, tw, getBool(n, "flipNormals", false) ? 1 : 0, bsdf));
            return;
        }
        std::string file = getString(n, "filename", "");
        if (file.empty()) throw std::runtime_error(n.get("type") + " shape: missing 'filename'");
        if (file[0] != '/') file = baseDir + "/" + file;
        float tw[16]; toFloat(getTransform(n, "toWorld"), tw);
        if (dry) {
            std::ifstream probe(file, std::ios::binary);
            note(std::string("shape ") + n.get("type") + " \"" + file + "\"" + (probe ? "" : " (file missing)"));
            if (isObj && child(n, "float", "maxSmoothAngle")) throw std::runtime_error("obj: 'maxSmoothAngle' is not supported");
            return;
        }
        if (isObj) {
            if (child(n, "float", "maxSmoothAngle")) throw std::runtime_error("obj: 'maxSmoothAngle' is not supported");
            check(cudapath_add_mesh_file(ctx, file.c_str(), tw, getBool(n, "faceNormals", false) ? 1 : 0, getBool(n, "flipNormals", false) ? 1 : 0, bsdf));
            return;
        }
        check(cudapath_add_hair_file(ctx, file.c_str(), (float) getFloat(n, "radius", 0.025), (float) getFloat(n, "angleThreshold", 1.0),
                                     (float) getFloat(n, "reduction", 0.0), tw, bsdf));
    }
    void loadSensor(const Node &n) {
        if (n.get("type") != "perspective") throw std::runtime_error("sensor plugin \"" + n.get("type") + "\" is outside the hair hot path (supported: perspective)");
        int w = 768, h = 576, filter = 2; float fparam = 0; bool alpha = false;   // film.cpp defaults; hdrfilm/ldrfilm default rfilter is gaussian
        for (auto &c : n.children) {
            if (c->tag == "sampler") {
                spp = (uint32_t) getInt(*c, "sampleCount", 4);
                // sampler-faithful mode (cudapath_set_sampler(ctx, 2, 0) before loading): `sobol` is reproduced, anything else cannot be
                if (!dry && cudapath_get_sampler(ctx) == 2) {
                    if (c->get("type") != "sobol") throw std::runtime_error("sampler plugin \"" + c->get("type") + "\" cannot be reproduced number for number on this path (only `sobol`: "
                                                                             "`independent` hands every worker one sequential stream whose position depends on all paths traced before)");
                    check(cudapath_set_sampler(ctx, 1, (uint64_t) getInt(*c, "scramble", 0)));
                }
            }
            if (c->tag == "film") {
                w = (int) getInt(*c, "width", 768); h = (int) getInt(*c, "height", 576);
                std::string pf = lower(getString(*c, "pixelFormat", c->get("type") == "hdrfilm" ? "rgb" : "rgb"));
                alpha = pf == "rgba" || pf == "luminancealpha" || pf == "spectrumalpha";
                if (c->get("type") != "ldrfilm" && c->get("type") != "hdrfilm") throw std::runtime_error("film plugin \"" + c->get("type") + "\" is not supported (ldrfilm, hdrfilm)");
                if (lower(getString(*c, "tonemapMethod", "gamma")) != "gamma") throw std::runtime_error("ldrfilm: only tonemapMethod=gamma is supported");
                if (!dry) check(cudapath_set_film_output(ctx, c->get("type") == "hdrfilm" ? 1 : 0, (float) getFloat(*c, "gamma", -1.0), (float) getFloat(*c, "exposure", 0.0)));
                if (child(*c, "integer", "cropWidth") || child(*c, "integer", "cropOffsetX")) throw std::runtime_error("film crop windows are not supported");
                for (auto &f : c->children) if (f->tag == "rfilter") {
                    const std::string t = f->get("type");
                    if (t == "tent") { filter = 0; fparam = 0; }          // TentFilter has a fixed radius of 1 (tent.cpp:34); a `radius` property is ignored there as well
                    else if (t == "box") filter = 1;
                    else if (t == "gaussian") { filter = 2; fparam = (float) getFloat(*f, "stddev", 0.5); }
                    else throw std::runtime_error("rfilter plugin \"" + t + "\" is not supported (tent, box, gaussian)");
                }
            }
        }
        const double aspect = (double) w / h;
        double fov = getFloat(n, "fov", -1);
        if (fov < 0) {
            std::string f = getString(n, "focalLength", "50mm");
            if (f.size() > 2 && f.substr(f.size() - 2) == "mm") f = f.substr(0, f.size() - 2);
            const double diag = 2 * 180 / M_PI * std::atan(std::sqrt(36.0 * 36 + 24 * 24) / (2 * num(f)));
            // diagonal fov -> x fov (sensor.cpp setDiagonalFov): tan(x/2) = tan(d/2) * aspect / sqrt(1 + aspect^2)
            fov = 2 * 180 / M_PI * std::atan(std::tan(0.5 * diag * M_PI / 180) * aspect / std::sqrt(1 + aspect * aspect));
        } else {
            std::string axis = lower(getString(n, "fovAxis", "x"));
            if (axis == "smaller") axis = aspect > 1 ? "y" : "x"; else if (axis == "larger") axis = aspect > 1 ? "x" : "y";
            if (axis == "y") fov = 2 * 180 / M_PI * std::atan(std::tan(0.5 * fov * M_PI / 180) * aspect);
            else if (axis == "diagonal") fov = 2 * 180 / M_PI * std::atan(std::tan(0.5 * fov * M_PI / 180) * aspect / std::sqrt(1 + aspect * aspect));
            else if (axis != "x") throw std::runtime_error("The 'fovAxis' parameter must be set to one of 'smaller', 'larger', 'diagonal', 'x', or 'y'!");
        }
        float tw[16]; toFloat(getTransform(n, "toWorld"), tw);
        check((dry ? note("sensor perspective") : cudapath_set_camera_perspective(ctx, tw, (float) fov, (float) getFloat(n, "nearClip", 1e-2), (float) getFloat(n, "farClip", 1e4), w, h)));
        check((dry ? note("film") : cudapath_set_film(ctx, filter, fparam, alpha ? 1 : 0)));
        haveSensor = true;
    }
    void loadEmitter(const Node &n) {
        if (n.get("type") == "envmap") {          // src/emitters/envmap.cpp:100-190 with a Radiance RGBE file
            std::string file = getString(n, "filename", "");
            if (file.empty()) throw std::runtime_error("envmap: the 'filename' parameter is required");
            if (file[0] != '/') file = baseDir + "/" + file;
            if (getFloat(n, "gamma", 0.0) != 0.0) throw std::runtime_error("envmap: the 'gamma' override is not supported");
            if (child(n, "float", "intensityScale")) throw std::runtime_error("The 'intensityScale' parameter has been deprecated and is now called scale.");
            float tw[16]; toFloat(getTransform(n, "toWorld"), tw);
            if (dry) {
                std::ifstream probe(file, std::ios::binary);
                note(std::string("emitter envmap \"") + file + "\"" + (probe ? "" : " (file missing)"));
            } else check(cudapath_set_envmap_file(ctx, file.c_str(), tw, (float) getFloat(n, "scale", 1.0)));
            return;
        }
        if (n.get("type") != "sunsky") throw std::runtime_error("emitter plugin \"" + n.get("type") + "\" is outside the hair hot path (supported: sunsky, envmap with an .hdr file; raw environment maps go through cudapath_set_envmap)");
        if (child(n, "transform", "toWorld")) throw std::runtime_error("sunsky: a toWorld transform is not supported");
        auto sd = child(n, "vector", "sunDirection");
        if (!sd) throw std::runtime_error("sunsky: only the 'sunDirection' form is supported (no date/time/location)");
        float dir[3] = {(float) (sd->has("x") ? num(sd->get("x")) : 0), (float) (sd->has("y") ? num(sd->get("y")) : 0), (float) (sd->has("z") ? num(sd->get("z")) : 0)};
        float albedo[3]; getColor(n, "albedo", 0.2f, albedo);
        const double scale = getFloat(n, "scale", 1.0);
        check((dry ? note("emitter sunsky") : cudapath_set_sunsky(ctx, (float) getFloat(n, "turbidity", 3.0), albedo, dir, (float) getFloat(n, "skyScale", scale), (float) getFloat(n, "sunScale", scale),
                                  (float) getFloat(n, "sunRadiusScale", 1.0), (int) getInt(n, "resolution", 512))));
    }
    void loadIntegrator(const Node &n) {
        const std::string t = n.get("type");
        if (t != "path" && t != "cudapath") throw std::runtime_error("integrator plugin \"" + t + "\" is outside the hair hot path (supported: path, cudapath)");
        check((dry ? note("integrator path") : cudapath_set_integrator(ctx, (int) getInt(n, "maxDepth", -1), (int) getInt(n, "rrDepth", 5), getBool(n, "strictNormals", false) ? 1 : 0,
                                      getBool(n, "hideEmitters", false) ? 1 : 0)));
        haveIntegrator = true;
    }
    void load(Node &root) {
        if (root.tag != "scene") throw std::runtime_error("the root element must be <scene>");
        if (!root.has("version")) throw std::runtime_error("The requested scene cannot be loaded since it is missing a version attribute");   // scenehandler.cpp:228-250
        substitute(root);
        for (auto &c : root.children) if (c->tag == "bsdf") loadBsdf(*c);          // ids first: <ref> may precede the definition's use
        for (auto &c : root.children) {
            if (c->tag == "integrator") loadIntegrator(*c);
            else if (c->tag == "sensor") loadSensor(*c);
            else if (c->tag == "shape") loadShape(*c);
            else if (c->tag == "emitter") loadEmitter(*c);
            else if (c->tag == "bsdf" || c->tag == "default") continue;
            else throw std::runtime_error("unsupported scene element <" + c->tag + ">");
        }
        if (!haveSensor) throw std::runtime_error("the scene has no sensor");
        if (!haveIntegrator) check((dry ? note("integrator path (default)") : cudapath_set_integrator(ctx, -1, 5, 0, 0)));
    }
};

} // namespace

extern "C" int cudapath_set_error_message(const char *msg);

static int load_or_validate(cudapath_ctx *ctx, bool dry, const char *filename, const char *defines, uint32_t *out_spp, std::string *outReport);

extern "C" int cudapath_load_scene_xml(cudapath_ctx *ctx, const char *filename, const char *defines, uint32_t *out_spp) {
    if (!ctx) { cudapath_set_error_message("null argument"); return -1; }
    return load_or_validate(ctx, false, filename, defines, out_spp, nullptr);
}

extern "C" int cudapath_validate_scene_xml(const char *filename, const char *defines, char *report, size_t report_size) {
    std::string rep; uint32_t spp = 0;
    const int rc = load_or_validate(nullptr, true, filename, defines, &spp, &rep);
    if (rc == 0) rep += "sampleCount " + std::to_string(spp) + "\n";
    if (report && report_size) { const size_t n = std::min(rep.size(), report_size - 1); std::memcpy(report, rep.data(), n); report[n] = 0; }
    return rc;
}

static int load_or_validate(cudapath_ctx *ctx, bool dry, const char *filename, const char *defines, uint32_t *out_spp, std::string *outReport) {
    try {
        if (!filename) throw std::runtime_error("null argument");
        std::ifstream f(filename, std::ios::binary);
        if (!f) throw std::runtime_error(std::string("cannot open scene file \"") + filename + "\"");
        std::stringstream ss; ss << f.rdbuf();
        const std::string src = ss.str();
        Parser p(src);
        std::unique_ptr<Node> root = p.document();
        Loader L; L.ctx = ctx; L.dry = dry;
        std::string fn(filename); size_t slash = fn.find_last_of('/');
        L.baseDir = slash == std::string::npos ? "." : fn.substr(0, slash);
        if (defines) {
            std::string d(defines); size_t b = 0;
            while (b < d.size()) {
                size_t e = d.find(';', b); if (e == std::string::npos) e = d.size();
                std::string kv = d.substr(b, e - b); size_t eq = kv.find('=');
                if (!kv.empty()) { if (eq == std::string::npos) throw std::runtime_error("defines must be name=value pairs"); L.defines[kv.substr(0, eq)] = kv.substr(eq + 1); }
                b = e + 1;
            }
        }
        L.load(*root);
        if (out_spp) *out_spp = L.spp;
        if (!dry) {   // the job the file describes (film size x sampleCount over the devices of the context) picks the build effort
            int w = 0, h = 0;
            if (cudapath_film_size(ctx, &w, &h) == 0 && w > 0 && h > 0)
                cudapath_set_job_size_hint(ctx, (uint64_t) w * (uint64_t) h * (uint64_t) L.spp / (uint64_t) std::max(1, cudapath_device_count(ctx)));
        }
        if (outReport) for (auto &r : L.report) *outReport += r + "\n";
        return 0;
    } catch (const std::exception &e) {
        cudapath_set_error_message(e.what());
        return -1;
    }
}
