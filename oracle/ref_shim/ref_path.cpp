// oracle/ref_shim/ref_path.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry point around the reference's path integrator: src/integrators/path/path.cpp is compiled UNMODIFIED against
// fake_path/path_shim.h (oracle/Makefile), and MIPathTracer::Li is called for one camera ray with the scene components supplied
// through the callback table.  Output: oracle/_ref/libref_path.so.
#include "path_shim.h"
using namespace mitsuba;
extern "C" void *ref_create_MIPathTracer(const Properties *);
extern "C" {
// One call of MIPathTracer::Li as renderBlock makes it (integrator.cpp:157-181): rRec.newQuery(queryType), Li(sensorRay, rRec).
// queryHasOpacity: ESensorRay with (film has alpha) or without EOpacity.  Returns the path depth counter at exit.
int ref_path_li(const RefPathCallbacks *cb, int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int queryHasOpacity,
                const float o[3], const float d[3], float mint, float maxt, const float rxDir[3], const float ryDir[3], float outLi[3], float *outAlpha) {
    Properties props; props.maxDepth = maxDepth; props.rrDepth = rrDepth; props.strictNormals = strictNormals != 0; props.hideEmitters = hideEmitters != 0;
    MonteCarloIntegrator *integrator = (MonteCarloIntegrator *) ref_create_MIPathTracer(&props);
    Scene scene; scene.cb = cb; scene.env.cb = cb;
    RadianceQueryRecord rRec(&scene, nullptr);
    int queryType = RadianceQueryRecord::ESensorRay;
    if (!queryHasOpacity) queryType &= ~RadianceQueryRecord::EOpacity;
    rRec.newQuery(queryType, nullptr);
    RayDifferential ray; ray.o = Point(o[0], o[1], o[2]); ray.d = Vector(d[0], d[1], d[2]); ray.mint = mint; ray.maxt = maxt; ray.time = 0;
    ray.rxDirection = Vector(rxDir[0], rxDir[1], rxDir[2]); ray.ryDirection = Vector(ryDir[0], ryDir[1], ryDir[2]); ray.hasDifferentials = true;
    const Spectrum L = integrator->Li(ray, rRec);
    outLi[0] = L.s[0]; outLi[1] = L.s[1]; outLi[2] = L.s[2]; *outAlpha = rRec.alpha;
    const int depth = rRec.depth;
    delete integrator;
    return depth;
}
}
