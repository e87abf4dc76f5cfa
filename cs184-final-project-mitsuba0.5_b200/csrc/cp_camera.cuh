// cp_camera.cuh -- perspective sensor ray generation on the device (sm_100a).
//
// Replaces PerspectiveCameraImpl::sampleRayDifferential (src/sensors/perspective.cpp:271-298); the
// sampleToCamera matrix and the near-plane differentials (perspective.cpp:126-160) are prepared on the host.
#pragma once
#include "cp_scene.cuh"

namespace cp {

__device__ __forceinline__ V3 xfm_point(const float *m, const V3 &p) {   // Transform::operator()(Point): divides by w
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return V3(x, y, z);
    return V3(x, y, z) / w;
}
__device__ __forceinline__ V3 xfm_vector(const float *m, const V3 &v) {
    return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

struct CameraRay { V3 o, d, rx, ry; float mint, maxt; };
__device__ __forceinline__ CameraRay camera_ray(const CameraDev &C, float px, float py, float diffScale) {
    CameraRay r;
    V3 nearP = xfm_point(C.s2c, V3(px * C.invResX, py * C.invResY, 0.0f));
    V3 d = normalize(nearP);
    float invZ = 1.0f / d.z;
    r.mint = C.nearClip * invZ; r.maxt = C.farClip * invZ;
    r.o = xfm_point(C.toWorld, V3(0.0f));
    r.d = xfm_vector(C.toWorld, d);
    V3 rx = xfm_vector(C.toWorld, normalize(nearP + V3(C.dx[0], C.dx[1], C.dx[2])));
    V3 ry = xfm_vector(C.toWorld, normalize(nearP + V3(C.dy[0], C.dy[1], C.dy[2])));
    // RayDifferential::scaleDifferential(1/sqrt(spp)) (integrator.cpp:143-144,178)
    r.rx = r.d + (rx - r.d) * diffScale; r.ry = r.d + (ry - r.d) * diffScale;
    return r;
}

} // namespace cp
