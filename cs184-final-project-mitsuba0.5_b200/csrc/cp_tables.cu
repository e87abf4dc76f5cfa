// cp_tables.cu -- load-time tables built on the device (sm_100a).
//
// Replaces (reference file:line):
//   MarschnerDiffuse::precomputeAzimuthalDistributions   src/bsdfs/marschner_diffuse.cpp:751-847
//   Azimuthal::Azimuthal (weights, +-1 dilation)         src/bsdfs/marschner_diffuse.cpp:41-64
//   InterpolatedDistribution1D ctor (row pdf/cdf/sums)   src/bsdfs/InterpolatedDistribution1D.hpp:37-67
//   EnvironmentMap texel storage + configure()           src/emitters/envmap.cpp:102-103, 260-329
// Every sum is accumulated in the reference's order (one thread walks the 140 quadrature nodes / the 64 or
// 512 row entries sequentially), so the tables agree with a CPU evaluation to the last few ulp.  These
// kernels run once per material / environment map; transcendental functions are evaluated in fp64 and
// rounded once so that the result does not depend on which fp32 libm produced it.
#include "cp_host.h"
#include <cuda_fp16.h>

namespace cp {

__device__ __forceinline__ float t_exp(float x) { return (float) exp((double) x); }
__device__ __forceinline__ float t_cos(float x) { return (float) cos((double) x); }
__device__ __forceinline__ float t_sin(float x) { return (float) sin((double) x); }
__device__ __forceinline__ float t_asin(float x) { return (float) asin((double) x); }

// marschner_diffuse.cpp:301-315
__device__ float gaussian_g(float beta, float theta) { return t_exp(-theta * theta / (2.0f * beta * beta)) / (sqrtf(2.0f * kPi) * beta); }
__device__ float detector_D(float beta, float phi) {
    float result = 0.0f, delta, shift = 0.0f;
    do {
        delta = gaussian_g(beta, phi + shift) + gaussian_g(beta, phi - shift - 2 * kPi);
        result += delta;
        shift += 2 * kPi;
    } while (delta > 1e-4f);
    return result;
}
__global__ void k_detector_table(float beta, float *Ds) { // :773-779 (the same beta_R table serves p = 0,1,2)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2048) Ds[i] = detector_D(beta, i / (2048 - 1.0f) * 2 * kPi);
}

__device__ __forceinline__ float approxD(const float *__restrict__ Ds, float phi) { // :782-788
    float u = (float) fabs(phi * (1.0 / (double) (2 * kPi) * (2048 - 1)));
    int x0 = int(u), x1 = x0 + 1;
    u -= x0;
    return Ds[x0 % 2048] * (1.0f - u) + Ds[x1 % 2048] * u;
}
__device__ __forceinline__ float lobePhi(float gammaI, float gammaT, int p) { return 2.0f * p * gammaT - 2.0f * gammaI + p * kPi; } // :317-319

// grid = 64 rows (cos theta_d = y/63), block = 192 threads.  out: 3 x 64 x 64 float4
__global__ void k_azimuthal_tables(float eta, float3 sigmaA, const float *__restrict__ points, const float *__restrict__ weights,
                                   const float *__restrict__ Ds, float4 *tab) {
    __shared__ float gammaI[140], gammaT[140], fres[140], absR[140], absG[140], absB[140], wgt[140];
    const int y = blockIdx.x, tid = threadIdx.x;
    const float cosHalfAngle = y / (64 - 1.0f);
    if (tid < 140) {
        float iorPrime = sqrtf(eta * eta - (1.0f - cosHalfAngle * cosHalfAngle)) / cosHalfAngle;
        float cosThetaT = sqrtf(1.0f - (1.0f - cosHalfAngle * cosHalfAngle) * (1.0f / eta) * (1.0f / eta));
        float rc = 1.0f / cosThetaT;                                   // Vector3f / float multiplies by the reciprocal
        float sx = sigmaA.x * rc, sy = sigmaA.y * rc, sz = sigmaA.z * rc;
        float gi = t_asin(points[tid]);
        float gt = t_asin(clampf(points[tid] / iorPrime, -1.0f, 1.0f));
        gammaI[tid] = gi; gammaT[tid] = gt; wgt[tid] = weights[tid];
        // Fresnel with the arguments in the reference's (swapped) order: cosThetaI = 1/eta, eta = cos(theta_d) cos(gamma_i)
        fres[tid] = fresnelDielectricExt(1.0f / eta, cosHalfAngle * t_cos(gi));
        float c = t_cos(gt);
        absR[tid] = t_exp(-sx * 2.0f * c); absG[tid] = t_exp(-sy * 2.0f * c); absB[tid] = t_exp(-sz * 2.0f * c);
    }
    __syncthreads();
    if (tid < 64) {
        const int phiI = tid;
        float phi = kPi * 2 * phiI / (64 - 1.0f);
        float integralR = 0.0f, ttx = 0, tty = 0, ttz = 0, trx = 0, try_ = 0, trz = 0;
        for (int i = 0; i < 140; ++i) {
            float fR = fres[i], Tx = absR[i], Ty = absG[i], Tz = absB[i];
            float k = (1.0f - fR) * (1.0f - fR);
            float ax = k * Tx, ay = k * Ty, az = k * Tz;                // A_TT
            float bx = ax * fR * Tx, by = ay * fR * Ty, bz = az * fR * Tz; // A_TRT = A_TT * f * T
            float w0 = wgt[i] * approxD(Ds, phi - lobePhi(gammaI[i], gammaT[i], 0));
            float w1 = wgt[i] * approxD(Ds, phi - lobePhi(gammaI[i], gammaT[i], 1));
            float w2 = wgt[i] * approxD(Ds, phi - lobePhi(gammaI[i], gammaT[i], 2));
            integralR += w0 * fR;
            ttx += w1 * ax; tty += w1 * ay; ttz += w1 * az;
            trx += w2 * bx; try_ += w2 * by; trz += w2 * bz;
        }
        const int o = phiI + y * 64;
        float r = 0.5f * integralR;
        tab[o] = make_float4(r, r, r, 0.0f);
        tab[4096 + o] = make_float4(0.5f * ttx, 0.5f * tty, 0.5f * ttz, 0.0f);
        tab[8192 + o] = make_float4(0.5f * trx, 0.5f * try_, 0.5f * trz, 0.0f);
    }
}

// grid = 3 lobes, block = 64 threads (one row each for the cdf pass)
__global__ void k_azimuthal_sampler(const float4 *__restrict__ tab, float *pdf, float *cdf, float *sums) {
    __shared__ float w0[64 * 64], w1[64 * 64];
    const int lobe = blockIdx.x, tid = threadIdx.x;
    const float4 *t = tab + lobe * 4096;
    for (int i = tid; i < 4096; i += 64) { float4 v = t[i]; w0[i] = fmaxf(v.x, fmaxf(v.y, v.z)); }
    __syncthreads();
    // dilation in x: forward pass max(w[x], w[x+1]) then backward pass max(w[x], w[x-1]) (marschner_diffuse.cpp:50-55)
    for (int i = tid; i < 4096; i += 64) { int x = i & 63; w1[i] = x < 63 ? fmaxf(w0[i], w0[i + 1]) : w0[i]; }
    __syncthreads();
    for (int i = tid; i < 4096; i += 64) { int x = i & 63; w0[i] = x > 0 ? fmaxf(w1[i], w1[i - 1]) : w1[i]; }
    __syncthreads();
    // dilation in y (:56-61)
    for (int i = tid; i < 4096; i += 64) { int yy = i >> 6; w1[i] = yy < 63 ? fmaxf(w0[i], w0[i + 64]) : w0[i]; }
    __syncthreads();
    for (int i = tid; i < 4096; i += 64) { int yy = i >> 6; w0[i] = yy > 0 ? fmaxf(w1[i], w1[i - 64]) : w1[i]; }
    __syncthreads();
    // InterpolatedDistribution1D ctor, one distribution (row) per thread
    const int dist = tid;
    float *p = pdf + lobe * 4096 + dist * 64, *c = cdf + lobe * 64 * 65 + dist * 65;
    c[0] = 0.0f;
    for (int x = 0; x < 64; ++x) c[x + 1] = w0[dist * 64 + x] + c[x];
    float sum = c[64];
    sums[lobe * 64 + dist] = sum;
    if (sum < 1e-4f) {
        float ratio = 1.0f / 64;
        for (int x = 0; x < 64; ++x) { p[x] = ratio; c[x] = x * ratio; }
    } else {
        float scale = 1.0f / sum;
        for (int x = 0; x < 64; ++x) { p[x] = w0[dist * 64 + x] * scale; c[x] *= scale; }
    }
    c[64] = 1.0f;
}

#define CKT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return false; } } while (0)

bool build_marschner_tables(float eta, float betaR, const float sigmaA[3], const float *glPoints140, const float *glWeights140,
                            cudaStream_t stream, MarschnerTables &out, std::string &err) {
    float *d_pts = nullptr, *d_wts = nullptr, *d_Ds = nullptr;
    CKT(dev_alloc(&d_pts, 140 * 4)); CKT(dev_alloc(&d_wts, 140 * 4)); CKT(dev_alloc(&d_Ds, 2048 * 4));
    CKT(dev_alloc(&out.tab, sizeof(float4) * 3 * 4096)); CKT(dev_alloc(&out.pdf, 4 * 3 * 4096));
    CKT(dev_alloc(&out.cdf, 4 * 3 * 64 * 65)); CKT(dev_alloc(&out.sums, 4 * 3 * 64));
    CKT(cudaMemcpyAsync(d_pts, glPoints140, 140 * 4, cudaMemcpyHostToDevice, stream));
    CKT(cudaMemcpyAsync(d_wts, glWeights140, 140 * 4, cudaMemcpyHostToDevice, stream));
    k_detector_table<<<8, 256, 0, stream>>>(betaR, d_Ds);
    k_azimuthal_tables<<<64, 192, 0, stream>>>(eta, make_float3(sigmaA[0], sigmaA[1], sigmaA[2]), d_pts, d_wts, d_Ds, out.tab);
    k_azimuthal_sampler<<<3, 64, 0, stream>>>(out.tab, out.pdf, out.cdf, out.sums);
    CKT(cudaStreamSynchronize(stream));
    CKT(cudaGetLastError());
    dev_free(d_pts); dev_free(d_wts); dev_free(d_Ds);
    return true;
}

// ------------------------------------------------------------------------------------------ envmap
__global__ void k_env_quantize(const float *__restrict__ rgb, int n, float4 *texels) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r = fmaxf(rgb[3 * i], 0.0f), g = fmaxf(rgb[3 * i + 1], 0.0f), b = fmaxf(rgb[3 * i + 2], 0.0f); // mipmap.h:232-240 clamps negatives
    texels[i] = make_float4(__half2float(__float2half_rn(r)), __half2float(__float2half_rn(g)), __half2float(__float2half_rn(b)), 0.0f);
}
void quantize_texels(const float *d_rgb, int n, float4 *d_texels, cudaStream_t stream) { if (n > 0) k_env_quantize<<<(n + 255) / 256, 256, 0, stream>>>(d_rgb, n, d_texels); }
// one thread per row: conditional CDF over luminance (envmap.cpp:284-299)
__global__ void k_env_rows(const float4 *__restrict__ texels, int w, int h, float *cdfCols, float *colSums) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    float *c = cdfCols + (size_t) y * (w + 1);
    float colSum = 0;
    c[0] = 0;
    for (int x = 0; x < w; ++x) {
        float4 v = texels[(size_t) y * w + x];
        colSum += luminance(V3(v.x, v.y, v.z));
        c[x + 1] = colSum;
    }
    float normalization = 1.0f / colSum;
    for (int x = 1; x < w; ++x) c[x] *= normalization;
    c[w] = 1.0f;
    colSums[y] = colSum;
}
// single thread: marginal CDF weighted by sin(theta) (envmap.cpp:282-315)
__global__ void k_env_marginal(const float *__restrict__ colSums, int w, int h, float *cdfRows, float *rowWeights, float *outNorm) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float rowSum = 0.0f;
    cdfRows[0] = 0;
    for (int y = 0; y < h; ++y) {
        float weight = t_sin((y + 0.5f) * kPi / h);
        rowWeights[y] = weight;
        rowSum += colSums[y] * weight;
        cdfRows[y + 1] = rowSum;
    }
    float normalization = 1.0f / rowSum;
    for (int y = 1; y < h; ++y) cdfRows[y] *= normalization;
    cdfRows[h] = 1.0f;
    outNorm[0] = 1.0f / (rowSum * (2 * kPi / w) * (kPi / h));
    outNorm[1] = rowSum;
}

bool build_env_tables(const float *d_rgb, int w, int h, cudaStream_t stream, EnvTables &out, std::string &err) {
    float *d_colSums = nullptr, *d_norm = nullptr;
    const int n = w * h;
    CKT(dev_alloc(&out.texels, sizeof(float4) * (size_t) n)); CKT(dev_alloc(&out.cdfCols, 4 * (size_t) (w + 1) * h));
    CKT(dev_alloc(&out.cdfRows, 4 * (size_t) (h + 1))); CKT(dev_alloc(&out.rowWeights, 4 * (size_t) h));
    CKT(dev_alloc(&d_colSums, 4 * (size_t) h)); CKT(dev_alloc(&d_norm, 8));
    k_env_quantize<<<(n + 255) / 256, 256, 0, stream>>>(d_rgb, n, out.texels);
    k_env_rows<<<(h + 63) / 64, 64, 0, stream>>>(out.texels, w, h, out.cdfCols, d_colSums);
    k_env_marginal<<<1, 32, 0, stream>>>(d_colSums, w, h, out.cdfRows, out.rowWeights, d_norm);
    float hn[2] = {0, 0};
    CKT(cudaMemcpyAsync(hn, d_norm, 8, cudaMemcpyDeviceToHost, stream));
    CKT(cudaStreamSynchronize(stream));
    CKT(cudaGetLastError());
    dev_free(d_colSums); dev_free(d_norm);
    if (!(hn[1] > 0) || !isfinite(hn[1])) { err = "The environment map is completely black or contains invalid values"; return false; }
    out.normalization = hn[0];
    return true;
}

} // namespace cp
