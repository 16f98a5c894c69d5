// Synthetic translating / deforming Gaussian-blob stacks, generated on the device (benchmark
// utility).  One blob per 16^3-voxel cell (16^2 in 2D), all parameters derived
// from a counter-based hash of (seed, cell), so any (t, z) shard of a time-lapse can be generated
// independently and reproducibly.  Internal header.
#pragma once
#include "common.cuh"

namespace of3d {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ float u01(uint64_t h, int part) {  // two 24-bit uniforms per hash
    return (float)((h >> (part ? 40 : 8)) & 0xFFFFFF) * (1.0f / 16777216.0f);
}

constexpr int kCell = 16;

// intensity contribution of the blob owned by cell (cz,cy,cx) at voxel (z,y,x), time t
__device__ __forceinline__ float blob_at(uint64_t seed, int64_t cz, int64_t cy, int64_t cx, float z, float y, float x, float t,
                                         bool is2d) {
    const uint64_t key = mix64(seed ^ mix64((uint64_t)cz * 0x1000003ull + (uint64_t)cy) ^ mix64((uint64_t)cx + 0x51ull));
    const uint64_t h0 = mix64(key), h1 = mix64(key + 1), h2 = mix64(key + 2), h3 = mix64(key + 3), h4 = mix64(key + 4),
                   h5 = mix64(key + 5);
    // centre: cell centre + jitter (+-4) + bounded oscillation (amplitude <= 2 px, speed <= ~0.7 px/frame)
    const float w0 = 0.15f + 0.2f * u01(h3, 0);  // rad/frame
    const float ax = 2.0f * u01(h1, 0), ay = 2.0f * u01(h1, 1), az = 2.0f * u01(h2, 0);
    const float px = 6.2831853f * u01(h2, 1), py = 6.2831853f * u01(h3, 1), pz = 6.2831853f * u01(h4, 0);
    const float bx = (cx + 0.5f) * kCell + 8.0f * (u01(h0, 0) - 0.5f) + ax * __sinf(w0 * t + px);
    const float by = (cy + 0.5f) * kCell + 8.0f * (u01(h0, 1) - 0.5f) + ay * __sinf(w0 * t + py);
    const float bz = is2d ? 0.0f : (cz + 0.5f) * kCell + 8.0f * (u01(h4, 1) - 0.5f) + az * __sinf(w0 * t + pz);
    const float amp = 200.0f + 1300.0f * u01(h5, 0);
    const float s = (1.8f + 1.0f * u01(h5, 1)) * (1.0f + 0.1f * __sinf(0.11f * t + px));  // breathing width ("deforming")
    const float dx = x - bx, dy = y - by, dz = z - bz;
    const float r2 = dx * dx + dy * dy + dz * dz;
    return amp * __expf(-r2 / (2.0f * s * s));
}

__global__ void __launch_bounds__(256) synth_blobs_kernel(uint16_t* __restrict__ out, int64_t nt, int64_t nz, int64_t ny, int64_t nx,
                                                          int64_t t0, int64_t z0, uint64_t seed) {
    const int64_t nvol = nz * ny * nx, n = nt * nvol;
    const bool is2d = (nz == 1 && z0 == 0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ti = i / nvol, r = i - ti * nvol;
        const int64_t zi = r / (ny * nx), yi = (r / nx) % ny, xi = r % nx;
        const float t = (float)(t0 + ti), z = (float)(z0 + zi), y = (float)yi, x = (float)xi;
        // the 2 cells per axis whose blobs can reach this voxel (reach 16 from a cell centre)
        const int64_t gx = (xi + kCell / 2) / kCell - 1, gy = (yi + kCell / 2) / kCell - 1, gz = (z0 + zi + kCell / 2) / kCell - 1;
        float v = 100.0f;
        for (int a = 0; a < (is2d ? 1 : 2); ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) v += blob_at(seed, is2d ? 0 : gz + a, gy + b, gx + c, z, y, x, t, is2d);
        // Gaussian noise sigma 5 (Box-Muller on a per-sample hash)
        const uint64_t hn = mix64(seed * 0x2545F4914F6CDD1Dull + (uint64_t)((t0 + ti) * 0x100000001B3ull) +
                                  (uint64_t)(z0 + zi) * (uint64_t)(ny * nx) + (uint64_t)(yi * nx + xi));
        const float u1 = fmaxf(u01(hn, 0), 5.9604645e-8f), u2 = u01(hn, 1);
        v += 5.0f * sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
        out[i] = (uint16_t)fminf(fmaxf(rintf(v), 0.0f), 65535.0f);
    }
}

static inline void launch_synth_blobs(uint16_t* out, int64_t nt, int64_t nz, int64_t ny, int64_t nx, int64_t t0, int64_t z0,
                                      uint64_t seed, int sm_count, cudaStream_t stream) {
    const int64_t n = nt * nz * ny * nx;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count * 64));
    synth_blobs_kernel<<<grid, 256, 0, stream>>>(out, nt, nz, ny, nx, t0, z0, seed);
}

}  // namespace of3d
