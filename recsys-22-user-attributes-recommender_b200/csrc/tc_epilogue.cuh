// Epilogue helpers shared by the tcgen05 dense-layer kernels (gemm_tc.cu, ffn_fused_tc.cu): bf16 packing and the warp-private
// staging tiles that turn per-row (TMEM lane = output row) accesses into coalesced 64-byte row pieces.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// ---- coalesced epilogue traffic -------------------------------------------------------------------------------------------
// A TMEM lane is an output row, so each epilogue thread owns a row and would touch memory in 16-byte pieces that are a whole
// row apart from its neighbours' -- 32 different cache lines per warp instruction.  With a warp-private 2 KB staging tile
// (32 rows x 64 bytes, 16-byte units XOR-swizzled) the warp moves 64-byte row pieces instead: lanes 4r..4r+3 handle row r of
// eight rows per instruction, i.e. whole sectors, 4x fewer L1 wavefronts.  stage == NULL: per-thread accesses (per-tile kernel).
struct WarpTile {
    uint8_t* stage;      // 2 KB, private to the warp, or NULL
    int lane;
    int rows_valid;      // rows [0, rows_valid) of the warp's 32 rows exist
};
__device__ __forceinline__ uint32_t wt_off(int r, int u) { return (uint32_t)(r * 64 + ((u ^ ((r >> 1) & 3)) << 4)); }
// every lane stores its own row's 64 bytes (`mine`) to gbase + lane * pitch
__device__ __forceinline__ void tile_store64(const WarpTile& t, const uint4 (&mine)[4], uint8_t* gbase, size_t pitch) {
    if (t.stage == nullptr) {
        if (t.lane < t.rows_valid) {
#pragma unroll
            for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(gbase + (size_t)t.lane * pitch + u * 16) = mine[u];
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(t.stage + wt_off(t.lane, u)) = mine[u];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = j * 8 + t.lane / 4, u = t.lane % 4;
        const uint4 w = *reinterpret_cast<const uint4*>(t.stage + wt_off(r, u));
        if (r < t.rows_valid) *reinterpret_cast<uint4*>(gbase + (size_t)r * pitch + u * 16) = w;
    }
    __syncwarp();
}
// every lane receives its own row's 64 bytes from gbase + lane * pitch (zeros for rows that do not exist)
__device__ __forceinline__ void tile_load64(const WarpTile& t, uint4 (&mine)[4], const uint8_t* gbase, size_t pitch) {
    if (t.stage == nullptr) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            mine[u] = t.lane < t.rows_valid ? __ldg(reinterpret_cast<const uint4*>(gbase + (size_t)t.lane * pitch + u * 16)) : make_uint4(0u, 0u, 0u, 0u);
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = j * 8 + t.lane / 4, u = t.lane % 4;
        const uint4 w = r < t.rows_valid ? __ldg(reinterpret_cast<const uint4*>(gbase + (size_t)r * pitch + u * 16)) : make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(t.stage + wt_off(r, u)) = w;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) mine[u] = *reinterpret_cast<const uint4*>(t.stage + wt_off(t.lane, u));
    __syncwarp();
}
__device__ __forceinline__ void pack32_bf16(const float (&v)[32], uint4 (&w)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        w[u].x = pack_bf16(v[8 * u], v[8 * u + 1]); w[u].y = pack_bf16(v[8 * u + 2], v[8 * u + 3]);
        w[u].z = pack_bf16(v[8 * u + 4], v[8 * u + 5]); w[u].w = pack_bf16(v[8 * u + 6], v[8 * u + 7]);
    }
}

