// Shared device utilities for the agdiff_b200 kernels (sm_100a).
//
// Every dense contraction on the hot path is a [128 rows x K] x [K x N] product whose rows are
// edges (or atoms) and whose weights are shared by all rows, so all kernels use one tile shape:
//   * 128 rows per CTA tile, 256 threads, each thread an 8x8 (N=128) or 8x4 (N=64) register block;
//   * the activation tile lives in shared memory TRANSPOSED (As[k][m], leading dim 132) so the
//     per-k operand reads are two conflict-free LDS.128 per thread and a layer's output can be
//     written back in place as the next layer's input (the whole MLP chain stays on chip);
//   * weights ([K][N] row-major, pre-transposed/folded on the host) stream from L2 through a
//     double-buffered cp.async ring of 16-row chunks.
// fp32 FFMA throughout: plain TF32 fails the rtol 1e-4 parity bar.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace agd {

constexpr int HID = 128;        // hidden_dim
constexpr int TM = 128;         // rows per tile
constexpr int LDA = TM + 4;     // padded leading dimension of the transposed activation tile
constexpr int NT = 256;         // threads per CTA of the tile kernels
constexpr int KC = 16;          // weight rows per cp.async chunk
constexpr int AS_FLOATS = HID * LDA;          // one activation tile (K up to 128)
constexpr int WS_FLOATS = 2 * KC * 128;       // weight ring
constexpr float LN2F = 0.69314718055994530942f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// thread -> (tx, ty) in the 16x16 register-block grid.  Within a warp tx spans 8 and ty spans 4
// consecutive values, so the per-k loads touch 128 B (weights) + 64 B (activations): 1 wavefront each.
struct TileCoord {
  int tx, ty;
};
__device__ __forceinline__ TileCoord tile_coord() {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  TileCoord c;
  c.tx = (l & 7) + 8 * (w & 1);
  c.ty = (l >> 3) + 4 * (w >> 1);
  return c;
}
__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)); }
__device__ __forceinline__ int tile_col(int tx, int j) { return (j >> 2) * 64 + tx * 4 + (j & 3); }

// acc[8][N/16] (+)= As^T[128 x K] * Wt[K x N].  Wt is global, [K][N] row-major, 16-byte aligned.
// Begins with a block barrier after the first weight chunk lands (so As written by any thread just
// before the call is visible) and ends with one (so As / Ws may be overwritten right after).
template <int K, int N, bool ACCUM>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Wt, const float* __restrict__ As, float* Ws,
                                          float (&acc)[8][N / 16], int tx, int ty) {
  static_assert(N == 64 || N == 128, "N must be 64 or 128");
  static_assert(K % KC == 0, "K must be a multiple of KC");
  constexpr int NG = N / 64;
  constexpr int NCH = K / KC;
  constexpr int V4 = KC * N / 4;  // float4 per chunk
  static_assert(V4 % NT == 0, "chunk must divide over the CTA");
  if (!ACCUM) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < N / 16; ++j) acc[i][j] = 0.f;
  }
  const int tid = threadIdx.x;
  {
    const float4* g = reinterpret_cast<const float4*>(Wt);
    float4* s = reinterpret_cast<float4*>(Ws);
#pragma unroll
    for (int i = 0; i < V4 / NT; ++i) cp_async16(s + tid + i * NT, g + tid + i * NT);
    cp_async_commit();
  }
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) {
      const float4* g = reinterpret_cast<const float4*>(Wt + (size_t)(c + 1) * KC * N);
      float4* s = reinterpret_cast<float4*>(Ws + ((c + 1) & 1) * KC * N);
#pragma unroll
      for (int i = 0; i < V4 / NT; ++i) cp_async16(s + tid + i * NT, g + tid + i * NT);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* w = Ws + (c & 1) * KC * N;
    const float* a = As + c * KC * LDA;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(a + kk * LDA + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(a + kk * LDA + 64 + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(w + kk * N + g * 64 + tx * 4);
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][g * 4 + j] = fmaf(av[i], bv[j], acc[i][g * 4 + j]);
      }
    }
    __syncthreads();
  }
}

// Write a register block back as the next layer's transposed input: As[n][m] = f(acc, m, n).
template <int N, class F>
__device__ __forceinline__ void tile_store_smem(const float (&acc)[8][N / 16], float* As, int tx, int ty, F f) {
#pragma unroll
  for (int j = 0; j < N / 16; ++j) {
    const int n = tile_col(tx, j);
    float4 lo, hi;
    lo.x = f(acc[0][j], ty * 4 + 0, n);
    lo.y = f(acc[1][j], ty * 4 + 1, n);
    lo.z = f(acc[2][j], ty * 4 + 2, n);
    lo.w = f(acc[3][j], ty * 4 + 3, n);
    hi.x = f(acc[4][j], 64 + ty * 4 + 0, n);
    hi.y = f(acc[5][j], 64 + ty * 4 + 1, n);
    hi.z = f(acc[6][j], 64 + ty * 4 + 2, n);
    hi.w = f(acc[7][j], 64 + ty * 4 + 3, n);
    *reinterpret_cast<float4*>(As + n * LDA + ty * 4) = lo;
    *reinterpret_cast<float4*>(As + n * LDA + 64 + ty * 4) = hi;
  }
}

// Write a register block to a row-major global matrix G[row0 + m][col0 + n] (leading dim ld).
template <int N, class F>
__device__ __forceinline__ void tile_store_global(const float (&acc)[8][N / 16], float* __restrict__ G, int64_t row0,
                                                  int64_t n_rows, int ld, int col0, int tx, int ty, F f) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = tile_row(ty, i);
    const int64_t r = row0 + m;
    if (r < n_rows) {
#pragma unroll
      for (int g = 0; g < N / 64; ++g) {
        const int n = g * 64 + tx * 4;
        float4 v;
        v.x = f(acc[i][g * 4 + 0], m, n + 0);
        v.y = f(acc[i][g * 4 + 1], m, n + 1);
        v.z = f(acc[i][g * 4 + 2], m, n + 2);
        v.w = f(acc[i][g * 4 + 3], m, n + 3);
        *reinterpret_cast<float4*>(G + r * ld + col0 + n) = v;
      }
    }
  }
}

// Load rows [row0, row0+128) x cols [col0, col0+K) of a row-major global matrix into As[k][m].
// Lanes walk consecutive rows (conflict-free transposed stores); the 16-byte pieces of one row are
// fetched by consecutive warps, so every 32-byte sector is consumed through L1.
template <int K>
__device__ __forceinline__ void tile_load_T(const float* __restrict__ G, int64_t row0, int64_t n_rows, int ld, int col0,
                                            float* As) {
  constexpr int ITERS = TM * (K / 4) / NT;
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    const int idx = it * NT + threadIdx.x;
    const int m = idx & (TM - 1);
    const int kq = idx >> 7;
    const int64_t r = row0 + m;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n_rows) v = __ldg(reinterpret_cast<const float4*>(G + r * ld + col0) + kq);
    As[(kq * 4 + 0) * LDA + m] = v.x;
    As[(kq * 4 + 1) * LDA + m] = v.y;
    As[(kq * 4 + 2) * LDA + m] = v.z;
    As[(kq * 4 + 3) * LDA + m] = v.w;
  }
}

// As[k][m] = X[src[m]][k] * X[dst[m]][k]  (assemble_atom_pair_feature, common.py:106-109)
__device__ __forceinline__ void tile_load_pair_T(const float* __restrict__ X, const int* s_src, const int* s_dst,
                                                 float* As) {
  constexpr int ITERS = TM * (HID / 4) / NT;
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    const int idx = it * NT + threadIdx.x;
    const int m = idx & (TM - 1);
    const int kq = idx >> 7;
    const float4 a = __ldg(reinterpret_cast<const float4*>(X + (size_t)s_src[m] * HID) + kq);
    const float4 b = __ldg(reinterpret_cast<const float4*>(X + (size_t)s_dst[m] * HID) + kq);
    As[(kq * 4 + 0) * LDA + m] = a.x * b.x;
    As[(kq * 4 + 1) * LDA + m] = a.y * b.y;
    As[(kq * 4 + 2) * LDA + m] = a.z * b.z;
    As[(kq * 4 + 3) * LDA + m] = a.w * b.w;
  }
}

// ---------------------------------------------------------------- scalar math (precise variants)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// ShiftedSoftplus with learnable beta, schnet.py:77-80 (F.softplus threshold 20)
__device__ __forceinline__ float ssp(float x, float beta) {
  const float y = beta * x;
  const float sp = (y > 20.0f) ? y : log1pf(expf(y));
  return sp - LN2F;
}
__device__ __forceinline__ float leaky02(float x) { return x > 0.f ? x : 0.2f * x; }
// torch.relu propagates NaN (fmaxf would swallow it and hide a diverged trajectory from the NaN guard)
__device__ __forceinline__ float relu_(float x) { return x < 0.f ? 0.f : x; }

// CFConv cutoff envelope times the learnable distance weight, schnet.py:136-149.
// dw = [w1(32) | b1(32) | w2(32) | b2]
__device__ __forceinline__ float cfconv_edge_weight(float d, const float* __restrict__ dw, float cutoff, int smooth) {
  float z = __ldg(dw + 96);
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const float hj = fmaf(__ldg(dw + j), d, __ldg(dw + 32 + j));
    z = fmaf(__ldg(dw + 64 + j), relu_(hj), z);
  }
  const float lw = sigmoidf_(z);
  float C;
  if (smooth) {
    C = 0.5f * (cosf(d * 3.14159265358979323846f / cutoff) + 1.0f);
    C = (d <= cutoff) ? C : 0.f;
  } else {
    const float t = d - cutoff;
    C = expf(-(t * t) / (2.0f * cutoff * cutoff));
  }
  C = (d <= cutoff && d >= 0.f) ? C : 0.f;
  return lw * C;
}

// same, with the distance MLP already staged in shared memory
__device__ __forceinline__ float cfconv_edge_weight_smem(float d, const float* dw, float cutoff, int smooth) {
  float z = dw[96];
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const float hj = fmaf(dw[j], d, dw[32 + j]);
    z = fmaf(dw[64 + j], relu_(hj), z);
  }
  const float lw = sigmoidf_(z);
  float C;
  if (smooth) {
    C = 0.5f * (cosf(d * 3.14159265358979323846f / cutoff) + 1.0f);
    C = (d <= cutoff) ? C : 0.f;
  } else {
    const float t = d - cutoff;
    C = expf(-(t * t) / (2.0f * cutoff * cutoff));
  }
  C = (d <= cutoff && d >= 0.f) ? C : 0.f;
  return lw * C;
}

__device__ __forceinline__ float block_rows_tail_guard(int64_t r, int64_t n) { return r < n ? 1.f : 0.f; }

}  // namespace agd
