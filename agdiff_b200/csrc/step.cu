// The Langevin/DDPM position update (dualenc.py:506-545) as ONE kernel per step:
// eq_transform of the local and (when sigma < global_start_sigma) global edge scores
// (geometry.py:9-17), clip_norm (dualenc.py:586-589), the noise/position update, the NaN guard,
// center_pos (dualenc.py:581-583) and the optional clamp.  One CTA per molecule; each atom sums
// its CSC in-edge segments and its canonical out-edge segments (four threads, one per segment), so
// there are no atomics and the result is bit-reproducible (and independent of how molecules are
// sharded over GPUs).
#include "common.cuh"
#include "kernels.h"

namespace agd {

constexpr int STEP_THREADS = 256;                 // four threads per atom
constexpr int STEP_ATOMS = STEP_THREADS / 4;      // atoms per sweep of the CTA
// (atoms per quad of threads: 4 for molecules of <= 256 atoms, 8 up to AGD_MAX_MOL_ATOMS = 512 - template parameter of the kernel)

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

// three standard normals for (seed, molecule id, atom-in-molecule, step): Philox4x32-10 + Box-Muller
__device__ __forceinline__ void normal3(uint64_t seed, int64_t gid, int atom, int step, float& z0, float& z1, float& z2) {
  uint32_t c[4] = {(uint32_t)atom, (uint32_t)step, (uint32_t)(gid & 0xffffffff), (uint32_t)((uint64_t)gid >> 32)};
  philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  const float u0 = ((c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = ((c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u1, &s0, &c0);
  sincosf(6.283185307179586f * u3, &s1, &c1);
  z0 = r0 * c0;
  z1 = r0 * s0;
  z2 = r1 * c1;
  (void)s1;
}

struct StepArgs {
  float* pos;
  const int* mol_ptr;
  const int64_t* mol_gid;
  int n_atoms;
  const int *in_ptr, *e_src, *e_type;
  const float *e_len, *s_csc;
  const int *out_ptr, *c_dst, *c_type;
  const float *c_len, *s_canon;
  const int *lin_ptr, *lsrc;
  const float *llen, *sl_csc;
  const int *lout_ptr, *lcdst;
  const float *lclen, *sl_canon;
  const float* sched;
  int* counters;
  int* nan_mol;   // [n_mols] first step at which the molecule's positions became NaN (INT_MAX: never)
  StepParams p;
};

__device__ __forceinline__ void clip3(float& x, float& y, float& z, float limit) {
  const float nrm = sqrtf(x * x + y * y + z * z);
  if (nrm > limit) {
    const float f = limit / nrm;
    x *= f; y *= f; z *= f;
  }
}

// One CTA per molecule, FOUR threads per atom: lane part p of an atom's quad sums one of its four edge segments (local in,
// local out, global in, global out; each in CSC / canonical order), the quad combines them in a fixed order with shuffles -
// eq = (in + out) per branch - so the result is bit-reproducible and does not depend on batch composition, while the four
// dependent-load chains of an atom run side by side (the one-thread-per-atom version was latency-bound at 6 % issue rate).
template <int STEP_MAX_PER_THREAD>
__global__ void __launch_bounds__(STEP_THREADS) langevin_step_kernel(const StepArgs a) {
  __shared__ float red[3][STEP_THREADS / 32];
  __shared__ int s_bad;
  const int m = blockIdx.x;
  const int a0 = a.mol_ptr[m], n = a.mol_ptr[m + 1] - a0;
  const int tid = threadIdx.x, part = tid & 3, slot = tid >> 2;
  const int step = a.counters[1];
  const float sigma = a.sched[4 * step], step_size = a.sched[4 * step + 1], nscale = a.sched[4 * step + 2];
  const float* pos = a.pos;
  if (tid == 0) s_bad = 0;
  float nx[STEP_MAX_PER_THREAD], ny[STEP_MAX_PER_THREAD], nz[STEP_MAX_PER_THREAD];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  bool bad = false;
#pragma unroll
  for (int q = 0; q < STEP_MAX_PER_THREAD; ++q) {
    const int i = slot + q * STEP_ATOMS;
    nx[q] = ny[q] = nz[q] = 0.f;
    const bool active = i < n;                       // (whole quads are active or not: shuffles below stay converged per quad)
    const int at = a0 + (active ? i : 0);
    const float px = pos[3 * (size_t)at], py = pos[3 * (size_t)at + 1], pz = pos[3 * (size_t)at + 2];
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (active) {
      // Each part walks one sorted segment.  The walk is latency-bound (index -> position, two dependent L2 round trips per
      // edge), so edges are taken four at a time: all loads of a batch are issued before the first use, the sum is then
      // accumulated edge by edge in segment order - the same arithmetic in the same order as a one-edge-at-a-time loop.
      const int* ptr = (part == 0) ? a.lin_ptr : (part == 1) ? a.lout_ptr : (part == 2) ? a.in_ptr : a.out_ptr;
      const int* other = (part == 0) ? a.lsrc : (part == 1) ? a.lcdst : (part == 2) ? a.e_src : a.c_dst;
      const float* elen = (part == 0) ? a.llen : (part == 1) ? a.lclen : (part == 2) ? a.e_len : a.c_len;
      const float* esc = (part == 0) ? a.sl_csc : (part == 1) ? a.sl_canon : (part == 2) ? a.s_csc : a.s_canon;
      const int* etype = (part == 2) ? a.e_type : (part == 3) ? a.c_type : nullptr;   // global parts skip the local edges (type > 0)
      const bool in_seg = (part & 1) == 0;   // in-segments: the atom is edge_index[1] -> subtract (pos[src] - p); out-segments: add (p - pos[dst])
      const int e0 = (part < 2 || a.p.use_global) ? ptr[at] : 0, e1 = (part < 2 || a.p.use_global) ? ptr[at + 1] : 0;
      for (int e = e0; e < e1; e += 4) {
        int j[4];
        float inv[4], sc[4], qx[4], qy[4], qz[4];
        bool use[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ee = (e + u < e1) ? e + u : e1 - 1;
          use[u] = (e + u < e1) && !(etype && etype[ee] > 0);
          j[u] = other[ee];
          inv[u] = elen[ee];
          sc[u] = esc[ee];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          qx[u] = pos[3 * (size_t)j[u]]; qy[u] = pos[3 * (size_t)j[u] + 1]; qz[u] = pos[3 * (size_t)j[u] + 2];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!use[u]) continue;
          const float iv = 1.0f / inv[u];
          if (in_seg) {
            vx -= (iv * (qx[u] - px)) * sc[u];
            vy -= (iv * (qy[u] - py)) * sc[u];
            vz -= (iv * (qz[u] - pz)) * sc[u];
          } else {
            vx += (iv * (px - qx[u])) * sc[u];
            vy += (iv * (py - qy[u])) * sc[u];
            vz += (iv * (pz - qz[u])) * sc[u];
          }
        }
      }
    }
    // quad combine: parts (0,1) -> local sum, parts (2,3) -> global sum; every lane of the quad ends up with both
    const float ox = __shfl_xor_sync(0xffffffffu, vx, 1), oy = __shfl_xor_sync(0xffffffffu, vy, 1), oz = __shfl_xor_sync(0xffffffffu, vz, 1);
    const float bx = (part & 1) ? ox + vx : vx + ox, by = (part & 1) ? oy + vy : vy + oy, bz = (part & 1) ? oz + vz : vz + oz;   // in + out
    const float qx = __shfl_xor_sync(0xffffffffu, bx, 2), qy = __shfl_xor_sync(0xffffffffu, by, 2), qz = __shfl_xor_sync(0xffffffffu, bz, 2);
    float lx = (part & 2) ? qx : bx, ly = (part & 2) ? qy : by, lz = (part & 2) ? qz : bz;
    float gx = (part & 2) ? bx : qx, gy = (part & 2) ? by : qy, gz = (part & 2) ? bz : qz;
    if (!active || part != 0) continue;              // lane 0 of the quad finishes the atom
    if (a.p.clip_local >= 0.f) clip3(lx, ly, lz, a.p.clip_local);
    if (a.p.use_global) clip3(gx, gy, gz, a.p.clip);
    else gx = gy = gz = 0.f;
    const float ex = lx + gx * a.p.w_global, ey = ly + gy * a.p.w_global, ez = lz + gz * a.p.w_global;
    float z0, z1, z2;
    if (a.p.noise) {
      const float* nz_ = a.p.noise + ((size_t)step * a.n_atoms + at) * 3;
      z0 = nz_[0]; z1 = nz_[1]; z2 = nz_[2];
    } else {
      normal3(a.p.seed, a.mol_gid[m], i, step + a.p.step_offset, z0, z1, z2);
    }
    const float x1 = px + (step_size * ex) / sigma + z0 * nscale;
    const float y1 = py + (step_size * ey) / sigma + z1 * nscale;
    const float z1n = pz + (step_size * ez) / sigma + z2 * nscale;
    bad = bad || isnan(x1) || isnan(y1) || isnan(z1n);
    nx[q] = x1; ny[q] = y1; nz[q] = z1n;
    sx += x1; sy += y1; sz += z1n;
  }
  if (bad) atomicMin(&a.counters[2], step);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = sx; red[1][tid >> 5] = sy; red[2][tid >> 5] = sz;
  }
  const int any_bad = __syncthreads_or(bad ? 1 : 0);   // also: every thread has finished reading the old positions
  if (any_bad && tid == 0) atomicMin(&a.nan_mol[m], step);
  float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
  for (int w = 0; w < STEP_THREADS / 32; ++w) { cx += red[0][w]; cy += red[1][w]; cz += red[2][w]; }
  const float cnt = (float)(n > 0 ? n : 1);
  cx /= cnt; cy /= cnt; cz /= cnt;
  if (part != 0) return;
#pragma unroll
  for (int q = 0; q < STEP_MAX_PER_THREAD; ++q) {
    const int i = slot + q * STEP_ATOMS;
    if (i >= n) continue;
    const int at = a0 + i;
    float x = nx[q] - cx, y = ny[q] - cy, z = nz[q] - cz;
    if (a.p.clip_pos >= 0.f) {
      x = fminf(fmaxf(x, -a.p.clip_pos), a.p.clip_pos);
      y = fminf(fmaxf(y, -a.p.clip_pos), a.p.clip_pos);
      z = fminf(fmaxf(z, -a.p.clip_pos), a.p.clip_pos);
    }
    a.pos[3 * (size_t)at] = x; a.pos[3 * (size_t)at + 1] = y; a.pos[3 * (size_t)at + 2] = z;
    if (a.p.traj) {
      float* t = a.p.traj + ((size_t)step * a.n_atoms + at) * 3;
      t[0] = x; t[1] = y; t[2] = z;
    }
  }
}

__global__ void advance_step_kernel(int* counters) { counters[1] += 1; }

void launch_step(const LaunchCtx& c, const BatchDev& b, float* pos, const StepParams& p) {
  StepArgs a{};
  a.pos = pos;
  a.mol_ptr = b.mol_ptr;
  a.mol_gid = b.mol_gid;
  a.n_atoms = b.n_atoms;
  a.in_ptr = b.in_ptr; a.e_src = b.e_src; a.e_type = b.e_type; a.e_len = b.e_len; a.s_csc = b.s_csc;
  a.out_ptr = b.out_ptr; a.c_dst = b.c_dst; a.c_type = b.c_type; a.c_len = b.c_len; a.s_canon = b.s_canon;
  a.lin_ptr = b.lc_in_ptr; a.lsrc = b.lc_src; a.llen = b.lc_len; a.sl_csc = b.sl_csc;
  a.lout_ptr = b.lc_out_ptr; a.lcdst = b.lc_cdst; a.lclen = b.lcc_len; a.sl_canon = b.sl_canon;
  a.sched = b.sched;
  a.counters = b.counters;
  a.nan_mol = b.nan_mol;
  a.p = p;
  if (b.mw <= 8) langevin_step_kernel<256 / STEP_ATOMS><<<b.n_mols, STEP_THREADS, 0, c.stream>>>(a);
  else langevin_step_kernel<AGD_MAX_MOL_ATOMS / STEP_ATOMS><<<b.n_mols, STEP_THREADS, 0, c.stream>>>(a);
  note_launch(c, "step.langevin");
}

void launch_advance(const LaunchCtx& c, const BatchDev& b) {
  advance_step_kernel<<<1, 1, 0, c.stream>>>(b.counters);
  note_launch(c, "step.advance");
}

// stand-alone eq_transform over an arbitrary edge list (microbenchmark iii / module parity):
// out[row] += dd*s, out[col] -= dd*s with fp32 atomics, like torch_scatter does.
__global__ void eq_transform_kernel(const float* __restrict__ score, const float* __restrict__ pos,
                                    const int* __restrict__ src, const int* __restrict__ dst,
                                    const float* __restrict__ len, int64_t n_edges, float* __restrict__ out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = src[e], c = dst[e];
    const float inv = 1.0f / len[e], s = score[e];
    const float dx = (inv * (pos[3 * (size_t)r] - pos[3 * (size_t)c])) * s;
    const float dy = (inv * (pos[3 * (size_t)r + 1] - pos[3 * (size_t)c + 1])) * s;
    const float dz = (inv * (pos[3 * (size_t)r + 2] - pos[3 * (size_t)c + 2])) * s;
    atomicAdd(out + 3 * (size_t)r, dx); atomicAdd(out + 3 * (size_t)r + 1, dy); atomicAdd(out + 3 * (size_t)r + 2, dz);
    atomicAdd(out + 3 * (size_t)c, -dx); atomicAdd(out + 3 * (size_t)c + 1, -dy); atomicAdd(out + 3 * (size_t)c + 2, -dz);
  }
}

void launch_eq_transform(cudaStream_t s, const float* score, const float* pos, const int* src, const int* dst,
                         const float* len, int64_t n_edges, int n_nodes, float* out) {
  cudaMemsetAsync(out, 0, sizeof(float) * 3 * (size_t)n_nodes, s);
  if (n_edges <= 0) return;
  int64_t blocks = (n_edges + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  eq_transform_kernel<<<(int)blocks, 256, 0, s>>>(score, pos, src, dst, len, n_edges, out);
}

}  // namespace agd
