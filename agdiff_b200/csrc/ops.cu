// Stand-alone versions of the two gather -> segment-sum patterns that only exist inside larger kernels on the product path,
// exported for the roofline microbenchmarks (BASELINE.json config 4) and module-level parity tests:
//   gin_message_kernel             the GIN aggregation of tc_gin_kernel's gather phase (tc_node.cu)        gin.py:76-96
//   eq_transform_segments_kernel   the deterministic in-/out-segment form of eq_transform used by          geometry.py:9-17
//                                  langevin_step_kernel (step.cu): no atomics, sums in sorted-segment order
#include "common.cuh"
#include "kernels.h"

namespace agd {

// out_i = (1 + eps) * x_i + sum_{e in in(i)} relu(x[src_e] + ea_e); one warp per node, lanes across the 128 columns (a row = one
// coalesced 512 B request), 8 edges in flight - the same loop, in the same summation order, as the gather of tc_gin_kernel.
__global__ void __launch_bounds__(256) gin_message_kernel(const float* __restrict__ x, const float* __restrict__ ea,
                                                          const int* __restrict__ ea_idx,
                                                          const int* __restrict__ src, const int* __restrict__ in_ptr, int n_nodes,
                                                          float ope_host, const float* __restrict__ ope_dev, float* __restrict__ out) {
  const float ope = ope_dev ? __ldg(ope_dev) : ope_host;   // 1 + eps: a model weight on the product path
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; node < n_nodes; node += warps) {
    const float4 self = __ldg(reinterpret_cast<const float4*>(x + (size_t)node * HID) + lane);
    const int e0 = __ldg(in_ptr + node), e1 = __ldg(in_ptr + node + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; e += 8) {
      const int n = (e1 - e < 8) ? e1 - e : 8;
      float4 xv[8], ev[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < n) {
          xv[u] = __ldg(reinterpret_cast<const float4*>(x + (size_t)__ldg(src + e + u) * HID) + lane);
          ev[u] = ea_idx ? __ldg(reinterpret_cast<const float4*>(ea + (size_t)__ldg(ea_idx + e + u) * HID) + lane)
                         : __ldcs(reinterpret_cast<const float4*>(ea + (size_t)(e + u) * HID) + lane);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (u < n) {
          acc.x += relu_(xv[u].x + ev[u].x); acc.y += relu_(xv[u].y + ev[u].y);
          acc.z += relu_(xv[u].z + ev[u].z); acc.w += relu_(xv[u].w + ev[u].w);
        }
      }
    }
    acc.x = fmaf(ope, self.x, acc.x); acc.y = fmaf(ope, self.y, acc.y);
    acc.z = fmaf(ope, self.z, acc.z); acc.w = fmaf(ope, self.w, acc.w);
    reinterpret_cast<float4*>(out + (size_t)node * HID)[lane] = acc;
  }
}

void launch_gin_message(cudaStream_t s, const float* x, const float* ea, const int* ea_idx, const int* src, const int* in_ptr, int n_nodes, float eps,
                        const float* one_plus_eps_dev, float* out) {
  if (n_nodes <= 0) return;
  int64_t blocks = ((int64_t)n_nodes + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gin_message_kernel<<<(int)blocks, 256, 0, s>>>(x, ea, ea_idx, src, in_ptr, n_nodes, 1.0f + eps, one_plus_eps_dev, out);
}

// out_i = sum_{e: row_e = i} dd_e s_e - sum_{e: col_e = i} dd_e s_e with dd_e = (pos[row_e] - pos[col_e]) / |.|: two threads per
// atom, one walking the atom's out-segment (edges sorted by row; the other end is col), one its in-segment (edges sorted by
// col; the other end is row), combined by a shuffle - the per-atom part of langevin_step_kernel without the update.
__global__ void __launch_bounds__(256) eq_transform_segments_kernel(const float* __restrict__ pos, const float* __restrict__ s_out,
                                                                    const int* __restrict__ col_of_out, const int* __restrict__ out_ptr,
                                                                    const float* __restrict__ s_in, const int* __restrict__ row_of_in,
                                                                    const int* __restrict__ in_ptr, int n_nodes, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 1, part = t & 1;
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (i < n_nodes) {
    const float px = pos[3 * (size_t)i], py = pos[3 * (size_t)i + 1], pz = pos[3 * (size_t)i + 2];
    const int* ptr = part ? in_ptr : out_ptr;
    const int* other = part ? row_of_in : col_of_out;
    const float* sc = part ? s_in : s_out;
    const int e0 = ptr[i], e1 = ptr[i + 1];
    for (int e = e0; e < e1; ++e) {
      const int j = other[e];
      // dd = (pos[row] - pos[col]) / len: row = i on the out-segment, row = j on the in-segment
      float dx = px - pos[3 * (size_t)j], dy = py - pos[3 * (size_t)j + 1], dz = pz - pos[3 * (size_t)j + 2];
      const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
      const float w = sc[e] * inv;   // the in-segment contributes -(pos[j] - pos[i]) * s = (pos[i] - pos[j]) * s: same sign
      ax = fmaf(dx, w, ax); ay = fmaf(dy, w, ay); az = fmaf(dz, w, az);
    }
  }
  ax += __shfl_xor_sync(0xffffffffu, ax, 1);
  ay += __shfl_xor_sync(0xffffffffu, ay, 1);
  az += __shfl_xor_sync(0xffffffffu, az, 1);
  if (i < n_nodes && part == 0) {
    out[3 * (size_t)i] = ax; out[3 * (size_t)i + 1] = ay; out[3 * (size_t)i + 2] = az;
  }
}

void launch_eq_transform_segments(cudaStream_t s, const float* pos, const float* s_out, const int* col_of_out, const int* out_ptr,
                                  const float* s_in, const int* row_of_in, const int* in_ptr, int n_nodes, float* out) {
  if (n_nodes <= 0) return;
  const int blocks = (int)(((int64_t)n_nodes * 2 + 255) / 256);
  eq_transform_segments_kernel<<<blocks, 256, 0, s>>>(pos, s_out, col_of_out, out_ptr, s_in, row_of_in, in_ptr, n_nodes, out);
}

// Pair mode of the local branch (api.cu: run_local_branch): the edge encoder and the pair MLP ran once per undirected pair; this
// hands their per-pair results (edge length, score) to every directed local edge, in CSC order and in canonical order, which is
// what the step kernel and forward's edge_inv_local read.
__global__ void local_pairs_expand_kernel(const int* __restrict__ pair_of, const int* __restrict__ canon, int n_local,
                                          const float* __restrict__ len_pair, const float* __restrict__ s_pair,
                                          float* __restrict__ len_csc, float* __restrict__ len_canon, float* __restrict__ s_csc,
                                          float* __restrict__ s_canon) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_local; e += gridDim.x * blockDim.x) {
    const int p = pair_of[e], cp = canon[e];
    const float l = len_pair[p], sc = s_pair[p];
    len_csc[e] = l;
    len_canon[cp] = l;
    s_csc[e] = sc;
    s_canon[cp] = sc;
  }
}

void launch_local_pairs_expand(const LaunchCtx& c, const BatchDev& b) {
  if (b.n_local <= 0) return;
  int64_t blocks = ((int64_t)b.n_local + 255) / 256;
  if (blocks > c.num_sms * 8) blocks = c.num_sms * 8;
  local_pairs_expand_kernel<<<(int)blocks, 256, 0, c.stream>>>(b.lp_of, b.lc_canon, b.n_local, b.lp_len, b.lp_s, b.lc_len, b.lcc_len,
                                                               b.sl_csc, b.sl_canon);
  note_launch(c, "local.expand_pairs");
}

// rows of a per-pair tensor in local-edge order (debug fetch of edge_attr in pair mode)
__global__ void gather_rows128_kernel(const float* __restrict__ src, const int* __restrict__ idx, int n, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps)
    reinterpret_cast<float4*>(dst + (size_t)r * HID)[lane] = __ldg(reinterpret_cast<const float4*>(src + (size_t)idx[r] * HID) + lane);
}

void launch_gather_rows128(cudaStream_t s, const float* src, const int* idx, int n, float* dst) {
  if (n <= 0) return;
  int64_t blocks = ((int64_t)n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gather_rows128_kernel<<<(int)blocks, 256, 0, s>>>(src, idx, n, dst);
}

// ---------------------------------------------------------------------------------------------------------------- COV / MAT
// RMSD after optimal superposition (proper rotations only, like RDKit's alignment behind get_best_rmsd, utils/chem.py) for every
// (reference conformer, generated conformer) pair of one molecule: covmat.py:16-34.  One warp per pair; lanes stride over the
// selected atoms accumulating centroids, the 3 x 3 cross-covariance and the two inner products in fp64; lane 0 then gets the
// largest eigenvalue of Horn's 4 x 4 quaternion key matrix (Jacobi rotations, fp64): rmsd^2 = (Ga + Gb - 2 lambda_max) / n.  No symmetry permutations (RDKit tries the molecule's automorphisms): an upper bound of
// GetBestRMS, equal to it for molecules without non-trivial heavy-atom automorphisms.
__global__ void __launch_bounds__(256) kabsch_rmsd_kernel(const float* __restrict__ ref, const float* __restrict__ gen,
                                                          const int* __restrict__ sel, int n_sel, int n_atoms, int n_ref, int n_gen,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pair < n_ref * n_gen; pair += warps) {
    const int ir = pair / n_gen, ig = pair - ir * n_gen;
    const float* A = ref + (size_t)ir * n_atoms * 3;
    const float* B = gen + (size_t)ig * n_atoms * 3;
    double v[17];   // sum a (3), sum b (3), sum a_i b_j (9), sum |a|^2, sum |b|^2
#pragma unroll
    for (int k = 0; k < 17; ++k) v[k] = 0.0;
    for (int t = lane; t < n_sel; t += 32) {
      const int at = sel ? sel[t] : t;
      const double a[3] = {A[3 * at], A[3 * at + 1], A[3 * at + 2]};
      const double b[3] = {B[3 * at], B[3 * at + 1], B[3 * at + 2]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        v[i] += a[i];
        v[3 + i] += b[i];
        v[15] += a[i] * a[i];
        v[16] += b[i] * b[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) v[6 + 3 * i + j] += a[i] * b[j];
      }
    }
#pragma unroll
    for (int k = 0; k < 17; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) {
      const double n = (double)n_sel;
      double S[3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[i][j] = v[6 + 3 * i + j] - v[i] * v[3 + j] / n;   // centred cross-covariance
      const double Ga = v[15] - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / n;
      const double Gb = v[16] - (v[3] * v[3] + v[4] * v[4] + v[5] * v[5]) / n;
      const double Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2], Syx = S[1][0], Syy = S[1][1], Syz = S[1][2], Szx = S[2][0],
                   Szy = S[2][1], Szz = S[2][2];
      // Horn's key matrix; its largest eigenvalue by cyclic Jacobi rotations.  (Newton on the characteristic polynomial - the
      // usual QCP shortcut - loses half the digits when the top eigenvalue is degenerate, which is the case for every planar
      // molecule and every 3-atom selection, and the square root below turns 1e-8 into 1e-4 A.)
      double K[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                        {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                        {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                        {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
      for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 4; ++p) {
          diag += K[p][p] * K[p][p];
          for (int q = p + 1; q < 4; ++q) off += K[p][q] * K[p][q];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 3; ++p)
          for (int q = p + 1; q < 4; ++q) {
            const double apq = K[p][q];
            if (apq == 0.0) continue;
            const double theta = (K[q][q] - K[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            for (int k = 0; k < 4; ++k) {   // K <- J^T K J, columns then rows
              const double kp = K[k][p], kq = K[k][q];
              K[k][p] = c * kp - sn * kq;
              K[k][q] = sn * kp + c * kq;
            }
            for (int k = 0; k < 4; ++k) {
              const double pk = K[p][k], qk = K[q][k];
              K[p][k] = c * pk - sn * qk;
              K[q][k] = sn * pk + c * qk;
            }
          }
      }
      double l = K[0][0];
      for (int p = 1; p < 4; ++p) l = (K[p][p] > l || !(l == l)) ? K[p][p] : l;
      double r2 = (Ga + Gb - 2.0 * l) / n;
      if (!(r2 > 0.0)) r2 = (r2 != r2) ? r2 : 0.0;   // rounding below zero for identical conformers; NaN inputs stay NaN
      out[pair] = (float)sqrt(r2);
    }
  }
}

void launch_kabsch_rmsd(cudaStream_t s, const float* ref, const float* gen, const int* sel, int n_sel, int n_atoms, int n_ref, int n_gen,
                        float* out) {
  const int64_t pairs = (int64_t)n_ref * n_gen;
  if (pairs <= 0) return;
  int64_t blocks = (pairs + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  kabsch_rmsd_kernel<<<(int)blocks, 256, 0, s>>>(ref, gen, sel, n_sel, n_atoms, n_ref, n_gen, out);
}

}  // namespace agd
