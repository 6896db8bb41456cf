// Reduced camera system: per point eliminate the 3x3 block and accumulate
//   S_jk += [j==k] Jc_j^T Jc_j - Y_j Y_k^T ,  rhs_j -= Y_j z ,   Y_j = (Jc_j^T Jp_j) L^-T
// over all camera pairs (j >= k) that see the point (V + lam Dp^2 = L L^T, z = L^-1 g_p).
// This is the FP64-pipe-bound kernel of the iteration: ~363 DFMA
// per (point, camera pair).  J is never materialised; Y lives in shared memory only.
//
// Decomposition.  Cameras are grouped in duos; a half-warp owns one 2x2 duo block
// (4 camera pairs) with its 4 x 9 accumulators in registers: lane (rr,cc) of the 4x4
// half-warp grid holds rows 3rr..3rr+2 x cols 3cc..3cc+2 of every 11x11 (padded to 12)
// pair block.  A CTA ("kind") owns up to 32 duo blocks and streams a slice of the points:
//   phase 1: threads = (point, camera slot) -> Jacobian blocks -> Y (and Jc for the
//            diagonal pairs) into shared memory, K-major [slice][row-group][4];
//   phase 2: half-warps run the rank-3 (+ rank-2 on the diagonal) updates for the pairs
//            whose two cameras both see the point (visibility mask test, no divergence
//            inside a half-warp).
// Partials per point-slice are written without atomics and summed in a fixed order.
#pragma once
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace lcba {

constexpr int SCHUR_MAX_HW = 24;     // half-warps (duo blocks) per CTA
constexpr int YS_LD = 80;            // doubles per (point, slot): 5 K-slices x 4 row groups x 4

struct SchurHw {          // one duo block = 2x2 camera pairs
  int8_t s[4];            // shared-memory slots of cams j0, j1, k0, k1
  int8_t c[4];            // camera ids   j0, j1, k0, k1 (0 when absent; see valid)
  uint8_t valid;          // bit (2*ja + kb): pair (j_ja, k_kb) is a lower-triangle pair
  uint8_t diag;           // same bit layout: pair is a diagonal pair (same camera)
  uint8_t pad[6];
};

struct SchurKind {
  int nslots, nhw, hw_base, threads;
  uint8_t slot_cam[LCBA_MAX_CAMERAS];
};

struct SchurPlan {
  int C = 0, nkinds = 0, nslices = 0, pc = 0, max_threads = 0, max_slots = 0;
  int npairs = 0;
  size_t part_stride = 0;   // doubles per slice partial: npairs*121 + 11*C
  size_t smem_bytes = 0;
  std::vector<SchurKind> kinds;
  std::vector<SchurHw> hws;
};

inline int pair_index(int j, int k) { return j * (j + 1) / 2 + k; }

inline SchurPlan make_schur_plan(int C, int sm_count, size_t smem_limit) {
  SchurPlan pl;
  pl.C = C;
  const int nb = (C + 1) / 2;
  const int nblocks = nb * (nb + 1) / 2;
  pl.nkinds = (nblocks + SCHUR_MAX_HW - 1) / SCHUR_MAX_HW;
  const int bpk = (nblocks + pl.nkinds - 1) / pl.nkinds;
  pl.npairs = C * (C + 1) / 2;
  pl.part_stride = (size_t)pl.npairs * 121 + (size_t)NCP * C;
  int b = 0, bj = 0, bk = 0;
  for (int kd = 0; kd < pl.nkinds; ++kd) {
    SchurKind K{};
    K.hw_base = (int)pl.hws.size();
    bool used[LCBA_MAX_CAMERAS] = {false};
    std::vector<std::pair<int, int>> blks;
    for (int i = 0; i < bpk && b < nblocks; ++i, ++b) {
      blks.push_back({bj, bk});
      for (int d = 0; d < 2; ++d) {
        if (2 * bj + d < C) used[2 * bj + d] = true;
        if (2 * bk + d < C) used[2 * bk + d] = true;
      }
      if (++bk > bj) { bk = 0; ++bj; }
    }
    int slot_of[LCBA_MAX_CAMERAS];
    K.nslots = 0;
    for (int c = 0; c < C; ++c)
      if (used[c]) { slot_of[c] = K.nslots; K.slot_cam[K.nslots++] = (uint8_t)c; }
    for (auto& bl : blks) {
      SchurHw h{};
      const int cj[2] = {2 * bl.first, 2 * bl.first + 1};
      const int ck[2] = {2 * bl.second, 2 * bl.second + 1};
      for (int d = 0; d < 2; ++d) {
        h.c[d] = (int8_t)(cj[d] < C ? cj[d] : 0);
        h.s[d] = (int8_t)(cj[d] < C ? slot_of[cj[d]] : 0);
        h.c[2 + d] = (int8_t)(ck[d] < C ? ck[d] : 0);
        h.s[2 + d] = (int8_t)(ck[d] < C ? slot_of[ck[d]] : 0);
      }
      for (int ja = 0; ja < 2; ++ja)
        for (int kb = 0; kb < 2; ++kb)
          if (cj[ja] < C && ck[kb] < C && ck[kb] <= cj[ja]) {
            h.valid |= (uint8_t)(1u << (2 * ja + kb));
            if (ck[kb] == cj[ja]) h.diag |= (uint8_t)(1u << (2 * ja + kb));
          }
      pl.hws.push_back(h);
    }
    K.nhw = (int)blks.size();
    K.threads = std::max(64, 32 * ((K.nhw + 1) / 2));
    pl.max_threads = std::max(pl.max_threads, K.threads);
    pl.max_slots = std::max(pl.max_slots, K.nslots);
    pl.kinds.push_back(K);
  }
  pl.nslices = std::max(1, sm_count / pl.nkinds);
  const size_t fixed = (size_t)C * CAMTAB * 8 + 64;
  const size_t per_pt = (size_t)pl.max_slots * YS_LD * 8 + 4 * 8 + 8;
  long pc = (long)((smem_limit - fixed) / per_pt);
  pl.pc = (int)std::max(1L, std::min(32L, pc));
  pl.smem_bytes = fixed + per_pt * pl.pc;
  return pl;
}

__device__ __forceinline__ void ld3(const double* __restrict__ p, double (&v)[3]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = p[2];
}

template <bool SUB>
__device__ __forceinline__ void outer9(double (&acc)[9], const double (&a)[3], const double (&b)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      acc[3 * i + j] = SUB ? fma(-a[i], b[j], acc[3 * i + j]) : fma(a[i], b[j], acc[3 * i + j]);
}

__global__ void __maxnreg__(168)
k_schur(const double* __restrict__ tab, const double* __restrict__ pts,
        const double* __restrict__ wgt, const uint32_t* __restrict__ obs_start,
        const unsigned long long* __restrict__ mask, const double* __restrict__ Lz, long long P,
        long long N, int C, const SchurKind* __restrict__ kinds, const SchurHw* __restrict__ hws,
        int PC, int nslices, size_t part_stride, int npairs, double* __restrict__ part) {
  extern __shared__ __align__(16) double s_dyn[];
  const SchurKind& K = kinds[blockIdx.y];
  const int nthreads = K.threads;
  const int tid = threadIdx.x;
  if (tid >= nthreads) return;          // uniform per warp (threads multiple of 32)
  const int nslots = K.nslots;
  double* s_Y = s_dyn;                                        // PC * nslots * YS_LD (16B aligned)
  double* s_z = s_Y + (size_t)PC * nslots * YS_LD;            // PC * 4
  unsigned long long* s_mask = reinterpret_cast<unsigned long long*>(s_z + PC * 4);   // PC
  double* s_tab = reinterpret_cast<double*>(s_mask + PC);     // C * CAMTAB
  for (int i = tid; i < C * CAMTAB; i += nthreads) s_tab[i] = tab[i];

  const int hw = tid >> 4, l16 = tid & 15, rr = l16 >> 2, cc = l16 & 3;
  SchurHw D{};
  if (hw < K.nhw) D = hws[K.hw_base + hw];
  const unsigned valid = D.valid, diag = D.diag;

  double acc[4][9];
  double rh[2][3];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[i][j] = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) rh[i][j] = 0.0;

  // slice boundaries balanced by observation count
  const int slice = blockIdx.x;
  long long pa, pb;
  {
    const unsigned long long ta = (unsigned long long)N * slice / nslices;
    const unsigned long long tb = (unsigned long long)N * (slice + 1) / nslices;
    pa = (slice == 0) ? 0 : lower_bound_u32(obs_start, P, ta);
    pb = (slice == nslices - 1) ? P : lower_bound_u32(obs_start, P, tb);
  }
  // named barrier over the active threads only
  auto bar = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); };
  bar();

  for (long long q0 = pa; q0 < pb; q0 += PC) {
    const int npc = (int)min((long long)PC, pb - q0);
    // ---------------- phase 1: Y (and Jc) for every (point, camera slot) ----------------
    for (int idx = tid; idx < npc * nslots; idx += nthreads) {
      const int q = idx / nslots, s = idx - q * nslots;
      const long long p = q0 + q;
      const unsigned long long m = mask[p];
      const int c = K.slot_cam[s];
      if ((m >> c) & 1ull) {
        const long long o = (long long)obs_start[p] + __popcll(m & ((1ull << c) - 1ull));
        const double w = wgt ? wgt[o] : 1.0;
        ObsLin L;
        obs_linearize<false>(s_tab + c * CAMTAB, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], 0.0,
                             0.0, w, L);
        const double* li = Lz + p * 9;
        double Q[2][3];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          Q[i][0] = L.Jp[i][0] * li[0];
          Q[i][1] = fma(L.Jp[i][0], li[1], L.Jp[i][1] * li[2]);
          Q[i][2] = fma(L.Jp[i][0], li[3], fma(L.Jp[i][1], li[4], L.Jp[i][2] * li[5]));
        }
        double* Y = s_Y + ((size_t)q * nslots + s) * YS_LD;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
          for (int a = 0; a < 9; ++a)
            Y[k * 16 + (a / 3) * 4 + (a % 3)] = fma(L.Jc[0][a], Q[0][k], L.Jc[1][a] * Q[1][k]);
          Y[k * 16 + 12] = w * Q[0][k];
          Y[k * 16 + 13] = w * Q[1][k];
          Y[k * 16 + 14] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
          for (int a = 0; a < 9; ++a) Y[(3 + i) * 16 + (a / 3) * 4 + (a % 3)] = L.Jc[i][a];
          Y[(3 + i) * 16 + 12] = (i == 0) ? w : 0.0;
          Y[(3 + i) * 16 + 13] = (i == 1) ? w : 0.0;
          Y[(3 + i) * 16 + 14] = 0.0;
        }
      }
    }
    for (int q = tid; q < npc; q += nthreads) {
      s_mask[q] = mask[q0 + q];
      const double* li = Lz + (q0 + q) * 9;
      s_z[q * 4] = li[6]; s_z[q * 4 + 1] = li[7]; s_z[q * 4 + 2] = li[8];
    }
    bar();
    // ---------------- phase 2: rank-3 updates per visible camera pair -------------------
    if (valid) {
      for (int q = 0; q < npc; ++q) {
        const unsigned long long m = s_mask[q];
        const unsigned vj0 = (unsigned)(m >> D.c[0]) & 1u, vj1 = (unsigned)(m >> D.c[1]) & 1u;
        const unsigned vk0 = (unsigned)(m >> D.c[2]) & 1u, vk1 = (unsigned)(m >> D.c[3]) & 1u;
        const unsigned pm = valid & ((vj0 & vk0) | ((vj0 & vk1) << 1) | ((vj1 & vk0) << 2) |
                                     ((vj1 & vk1) << 3));
        if (!pm) continue;
        const double* Yq = s_Y + (size_t)q * nslots * YS_LD;
        const double* Yj0 = Yq + D.s[0] * YS_LD + rr * 4;
        const double* Yj1 = Yq + D.s[1] * YS_LD + rr * 4;
        const double* Yk0 = Yq + D.s[2] * YS_LD + cc * 4;
        const double* Yk1 = Yq + D.s[3] * YS_LD + cc * 4;
        const double* z = s_z + q * 4;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double a0[3], a1[3], b0[3], b1[3];
          if (pm & 3u) ld3(Yj0 + k * 16, a0);
          if (pm & 12u) ld3(Yj1 + k * 16, a1);
          if (pm & 5u) ld3(Yk0 + k * 16, b0);
          if (pm & 10u) ld3(Yk1 + k * 16, b1);
          if (pm & 1u) outer9<true>(acc[0], a0, b0);
          if (pm & 2u) outer9<true>(acc[1], a0, b1);
          if (pm & 4u) outer9<true>(acc[2], a1, b0);
          if (pm & 8u) outer9<true>(acc[3], a1, b1);
          if (cc == 0) {
            const double zk = z[k];
            if (pm & diag & 1u) {
#pragma unroll
              for (int i = 0; i < 3; ++i) rh[0][i] = fma(-a0[i], zk, rh[0][i]);
            }
            if (pm & diag & 8u) {
#pragma unroll
              for (int i = 0; i < 3; ++i) rh[1][i] = fma(-a1[i], zk, rh[1][i]);
            }
          }
        }
        if (pm & diag) {   // U_j = sum Jc^T Jc on the diagonal pairs (K-slices 3,4)
#pragma unroll
          for (int k = 3; k < 5; ++k) {
            double a[3], b[3];
            if (pm & diag & 1u) { ld3(Yj0 + k * 16, a); ld3(Yk0 + k * 16, b); outer9<false>(acc[0], a, b); }
            if (pm & diag & 8u) { ld3(Yj1 + k * 16, a); ld3(Yk1 + k * 16, b); outer9<false>(acc[3], a, b); }
          }
        }
      }
    }
    bar();
  }
  // ---------------- write the slice partial (every lower-triangle entry exactly once) ----
  if (valid) {
    double* out = part + (size_t)slice * part_stride;
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      if (!((valid >> pr) & 1u)) continue;
      const int cj = D.c[pr >> 1], ck = D.c[2 + (pr & 1)];
      double* blk = out + (size_t)(cj * (cj + 1) / 2 + ck) * 121;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int a = 3 * rr + i, b = 3 * cc + j;
          if (a < NCP && b < NCP) blk[a * NCP + b] = acc[pr][3 * i + j];
        }
    }
    if (cc == 0) {
      double* rout = out + (size_t)npairs * 121;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        if (!((diag >> (3 * d)) & 1u)) continue;
        const int cj = D.c[d];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (3 * rr + i < NCP) rout[cj * NCP + 3 * rr + i] = rh[d][i];
      }
    }
  }
}

// S (n x n, both triangles) = reduced blocks + lam * diag(scl_c^2) ; rhs = g_c + rhs_red
__global__ void k_assemble_S(const double* __restrict__ red, int C, int npairs,
                             const double* __restrict__ camsum, const double* __restrict__ scl_c,
                             double lam, double* __restrict__ S, double* __restrict__ rhs) {
  const int n = C * NCP;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * n) return;
  const int r = (int)(idx / n), c = (int)(idx % n);
  const int hi = max(r, c), lo = min(r, c);
  const int j = hi / NCP, a = hi % NCP, k = lo / NCP, b = lo % NCP;
  double v;
  if (j == k) {
    // diagonal block is stored in full (both halves computed from the same products)
    v = red[(size_t)(j * (j + 1) / 2 + j) * 121 + (r % NCP) * NCP + (c % NCP)];
  } else {
    v = red[(size_t)(j * (j + 1) / 2 + k) * 121 + a * NCP + b];
  }
  if (r == c) {
    const double s = scl_c[r];
    v = fma(lam * s, s, v);
    rhs[r] = camsum[(r / NCP) * 22 + (r % NCP)] + red[(size_t)npairs * 121 + r];
  }
  S[idx] = v;
}

}  // namespace lcba
