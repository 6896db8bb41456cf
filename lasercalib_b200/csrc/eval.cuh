// Observation-major evaluation kernels: PySBA.rotate / project / fun (pySBA.py:61-101),
// the analytic Jacobian blocks that replace scipy's 3-point finite difference
// (scipy/optimize/_numdiff.py:770) and the sparsity pattern (pySBA.py:103-118).
#pragma once
#include "common.cuh"

namespace lcba {

// ---- camera tables -----------------------------------------------------------------
__global__ void k_cam_tables(const double* __restrict__ cams, double* __restrict__ tab, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) cam_table_build(cams + (size_t)c * NCP, tab + (size_t)c * CAMTAB);
}

__device__ __forceinline__ void load_tables_smem(const double* __restrict__ tab, double* s_tab,
                                                 int C) {
  for (int i = threadIdx.x; i < C * CAMTAB; i += blockDim.x) s_tab[i] = tab[i];
}

// ---- row-wise API kernels (arbitrary gathered rows; not the solver's hot loop) -----
__global__ void k_rotate_rows(const double* __restrict__ pts, const double* __restrict__ rv,
                              double* __restrict__ out, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double ox, oy, oz;
  rotate_row(rv[3 * i], rv[3 * i + 1], rv[3 * i + 2], pts[3 * i], pts[3 * i + 1], pts[3 * i + 2],
             ox, oy, oz);
  out[3 * i] = ox;
  out[3 * i + 1] = oy;
  out[3 * i + 2] = oz;
}

__global__ void k_project_rows(const double* __restrict__ pts, const double* __restrict__ cams,
                               double* __restrict__ out, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double* cm = cams + i * NCP;
  double x, y, z;
  rotate_row(cm[0], cm[1], cm[2], pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y, z);
  x += cm[3];
  y += cm[4];
  z += cm[5];
  x /= z;
  y /= z;
  const double n = x * x + y * y;
  const double r = 1.0 + cm[7] * n + cm[8] * n * n;   // pySBA.py:86
  const double rf = r * cm[6];
  out[2 * i] = x * rf + cm[9];
  out[2 * i + 1] = y * rf + cm[10];
}

// ---- fun: residual (+ cost) over the resident, point-major observation stream ------
// r_out may be null. perm (original index of sorted observation i) may be null (identity).
// Block partial of sum(r^2) goes to part[blockIdx.x].
__global__ void __launch_bounds__(256)
k_residual(const double* __restrict__ tab, const double* __restrict__ pts,
           const double2* __restrict__ uv, const uint8_t* __restrict__ cam,
           const int32_t* __restrict__ pt, const double* __restrict__ wgt,
           const int32_t* __restrict__ perm, long long N, int C, double2* __restrict__ r_out,
           double* __restrict__ part) {
  extern __shared__ double s_dyn[];
  double* s_tab = s_dyn;
  __shared__ double s_red[32];
  load_tables_smem(tab, s_tab, C);
  __syncthreads();
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int c = cam[i];
    const long long p = pt[i];
    const double2 o = uv[i];
    const double w = wgt ? wgt[i] : 1.0;
    double pu, pv;
    project_tab(s_tab + c * CAMTAB, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], pu, pv);
    const double ru = w * (pu - o.x), rv = w * (pv - o.y);
    acc = fma(ru, ru, fma(rv, rv, acc));
    if (r_out) {
      const long long dst = perm ? (long long)perm[i] : i;
      r_out[dst] = make_double2(ru, rv);
    }
  }
  const double s = block_sum(acc, s_red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// ---- residual + Jacobian blocks materialised to HBM (metric M1) --------------------
// One thread per observation; each warp stages its 32 x (2x11 | 2x3) blocks in shared memory in
// the OUTPUT layout (conflict-free 128-bit stores: lane strides of 176 B and 48 B) and lane 0
// sends them out as two TMA bulk stores (cp.async.bulk shared -> global: 5632 B + 1536 B,
// contiguous in Jc / Jp).  264 B/obs algorithmic.  Unsorted input (perm) scatters per row instead.
constexpr int JB_THREADS = 256;
constexpr int JB_STAGE = 22 + 6;   // doubles per observation in the staging tile
constexpr int JB_LD = 32;           // observations per warp tile

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned src = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(JB_THREADS)
k_jacobian_blocks(const double* __restrict__ tab, const double* __restrict__ pts,
                  const double2* __restrict__ uv, const uint8_t* __restrict__ cam,
                  const int32_t* __restrict__ pt, const double* __restrict__ wgt,
                  const int32_t* __restrict__ perm, long long N, int C,
                  double2* __restrict__ r_out, double* __restrict__ Jc_out,
                  double* __restrict__ Jp_out) {
  extern __shared__ __align__(16) double s_dyn[];
  double* s_tab = s_dyn;                                         // C*CAMTAB
  double* s_stage = s_dyn + ((C * CAMTAB + 1) & ~1);             // warps * JB_LD * JB_STAGE
  load_tables_smem(tab, s_tab, C);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* sc = s_stage + (size_t)wid * JB_LD * JB_STAGE;         // [32][22]
  double* sp = sc + JB_LD * 22;                                  // [32][6]
  const long long nwarps_total = (long long)gridDim.x * (JB_THREADS / 32);
  const long long ngroups = (N + 31) / 32;
  for (long long g = (long long)blockIdx.x * (JB_THREADS / 32) + wid; g < ngroups;
       g += nwarps_total) {
    const long long i = g * 32 + lane;
    const bool active = i < N;
    if (active) {
      const int c = cam[i];
      const long long p = pt[i];
      const double2 o = uv[i];
      const double w = wgt ? wgt[i] : 1.0;
      ObsLin L;
      obs_linearize<true>(s_tab + c * CAMTAB, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], o.x, o.y,
                          w, L);
      if (r_out) r_out[perm ? (long long)perm[i] : i] = make_double2(L.ru, L.rv);
      double2* jc = reinterpret_cast<double2*>(sc + lane * 22);
      jc[0] = make_double2(L.Jc[0][0], L.Jc[0][1]);
      jc[1] = make_double2(L.Jc[0][2], L.Jc[0][3]);
      jc[2] = make_double2(L.Jc[0][4], L.Jc[0][5]);
      jc[3] = make_double2(L.Jc[0][6], L.Jc[0][7]);
      jc[4] = make_double2(L.Jc[0][8], w);
      jc[5] = make_double2(0.0, L.Jc[1][0]);
      jc[6] = make_double2(L.Jc[1][1], L.Jc[1][2]);
      jc[7] = make_double2(L.Jc[1][3], L.Jc[1][4]);
      jc[8] = make_double2(L.Jc[1][5], L.Jc[1][6]);
      jc[9] = make_double2(L.Jc[1][7], L.Jc[1][8]);
      jc[10] = make_double2(0.0, w);
      double2* jp = reinterpret_cast<double2*>(sp + lane * 6);
      jp[0] = make_double2(L.Jp[0][0], L.Jp[0][1]);
      jp[1] = make_double2(L.Jp[0][2], L.Jp[1][0]);
      jp[2] = make_double2(L.Jp[1][1], L.Jp[1][2]);
    }
    const long long base = g * 32;
    const int nvalid = (int)min((long long)32, N - base);
    if (!perm) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        bulk_store(Jc_out + base * 22, sc, (unsigned)(nvalid * 22 * 8));
        bulk_store(Jp_out + base * 6, sp, (unsigned)(nvalid * 6 * 8));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // tile may be rewritten
      }
    } else {
      __syncwarp();
      for (int e = lane; e < nvalid * 22; e += 32)
        Jc_out[(long long)perm[base + e / 22] * 22 + (e % 22)] = sc[e];
      for (int e = lane; e < nvalid * 6; e += 32)
        Jp_out[(long long)perm[base + e / 6] * 6 + (e % 6)] = sp[e];
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- 3-D initialisation: Unproject (lasercalib/rigid_body.py:205-243) --------------------
// OpenCV's iterative undistortPoints (5 fixed-point iterations, the default TermCriteria) for
// the (k1, k2, p1, p2, k3) model, then ray / plane z = Z intersection in world coordinates.
// par: K row-major (9), dist (5), R row-major (9), t (3).  Z has nZ = 1 or M entries.
__global__ void k_unproject(const double2* __restrict__ uv, const double* __restrict__ Z, int nZ,
                            const double* __restrict__ par, long long M,
                            double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double fx = par[0], fy = par[4], cx = par[2], cy = par[5];
  const double k1 = par[9], k2 = par[10], p1 = par[11], p2 = par[12], k3 = par[13];
  const double* R = par + 14;
  const double* t = par + 23;
  const double2 p = uv[i];
  const double x0 = (p.x - cx) / fx, y0 = (p.y - cy) / fy;
  double x = x0, y = y0;
  for (int it = 0; it < 5; ++it) {
    const double r2 = x * x + y * y;
    const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
    if (icdist < 0.0) { x = x0; y = y0; break; }
    const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
    const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  // the reference maps back through P = K and normalises again: (x*fx + cx - cx)/fx
  const double u = ((x * fx + cx) - cx) / fx, v = ((y * fy + cy) - cy) / fy;
  // left = R^T [u v 1],  right0 = R^T t
  const double l2 = R[2] * u + R[5] * v + R[8];
  const double r02 = R[2] * t[0] + R[5] * t[1] + R[8] * t[2];
  const double zw = Z[nZ == 1 ? 0 : i];
  const double zc = (zw + r02) / l2;
  const double a0 = u * zc - t[0], a1 = v * zc - t[1], a2 = zc - t[2];
  out[3 * i] = R[0] * a0 + R[3] * a1 + R[6] * a2;
  out[3 * i + 1] = R[1] * a0 + R[4] * a1 + R[7] * a2;
  out[3 * i + 2] = R[2] * a0 + R[5] * a1 + R[8] * a2;
}

// ---- bundle_adjustment_sparsity (pySBA.py:103-118) ---------------------------------
// 14 sorted column indices per row, rows 2i and 2i+1 identical.  One thread per entry.
__global__ void k_sparsity_indices(const long long* __restrict__ cam_idx,
                                   const long long* __restrict__ pt_idx, long long N, int C,
                                   int32_t* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * 28) return;
  const long long i = e / 28;
  const int j = (int)(e % 14);
  out[e] = (j < NCP) ? (int32_t)(cam_idx[i] * NCP + j)
                     : (int32_t)((long long)C * NCP + pt_idx[i] * 3 + (j - NCP));
}

}  // namespace lcba
