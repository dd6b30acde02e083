// k_validity: the reference's post-sampling validity filter, on the device.
//   chemeleon/scripts/evaluate.py:177-189 (test_valid): max(lattice.abc) > 60 A  -> invalid;
//       min over the positive entries of the periodic distance matrix < 0.5 A      -> invalid
//   chemeleon/scripts/sample_target_composition.py:57-62: max(abc) > 60 -> skip; reduced
//       composition != target -> skip
// One block per crystal.  The distance matrix of the reference is pymatgen's minimum-image
// distance (Structure.distance_matrix); here every pair (i < j) is scanned over the images
// m in [-R_k, R_k]^3 after wrapping the fractional difference to [-1/2, 1/2), with
// R_k = max(1, ceil(thr / h_k + 1/2)) and h_k the spacing of the lattice planes normal to
// reciprocal axis k: |(d + m) L| >= |d_k + m_k| h_k, so no image outside that box can be closer
// than thr -- the threshold decision is exact for any cell shape, and the reported minimum is the
// true minimum-image distance whenever it is attained within the box (always for reduced cells).
// Exactly coincident atoms (distance 0) are skipped, as `dist_mat[dist_mat > 0]` does.
#include "cb2_internal.cuh"

namespace cb2 {

constexpr int VF_THREADS = 128;
constexpr int VF_RMAX = 6;

__global__ void __launch_bounds__(VF_THREADS) k_validity(const int64_t *__restrict__ a, const float *__restrict__ x,
                                                         const float *__restrict__ lat,
                                                         const int32_t *__restrict__ graph_off,
                                                         const int32_t *__restrict__ target, float max_len,
                                                         float thr, int32_t *__restrict__ flags,
                                                         float *__restrict__ min_dist, float *__restrict__ max_abc) {
  const int g = blockIdx.x, tid = threadIdx.x;
  __shared__ float L[9];
  __shared__ int R[3];
  __shared__ int hist[NTYPE];
  __shared__ float red[VF_THREADS];
  const int n0 = graph_off[g], n = graph_off[g + 1] - n0;
  if (tid < 9) L[tid] = lat[(int64_t)g * 9 + tid];
  for (int i = tid; i < NTYPE; i += VF_THREADS) hist[i] = 0;
  __syncthreads();
  if (tid == 0) {
    const float *A = L, *B = L + 3, *C = L + 6;
    float cab[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
    float cbc[3] = {B[1] * C[2] - B[2] * C[1], B[2] * C[0] - B[0] * C[2], B[0] * C[1] - B[1] * C[0]};
    float cca[3] = {C[1] * A[2] - C[2] * A[1], C[2] * A[0] - C[0] * A[2], C[0] * A[1] - C[1] * A[0]};
    const float vol = fabsf(A[0] * cbc[0] + A[1] * cbc[1] + A[2] * cbc[2]);
    const float nbc = sqrtf(cbc[0] * cbc[0] + cbc[1] * cbc[1] + cbc[2] * cbc[2]);
    const float nca = sqrtf(cca[0] * cca[0] + cca[1] * cca[1] + cca[2] * cca[2]);
    const float nab = sqrtf(cab[0] * cab[0] + cab[1] * cab[1] + cab[2] * cab[2]);
    const float hk[3] = {vol / nbc, vol / nca, vol / nab};     // plane spacings along a*, b*, c*
    for (int k = 0; k < 3; k++) {
      int r = VF_RMAX;
      if (hk[k] > 0.f && isfinite(hk[k])) {
        const float need = ceilf(thr / hk[k] + 0.5f);
        r = need < 1.f ? 1 : (need > (float)VF_RMAX ? VF_RMAX : (int)need);
      }
      R[k] = r;
    }
    const float la = sqrtf(A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
    const float lb = sqrtf(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
    const float lc = sqrtf(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
    max_abc[g] = fmaxf(la, fmaxf(lb, lc));
  }
  for (int i = tid; i < n; i += VF_THREADS) {
    int64_t z = a[n0 + i];
    if (z > 103 || z < 0) z = 0;                               // schema.py:60-62
    atomicAdd(&hist[(int)z], 1);
  }
  __syncthreads();
  // ---- all pairs i < j, all images in the box ----
  float best = INFINITY;
  const int n_pairs = n * (n - 1) / 2;
  for (int p = tid; p < n_pairs; p += VF_THREADS) {
    // p -> (i, j), i < j, row-major over the strict upper triangle
    int i = (int)((2.0f * n - 1.0f - sqrtf((2.0f * n - 1.0f) * (2.0f * n - 1.0f) - 8.0f * p)) * 0.5f);
    while (i > 0 && i * (2 * n - i - 1) / 2 > p) i--;
    while ((i + 1) * (2 * n - i - 2) / 2 <= p) i++;
    const int j = p - i * (2 * n - i - 1) / 2 + i + 1;
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float t = x[(int64_t)(n0 + j) * 3 + k] - x[(int64_t)(n0 + i) * 3 + k];
      d[k] = t - rintf(t);
    }
    for (int ma = -R[0]; ma <= R[0]; ma++)
      for (int mb = -R[1]; mb <= R[1]; mb++)
        for (int mc = -R[2]; mc <= R[2]; mc++) {
          const float fa = d[0] + ma, fb = d[1] + mb, fc = d[2] + mc;
          const float vx = fa * L[0] + fb * L[3] + fc * L[6];
          const float vy = fa * L[1] + fb * L[4] + fc * L[7];
          const float vz = fa * L[2] + fb * L[5] + fc * L[8];
          const float d2 = vx * vx + vy * vy + vz * vz;
          if (d2 > 0.f && d2 < best) best = d2;
        }
  }
  red[tid] = best;
  __syncthreads();
  for (int s = VF_THREADS / 2; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fminf(red[tid], red[tid + s]);
    __syncthreads();
  }
  if (tid == 0) {
    const float dmin = sqrtf(red[0]);                          // +inf when there is no positive distance
    min_dist[g] = dmin;
    int f = 0;
    if (!(max_abc[g] <= max_len)) f |= CB2_INVALID_LATTICE;
    if (dmin < thr) f |= CB2_INVALID_DISTANCE;
    if (target != nullptr) {
      int gc = 0;
      for (int z = 0; z < NTYPE; z++) {
        int c = hist[z];
        while (c) { const int t = gc % c; gc = c; c = t; }
      }
      bool same = gc > 0;
      for (int z = 0; z < NTYPE && same; z++) same = (hist[z] / (gc > 0 ? gc : 1)) == target[z];
      if (!same) f |= CB2_INVALID_COMPOSITION;
    }
    flags[g] = f;
  }
}

int launch_validity(const int64_t *a, const float *x, const float *lat, const int32_t *graph_off, int B,
                    const int32_t *target, float max_len, float thr, int32_t *flags, float *min_dist,
                    float *max_abc, cudaStream_t st) {
  if (B == 0) return CB2_OK;
  k_validity<<<B, VF_THREADS, 0, st>>>(a, x, lat, graph_off, target, max_len, thr, flags, min_dist, max_abc);
  CB2_LAUNCH_OK("k_validity");
  return CB2_OK;
}

}  // namespace cb2
