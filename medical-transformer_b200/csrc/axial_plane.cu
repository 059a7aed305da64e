// axial_plane.cu -- the two layout-changing ends of the axial op:
//   k_out_*        o_raw (strip-major sv | sve) -> BN(bn_output) affine + pair sum -> NCHW (+ AvgPool)
//                  (reference axialnet.py:178-187)
//   k_bwd_gather_* dout (NCHW, through AvgPool) -> strip-major g_s + the two bn_output backward sums
// Height-axis ops (strips run along H, NCHW runs along W) go through a 32x32 shared-memory
// transpose so that both the strip-major and the NCHW side move in full 128-byte lines; every
// other case is one element per thread.
#include "common.cuh"

// ---------------------------------------------------------------------------- k_out_direct
// one output element per thread; a CTA covers 256 consecutive NCHW outputs
__global__ void __launch_bounds__(256)
k_out_direct(Dims m, const float* __restrict__ coef, const float* __restrict__ oraw, float* __restrict__ out) {
  __shared__ float sa[512], sb[512];
  const int NO = m.no, L = m.L, st_ = m.stride, Ho = m.H / st_, Wo = m.W / st_, HW = Ho * Wo;
  const long total = (long)m.N * m.C * HW;
  const long first = (long)blockIdx.x * 256;
  const int plane0 = (int)(first / HW);
  long lastl = first + 255 < total - 1 ? first + 255 : total - 1;
  const int nplanes = (int)(lastl / HW) - plane0 + 1;
  for (int k = threadIdx.x; k < nplanes * NO; k += 256) {
    const int plane = plane0 + k / NO, s = k % NO;
    const int n = plane / m.C, c = plane - n * m.C;
    const float* tb = coef + (size_t)(n / (m.N / m.ngroups)) * m.ncoef_f + m.cf_o + 2 * (c * NO + s);
    sa[k] = tb[0]; sb[k] = tb[1];
  }
  __syncthreads();
  const long idx = first + threadIdx.x;
  if (idx < total) {
    const int plane = (int)(idx / HW), rem = (int)(idx - (long)plane * HW);
    const int ho = rem / Wo, wo = rem - ho * Wo;
    const int n = plane / m.C, c = plane - n * m.C;
    const int k = (plane - plane0) * NO;
    const float a0 = sa[k], a1 = NO == 2 ? sa[k + 1] : 0.f, bsum = sb[k] + (NO == 2 ? sb[k + 1] : 0.f);
    float acc = 0.f;
    for (int dh = 0; dh < st_; ++dh)
      for (int dw = 0; dw < st_; ++dw) {
        const int h = ho * st_ + dh, w = wo * st_ + dw;
        const int b = m.axis == 0 ? n * m.W + w : n * m.H + h;
        const int l = m.axis == 0 ? h : w;
        const float* o = oraw + ((size_t)b * NO * m.C + c * NO) * L + l;
        acc = fmaf(a0, __ldg(o), acc);
        if (NO == 2) acc = fmaf(a1, __ldg(o + L), acc);
      }
    out[idx] = fmaf(acc, 1.f / (float)(st_ * st_), bsum);
  }
}

// ---------------------------------------------------------------------------- k_out_t
// axis 0, stride 1: 32(h) x 32(w) tile of one (n, c) plane through smem
__global__ void __launch_bounds__(256)
k_out_t(Dims m, const float* __restrict__ coef, const float* __restrict__ oraw, float* __restrict__ out) {
  __shared__ float tile[32][33];
  __shared__ float sa[2], sb[2];
  const int NO = m.no, L = m.L;
  const int n = blockIdx.z / m.C, c = blockIdx.z - n * m.C;
  const int w0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  if (threadIdx.x < NO) {
    const float* tb = coef + (size_t)(n / (m.N / m.ngroups)) * m.ncoef_f + m.cf_o + 2 * (c * NO + threadIdx.x);
    sa[threadIdx.x] = tb[0]; sb[threadIdx.x] = tb[1];
  }
  __syncthreads();
  const float a0 = sa[0], a1 = NO == 2 ? sa[1] : 0.f, bsum = sb[0] + (NO == 2 ? sb[1] : 0.f);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int wl = ty + 8 * rr, h = h0 + tx, w = w0 + wl;
    float v = 0.f;
    if (h < m.H && w < m.W) {
      const float* o = oraw + ((size_t)(n * m.W + w) * NO * m.C + c * NO) * L + h;
      v = a0 * __ldg(o);
      if (NO == 2) v = fmaf(a1, __ldg(o + L), v);
    }
    tile[wl][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int hl = ty + 8 * rr, h = h0 + hl, w = w0 + tx;
    if (h < m.H && w < m.W) out[((size_t)blockIdx.z * m.H + h) * m.W + w] = tile[tx][hl] + bsum;
  }
}

// ---------------------------------------------------------------------------- backward gather
__device__ __forceinline__ void gather_reduce(const Dims& m, const MtAxialParams& p, float s1, float s2a, float s2b,
                                              int c, int grp, const double* stats, double* bstats, float* bcoef,
                                              unsigned* bcnt, int cta) {
  __shared__ float red[3][8];
  s1 = seg_sum(s1, 32); s2a = seg_sum(s2a, 32); s2b = seg_sum(s2b, 32);
  const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][wid] = s1; red[1][wid] = s2a; red[2][wid] = s2b; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int k = 0; k < nw; ++k) t += red[threadIdx.x][k];
    double* bs = stat_slot(m, bstats, grp, m.nstat_b, cta);
    if (threadIdx.x == 0) MT_GADD(m, bs + m.b_out_s1 + c, (double)t);
    else if (threadIdx.x == 1) MT_GADD(m, bs + m.b_out_s2 + c * m.no, (double)t);
    else if (m.no == 2) MT_GADD(m, bs + m.b_out_s2 + c * m.no + 1, (double)t);
  }
  if (last_cta_done(bcnt + 0, gridDim.x * gridDim.y * gridDim.z)) fin_bwd_out(m, p, stats, bstats, bcoef);
}

// grid (splits, C, groups): a CTA walks a chunk of the (group, channel) slab of npg*H*W full-resolution
// positions, 4 independent loads in flight per thread; one ticket per CTA (not per plane).
__global__ void __launch_bounds__(256)
k_bwd_gather_direct(Dims m, MtAxialParams p, const float* __restrict__ dout, const float* __restrict__ oraw,
                    float* __restrict__ gs, const double* __restrict__ stats, double* __restrict__ bstats,
                    float* __restrict__ bcoef, unsigned* __restrict__ bcnt, int chunk) {
  const int c = blockIdx.y, grp = blockIdx.z, npg = m.N / m.ngroups;
  const int NO = m.no, L = m.L, st_ = m.stride, Ho = m.H / st_, Wo = m.W / st_, HW = m.H * m.W;
  const int cnt = npg * HW, e0 = blockIdx.x * chunk, e1 = e0 + chunk < cnt ? e0 + chunk : cnt;
  const float inv = 1.f / (float)(st_ * st_);
  float s1 = 0.f, s2a = 0.f, s2b = 0.f;
  for (int base = e0; base < e1; base += 4 * blockDim.x) {
    float g[4], oa[4], ob[4];
    size_t go[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = base + u * blockDim.x + threadIdx.x;
      g[u] = 0.f; oa[u] = 0.f; ob[u] = 0.f; go[u] = 0;
      if (e < e1) {
        const int nl = e / HW, r = e - nl * HW, h = r / m.W, w = r - h * m.W, n = grp * npg + nl;
        g[u] = __ldg(dout + ((size_t)(n * m.C + c) * Ho + h / st_) * Wo + w / st_) * inv;
        const int b = m.axis == 0 ? n * m.W + w : n * m.H + h, l = m.axis == 0 ? h : w;
        go[u] = ((size_t)b * m.C + c) * L + l;
        const float* o = oraw + ((size_t)b * NO * m.C + c * NO) * L + l;
        oa[u] = __ldg(o);
        if (NO == 2) ob[u] = __ldg(o + L);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = base + u * blockDim.x + threadIdx.x;
      if (e < e1) { gs[go[u]] = g[u]; s1 += g[u]; s2a = fmaf(g[u], oa[u], s2a); s2b = fmaf(g[u], ob[u], s2b); }
    }
  }
  gather_reduce(m, p, s1, s2a, s2b, c, grp, stats, bstats, bcoef, bcnt, blockIdx.x + blockIdx.y);
}

// axis 0, stride 1: 32x32 tile transpose.  grid (ceil(W/32), ceil(H/32), N*C)
__global__ void __launch_bounds__(256)
k_bwd_gather_t(Dims m, MtAxialParams p, const float* __restrict__ dout, const float* __restrict__ oraw,
               float* __restrict__ gs, const double* __restrict__ stats, double* __restrict__ bstats,
               float* __restrict__ bcoef, unsigned* __restrict__ bcnt) {
  __shared__ float tile[32][33];
  const int NO = m.no, L = m.L;
  const int n = blockIdx.z / m.C, c = blockIdx.z - n * m.C;
  const int w0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int hl = ty + 8 * rr, h = h0 + hl, w = w0 + tx;
    tile[hl][tx] = (h < m.H && w < m.W) ? __ldg(dout + ((size_t)blockIdx.z * m.H + h) * m.W + w) : 0.f;
  }
  __syncthreads();
  float s1 = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int wl = ty + 8 * rr, h = h0 + tx, w = w0 + wl;
    if (h < m.H && w < m.W) {
      const float g = tile[tx][wl];
      const int b = n * m.W + w;
      gs[((size_t)b * m.C + c) * L + h] = g;
      const float* o = oraw + ((size_t)b * NO * m.C + c * NO) * L + h;
      s1 += g; s2a = fmaf(g, __ldg(o), s2a);
      if (NO == 2) s2b = fmaf(g, __ldg(o + L), s2b);
    }
  }
  gather_reduce(m, p, s1, s2a, s2b, c, n / (m.N / m.ngroups), stats, bstats, bcoef, bcnt, blockIdx.x + blockIdx.y + blockIdx.z);
}

// ---------------------------------------------------------------------------- host launchers
int launch_out(const Dims& m, const float* coef, const float* oraw, float* out, cudaStream_t st) {
  if (m.axis == 0 && m.stride == 1 && m.H >= 32 && m.W >= 32) {
    dim3 grid((m.W + 31) / 32, (m.H + 31) / 32, m.N * m.C);
    k_out_t<<<grid, 256, 0, st>>>(m, coef, oraw, out);
  } else {
    const long total = (long)m.N * m.C * (m.H / m.stride) * (m.W / m.stride);
    k_out_direct<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m, coef, oraw, out);
  }
  return 1;
}

int launch_bwd_gather(const Dims& m, const MtAxialParams& p, const float* dout, const float* oraw, float* gs,
                      const double* stats, double* bstats, float* bcoef, unsigned* bcnt, cudaStream_t st) {
  if (m.axis == 0 && m.stride == 1 && m.H >= 32 && m.W >= 32) {
    dim3 grid((m.W + 31) / 32, (m.H + 31) / 32, m.N * m.C);
    k_bwd_gather_t<<<grid, 256, 0, st>>>(m, p, dout, oraw, gs, stats, bstats, bcoef, bcnt);
  } else {       // small planes (everything is L2 resident; full 32x32 tiles would be mostly empty) and width-axis ops
    const int cnt = (m.N / m.ngroups) * m.H * m.W;
    // ~1024 CTAs in total: enough to fill the GPU, few enough that the CTA-done tickets (one same-address
    // atomic each; 8192 of them cost > 100 us on the batched LoGo branch) stay cheap
    long want = ((long)cnt * m.C * m.ngroups / 1024 + 255) / 256 * 256;
    const int chunk = want < 256 ? 256 : (want > 4096 ? 4096 : (int)want);
    int bs = chunk / 4; if (bs > 256) bs = 256; if (bs > cnt) bs = ((cnt + 31) / 32) * 32; if (bs < 32) bs = 32;
    dim3 grid((cnt + chunk - 1) / chunk, m.C, m.ngroups);
    k_bwd_gather_direct<<<grid, bs, 0, st>>>(m, p, dout, oraw, gs, stats, bstats, bcoef, bcnt, chunk);
  }
  return 1;
}
