// K1 fast path, second generation: periodic grid, first-order upwind advection (Flapping / journal-bearing
// demos, the ib2048 / slab8192 bench workloads, the 600^2 ensemble).  Same arithmetic as stencil.cu
// (equations.py:42-83, time_stepping.py:191-208, advection.py:120-124), organised around the Blackwell
// asynchronous copy engine:
//
//   * a CTA owns a strip of W = 4 * NT columns and MARCHES over `rpc` rows.  One PRODUCER warp streams whole
//     row segments of u, v, p into a ring of shared-memory row slots with bulk asynchronous copies
//     (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier: SASS UBLKCP + SYNCS);
//     the periodic y-halo is two extra 16-byte copies per row from the wrapped columns, the periodic
//     x-halo is just the wrapped row index.  No thread ever computes a load address per cell;
//   * NT CONSUMER threads own 4 consecutive columns each.  They wait on the slot's "full" barrier, read the
//     row (one LDS.128 + two LDS.32 per field), release the slot ("empty" barrier, one arrive per warp) and
//     keep rows i-1, i, i+1 in registers: three register sets rotate roles through a 3x unrolled loop, so
//     the window never moves through MOVs;
//   * x-face fluxes are carried from row to row; the arithmetic is folded so that a cell-step costs ~42
//     floating-point instructions: fluxes are computed DOUBLED (w2 = a + b has the sign of the face velocity
//     (a + b)/2) and every constant factor (dt, 1/h, nu/h^2, the halves) is folded into five coefficients.
//     Results differ from the unfused form by a few ulp per operation (bar: 1e-5 relative L2);
//   * outputs are written with 16-byte coalesced stores.
//
// The ring keeps depth * 3 * (W + 8) * 4 bytes in flight per CTA (2 CTAs x 50 KB per SM at 2048 columns),
// which covers the HBM latency-bandwidth product (~44 KB per SM at 6.5 TB/s and 1 us).
#include <stdlib.h>

#include "common.cuh"

namespace jaxib {
namespace {

constexpr int kRingMaxThreads = 544;  // 512 consumers + the producer warp
constexpr int kRingMaxDepth = 8;
constexpr int kBarBytes = 128;        // 2 * kRingMaxDepth mbarriers

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// bulk asynchronous copy global -> shared, completion signalled in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct RingParams {
  RowBands bands;  // row ranges this launch covers (one range [0, nx) for the whole grid)
  int nx, ny;
  int W;      // strip width in columns = 4 * consumer threads
  int rpc;    // rows per CTA
  int depth;  // ring slots
  // folded coefficients (S = dt for the predictor, 1 for the bare explicit terms):
  float c0;   // [predictor: 1] - 2 S nu (sx + sy)
  float cx;   // -S / (2 dx)     times the difference of DOUBLED x-face fluxes
  float cy;   // -S / (2 dy)
  float dxx;  // S nu / dx^2
  float dyy;  // S nu / dy^2
  float gx;   // 1 / dx          pressure gradient (stored pressure is dt p / rho: no dt, time_stepping.py:208)
  float gy;   // 1 / dy
  float fx;   // S force_x / rho
  float fy;
  float pk;   // S / rho         Brinkman term -kappa u / rho (penalty/util_funs.py:6-16)
};

struct Row {
  float u[6];  // columns j-1 .. j+4
  float v[6];
  float p[5];  // columns j .. j+4
};

__device__ __forceinline__ float flux2(float w2, float c_lo, float c_hi) { return w2 * (w2 > 0.0f ? c_lo : c_hi); }

template <bool PREDICT>
__global__ void __launch_bounds__(kRingMaxThreads, 1)
explicit_upwind_ring_kernel(RingParams P, const float* __restrict__ u, const float* __restrict__ v,
                            const float* __restrict__ p, const float* __restrict__ perm, float* __restrict__ out_u,
                            float* __restrict__ out_v, unsigned* __restrict__ done_counter) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  pdl_trigger();
  const int NT = blockDim.x - 32;  // consumer threads
  const int tid = threadIdx.x, lane = tid & 31;
  const int SW = P.W + 8;          // floats per field row slot: [4 halo | W | 4 halo]
  const int ny = P.ny, nx = P.nx, depth = P.depth;
  const int c0col = blockIdx.x * P.W;
  const int Ws = min(P.W, ny - c0col);
  // blockIdx.y -> (band, chunk of rpc rows inside the band)
  int band = 0;
  while (band + 1 < P.bands.n && (int)blockIdx.y >= P.bands.chunk0[band + 1]) ++band;
  const int i0 = P.bands.start[band] + ((int)blockIdx.y - P.bands.chunk0[band]) * P.rpc;
  const int i1 = min(i0 + P.rpc, P.bands.start[band] + P.bands.rows[band]);
  const int nload = i1 - i0 + 2;   // rows i0-1 .. i1
  const bool hasp = PREDICT && p != nullptr;
  const size_t boff = (size_t)blockIdx.z * nx * ny;
  const unsigned bar_full = smem_u32(ring_smem), bar_empty = bar_full + 8 * kRingMaxDepth;
  float* const slots = reinterpret_cast<float*>(ring_smem + kBarBytes);

  if (tid == 0) {
    for (int d = 0; d < depth; ++d) {
      mbar_init(bar_full + 8 * d, 1);          // the producer's arrive.expect_tx (+ the copies' bytes)
      mbar_init(bar_empty + 8 * d, NT >> 5);   // one arrive per consumer warp
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();  // nothing above touches global memory

  if (tid >= NT) {  // ---------------------------------------------------------------- producer warp
    if (lane == 0) {
      const float* ub = u + boff;
      const float* vb = v + boff;
      const float* pb = hasp ? p + boff : nullptr;
      const int lcol = c0col == 0 ? ny - 4 : c0col - 4;                      // wrapped y-halo sources
      const int rcol = c0col + Ws >= ny ? c0col + Ws - ny : c0col + Ws;
      const unsigned main_bytes = (unsigned)Ws * 4u;
      int slot = 0;
      unsigned phase = 0;
      int r = i0 - 1;
      if (r < 0) r += nx;
      for (int k = 0; k < nload; ++k) {
        mbar_wait(bar_empty + 8 * slot, phase ^ 1u);  // first pass over the ring: returns at once
        const bool wp = hasp && k >= 1;                // p is needed for rows i0 .. i1 only
        const unsigned full = bar_full + 8 * slot;
        mbar_expect_tx(full, (wp ? 3u : 2u) * (main_bytes + 32u));
        const unsigned dst = smem_u32(slots + (size_t)slot * 3 * SW);
        const size_t ro = (size_t)r * ny;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          if (f == 2 && !wp) break;
          const float* row = (f == 0 ? ub : (f == 1 ? vb : pb)) + ro;
          const unsigned d = dst + (unsigned)(f * SW) * 4u;
          bulk_g2s(d + 16u, row + c0col, main_bytes, full);
          bulk_g2s(d, row + lcol, 16u, full);
          bulk_g2s(d + 16u + main_bytes, row + rcol, 16u, full);
        }
        if (++r == nx) r = 0;
        if (++slot == depth) { slot = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------ consumers
  const int col = 4 * tid;  // first owned column within the strip
  const bool active = col < Ws;
  int slot = 0;
  unsigned phase = 0;
  auto load_row = [&](Row& R, bool withp) {
    mbar_wait(bar_full + 8 * slot, phase);
    const float* sl = slots + (size_t)slot * 3 * SW + col;  // sl[4] is this thread's first column
    {
      const float4 c = *reinterpret_cast<const float4*>(sl + 4);
      R.u[0] = sl[3]; R.u[1] = c.x; R.u[2] = c.y; R.u[3] = c.z; R.u[4] = c.w; R.u[5] = sl[8];
    }
    {
      const float* sv = sl + SW;
      const float4 c = *reinterpret_cast<const float4*>(sv + 4);
      R.v[0] = sv[3]; R.v[1] = c.x; R.v[2] = c.y; R.v[3] = c.z; R.v[4] = c.w; R.v[5] = sv[8];
    }
    if (withp) {
      const float* sp = sl + 2 * SW;
      const float4 c = *reinterpret_cast<const float4*>(sp + 4);
      R.p[0] = c.x; R.p[1] = c.y; R.p[2] = c.z; R.p[3] = c.w; R.p[4] = sp[8];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
    if (++slot == depth) { slot = 0; phase ^= 1u; }
  };

  float fxu[4], fxv[4];  // doubled x-face fluxes through face i - 1/2, carried from row to row
  size_t oidx = boff + (size_t)i0 * ny + c0col + col;  // running output index

  // one row: m = row i-1, c = row i, n = row i+1 (loaded here)
  auto step = [&](const Row& m, const Row& c, Row& n) {
    float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (perm != nullptr && active) k4 = __ldg(reinterpret_cast<const float4*>(perm + oidx));
    load_row(n, hasp);
    float fyu[5], fyv[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {  // y-face k lies between columns j-1+k and j+k
      fyu[k] = flux2(c.v[k] + n.v[k], c.u[k], c.u[k + 1]);
      fyv[k] = flux2(c.v[k] + c.v[k + 1], c.v[k], c.v[k + 1]);
    }
    float ou[4], ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = k + 1;
      const float fu = flux2(c.u[q] + n.u[q], c.u[q], n.u[q]);      // x-face i + 1/2
      const float fv = flux2(c.u[q] + c.u[q + 1], c.v[q], n.v[q]);
      float au = c.u[q] * P.c0;
      float av = c.v[q] * P.c0;
      au = fmaf(P.cx, fu - fxu[k], au);
      av = fmaf(P.cx, fv - fxv[k], av);
      au = fmaf(P.cy, fyu[q] - fyu[q - 1], au);
      av = fmaf(P.cy, fyv[q] - fyv[q - 1], av);
      au = fmaf(P.dxx, m.u[q] + n.u[q], au);
      av = fmaf(P.dxx, m.v[q] + n.v[q], av);
      au = fmaf(P.dyy, c.u[q - 1] + c.u[q + 1], au);
      av = fmaf(P.dyy, c.v[q - 1] + c.v[q + 1], av);
      au += P.fx;
      av += P.fy;
      if (hasp) {
        au = fmaf(-P.gx, n.p[k] - c.p[k], au);
        av = fmaf(-P.gy, c.p[k + 1] - c.p[k], av);
      }
      fxu[k] = fu;
      fxv[k] = fv;
      ou[k] = au;
      ov[k] = av;
    }
    if (perm != nullptr) {
      const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ou[k] = fmaf(-P.pk * kk[k], c.u[k + 1], ou[k]);
        ov[k] = fmaf(-P.pk * kk[k], c.v[k + 1], ov[k]);
      }
    }
    if (active) {
      *reinterpret_cast<float4*>(out_u + oidx) = make_float4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<float4*>(out_v + oidx) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
    oidx += ny;
  };

  Row A, B, C;
#pragma unroll
  for (int k = 0; k < 5; ++k) A.p[k] = B.p[k] = C.p[k] = 0.0f;
  load_row(A, false);  // row i0 - 1
  load_row(B, hasp);   // row i0
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    fxu[k] = flux2(A.u[k + 1] + B.u[k + 1], A.u[k + 1], B.u[k + 1]);
    fxv[k] = flux2(A.u[k + 1] + A.u[k + 2], A.v[k + 1], B.v[k + 1]);
  }
  for (int i = i0; i < i1; i += 3) {
    step(A, B, C);
    if (i + 1 < i1) step(B, C, A);
    if (i + 2 < i1) step(C, A, B);
  }
  if (done_counter != nullptr) {  // this warp's rows are stored: publish (see RowDeps in common.cuh)
    __syncwarp();
    __threadfence();
    if (lane == 0) atomicAdd(done_counter, 1u);
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

bool explicit_ring_eligible(const GridF& g) {
  return g.bc == JAXIB_BC_PERIODIC && g.convect == JAXIB_CONVECT_UPWIND && !(g.ny & 3) && g.ny >= 8 && g.nx >= 4;
}

// Returns true when this path handled the launch (rc holds the status).  bands == nullptr: the whole grid.
bool launch_explicit_ring(cudaStream_t s, const GridF& g, const float* u, const float* v, const float* p,
                          const float* perm, float* out_u, float* out_v, bool predictor, int* rc, const RowBands* bands,
                          unsigned* done_counter, unsigned* done_increment) {
  if (g.bc != JAXIB_BC_PERIODIC || g.convect != JAXIB_CONVECT_UPWIND || (g.ny & 3) || g.ny < 8 || g.nx < 4) return false;
  const uintptr_t align = reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v) |
                          reinterpret_cast<uintptr_t>(out_u) | reinterpret_cast<uintptr_t>(out_v) |
                          reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(perm);
  if (align & 15) return false;
  if (done_increment) *done_increment = 0;
  static int mode = -1;  // JAXIB_K1=march selects the first-generation register-marching kernel (A/B runs)
  if (mode < 0) { const char* e = getenv("JAXIB_K1"); mode = (e && !strcmp(e, "march")) ? 0 : 1; }
  if (!mode) return false;
  static int sms = 0, max_smem = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(explicit_upwind_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(explicit_upwind_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  }
  // strip geometry: strips of at most `wmax` columns, W a multiple of 128 (whole consumer warps)
  const int wmax = env_int("JAXIB_K1_W", 1024);
  const int strips = (g.ny + wmax - 1) / wmax;
  int W = ((g.ny + strips - 1) / strips + 127) / 128 * 128;
  if (W > 2048) W = 2048;
  const int nstrips = (g.ny + W - 1) / W;
  const int threads = W / 4 + 32;
  int depth = env_int("JAXIB_K1_DEPTH", 4);
  if (depth < 2) depth = 2;
  if (depth > kRingMaxDepth) depth = kRingMaxDepth;
  size_t smem = kBarBytes + (size_t)depth * 3 * (W + 8) * sizeof(float);
  while (smem > (size_t)max_smem && depth > 2) { --depth; smem = kBarBytes + (size_t)depth * 3 * (W + 8) * sizeof(float); }
  if (smem > (size_t)max_smem) return false;
  int per_sm = 1;
  if (predictor) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, explicit_upwind_ring_kernel<true>, threads, smem);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, explicit_upwind_ring_kernel<false>, threads, smem);
  if (per_sm < 1) per_sm = 1;
  // rows per CTA: one CTA per resident slot when the grid is small enough, otherwise marches of <= 64 rows
  RowBands B;
  if (bands) B = *bands;
  else { B.n = 1; B.start[0] = 0; B.rows[0] = g.nx; }
  int band_rows = 0;
  for (int k = 0; k < B.n; ++k) band_rows += B.rows[k];
  if (band_rows <= 0) { *rc = JAXIB_OK; return true; }
  const long long slots = (long long)sms * per_sm;
  const long long row_strips = (long long)band_rows * nstrips * g.batch;
  int rpc = env_int("JAXIB_K1_RPC", 0);
  if (rpc <= 0) {
    rpc = (int)((row_strips + slots - 1) / slots);
    if (rpc < 8) rpc = 8;
    if (rpc > 64) rpc = 64;
    if (rpc > g.nx) rpc = g.nx;
    if (!bands) {
      const int groups = (g.nx + rpc - 1) / rpc;  // equalise the marches
      rpc = (g.nx + groups - 1) / groups;
    }
  }
  int chunks = 0;
  for (int k = 0; k < B.n; ++k) {
    B.chunk0[k] = chunks;
    chunks += (B.rows[k] + rpc - 1) / rpc;
  }
  B.chunk0[B.n] = chunks;
  RingParams P;
  P.bands = B;
  P.nx = g.nx; P.ny = g.ny; P.W = W; P.rpc = rpc; P.depth = depth;
  const float S = predictor ? g.dt : 1.0f;
  P.c0 = (predictor ? 1.0f : 0.0f) - 2.0f * S * g.nu * g.ssum;
  P.cx = -0.5f * S * g.inv_dx;
  P.cy = -0.5f * S * g.inv_dy;
  P.dxx = S * g.nu * g.sx;
  P.dyy = S * g.nu * g.sy;
  P.gx = g.inv_dx;
  P.gy = g.inv_dy;
  P.fx = S * g.fx;
  P.fy = S * g.fy;
  P.pk = S * g.inv_rho;
  dim3 grid(nstrips, chunks, g.batch);
  if (done_increment) *done_increment = (unsigned)(grid.x * grid.y * grid.z) * (unsigned)(W / 4 / 32);
  if (predictor)
    launch_pdl(1, explicit_upwind_ring_kernel<true>, grid, dim3(threads), smem, s, P, u, v, p, perm, out_u, out_v, done_counter);
  else
    launch_pdl(1, explicit_upwind_ring_kernel<false>, grid, dim3(threads), smem, s, P, u, v, (const float*)nullptr, perm,
               out_u, out_v, done_counter);
  *rc = check_launch("explicit_upwind_ring_kernel");
  return true;
}

}  // namespace jaxib
