// K5, second generation ("team" column kernel): complex FFT along x of a group of adjacent ky columns,
// division by the Laplacian eigenvalues (pressure.py:94-108 through jax_cfd pseudoinverse) and inverse FFT,
// in place.  Same mathematics as poisson_fast.cu / poisson.cu (forward DIF, diagonal in digit-reversed
// positions, exact inverse); the data movement is re-designed around what ncu showed for the first
// generation (profiles/r1i: 48 % issue utilisation, 9.0 M warp instructions, 129 of 148 SMs busy,
// every pass fenced by a 1024-thread barrier):
//
//   * a CTA owns `cpc` <= CS adjacent columns, chosen so that the grid fills the machine evenly
//     (2048^2: 1025 columns = 147 CTAs x 7 columns, one CTA per SM);
//   * first / last pass ("G" mapping): thread = (column slot c = t % CS, butterfly b = t / CS).  A warp
//     reads CS x 8 bytes of contiguous spectrum per row (whole 32-byte sectors) straight into registers,
//     transforms, and stores into shared memory -- no staging pass;
//   * inner passes ("team" mapping): the T = N/16 threads with t / T = column work on that column alone
//     and synchronise with a NAMED barrier of T threads, so the columns of a CTA drift apart and one
//     team's shared-memory phase overlaps another team's arithmetic (only two CTA-wide barriers remain);
//   * every thread moves 16 complex values per pass (one radix-16 or two radix-8 butterflies);
//   * shared-memory layout [column][p + (p >> PS)] (+ a column pitch that staggers the columns over the
//     banks): every access of every pass is conflict-free (tools/fft_model.py counts 2.0 wavefronts per
//     8-byte warp access, the minimum) AND every address is a per-thread base plus a compile-time constant,
//     so the index arithmetic of the XOR-swizzled first generation (3 integer ops per access) is gone;
//   * twiddle factors come from shared tables laid out [q][b] (lanes read consecutive entries).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "fft.cuh"
#include "poisson_fast.cuh"

namespace jaxib {
namespace {

constexpr float kCutoffT = 10.0f * 1.1920928955078125e-07f;
__device__ __forceinline__ float inv_eig_t(float l) { return fabsf(l) > kCutoffT ? __frcp_rn(l) : 0.0f; }

template <int R, bool INV>
__device__ __forceinline__ void dft_t(float2* v) {
  if (INV) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
  }
  Dft<R>::run(v);
  if (INV) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
  }
}

__device__ __forceinline__ void team_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int R1, int R2, int R3, int CS, int PS>
struct TeamGeom {
  static constexpr int N = R1 * R2 * R3;
  static constexpr int T = N / 16;        // threads per column team
  static constexpr int SUB1 = N / R1;     // stride of pass A
  static constexpr int SUB2 = SUB1 / R2;  // stride of pass B
  static constexpr int NBA = 16 / R1, NBB = 16 / R2, NBC = 16 / R3;  // butterflies per thread and pass
  static constexpr int COLPITCH = N + (N >> PS) + 16 / CS;
  static constexpr int TA = (R1 - 1) * SUB1, TB = (R2 - 1) * SUB2;   // twiddle table entries
  static constexpr size_t SMEM = (size_t)(CS * COLPITCH + TA + TB) * sizeof(float2);
  static_assert(SUB1 % (1 << PS) == 0, "pass-A stride must be a multiple of the padding period");
  static_assert(16 % R1 == 0 && 16 % R2 == 0 && 16 % R3 == 0, "radices 2..16");
  __host__ __device__ static constexpr int pad(int p) { return p + (p >> PS); }
};

// HV = 2: the CTA works as two HALVES of CS/2 columns that only meet at kernel start.  Half 1 issues its column
// loads after half 0 has issued all of its own (named-barrier arrive / sync), so the memory system serves half 0
// first: its transform runs while half 1's columns are still streaming in, and it stores while half 1 still
// computes -- the load / compute / store phases of the two halves overlap inside one SM (the kernel is a single
// wave: without this every SM loads, then computes, then stores in lockstep; ncu r2d: issue slots 50 % busy).
template <int R1, int R2, int R3, int CS, int PS, int HV>
__global__ void __launch_bounds__(CS * (R1 * R2 * R3 / 16), 1)
cols_team_kernel(FastColParams P, int cpc, float2* __restrict__ spec) {
  using G = TeamGeom<R1, R2, R3, CS, PS>;
  constexpr int N = G::N, T = G::T, SUB1 = G::SUB1, SUB2 = G::SUB2, CP = G::COLPITCH;
  constexpr int CSH = CS / HV, TH = CSH * T;  // columns / threads per half
  extern __shared__ __align__(16) float2 smt[];
  float2* const ta = smt + CS * CP;  // [q - 1][b]:  W_N^(b q)
  float2* const tb = ta + G::TA;     // [q - 1][b2]: W_N^(b2 q R1)
  const int t = threadIdx.x;
  pdl_trigger();
  // twiddle tables: ONE bulk asynchronous copy (UBLKCP) of the plan's pre-laid-out [TA | TB] table, counted on
  // an mbarrier; it is independent of the previous kernel (issued before the dependency wait) and lands while
  // the column loads are in flight
  __shared__ __align__(8) unsigned long long tbar;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tbar);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr unsigned bytes = (unsigned)((G::TA + G::TB) * sizeof(float2));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(ta)),
                 "l"(P.tw_team), "r"(bytes), "r"(bar)
                 : "memory");
  }
  // ---- G mapping (inside the thread's half)
  const int half = t / TH, th = t - half * TH;
  const int c = half * CSH + th % CSH, bt = th / CSH;
  const int gc = blockIdx.x * cpc + c;
  const bool gvalid = c < cpc && gc < P.ncols && gc < P.limit;
  float2* const gspec = spec + (size_t)blockIdx.y * P.nx * P.sp + gc;
  const size_t rs = (size_t)P.sp;
  float2* const gcol = smt + c * CP;
  pdl_wait();
  __syncthreads();  // the mbarrier initialisation is visible to every thread
  bool tables_ready = false;
#pragma unroll
  for (int j = 0; j < G::NBA; ++j) {  // pass A forward: global -> registers -> shared
    const int b = bt + T * j;
    float2 v[R1];
    if (HV == 2 && j == 0 && half == 1) asm volatile("bar.sync 11, %0;" ::"r"(CS * T) : "memory");  // after half 0's loads
#pragma unroll
    for (int r = 0; r < R1; ++r) v[r] = gvalid ? gspec[(size_t)(b + SUB1 * r) * rs] : make_float2(0.f, 0.f);
    if (HV == 2 && j == G::NBA - 1 && half == 0) asm volatile("bar.arrive 11, %0;" ::"r"(CS * T) : "memory");
    if (!tables_ready) {  // the table copy was issued long before: this wait is normally free
      unsigned ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar) : "memory");
      tables_ready = true;
    }
    Dft<R1>::run(v);
    float2* const dst = gcol + G::pad(b);
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      if (q > 0) v[q] = cmul(v[q], ta[(q - 1) * SUB1 + b]);
      dst[G::pad(SUB1 * q)] = v[q];
    }
  }
  if (HV == 2) team_sync(9 + half, TH); else __syncthreads();
  // ---- team mapping
  const int team = t / T, b = t - team * T;
  if (team < cpc) {
    float2* const col = smt + team * CP;
    const int tcol = blockIdx.x * cpc + team;
    const float ly = P.lam_y[tcol];
#pragma unroll
    for (int j = 0; j < G::NBB; ++j) {  // pass B forward
      const int bf = b + T * j;
      const int blk = bf / SUB2, b2 = bf - blk * SUB2;
      float2* const e = col + G::pad(blk * SUB1 + b2);
      float2 v[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) v[r] = e[G::pad(SUB2 * r)];
      Dft<R2>::run(v);
#pragma unroll
      for (int q = 0; q < R2; ++q) {
        if (q > 0) v[q] = cmul(v[q], tb[(q - 1) * SUB2 + b2]);
        e[G::pad(SUB2 * q)] = v[q];
      }
    }
    team_sync(team + 1, T);
#pragma unroll
    for (int j = 0; j < G::NBC; ++j) {  // pass C forward, diagonal, pass C inverse -- in registers
      const int p0 = (b + T * j) * R3;
      float2* const e = col + G::pad(p0);
      float2 v[R3];
#pragma unroll
      for (int r = 0; r < R3; ++r) v[r] = e[G::pad(r)];
      Dft<R3>::run(v);
      const float4* lx4 = reinterpret_cast<const float4*>(P.lam_x_pos + p0);
#pragma unroll
      for (int r = 0; r < R3; r += 4) {
        const float4 l4 = __ldg(lx4 + r / 4);
        v[r] = cscale(v[r], inv_eig_t(l4.x + ly));
        v[r + 1] = cscale(v[r + 1], inv_eig_t(l4.y + ly));
        v[r + 2] = cscale(v[r + 2], inv_eig_t(l4.z + ly));
        v[r + 3] = cscale(v[r + 3], inv_eig_t(l4.w + ly));
      }
      dft_t<R3, true>(v);
#pragma unroll
      for (int r = 0; r < R3; ++r) e[G::pad(r)] = v[r];
    }
    team_sync(team + 1, T);
#pragma unroll
    for (int j = 0; j < G::NBB; ++j) {  // pass B inverse
      const int bf = b + T * j;
      const int blk = bf / SUB2, b2 = bf - blk * SUB2;
      float2* const e = col + G::pad(blk * SUB1 + b2);
      float2 v[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        v[r] = e[G::pad(SUB2 * r)];
        if (r > 0) v[r] = cmulc(v[r], tb[(r - 1) * SUB2 + b2]);
      }
      dft_t<R2, true>(v);
#pragma unroll
      for (int r = 0; r < R2; ++r) e[G::pad(SUB2 * r)] = v[r];
    }
  }
  if (HV == 2) team_sync(9 + half, TH); else __syncthreads();
#pragma unroll
  for (int j = 0; j < G::NBA; ++j) {  // pass A inverse: shared -> registers -> global
    const int bb = bt + T * j;
    const float2* const src = gcol + G::pad(bb);
    float2 v[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      v[r] = src[G::pad(SUB1 * r)];
      if (r > 0) v[r] = cmulc(v[r], ta[(r - 1) * SUB1 + bb]);
    }
    dft_t<R1, true>(v);
    if (gvalid) {
#pragma unroll
      for (int r = 0; r < R1; ++r) gspec[(size_t)(bb + SUB1 * r) * rs] = v[r];
    }
  }
}

template <int R1, int R2, int R3, int CS, int PS>
int launch_cols_team_t(cudaStream_t s, const FastColParams& P, int batch, float2* spec) {
  using G = TeamGeom<R1, R2, R3, CS, PS>;
  static int sms = 0, halves = 1;  // measured at 2048^2: staggered halves 21.9 us, plain 20.6 us (JAXIB_COLS_HALVES=2 to retry)
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(cols_team_kernel<R1, R2, R3, CS, PS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    cudaFuncSetAttribute(cols_team_kernel<R1, R2, R3, CS, PS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    const char* eh = getenv("JAXIB_COLS_HALVES");
    if (eh && atoi(eh) == 2) halves = 2;
  }
  // columns per CTA: CS (whole 32-byte sectors per row; measured at 2048^2: 129 CTAs x 8 columns 20.8 us,
  // 147 CTAs x 7 columns 22.2 us -- the 56-byte row segments straddle sectors); fewer only for tiny problems
  int cpc = (P.ncols * batch + sms - 1) / sms;
  if (cpc < 1) cpc = 1;
  if (cpc > CS || cpc * 2 > CS) cpc = CS;
  static const char* e = getenv("JAXIB_COLS_CPC");
  if (e && atoi(e) >= 1 && atoi(e) <= CS) cpc = atoi(e);
  dim3 grid((P.ncols + cpc - 1) / cpc, batch);
  if (halves == 2 && (CS / 2) * G::T >= 32)
    launch_pdl(16, cols_team_kernel<R1, R2, R3, CS, PS, 2>, grid, dim3(CS * G::T), G::SMEM, s, P, cpc, spec);
  else
    launch_pdl(16, cols_team_kernel<R1, R2, R3, CS, PS, 1>, grid, dim3(CS * G::T), G::SMEM, s, P, cpc, spec);
  return check_launch("cols_team_kernel");
}

}  // namespace

bool team_cols_supported(int n) { return n == 512 || n == 1024 || n == 2048 || n == 4096; }
void team_radices(int n, int* r);

// pre-laid-out twiddle table of the team kernel for length n: [TA | TB] (see cols_team_kernel)
void team_twiddle_table(int n, std::vector<float2>* out) {
  int r[4];
  team_radices(n, r);
  const int R1 = r[0], R2 = r[1];
  const int sub1 = n / R1, sub2 = sub1 / R2;
  const double two_pi = 6.283185307179586476925286766559;
  auto w = [&](long long m) {
    m %= n;
    return make_float2((float)cos(two_pi * m / n), (float)-sin(two_pi * m / n));
  };
  out->clear();
  for (int q = 1; q < R1; ++q)
    for (int b = 0; b < sub1; ++b) out->push_back(w((long long)b * q));
  for (int q = 1; q < R2; ++q)
    for (int b2 = 0; b2 < sub2; ++b2) out->push_back(w((long long)b2 * q * R1));
}

// radix triples of the team kernel (the plan's eigenvalue table follows the same digit order)
void team_radices(int n, int* r) {
  switch (n) {
    case 512: r[0] = 8; r[1] = 8; r[2] = 8; break;
    case 1024: r[0] = 8; r[1] = 16; r[2] = 8; break;
    case 2048: r[0] = 16; r[1] = 16; r[2] = 8; break;
    case 4096: r[0] = 16; r[1] = 16; r[2] = 16; break;
    default: r[0] = r[1] = r[2] = 0;
  }
  r[3] = 0;
}

int launch_cols_team(cudaStream_t s, const FastColParams& P, int batch, float2* spec) {
  switch (P.nx) {
    case 512: return launch_cols_team_t<8, 8, 8, 8, 3>(s, P, batch, spec);
    case 1024: return launch_cols_team_t<8, 16, 8, 8, 4>(s, P, batch, spec);
    case 2048: return launch_cols_team_t<16, 16, 8, 8, 4>(s, P, batch, spec);
    case 4096: return launch_cols_team_t<16, 16, 16, 4, 4>(s, P, batch, spec);
    default: set_error("team column kernel: unsupported size %d", P.nx); return JAXIB_ERR_UNSUPPORTED;
  }
}

}  // namespace jaxib
