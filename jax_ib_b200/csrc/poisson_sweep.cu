// K5, third generation ("column sweep"): the x-direction of the pressure solve WITHOUT a transform.
//
// After the row transform along y (K4: packed real FFT on periodic grids, DCT-II in a channel) every stored
// column is a 1-D periodic problem along x for one y-eigenvalue ly (pressure.py:94-108, 111-130: the
// eigenvalues lx + ly of the second-difference operators, array_utils.py:168-182):
//     Q[i-1] - (2 + mu) Q[i] + Q[i+1] = dx^2 X[i],      mu = -ly dx^2 >= 0.
// The first two generations diagonalise it with a complex FFT along x, divide and transform back (5 N log N
// flops per point twice, five shared-memory round trips, 0.25 of the HBM roofline: profiles/r2d).  But a periodic
// tridiagonal Toeplitz system has a closed-form inverse, a two-sided geometric sequence: with rho in (0, 1)
// the root of rho + 1/rho = 2 + mu,
//     a[i] = X[i] + rho a[i-1]  (forward sweep),   b[i] = a[i] + rho b[i+1]  (backward sweep),   Q = -dx^2 rho b,
// both sweeps periodic.  That is 4 FMAs per point instead of ~110 flops, and it is the SAME linear operator as the
// reference's pseudo-inverse (the mu = 0 column, singular, is solved with the mean mode projected out exactly as
// the 10 eps cut-off of the pseudo-inverse does; tests compare against the oracle's FFT solve).
//
// Kernel: a CTA owns CS adjacent columns (CS x 8 bytes per row: whole sectors) over ALL nx rows; thread =
// (column, segment of RS = 16 consecutive rows), the 16 values live in registers.
//   1. local forward sweep from zero; e = a of the segment's last row, S = sum_r rho^r a[r] (Horner);
//   2. the carried values of all segments obey  Aout[s] = e_s + P_s Ain[s], Ain[s] = Aout[s-1]  (cyclic; P_s =
//      rho^rows) and  Bout[s] = T_s + P_s Bin[s], Bin[s] = Bout[s+1],  T_s = S_s + Ain[s] rho (1 - rho^2rows)/(1 - rho^2):
//      one warp per column composes these affine maps with a shuffle scan and closes the cycle with 1/(1 - rho^nx);
//   3. a_true[r] = a[r] + rho^(r+1) Ain, backward sweep from Bin, scale, store in place.
// The real part of column 0 (mu = 0) is solved by the last CTA of the grid with two prefix sums in double.
// Slab decomposition (nx_global > 0; a slab holds nloc rows of every column): the same algebra one level up.  Pass A
// (kSweepAggregate) runs steps 1-2 on the slab's rows with nothing carried in and stores the slab's own affine maps
// -- two floats per column and part -- into every peer's table; after ONE flag barrier pass B (kSweepApply) closes
// the cycle over the `world` slabs in every scanning warp (a few FMAs) and runs steps 1-3 with the carried values
// entering the slab.  b entering the slab from above is b of the next slab's first row, i.e. the spectrum row the
// inverse row kernel needs after its last one: no transpose of the spectrum, no second exchange.
// A channel stores two real DCT columns per complex column: real and imaginary parts simply carry their own
// constants (on periodic grids they are equal).  tools/sweep_model.py restates the arithmetic in NumPy float32:
// rel-L2 2.9e-7 against the float64 solution at 2048^2 (float32 FFT solve: 1.7e-7; on the GPU both pass the same bars, tests/test_gpu_parity.py).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "poisson_fast.cuh"

namespace jaxib {
namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int RS = 16;  // rows per thread

// y = e + P x: composition "first f, then g" = (g.e + g.P f.e, g.P f.P)
struct Affine {
  float e, P;
};
__device__ __forceinline__ Affine then(const Affine& f, const Affine& g) { return {fmaf(g.P, f.e, g.e), g.P * f.P}; }

// Solves one real channel (re or im) of one column for the warp: lane l owns segments [l m, (l+1) m).
// e[], S[]: this column's shared arrays in the layout [j][lane] (segment = lane m + j: for a fixed j the lanes read
// consecutive words -- a [segment] layout would put every lane on the same bank; measured 32-way conflicts on
// every access, profiles/r2v).  On return e[] holds Ain and S[] holds Bin of every segment.
// cyclic: the carried values close on themselves (one GPU).  Otherwise ain0 / bin_last are the values entering the
// block from below / above (0 for the aggregate pass); *tot_e, *tot_t return the block's own maps: the forward value
// leaving it and the backward value leaving it, both for ZERO carried-in values when ain0 = bin_last = 0.
__device__ __forceinline__ void carry_scan(float* e, float* S, int segs, int m, float Pseg, float Plast, float rG,
                                           float rGlast, float close, int lane, bool cyclic, float ain0_in,
                                           float bin_in, float* tot_e, float* tot_t) {
  const int s0 = min(lane * m, segs), s1 = min(s0 + m, segs);
  // ---- forward
  Affine f = {0.0f, 1.0f};
  for (int s = s0; s < s1; ++s) f = then(f, Affine{e[(s - s0) * 32 + lane], s == segs - 1 ? Plast : Pseg});
  Affine inc = f;  // inclusive scan over the lanes: map of segments [0, s1)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Affine up = {__shfl_up_sync(FULLM, inc.e, o), __shfl_up_sync(FULLM, inc.P, o)};
    if (lane >= o) inc = then(up, inc);
  }
  const float all_e = __shfl_sync(FULLM, inc.e, 31);
  *tot_e = all_e;
  const float ain0 = cyclic ? all_e * close : ain0_in;  // cyclic: value entering segment 0 = value leaving the last one
  Affine ex = {__shfl_up_sync(FULLM, inc.e, 1), __shfl_up_sync(FULLM, inc.P, 1)};
  float cur = lane == 0 ? ain0 : fmaf(ex.P, ain0, ex.e);
  // ---- walk forward; T_s replaces S_s
  Affine g = {0.0f, 1.0f};  // backward map of this lane's segments: Bout[s0] = g.e + g.P Bin[s1 - 1]
  float W = 1.0f;
  for (int s = s0; s < s1; ++s) {
    const bool last = s == segs - 1;
    const float Ps = last ? Plast : Pseg;
    const int at = (s - s0) * 32 + lane;
    const float es = e[at];
    e[at] = cur;                                             // Ain[s]
    const float T = fmaf(cur, last ? rGlast : rG, S[at]);
    S[at] = T;
    g.e = fmaf(T, W, g.e);
    W *= Ps;
    cur = fmaf(Ps, cur, es);
  }
  g.P = W;
  // ---- backward: inclusive scan from the last lane down: map of segments [s0, segs)
  Affine dec = g;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Affine dn = {__shfl_down_sync(FULLM, dec.e, o), __shfl_down_sync(FULLM, dec.P, o)};
    if (lane + o < 32) dec = then(dn, dec);
  }
  const float ball = __shfl_sync(FULLM, dec.e, 0);
  *tot_t = ball;
  const float bin_last = cyclic ? ball * close : bin_in;  // cyclic: value entering the last segment from above = Bout[0]
  Affine dx = {__shfl_down_sync(FULLM, dec.e, 1), __shfl_down_sync(FULLM, dec.P, 1)};
  float b = lane == 31 ? bin_last : fmaf(dx.P, bin_last, dx.e);
  for (int s = s1 - 1; s >= s0; --s) {
    const int at = (s - s0) * 32 + lane;
    const float T = S[at];
    S[at] = b;                                               // Bin[s]
    b = fmaf(s == segs - 1 ? Plast : Pseg, b, T);
  }
}

template <int NTH>
__device__ __forceinline__ double cta_sum(double v, double* red, int t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// in-place scan of q[0..n) by the whole CTA (every thread a contiguous piece, piece totals by shuffles)
__device__ void cta_scan(double* q, int n, bool inclusive, double* red, int t) {
  const int nth = blockDim.x;
  const int per = (n + nth - 1) / nth;
  const int lo = min(t * per, n), hi = min(lo + per, n);
  __syncthreads();  // q was written with another thread mapping
  double s = 0.0;
  for (int i = lo; i < hi; ++i) s += q[i];
  double inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(FULLM, inc, o);
    if ((t & 31) >= o) inc += y;
  }
  __syncthreads();
  if ((t & 31) == 31) red[t >> 5] = inc;
  __syncthreads();
  double acc = inc - s;
  for (int w = 0; w < (t >> 5); ++w) acc += red[w];
  for (int i = lo; i < hi; ++i) {
    const double x = q[i];
    if (inclusive) { acc += x; q[i] = acc; } else { q[i] = acc; acc += x; }
  }
  __syncthreads();
}

// blockDim = CS * segs_padded (a multiple of 32); dynamic shared memory: see launch_cols_sweep_t
// Solves Q[i-1] - 2 Q[i] + Q[i+1] = dx2 (X[i] - mean X), mean Q = 0 in place in q[0..n) (shared, double): two prefix sums
__device__ void solve_mean_mode(double* q, int n, double dx2, double* red, int t) {
  const int nth = blockDim.x;
  double sum = 0.0;
  for (int i = t; i < n; i += nth) sum += q[i];
  const double mean = cta_sum<0>(sum, red, t) / n;
  for (int i = t; i < n; i += nth) q[i] = (q[i] - mean) * dx2;
  cta_scan(q, n, true, red, t);   // s_i = sum_{j <= i}
  sum = 0.0;
  for (int i = t; i < n; i += nth) sum += q[i];
  const double smean = cta_sum<0>(sum, red, t) / n;
  for (int i = t; i < n; i += nth) q[i] -= smean;  // d_i = Q[i+1] - Q[i]
  cta_scan(q, n, false, red, t);  // Q_i - Q_0
  sum = 0.0;
  for (int i = t; i < n; i += nth) sum += q[i];
  const double qmean = cta_sum<0>(sum, red, t) / n;
  for (int i = t; i < n; i += nth) q[i] -= qmean;
  __syncthreads();
}

constexpr int kSweepSingle = 0, kSweepAggregate = 1, kSweepApply = 2;

// slab table of one GPU: [world][ncols_pad][2 parts][2] floats (forward map, backward map of each slab) followed by
// the real part of column 0 of the WHOLE grid, [nx_global] floats
__device__ __forceinline__ float* slab_entry(float* table, int ncols_pad, int d, int col, int part) {
  return table + (((size_t)d * ncols_pad + col) * 2 + part) * 2;
}

// lane k < 9 fetches row k of the constant table for (column, part); broadcast later by shuffles
__device__ __forceinline__ float job_const(const ColSweepParams& P, int col, int part, int lane) {
  return lane < 9 ? __ldg(reinterpret_cast<const float*>(P.cst) + ((size_t)lane * P.cst_pitch + col) * 2 + part) : 0.0f;
}
__device__ __forceinline__ float gain_of(const ColSweepParams& P, int col, int part) {
  return __ldg(reinterpret_cast<const float*>(P.cst) + ((size_t)P.cst_pitch + col) * 2 + part);
}
// lane d < world fetches slab d's pair of maps for (column, part) from this GPU's table
__device__ __forceinline__ float2 job_entry(const ColSweepParams& P, int col, int part, int lane) {
  if (lane >= P.world) return make_float2(0.f, 0.f);
  return *reinterpret_cast<const float2*>(slab_entry(P.slab_table[P.rank], P.ncols_pad, lane, col, part));
}

// FULL: nx is a multiple of RS (every segment has RS rows: no per-row predicates)
template <int CS, int MODE, bool FULL>
__global__ void __launch_bounds__(1024, 1)
cols_sweep_kernel(ColSweepParams P, float2* __restrict__ spec, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char sw_raw[];
  __shared__ double red[32];
  pdl_trigger();
  const int t = threadIdx.x;
  const int nx = P.nx;
  spec += (size_t)blockIdx.y * nx * P.sp;
  if (blockIdx.x == 0) {
    // ---- real part of column 0 (mu = 0); CTA 0, i.e. scheduled with the first wave: it is the longest CTA: solved whole, by every slab for itself in the decomposed case
    double* const q = reinterpret_cast<double*>(sw_raw);
    const int nth = blockDim.x;
    pdl_wait();
    if (MODE == kSweepAggregate) {  // every slab gets this slab's rows of the column
      for (int i = t; i < nx; i += nth) {
        const float x = spec[(size_t)i * P.sp].x;
        for (int d = 0; d < P.world; ++d) (P.slab_table[d] + (size_t)P.world * P.ncols_pad * 4)[P.rank * nx + i] = x;
      }
      return;
    }
    const int n = MODE == kSweepApply ? P.nx_global : nx;
    if (MODE == kSweepApply) {
      const float* col0 = P.slab_table[P.rank] + (size_t)P.world * P.ncols_pad * 4;
      for (int i = t; i < n; i += nth) q[i] = col0[i];
    } else {
      for (int i = t; i < n; i += nth) q[i] = spec[(size_t)i * P.sp].x;
    }
    solve_mean_mode(q, n, P.dx2, red, t);
    const int i0 = MODE == kSweepApply ? P.rank * nx : 0;
    for (int i = t; i < nx; i += nth) spec[(size_t)i * P.sp].x = (float)(q[i0 + i] * P.out_scale);
    if (MODE == kSweepApply && t == 0)  // the row after the slab (the inverse row kernel differences q forward in x)
      spec[(size_t)nx * P.sp].x = (float)(q[(i0 + nx) % n] * P.out_scale);
    return;
  }
  const int segs = P.segs;
  const int c = t % CS, seg = t / CS;
  const int strip = blockIdx.x - 1;  // CTA 0 is the column-0 solver
  const int col = strip * CS + c;
  const bool active = seg < segs && col < P.ncols;
  // per-column constants (plan tables): before the dependency wait
  float2 rho = make_float2(0.f, 0.f), gain = rho;
  if (active) { rho = __ldg(P.cst + col); gain = __ldg(P.cst + P.cst_pitch + col); }
  const int r0 = seg * RS;
  const int rows = active ? min(RS, nx - r0) : 0;
  float2* const base = spec + (size_t)r0 * P.sp + col;
  // shared arrays [re | im][column][j][lane] (+ 1 word per column against bank conflicts of the column index);
  // segment = lane m + j
  const int m = (segs + 31) / 32;
  const int cp1 = 32 * m + 1;                          // words per column
  // TMA path (P.use_tma, opt-in: see the launcher for the measurements): the strip is LOADED by tensor copies, one
  // [RS rows x CS columns] box per segment (cp.async.bulk.tensor.2d, UTMALDG); the copy engine generates the strided
  // row requests itself.  Boxes clip at the tensor's edges: partial strips and a partial last segment need no
  // special case.  The results are stored plainly (the loads showed no gain, so the store side was not built).
  const bool use_tma = P.use_tma != 0;
  float2* const tile = reinterpret_cast<float2*>(sw_raw);  // [segs][RS][CS]
  const size_t tile_floats = use_tma ? (size_t)segs * RS * CS * 2 : 0;
  float* const sE = reinterpret_cast<float*>(sw_raw) + tile_floats;  // e, later Ain
  float* const sS = sE + 2 * CS * cp1;                 // S, later Bin
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(sS + 2 * CS * cp1);
  const unsigned bar = (unsigned)__cvta_generic_to_shared(mbar);
  if (use_tma) {
    if (t == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const int at = c * cp1 + (seg % m) * 32 + seg / m;
  const int trow0 = (int)blockIdx.y * nx;  // first row of this batch member in the tensor
  // this warp's scanning job (step 2): its nine per-column constants, one per lane, fetched now (plan tables) ...
  const int wj = t >> 5, lanej = t & 31;
  float kconst = 0.0f;
  float2 entry = make_float2(0.f, 0.f);
  {
    const int colj = strip * CS + (wj >> 1);
    if (wj < 2 * CS && colj < P.ncols) kconst = job_const(P, colj, wj & 1, lanej);
  }
  pdl_wait();
  if (MODE == kSweepApply) {  // ... and the slabs' maps of that column (written by pass A of every slab), one slab per lane
    const int colj = strip * CS + (wj >> 1);
    if (wj < 2 * CS && colj < P.ncols) entry = job_entry(P, colj, wj & 1, lanej);
  }
  float2 a[RS];
  if (use_tma) {
    constexpr unsigned kBoxBytes = RS * CS * sizeof(float2);
    if (t == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)segs * kBoxBytes) : "memory");
    if (t < segs)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(tile + (size_t)t * RS * CS)),
                   "l"(&tmap), "r"(2 * strip * CS), "r"(trow0 + t * RS), "r"(bar)
                   : "memory");
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(bar) : "memory");
  }
  const float2* const tsrc = tile + (size_t)seg * RS * CS + c;
#pragma unroll
  for (int r = 0; r < RS; ++r) {
    if (FULL) a[r] = make_float2(0.f, 0.f);
    if (FULL ? active : r < rows) a[r] = use_tma ? tsrc[r * CS] : base[(size_t)r * P.sp];
  }
  // ---- 1. local forward sweep, Horner sum
  float2 ap = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < RS; ++r)
    if (FULL || r < rows) {
      ap.x = fmaf(rho.x, ap.x, a[r].x);
      ap.y = fmaf(rho.y, ap.y, a[r].y);
      a[r] = ap;
    }
  float2 h = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = RS - 1; r >= 0; --r)
    if (FULL || r < rows) {
      h.x = fmaf(rho.x, h.x, a[r].x);
      h.y = fmaf(rho.y, h.y, a[r].y);
    }
  if (active) {
    sE[at] = ap.x; sE[CS * cp1 + at] = ap.y;
    sS[at] = h.x;  sS[CS * cp1 + at] = h.y;
  }
  __syncthreads();
  // ---- 2. carried values: warp w solves (column w >> 1, part w & 1)
  {
    const int nwarps = blockDim.x >> 5;
    for (int job = wj; job < 2 * CS; job += nwarps) {
      const int cj = job >> 1, part = job & 1;
      const int colj = strip * CS + cj;
      if (colj >= P.ncols) continue;
      float kc = kconst;
      float2 en = entry;
      if (job != wj) {  // more jobs than warps (small CTAs): load this job's words now
        kc = job_const(P, colj, part, lanej);
        if (MODE == kSweepApply) en = job_entry(P, colj, part, lanej);
      }
      const float Pseg = __shfl_sync(FULLM, kc, 2), Plast = __shfl_sync(FULLM, kc, 3), close = __shfl_sync(FULLM, kc, 4);
      const float rG = __shfl_sync(FULLM, kc, 5), rGl = __shfl_sync(FULLM, kc, 6);
      float ain0 = 0.0f, bin = 0.0f, tot_e, tot_t;
      if (MODE == kSweepApply) {
        // close the cycle over the slabs (the same recurrences one level up; every lane does the few steps itself,
        // lane d holds slab d's pair of maps)
        const float Pb = __shfl_sync(FULLM, kc, 7), rGb = __shfl_sync(FULLM, kc, 8);
        const int W = P.world, me = P.rank;
        float acc = 0.0f;
        for (int j = 0; j < W; ++j) {  // slabs me, me+1, ..., me-1: the forward value entering slab `me`
          const int d = me + j < W ? me + j : me + j - W;
          acc = fmaf(Pb, acc, __shfl_sync(FULLM, en.x, d));
        }
        ain0 = acc * close;
        // backward: T_d = T0_d + rGb Ain_d with Ain of every slab from walking the ring once more; lane j keeps T
        // of slab me + j
        float ain_d = ain0, Tmine = 0.0f;
        for (int j = 0; j < W; ++j) {
          const int d = me + j < W ? me + j : me + j - W;
          const float Ed = __shfl_sync(FULLM, en.x, d), T0d = __shfl_sync(FULLM, en.y, d);
          if (lanej == j) Tmine = fmaf(rGb, ain_d, T0d);
          ain_d = fmaf(Pb, ain_d, Ed);
        }
        // slabs me, me-1, ..., me+1 (ring positions 0, W-1, ..., 1): the backward value entering slab `me` from above
        float bacc = 0.0f;
        for (int j = 0; j < W; ++j) bacc = fmaf(Pb, bacc, __shfl_sync(FULLM, Tmine, j == 0 ? 0 : W - j));
        bin = bacc * close;
      }
      carry_scan(sE + (part * CS + cj) * cp1, sS + (part * CS + cj) * cp1, segs, m, Pseg, Plast, rG, rGl, close, lanej,
                 MODE == kSweepSingle, ain0, bin, &tot_e, &tot_t);
      if (MODE == kSweepAggregate && lanej == 0) {  // this slab's maps into every slab's table (NVLink stores)
        for (int d = 0; d < P.world; ++d)
          *reinterpret_cast<float2*>(slab_entry(P.slab_table[d], P.ncols_pad, P.rank, colj, part)) = make_float2(tot_e, tot_t);
      }
      if (MODE == kSweepApply && lanej == 0 && !(colj == 0 && part == 0)) {
        // spectrum row after the slab = gain * (b entering from above)
        reinterpret_cast<float*>(spec + (size_t)nx * P.sp + colj)[part] = bin * gain_of(P, colj, part);
      }
    }
  }
  if (MODE == kSweepAggregate) return;
  __syncthreads();
  if (!active) return;
  // ---- 3. carried part of the forward sweep, backward sweep, scale, store
  // (Tried: the column tile moved by one bulk asynchronous copy per row, CS x 8 bytes each, into a bank-staggered
  // shared tile and back -- 16.5 -> 20.1 us at 2048^2, 1205 -> 1469 us at 8192^2 with 16-byte rows: the copy engine
  // is slow on such short transfers, profiles/r3d.)
  const float2 ain = make_float2(sE[at], sE[CS * cp1 + at]);
  float2 b = make_float2(sS[at], sS[CS * cp1 + at]);
  // rho^n from rho, rho^2, rho^4, rho^8 (at most 4 roundings; n is a compile-time constant after unrolling)
  const float2 p1 = rho, p2 = make_float2(p1.x * p1.x, p1.y * p1.y), p4 = make_float2(p2.x * p2.x, p2.y * p2.y),
               p8 = make_float2(p4.x * p4.x, p4.y * p4.y);
  auto power = [&](int n) {
    if (n == 16) return make_float2(p8.x * p8.x, p8.y * p8.y);
    float2 v = make_float2(1.f, 1.f);
    if (n & 8) v = p8;
    if (n & 4) v = (n & 8) ? make_float2(v.x * p4.x, v.y * p4.y) : p4;
    if (n & 2) v = (n & 12) ? make_float2(v.x * p2.x, v.y * p2.y) : p2;
    if (n & 1) v = (n & 14) ? make_float2(v.x * p1.x, v.y * p1.y) : p1;
    return v;
  };
  const bool skip_re = col == 0;  // the last CTA owns the real part of column 0
#pragma unroll
  for (int r = RS - 1; r >= 0; --r)
    if (FULL ? active : r < rows) {
      const float2 pw = power(r + 1);
      b.x = fmaf(rho.x, b.x, fmaf(pw.x, ain.x, a[r].x));
      b.y = fmaf(rho.y, b.y, fmaf(pw.y, ain.y, a[r].y));
      float2* const o = base + (size_t)r * P.sp;
      if (skip_re) o->y = gain.y * b.y;
      else *o = make_float2(gain.x * b.x, gain.y * b.y);
    }
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}

// rank-2 float32 view of the spectrum: [rows][2 sp] floats, box = [RS rows][2 CS floats]
bool make_strip_map(CUtensorMap* map, float2* spec, int sp, long long rows, int cs) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc || cs < 2) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)2 * sp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)sp * sizeof(float2)};
  const cuuint32_t box[2] = {(cuuint32_t)(2 * cs), (cuuint32_t)RS};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, spec, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CS, int MODE>
int launch_cols_sweep_t(cudaStream_t s, const ColSweepParams& P0, int batch, float2* spec) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(cols_sweep_kernel<CS, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cols_sweep_kernel<CS, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  ColSweepParams P = P0;
  int threads = ((CS * P.segs + 31) / 32) * 32;
  // Tensor-copy loads (JAXIB_SWEEP_TMA=1; off by default).  Measured on B200 (profiles/r3j): bit-identical results, no gain
  // -- 2048^2 16.58 vs 16.56 us, a 1024-row slab of 8192 columns (128-byte boxes) 27.1 vs 25.5 us, 4096^2 (32-byte
  // boxes) 282 vs 274 us per step, 8192^2 on one GPU (16-byte boxes) 1302 vs 1197 us: the strip's rows are too short
  // for the copy engine to beat plain loads.  A partial last segment clips at the tensor's last row, which only is the
  // grid's last row for batch = 1.
  const char* env_tma = getenv("JAXIB_SWEEP_TMA");
  const bool tma_on = env_tma && atoi(env_tma) == 1;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  P.use_tma = 0;
  if (tma_on && CS >= 2 && (P.nx % RS == 0 || batch == 1) && make_strip_map(&map, spec, P.sp, (long long)batch * P.nx, CS)) P.use_tma = 1;
  size_t smem = (size_t)4 * CS * (32 * ((P.segs + 31) / 32) + 1) * sizeof(float) + 16;
  if (P.use_tma) smem += (size_t)P.segs * RS * CS * sizeof(float2);
  const size_t col0 = (size_t)(MODE == kSweepApply ? P.nx_global : P.nx) * sizeof(double);  // the column-0 CTA needs nx doubles
  if (col0 > smem) smem = col0;
  if (smem > 200 * 1024) { set_error("column sweep: %zu bytes of shared memory", smem); return JAXIB_ERR_UNSUPPORTED; }
  dim3 grid((P.ncols + CS - 1) / CS + 1, batch);
  if (P.nx % RS == 0) launch_pdl(16, cols_sweep_kernel<CS, MODE, true>, grid, dim3(threads), smem, s, P, spec, map);
  else launch_pdl(16, cols_sweep_kernel<CS, MODE, false>, grid, dim3(threads), smem, s, P, spec, map);
  return check_launch("cols_sweep_kernel");
}

template <int MODE>
int launch_cols_sweep_m(cudaStream_t s, const ColSweepParams& P, int batch, float2* spec) {
  // columns per CTA: whole 32-byte sectors per row where the thread budget allows (threads = CS * segs <= 1024)
  if (P.segs <= 0 || P.segs > 1024) { set_error("column sweep: nx=%d unsupported", P.nx); return JAXIB_ERR_UNSUPPORTED; }
  int cs = 1024 / P.segs;
  if (const char* e = getenv("JAXIB_SWEEP_CS")) { const int v = atoi(e); if (v >= 1 && v * P.segs <= 1024) cs = v; }
  if (cs >= 16) return launch_cols_sweep_t<16, MODE>(s, P, batch, spec);
  if (cs >= 8) return launch_cols_sweep_t<8, MODE>(s, P, batch, spec);
  if (cs >= 4) return launch_cols_sweep_t<4, MODE>(s, P, batch, spec);
  if (cs >= 2) return launch_cols_sweep_t<2, MODE>(s, P, batch, spec);
  return launch_cols_sweep_t<1, MODE>(s, P, batch, spec);
}

}  // namespace

bool cols_sweep_supported(int nx) { return nx >= 2 && (nx + RS - 1) / RS <= 1024 && (size_t)nx * sizeof(double) <= 160 * 1024; }

// Table layout [9][pitch] float2: rho, gain, rho^RS, rho^rows_last, 1/(1 - rho^nx_global), rho (1 - rho^2RS)/(1 - rho^2),
// same for the last segment, and the last two for the whole block of nx rows (slab level).  ly_re / ly_im: y-eigenvalues of the real and imaginary part of each stored column
// (equal on periodic grids).  out_scale: the row kernels normalise by 1/nx (they were written for an unnormalised
// forward + inverse column FFT), so the solution is stored times nx.
void cols_sweep_tables(int nx, int nx_global, int ncols, double dx, const double* ly_re, const double* ly_im,
                       std::vector<float2>* out, int* pitch) {
  const int pt = (ncols + 3) & ~3;
  *pitch = pt;
  const int nxg = nx_global > 0 ? nx_global : nx;  // slab: nx rows here, nxg in the whole column
  out->assign((size_t)9 * pt, make_float2(0.f, 0.f));
  const int segs = (nx + RS - 1) / RS, last = nx - (segs - 1) * RS;
  for (int c = 0; c < ncols; ++c) {
    float v[9][2];
    for (int part = 0; part < 2; ++part) {
      const double mu = -(part ? ly_im[c] : ly_re[c]) * dx * dx;
      if (!(mu > 1e-14)) {  // singular column (mean mode): the dedicated CTA solves it; the sweep passes it through
        for (int k = 0; k < 9; ++k) v[k][part] = 0.0f;
        v[4][part] = 1.0f;
        continue;
      }
      const double r = 1.0 / (1.0 + 0.5 * mu + sqrt(mu + 0.25 * mu * mu));
      v[0][part] = (float)r;
      v[1][part] = (float)(-dx * dx * r * nxg);
      v[2][part] = (float)pow(r, RS);
      v[3][part] = (float)pow(r, last);
      v[4][part] = (float)(1.0 / (1.0 - pow(r, nxg)));
      v[5][part] = (float)(r * (1.0 - pow(r, 2 * RS)) / (1.0 - r * r));
      v[6][part] = (float)(r * (1.0 - pow(r, 2 * last)) / (1.0 - r * r));
      v[7][part] = (float)pow(r, nx);                                     // the whole block (slab level)
      v[8][part] = (float)(r * (1.0 - pow(r, 2.0 * nx)) / (1.0 - r * r));
    }
    for (int k = 0; k < 9; ++k) (*out)[(size_t)k * pt + c] = make_float2(v[k][0], v[k][1]);
  }
}

// mode 0: whole grid on this GPU; 1 / 2: passes A / B of the slab decomposition (see the header comment)
int launch_cols_sweep(cudaStream_t s, const ColSweepParams& P, int batch, float2* spec, int mode) {
  if (mode == kSweepSingle) return launch_cols_sweep_m<kSweepSingle>(s, P, batch, spec);
  if (P.world < 1 || P.world > kMaxPeers || P.nx_global != P.nx * P.world || batch != 1) {
    set_error("column sweep: bad slab geometry (nx=%d world=%d nx_global=%d)", P.nx, P.world, P.nx_global);
    return JAXIB_ERR_INVALID;
  }
  if ((size_t)P.nx_global * sizeof(double) > 160 * 1024) { set_error("column sweep: nx_global=%d too long", P.nx_global); return JAXIB_ERR_UNSUPPORTED; }
  return mode == kSweepAggregate ? launch_cols_sweep_m<kSweepAggregate>(s, P, batch, spec)
                                 : launch_cols_sweep_m<kSweepApply>(s, P, batch, spec);
}

size_t cols_sweep_slab_table_bytes(int world, int ncols_pad, int nx_global) {
  return ((size_t)world * ncols_pad * 4 + (size_t)nx_global) * sizeof(float);
}

}  // namespace jaxib
