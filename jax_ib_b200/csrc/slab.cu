// Slab-decomposed step over peer-mapped GPU memory (jaxib_slab_step, see include/jaxib_b200.h).
//
// One simulation, `world` GPUs, one process per GPU.  Every slab keeps nloc = nx / world rows (+ halo
// rows of its neighbours) and the buffers every OTHER slab writes into are mapped into each process
// (CUDA VMM peer mappings, made by the host side: jax_ib_b200/slab.py uses torch symmetric memory).
// The kernels of a step store their results where the NEXT kernel will read them -- on whichever GPU
// that is -- so the "collectives" of the decomposition are NVLink stores issued tile by tile from inside
// the compute kernels, and all that is left between the phases is a flag barrier:
//
//   phase 0  K1 predictor on the extended slab;  K2 forces of the markers near the slab (the halo is
//            2R + 4 rows deep, so their interpolation windows are local: no exchange)  | --
//   phase 1  K3 spreading;  K4 divergence + row FFT: spectrum column block d of every row is stored
//            into slab d's column buffer (the all-to-all transpose)                  | barrier 1
//   phase 2  K5 column FFT x 1/lambda x inverse on the local columns, in place; a push kernel then
//            stores row R of the result into the spectrum buffer of slab R / nloc in whole row chunks
//            (the transpose back; the first row of every slab also goes to the previous slab: K6
//            needs q of the next row)                                                | barrier 2
//   phase 3  K6 inverse row FFT + projection + pressure update: owned rows in place, boundary rows
//            also into the neighbours' halo rows (the halo exchange)                 | barrier 3
//
// Since round 2 the default x-solve is the column SWEEP (poisson_sweep.cu) and phases 1-2 read:
//   phase 1  K3; K4 into the LOCAL spectrum; sweep pass A: the slab's carried-value maps (2 floats per
//            column and part) + its rows of column 0 into every slab's table      | barrier 1
//   phase 2  sweep pass B on the local rows (also writes the spectrum row after the slab)  | --
// i.e. no transposes, two barriers per step.  JAXIB_SLAB_FFT=1 selects the transpose path above.
//
// Barrier = one tiny kernel per GPU: a system-scope release store of the step number into every peer's
// flag word, then an acquire spin on the own flag words (bounded: a lost peer sets an error word
// instead of hanging the GPU).  Reuse hazards: a buffer written in phase k of step n+1 was last read
// in phase k+1 of step n, and every slab passes barrier k+1.. of step n only after that read finished.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace jaxib {
namespace {

struct BarrierArgs {
  uint32_t* flags[kMaxPeers];
  int world, rank, slot;
  uint32_t epoch;
  long long timeout_cycles;
};

__global__ void slab_barrier_kernel(BarrierArgs a) {
  pdl_trigger();
  pdl_wait();
  const int t = threadIdx.x;
  if (t >= a.world) return;
  // everything this GPU stored before (earlier kernels of the stream, local or over NVLink) becomes
  // visible system-wide before the flag does
  __threadfence_system();
  uint32_t* remote = a.flags[t] + a.slot * kMaxPeers + a.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(a.epoch) : "memory");
  const uint32_t* mine = a.flags[a.rank] + a.slot * kMaxPeers + t;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - a.epoch) >= 0) break;
    if (clock64() - t0 > a.timeout_cycles) {  // a peer never arrived: record it and let the stream drain
      a.flags[a.rank][4 * kMaxPeers] = a.epoch;
      break;
    }
  }
  __threadfence_system();
}

int enqueue_barrier(cudaStream_t s, const jaxib_slab& sl, int slot, uint32_t epoch) {
  BarrierArgs a;
  memset(&a, 0, sizeof(a));
  for (int d = 0; d < sl.world; ++d) a.flags[d] = sl.flags[d];
  a.world = sl.world; a.rank = sl.rank; a.slot = slot; a.epoch = epoch;
  // Give-up time of the spin: long enough for any host-side hiccup of a peer (first kinematics evaluation,
  // GC, a profiler sync), short enough that a dead peer cannot hang the GPU.  JAXIB_SLAB_TIMEOUT_MS overrides.
  // A give-up is FATAL for the results: the error word stays set until the host clears it and every Python
  // entry point that hands fields back raises on it (slab.py: PeerRank.check).
  static long long cycles = 0;
  if (!cycles) {
    const char* e = getenv("JAXIB_SLAB_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 20000;
    cycles = (ms < 1 ? 1 : ms) * 1900000LL;  // ~1.9 GHz
  }
  a.timeout_cycles = cycles;
  launch_pdl(64, slab_barrier_kernel, dim3(1), dim3(32), 0, s, a);
  return check_launch("slab_barrier_kernel");
}

// side stream + events for the overlapped push (one per process: one GPU per process)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
};
SideStream& side_stream() {
  static SideStream st;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking) != cudaSuccess) st.stream = nullptr;
    if (st.stream && (cudaEventCreateWithFlags(&st.ev_a, cudaEventDisableTiming) != cudaSuccess ||
                      cudaEventCreateWithFlags(&st.ev_b, cudaEventDisableTiming) != cudaSuccess ||
                      cudaEventCreateWithFlags(&st.ev_c, cudaEventDisableTiming) != cudaSuccess ||
                      cudaEventCreateWithFlags(&st.ev_d, cudaEventDisableTiming) != cudaSuccess))
      st.stream = nullptr;
    cudaGetLastError();
  }
  return st;
}

unsigned magic_of(int d) { return (unsigned)((0x100000000ULL + (unsigned)d - 1) / (unsigned)d); }

// The transpose back: row R of this slab's solved columns [nx][cpr] goes to the spectrum buffer
// [nloc + 1][sp] of slab R / nloc at column offset rank * cpr; the first row of every slab is also the
// row after the last one of the previous slab (K6 differences q forward in x).  One warp moves one row
// chunk (cpr * 8 bytes, contiguous on both sides) with 16-byte accesses, all loads before the stores:
// whole chunks make full-size NVLink write packets (K5's own last pass stores 16 bytes per row, and a
// pull by K6 was measured 2-3x slower: NVLink reads are latency-bound, posted writes are not).
struct PushArgs {
  float2* spec[kMaxPeers];
  const float2* cols;
  int world, nloc, nx, cpr, sp, col_offset;
  unsigned magic;
  int c_begin, c_end;  // column sub-range of the cpr local columns moved by this launch
};

constexpr int kPushMaxVec = 12;  // float4 per lane per row chunk: cpr <= 2 * 32 * 12 = 768 columns per pass

__global__ void __launch_bounds__(256) slab_push_kernel(PushArgs a) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = (a.c_end - a.c_begin) >> 1;  // float4 per row chunk
  for (int R = warp; R < a.nx; R += nwarps) {
    const int d = (int)__umulhi((unsigned)R, a.magic);
    const int l = R - d * a.nloc;
    const float4* src = reinterpret_cast<const float4*>(a.cols + (size_t)R * a.cpr + a.c_begin);
    float4* dst = reinterpret_cast<float4*>(a.spec[d] + (size_t)l * a.sp + a.col_offset + a.c_begin);
    float4* dst2 = nullptr;
    if (l == 0)
      dst2 = reinterpret_cast<float4*>(a.spec[d == 0 ? a.world - 1 : d - 1] + (size_t)a.nloc * a.sp + a.col_offset + a.c_begin);
    for (int base = 0; base < nvec; base += 32 * kPushMaxVec) {
      float4 v[kPushMaxVec];
#pragma unroll
      for (int k = 0; k < kPushMaxVec; ++k) {
        const int e = base + lane + 32 * k;
        if (e < nvec) v[k] = __ldcs(src + e);
      }
#pragma unroll
      for (int k = 0; k < kPushMaxVec; ++k) {
        const int e = base + lane + 32 * k;
        if (e < nvec) {
          dst[e] = v[k];
          if (dst2) dst2[e] = v[k];
        }
      }
    }
  }
}

}  // namespace
}  // namespace jaxib

using namespace jaxib;

extern "C" int jaxib_slab_step(jaxib_stream_t stream, const jaxib_plan* plan_rows, const jaxib_plan* plan_cols,
                               const jaxib_grid* grid, const jaxib_slab* slab, const jaxib_bodies* bodies,
                               const float* body_table, int32_t nsteps, int32_t phase_lo, int32_t phase_hi,
                               int32_t barriers, uint32_t* epoch, float* u_star, float* v_star) {
  int rc = check_grid(grid);
  if (rc) return rc;
  if (!plan_rows || !plan_cols || !slab || !epoch || !u_star || !v_star) {
    set_error("slab_step: null argument");
    return JAXIB_ERR_INVALID;
  }
  const jaxib_slab& sl = *slab;
  const int W = sl.world, me = sl.rank;
  if (W < 1 || W > JAXIB_MAX_PEERS || me < 0 || me >= W) { set_error("slab_step: bad world/rank %d/%d", W, me); return JAXIB_ERR_INVALID; }
  if (grid->bc != JAXIB_BC_PERIODIC || grid->batch != 1) {
    set_error("slab_step: periodic grids with batch = 1 only");
    return JAXIB_ERR_UNSUPPORTED;
  }
  const int nx = grid->nx, ny = grid->ny, n2 = ny / 2;
  if (nx % W) { set_error("slab_step: nx=%d not divisible by %d slabs", nx, W); return JAXIB_ERR_INVALID; }
  const int nloc = nx / W, halo = sl.halo, cpr = sl.cpr, sp = W * cpr;
  const int R = grid->ib_window > 0 ? grid->ib_window : 6;
  if (halo < 2 * R + 4 || halo > nloc) { set_error("slab_step: halo=%d must lie in [2 * ib_window + 4, nx / world]", halo); return JAXIB_ERR_INVALID; }
  if (cpr <= 0 || cpr % 8 || sp < n2 + 4) { set_error("slab_step: cpr=%d must be a multiple of 8 with world * cpr >= ny/2 + 4", cpr); return JAXIB_ERR_INVALID; }
  const GridF& gr = plan_grid(plan_rows);
  const GridF& gc = plan_grid(plan_cols);
  if (gr.nx != nloc || gr.ny != ny || gr.nx_global != nx || gc.nx != nx || gc.ny != ny) {
    set_error("slab_step: plans do not match the decomposition (rows plan %dx%d nx_global=%d, cols plan %dx%d)", gr.nx,
              gr.ny, gr.nx_global, gc.nx, gc.ny);
    return JAXIB_ERR_INVALID;
  }
  // Column sweep (poisson_sweep.cu): the x-direction is solved slab by slab with one small exchange of carried
  // values; JAXIB_SLAB_FFT=1 keeps the transpose + column FFT path (A/B runs, and the only path for plans
  // without the sweep tables).
  const char* env_fft = getenv("JAXIB_SLAB_FFT");
  const bool force_fft = env_fft && atoi(env_fft) == 1;
  const bool sweep = !force_fft && plan_col_sweep(plan_rows);
  if (!plan_fast_rows(plan_rows) || (!sweep && !plan_fast_cols(plan_cols))) {
    set_error("slab_step: needs ny/2 in {512..4096} and nx in {512..8192} (powers of two)");
    return JAXIB_ERR_UNSUPPORTED;
  }
  if (sweep && cols_sweep_slab_table_bytes(W, plan_col_sweep_pitch(plan_rows), nx) > (size_t)nx * cpr * sizeof(float2)) {
    set_error("slab_step: column buffer too small for the sweep tables");
    return JAXIB_ERR_INVALID;
  }
  const bool has_ib = bodies && bodies->n_markers > 0;
  if (has_ib && !body_table) { set_error("slab_step: bodies given without body_table"); return JAXIB_ERR_INVALID; }
  for (int d = 0; d < W; ++d)
    if (!sl.u[d] || !sl.v[d] || !sl.p[d] || !sl.colbuf[d] || !sl.specbuf[d] || !sl.flags[d] || (has_ib && !sl.markers[d])) {
      set_error("slab_step: null buffer for slab %d", d);
      return JAXIB_ERR_INVALID;
    }
  if (phase_lo < 0 || phase_hi > 3 || phase_lo > phase_hi || nsteps < 0) { set_error("slab_step: bad phase range / nsteps"); return JAXIB_ERR_INVALID; }

  cudaStream_t s = (cudaStream_t)stream;
  const int nxl = nloc + 2 * halo, own_lo = me * nloc, row0 = own_lo - halo;
  // K1 sees the extended slab as a periodic grid of its own
  jaxib_grid g_ext = *grid;
  g_ext.nx = nxl; g_ext.row0 = row0; g_ext.own_lo = g_ext.own_hi = 0; g_ext.nx_global = 0;
  const GridF gf_ext = make_gridf(g_ext);
  // the IB kernels see the un-wrapped rows (no periodic image of a marker: raw xp - X in the reference)
  const int seg_a = row0 < 0 ? -row0 : 0;
  const int seg_b = nx - row0 < nxl ? nx - row0 : nxl;
  jaxib_grid g_ib = *grid;
  g_ib.nx = seg_b - seg_a; g_ib.row0 = row0 + seg_a; g_ib.nx_global = 0;
  // forces are needed for every marker whose window touches u** rows [own_lo-1, own_hi) or v** rows
  // [own_lo, own_hi): u-cell row in [own_lo - R - 2, own_hi + R); their interpolation windows stay inside
  // the 2R + 4 halo rows, so every slab computes the forces it needs itself -- no exchange of forces
  g_ib.own_lo = me == 0 ? -2147483647 : own_lo - R - 2;
  g_ib.own_hi = me == W - 1 ? 2147483647 : own_lo + nloc + R;
  // first slab of several: the seam row nx-1 is fed by markers near the TOP of the domain; their windows
  // lie in the wrapped lower halo (global rows nx - halo .. nx - 1)
  jaxib_grid g_top = *grid;
  g_top.nx = halo; g_top.row0 = nx - halo; g_top.nx_global = 0;
  g_top.own_lo = nx - R - 3; g_top.own_hi = 2147483647;
  const GridF gf_top = make_gridf(g_top);
  const GridF gf_ib = make_gridf(g_ib);
  jaxib_grid g_seam = *grid;  // first slab: the wrapped row nx-1 below its first owned row (K4 reads it)
  g_seam.nx = 1; g_seam.row0 = nx - 1; g_seam.own_lo = g_seam.own_hi = 0; g_seam.nx_global = 0;
  const GridF gf_seam = make_gridf(g_seam);

  float* u = sl.u[me]; float* v = sl.v[me]; float* p = sl.p[me];
  const size_t own_off = (size_t)halo * ny;
  SpecScatter sc; memset(&sc, 0, sizeof(sc));
  PushArgs pa; memset(&pa, 0, sizeof(pa));
  HaloOut ho; memset(&ho, 0, sizeof(ho));
  sc.cpr = cpr; sc.magic = magic_of(cpr); sc.row_offset = own_lo;
  pa.cols = reinterpret_cast<const float2*>(sl.colbuf[me]);
  pa.world = W; pa.nloc = nloc; pa.nx = nx; pa.cpr = cpr; pa.sp = sp; pa.col_offset = me * cpr; pa.magic = magic_of(nloc);
  for (int d = 0; d < W; ++d) {
    sc.base[d] = reinterpret_cast<float2*>(sl.colbuf[d]);
    pa.spec[d] = reinterpret_cast<float2*>(sl.specbuf[d]);
  }
  const int prev = (me + W - 1) % W, next = (me + 1) % W;
  ho.halo = halo;
  float* const* fields[3] = {sl.u, sl.v, sl.p};
  for (int f = 0; f < 3; ++f) {
    ho.lo[f] = fields[f][prev] + (size_t)(halo + nloc) * ny;  // previous slab's upper halo
    ho.hi[f] = fields[f][next];                               // next slab's lower halo
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int push_ctas = 4 * sms;
  // split of the column transform into two launches of whole waves (0: one launch)
  int split_cols = 0;
  SideStream& side = side_stream();
  if (!getenv("JAXIB_NO_SPLIT") && side.stream) {
    int cols_per_cta = 2, per_sm = 1;
    plan_cols_geometry(plan_cols, &cols_per_cta, &per_sm);
    const int ctas = (cpr + cols_per_cta - 1) / cols_per_cta, slots = sms * per_sm;
    const int waves = (ctas + slots - 1) / slots;
    if (waves >= 2) split_cols = ((waves + 1) / 2) * slots * cols_per_cta;
    if (split_cols >= cpr) split_cols = 0;
  }
  const size_t state_stride = has_ib ? (size_t)bodies->n_bodies * JAXIB_BODY_STATE_STRIDE : 0;
  // host hint: the bodies (and their markers) that come near this slab during the block; everything else is skipped
  BodyRange range = {sl.body_first, sl.body_count, sl.marker_first, sl.marker_count};
  const BodyRange* rg = nullptr;
  if (has_ib && sl.body_count > 0 && sl.marker_count > 0) {
    if (sl.body_first < 0 || sl.body_first + sl.body_count > bodies->n_bodies || sl.marker_first < 0 ||
        sl.marker_first + sl.marker_count > bodies->n_markers) {
      set_error("slab_step: body / marker range hint outside the bodies");
      return JAXIB_ERR_INVALID;
    }
    rg = &range;
  }

  for (int n = 0; n < nsteps; ++n) {
    const uint32_t ep = ++(*epoch);
    const float* state = has_ib ? body_table + n * state_stride : nullptr;
    for (int ph = phase_lo; ph <= phase_hi; ++ph) {
      switch (ph) {
        case 0:
          if ((rc = launch_explicit(s, gf_ext, u, v, p, nullptr, u_star, v_star, true))) return rc;
          if (has_ib && own_lo == 0 && W > 1 && !sl.seam_idle) {
            // first slab of several: forces of the markers at the top of the domain (disjoint marker set) and,
            // in phase 1, the seam row they feed -- on the side stream, next to the main IB kernels, so that
            // slab 0 is not 20 us behind the others at every barrier
            cudaStream_t st = side.stream ? side.stream : s;
            if (side.stream) {
              cudaEventRecord(side.ev_c, s);  // K1 done
              cudaStreamWaitEvent(side.stream, side.ev_c, 0);
            }
            if ((rc = launch_ib_marker_force(st, gf_top, *bodies, state, u_star, v_star, sl.markers[me]))) return rc;
          }
          if (has_ib && (rc = launch_ib_marker_force(s, gf_ib, *bodies, state, u_star + (size_t)seg_a * ny,
                                                     v_star + (size_t)seg_a * ny, sl.markers[me], rg)))
            return rc;
          break;
        case 1:
          if (has_ib) {
            if ((rc = launch_ib_spread_bodies(s, gf_ib, *bodies, state, sl.markers[me], u_star + (size_t)seg_a * ny,
                                              v_star + (size_t)seg_a * ny, rg)))
              return rc;
            if (own_lo == 0 && !sl.seam_idle) {
              const bool aside = W > 1 && side.stream;  // the seam row only depends on the top-marker forces
              if ((rc = launch_ib_spread_bodies(aside ? side.stream : s, gf_seam, *bodies, state, sl.markers[me],
                                                u_star + (size_t)(halo - 1) * ny, v_star + (size_t)(halo - 1) * ny)))
                return rc;
              if (aside) {
                cudaEventRecord(side.ev_d, side.stream);
                cudaStreamWaitEvent(s, side.ev_d, 0);
              }
            }
          }
          if (sweep) {
            // row FFT into the LOCAL spectrum, then pass A of the column sweep: this slab's carried-value maps
            // (two floats per column and part) and its rows of column 0 go into every slab's table
            if ((rc = launch_rows_fwd(s, plan_rows, u_star + own_off, v_star + own_off, sl.specbuf[me], sp, nullptr))) return rc;
            if ((rc = launch_cols_sweep_slab(s, plan_rows, sl.specbuf[me], sp, W, me, sl.colbuf, 1))) return rc;
            break;
          }
          if ((rc = launch_rows_fwd(s, plan_rows, u_star + own_off, v_star + own_off, sl.specbuf[me], sp, &sc))) return rc;
          break;
        case 2:
          if (sweep) {  // pass B: close the cycle over the slabs, sweep, store rows 0 .. nloc of the local spectrum
            if ((rc = launch_cols_sweep_slab(s, plan_rows, sl.specbuf[me], sp, W, me, sl.colbuf, 2))) return rc;
            break;
          }
          if (split_cols > 0) {
            // K5 in two launches of whole waves; the push of the first part runs on a side stream under
            // the second K5 launch (the push is NVLink-bound and needs few SMs, K5 is latency-bound)
            if ((rc = launch_cols(s, plan_cols, sl.colbuf[me], cpr, split_cols, me * cpr, 0))) return rc;
            cudaEventRecord(side.ev_a, s);
            if ((rc = launch_cols(s, plan_cols, sl.colbuf[me], cpr, cpr - split_cols, me * cpr, split_cols))) return rc;
            cudaStreamWaitEvent(side.stream, side.ev_a, 0);
            pa.c_begin = 0; pa.c_end = split_cols;
            slab_push_kernel<<<push_ctas, 256, 0, side.stream>>>(pa);
            cudaEventRecord(side.ev_b, side.stream);
            pa.c_begin = split_cols; pa.c_end = cpr;
            launch_pdl(64, slab_push_kernel, dim3(push_ctas), dim3(256), 0, s, pa);
            cudaStreamWaitEvent(s, side.ev_b, 0);
          } else {
            if ((rc = launch_cols(s, plan_cols, sl.colbuf[me], cpr, cpr, me * cpr))) return rc;
            pa.c_begin = 0; pa.c_end = cpr;
            launch_pdl(64, slab_push_kernel, dim3(push_ctas), dim3(256), 0, s, pa);
          }
          if ((rc = check_launch("slab_push_kernel"))) return rc;
          break;
        case 3:
          if ((rc = launch_rows_inv(s, plan_rows, sl.specbuf[me], sp, u_star + own_off, v_star + own_off, p + own_off,
                                    u + own_off, v + own_off, p + own_off, nullptr, &ho)))
            return rc;
          break;
      }
      if (barriers && ph != 0 && !(sweep && ph == 2) && (rc = enqueue_barrier(s, sl, ph, ep))) return rc;
    }
  }
  return JAXIB_OK;
}
