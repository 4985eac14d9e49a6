// Shared-memory mixed-radix FFT building blocks (radices 2, 3, 4, 5, 8, 16), hand-written.
//
// Forward transforms are decimation-in-frequency, in place, leaving the spectrum in mixed-radix
// digit-reversed order; inverse transforms are the exact pass-by-pass inverse (decimation in time)
// and take digit-reversed input back to natural order.  Because the Poisson solve transforms
// forward, scales by the inverse Laplacian eigenvalue and transforms back, no reordering pass is
// ever needed: the diagonal is applied in digit-reversed positions through a position->frequency
// table held in the plan.
//
// Pass algebra (length-L block, radix R, b < L/R, input r <-> output q at the same location
// base + b + r*(L/R)):  y_q[b] = W_L^{b q} * DFT_R(x[b + (L/R) r])[q]  and sub-block q then holds the
// length-L/R problem for the frequencies congruent to q mod R, so
//   pos(k) = (k mod R) * (L/R) + pos'(k div R).
#pragma once
#include "common.cuh"

namespace jaxib {

// one padding slot per 16 complex values: breaks the power-of-two strides of the late passes
__host__ __device__ __forceinline__ int padi(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// ---- forward DFTs of length R in registers (W = exp(-2 pi i / R)) --------------------------------
template <int R>
struct Dft;

template <>
struct Dft<2> {
  __device__ static __forceinline__ void run(float2* v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <>
struct Dft<3> {
  __device__ static __forceinline__ void run(float2* v) {
    const float s = 0.86602540378443864676f;
    float2 t1 = cadd(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    float2 d = cscale(csub(v[1], v[2]), s);
    v[0] = cadd(v[0], t1);
    v[1] = make_float2(m.x + d.y, m.y - d.x);  // m - i d
    v[2] = make_float2(m.x - d.y, m.y + d.x);  // m + i d
  }
};

template <>
struct Dft<4> {
  __device__ static __forceinline__ void run(float2* v) {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
  }
};

template <>
struct Dft<5> {
  __device__ static __forceinline__ void run(float2* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    float2 t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    v[0] = cadd(v[0], cadd(t1, t2));
    v[1] = make_float2(m1.x + n1.y, m1.y - n1.x);  // m1 - i n1
    v[4] = make_float2(m1.x - n1.y, m1.y + n1.x);  // m1 + i n1
    v[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
    v[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
  }
};

template <>
struct Dft<8> {
  __device__ static __forceinline__ void run(float2* v) {
    const float h = 0.70710678118654752440f;
    float2 a[4] = {v[0], v[2], v[4], v[6]};
    float2 b[4] = {v[1], v[3], v[5], v[7]};
    Dft<4>::run(a);
    Dft<4>::run(b);
    b[1] = make_float2(h * (b[1].x + b[1].y), h * (b[1].y - b[1].x));    // * (1 - i)/sqrt2
    b[2] = mul_mi(b[2]);
    b[3] = make_float2(h * (b[3].y - b[3].x), -h * (b[3].x + b[3].y));   // * (-1 - i)/sqrt2
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = cadd(a[k], b[k]);
      v[k + 4] = csub(a[k], b[k]);
    }
  }
};

template <>
struct Dft<16> {
  __device__ static __forceinline__ void run(float2* v) {
    // n = 4 n1 + n2, k = k1 + 4 k2
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin(pi/8)
    const float h = 0.70710678118654752440f;
    float2 y[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      float2 t[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
      Dft<4>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < 4; ++k1) y[n2][k1] = t[k1];
    }
    // twiddles W16^{n2 k1}
    const float2 w1 = make_float2(c1, -s1), w2 = make_float2(h, -h), w3 = make_float2(s1, -c1);
    const float2 w6 = make_float2(-h, -h), w9 = make_float2(-c1, s1);
    y[1][1] = cmul(y[1][1], w1); y[1][2] = cmul(y[1][2], w2); y[1][3] = cmul(y[1][3], w3);
    y[2][1] = cmul(y[2][1], w2); y[2][2] = mul_mi(y[2][2]);   y[2][3] = cmul(y[2][3], w6);
    y[3][1] = cmul(y[3][1], w3); y[3][2] = cmul(y[3][2], w6); y[3][3] = cmul(y[3][3], w9);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      float2 t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
      Dft<4>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
    }
  }
};

template <>
struct Dft<32> {
  __device__ static __forceinline__ void run(float2* v) {
    // even / odd split: X[k] = E[k] + W32^k O[k], X[k + 16] = E[k] - W32^k O[k]
    float2 a[16], b[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) { a[n] = v[2 * n]; b[n] = v[2 * n + 1]; }
    Dft<16>::run(a);
    Dft<16>::run(b);
    const float c[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                         0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,
                         -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                         -0.92387953251128675613f, -0.98078528040323044913f};
    const float sn[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                          0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                          0.98078528040323044913f, 1.0f, 0.98078528040323044913f, 0.92387953251128675613f,
                          0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                          0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float2 t = k == 0 ? b[0] : cmul(b[k], make_float2(c[k], -sn[k]));
      v[k] = cadd(a[k], t);
      v[k + 16] = csub(a[k], t);
    }
  }
};

// One radix-R pass over `ntr` independent length-n transforms stored at buf + tr*pitch (padded).
// L is the current block length; tw is the table of W_N^m with N = n * tw_stride.
template <int R, bool INV>
__device__ __forceinline__ void fft_pass(float2* __restrict__ buf, int ntr, int pitch, int n, int L,
                                         const float2* __restrict__ tw, int tw_stride, int tid, int nthreads) {
  const int bpt = n / R;   // butterflies per transform
  const int sub = L / R;   // stride between the R inputs
  const int tstep = (n / L) * tw_stride;
  const int total = ntr * bpt;
  for (int t = tid; t < total; t += nthreads) {
    const int tr = t / bpt, bb = t - tr * bpt;
    const int blk = bb / sub, b = bb - blk * sub;
    float2* base = buf + tr * pitch;
    const int p0 = blk * L + b;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = base[padi(p0 + r * sub)];
    if (!INV) {
      Dft<R>::run(v);
      if (sub > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(tw + b * r * tstep));
      }
    } else {
      if (sub > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmulc(v[r], __ldg(tw + b * r * tstep));
      }
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
      Dft<R>::run(v);
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) base[padi(p0 + r * sub)] = v[r];
  }
}

template <bool INV>
__device__ __forceinline__ void fft_pass_dispatch(int R, float2* buf, int ntr, int pitch, int n, int L,
                                                  const float2* tw, int tw_stride, int tid, int nthreads) {
  switch (R) {
    case 2: fft_pass<2, INV>(buf, ntr, pitch, n, L, tw, tw_stride, tid, nthreads); break;
    case 3: fft_pass<3, INV>(buf, ntr, pitch, n, L, tw, tw_stride, tid, nthreads); break;
    case 4: fft_pass<4, INV>(buf, ntr, pitch, n, L, tw, tw_stride, tid, nthreads); break;
    case 5: fft_pass<5, INV>(buf, ntr, pitch, n, L, tw, tw_stride, tid, nthreads); break;
    case 8: fft_pass<8, INV>(buf, ntr, pitch, n, L, tw, tw_stride, tid, nthreads); break;
    case 16: fft_pass<16, INV>(buf, ntr, pitch, n, L, tw, tw_stride, tid, nthreads); break;
  }
}

// Forward DIF: natural order in, digit-reversed out.  Ends with a __syncthreads().
__device__ __forceinline__ void fft_forward(const FftPlan& pl, float2* buf, int ntr, int pitch, const float2* tw,
                                            int tw_stride) {
  int L = pl.n;
  for (int s = 0; s < pl.nstages; ++s) {
    fft_pass_dispatch<false>(pl.radix[s], buf, ntr, pitch, pl.n, L, tw, tw_stride, threadIdx.x, blockDim.x);
    L /= pl.radix[s];
    __syncthreads();
  }
}

// Inverse (unnormalised): digit-reversed in, natural order out.  Ends with a __syncthreads().
__device__ __forceinline__ void fft_inverse(const FftPlan& pl, float2* buf, int ntr, int pitch, const float2* tw,
                                            int tw_stride) {
  int L = 1;
  for (int s = pl.nstages - 1; s >= 0; --s) {
    L *= pl.radix[s];
    fft_pass_dispatch<true>(pl.radix[s], buf, ntr, pitch, pl.n, L, tw, tw_stride, threadIdx.x, blockDim.x);
    __syncthreads();
  }
}

}  // namespace jaxib
