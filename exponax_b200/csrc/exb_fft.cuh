// Shared-memory Stockham autosort FFT, mixed radix {8,4,2,3,5} + generic prime radix, any N.
// One CTA transforms `nlines` independent complex lines of length N that live in shared
// memory; data ping-pongs between two buffers, the result pointer is returned.
//
// Twiddles come from ONE table of the N-th roots of unity w[j] = exp(-2*pi*i*j/N) (computed in
// double precision on the host, rounded once to T), so no on-device sincos and no recurrence
// error.  DIR = -1 forward (exp(-i..)), +1 inverse (conjugated twiddles, unnormalised).
#pragma once
#include "exb_common.cuh"

namespace exb {

template <class T, int DIR> __device__ __forceinline__ cpx<T> twd(cpx<T> w) {
  return DIR < 0 ? w : cpx<T>(w.x, -w.y);
}
// multiply by -i (forward) / +i (inverse)
template <class T, int DIR> __device__ __forceinline__ cpx<T> rot90(cpx<T> a) {
  return DIR < 0 ? cpx<T>(a.y, -a.x) : cpx<T>(-a.y, a.x);
}

template <class T, int DIR> __device__ __forceinline__ void dft2(cpx<T>& a, cpx<T>& b) {
  cpx<T> t = a - b;
  a = a + b;
  b = t;
}

template <class T, int DIR> __device__ __forceinline__ void dft3(cpx<T>* v) {
  const T c = (T)-0.5, s = (T)0.86602540378443864676;
  cpx<T> t1 = v[1] + v[2];
  cpx<T> t2 = v[1] - v[2];
  cpx<T> m = axpy(c, t1, v[0]);
  // forward: X1 = m - i*s*t2 ; X2 = m + i*s*t2
  cpx<T> r = rot90<T, DIR>(s * t2);
  v[0] = v[0] + t1;
  v[1] = m + r;
  v[2] = m - r;
}

template <class T, int DIR> __device__ __forceinline__ void dft4(cpx<T>* v) {
  cpx<T> a0 = v[0] + v[2], a1 = v[0] - v[2];
  cpx<T> b0 = v[1] + v[3], b1 = rot90<T, DIR>(v[1] - v[3]);
  v[0] = a0 + b0;
  v[1] = a1 + b1;
  v[2] = a0 - b0;
  v[3] = a1 - b1;
}

template <class T, int DIR> __device__ __forceinline__ void dft5(cpx<T>* v) {
  // Winograd-style radix 5
  const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
  const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
  cpx<T> a1 = v[1] + v[4], b1 = v[1] - v[4];
  cpx<T> a2 = v[2] + v[3], b2 = v[2] - v[3];
  cpx<T> m1 = axpy(c2, a2, axpy(c1, a1, v[0]));
  cpx<T> m2 = axpy(c1, a2, axpy(c2, a1, v[0]));
  cpx<T> n1 = rot90<T, DIR>(axpy(s2, b2, s1 * b1));
  cpx<T> n2 = rot90<T, DIR>(axpy(-s1, b2, s2 * b1));
  v[0] = v[0] + a1 + a2;
  v[1] = m1 + n1;
  v[4] = m1 - n1;
  v[2] = m2 + n2;
  v[3] = m2 - n2;
}

template <class T, int DIR> __device__ __forceinline__ void dft8(cpx<T>* v) {
  const T h = (T)0.70710678118654752440;
  // three radix-2 layers (decimation in frequency), output in natural order
  cpx<T> a[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = v[i] + v[i + 4];
    a[i + 4] = v[i] - v[i + 4];
  }
  // twiddles on the lower half: w8^i, i=0..3 (forward: exp(-i*pi*i/4))
  //   a5 * w8   = h (a5 + rot90(a5)),   a7 * w8^3 = -h (a7 - rot90(a7))      (rot90 = * -i forward, * +i inverse)
  // written with whole-complex operators: for T = float each line is one packed add (the rotation is an operand
  // modifier of FADD2) + one packed multiply
  a[5] = h * (a[5] + rot90<T, DIR>(a[5]));
  a[6] = rot90<T, DIR>(a[6]);
  a[7] = (-h) * (a[7] - rot90<T, DIR>(a[7]));
  cpx<T> b[8];
  b[0] = a[0] + a[2];
  b[2] = a[0] - a[2];
  b[1] = a[1] + a[3];
  b[3] = rot90<T, DIR>(a[1] - a[3]);
  b[4] = a[4] + a[6];
  b[6] = a[4] - a[6];
  b[5] = a[5] + a[7];
  b[7] = rot90<T, DIR>(a[5] - a[7]);
  v[0] = b[0] + b[1];
  v[4] = b[0] - b[1];
  v[2] = b[2] + b[3];
  v[6] = b[2] - b[3];
  v[1] = b[4] + b[5];
  v[5] = b[4] - b[5];
  v[3] = b[6] + b[7];
  v[7] = b[6] - b[7];
}

template <class T, int R, int DIR> __device__ __forceinline__ void dftR(cpx<T>* v) {
  if (R == 2) dft2<T, DIR>(v[0], v[1]);
  if (R == 3) dft3<T, DIR>(v);
  if (R == 4) dft4<T, DIR>(v);
  if (R == 5) dft5<T, DIR>(v);
  if (R == 8) dft8<T, DIR>(v);
}

// One Stockham butterfly of radix R at position j (0 <= j < NR = N/R); `Ns` = product of the radices
// already applied, k = j % Ns, tunit = N / (Ns * R).  Element i of the line is at base[i * istr].
template <class T, int R, int DIR>
__device__ __forceinline__ void stockham_bfly(const cpx<T>* __restrict__ in, cpx<T>* __restrict__ out,
                                              int j, int k, int Ns, int NR, int tunit, int istr,
                                              const cpx<T>* __restrict__ tw) {
  cpx<T> v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = in[(size_t)(j + r * NR) * istr];
  if (Ns > 1) {
    const int tstep = tunit * k;
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = v[r] * twd<T, DIR>(tw[r * tstep]);
  }
  dftR<T, R, DIR>(v);
  const int j0 = (j - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) out[(size_t)(j0 + r * Ns) * istr] = v[r];
}

// generic (prime) radix p: O(p^2) per butterfly, reads its inputs straight from shared memory
template <class T, int DIR>
__device__ inline void stockham_bfly_generic(const cpx<T>* __restrict__ in, cpx<T>* __restrict__ out,
                                             int j, int Ns, int N, int p, int istr,
                                             const cpx<T>* __restrict__ tw) {
  const int NR = N / p;
  const int k = j % Ns;
  const int tstep = (N / (Ns * p)) * k;
  const int pstep = N / p;
  const int j0 = (j - k) * p + k;
  for (int m = 0; m < p; ++m) {
    cpx<T> acc((T)0, (T)0);
    for (int r = 0; r < p; ++r) {
      long long idx = ((long long)r * tstep + (long long)r * m * pstep) % N;
      acc = acc + in[(size_t)(j + r * NR) * istr] * twd<T, DIR>(tw[idx]);
    }
    out[(size_t)(j0 + m * Ns) * istr] = acc;
  }
}

// Transform `nlines` lines held in buffer A (scratch B) by all threads of the CTA.
//   element i of line l: buf[l * lstr + i * istr]
//   line_fastest: map consecutive threads to consecutive lines (column tiles, istr = #lines)
// Precondition: A is fully written and the CTA is synchronised.  Returns the buffer holding the
// result (synchronised).
template <class T, int DIR>
__device__ __noinline__ cpx<T>* fft_lines(cpx<T>* A, cpx<T>* B, int nlines, int lstr, int istr, bool line_fastest,
                             const FftDesc& fd, const cpx<T>* __restrict__ tw) {
  const int N = fd.N;
  int Ns = 1;
  for (int s = 0; s < fd.nst; ++s) {
    const int R = fd.radix[s];
    const int tunit = fd.tunit[s];
    const int NR = tunit * Ns;                       // N / R
    const unsigned m_ns = fd.m_ns[s], m_nr = fd.m_nr[s];
    const int total = nlines * NR;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
      int l, j;
      if (line_fastest) {
        j = q / nlines;
        l = q - j * nlines;
      } else {
        l = fastdiv(q, NR, m_nr);                    // (q * NR < 2^32: q < nlines * NR, a few thousand lines at most)
        j = q - l * NR;
      }
      const int k = j - fastdiv(j, Ns, m_ns) * Ns;   // j % Ns
      const cpx<T>* in = A + (size_t)l * lstr;
      cpx<T>* out = B + (size_t)l * lstr;
      switch (R) {
        case 8: stockham_bfly<T, 8, DIR>(in, out, j, k, Ns, NR, tunit, istr, tw); break;
        case 4: stockham_bfly<T, 4, DIR>(in, out, j, k, Ns, NR, tunit, istr, tw); break;
        case 2: stockham_bfly<T, 2, DIR>(in, out, j, k, Ns, NR, tunit, istr, tw); break;
        case 3: stockham_bfly<T, 3, DIR>(in, out, j, k, Ns, NR, tunit, istr, tw); break;
        case 5: stockham_bfly<T, 5, DIR>(in, out, j, k, Ns, NR, tunit, istr, tw); break;
        default: stockham_bfly_generic<T, DIR>(in, out, j, Ns, N, R, istr, tw); break;
      }
    }
    __syncthreads();
    cpx<T>* t = A;
    A = B;
    B = t;
    Ns *= R;
  }
  return A;
}

}  // namespace exb
