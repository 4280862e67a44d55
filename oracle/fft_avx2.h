/*
 * fft_avx2.h -- AVX2 + FMA complex DFT for the FAST flavour of the CPU oracle (libspf_oracle_fast.so, -DORC_FAST).
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  The strict oracle (libspf_oracle.so, -ffp-contract=off, plain C butterflies)
 * stays the parity checker; this flavour exists so that the CPU baseline bench.py reports (`cpu_baseline`,
 * `--impl reference`) is an honest stand-in for the reference's CPU path, which builds with +avx2,+fma
 * (/root/reference/.cargo/config.toml) and runs rustfft's AVX butterflies behind TwistedFft::{forward,reverse}
 * (sunscreen_tfhe/src/math/fft/negacyclic/mod.rs:96-122) plus hand-vectorised complex MADs
 * (sunscreen_tfhe/src/math/simd/x86_64/avx512.rs:15-79).  Own algorithm: Stockham autosort, radix 4, split re/im
 * arrays; the first stage (stride 1) is vectorised across butterflies with a 4x4 in-register transpose on the way
 * out, every later stage across the contiguous stride; the inverse transform is the forward one on swapped re/im.
 */
#ifndef ORC_FFT_AVX2_H
#define ORC_FFT_AVX2_H
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
  uint32_t m;      /* complex length, a power of 4 >= 16 */
  int stages;
  double *tw0;     /* stage 0 (s = 1): [6][m/4]  w1r w1i w2r w2i w3r w3i, contiguous in p */
  double *tw[8];   /* stage k >= 1: [n1][6] broadcast twiddles */
} orc_fastplan;

static inline int orc_fast_ok(uint32_t m) {
  if (m < 16) return 0;
  while (m > 1) { if (m & 3) return 0; m >>= 2; }
  return 1;
}

/* wr / wi: e^{-2 pi i j / m}, j < m */
static void orc_fastplan_build(orc_fastplan *fp, uint32_t m, const double *wr, const double *wi) {
  fp->m = m;
  uint32_t n1 = m / 4;
  fp->tw0 = (double *)aligned_alloc(64, sizeof(double) * 6 * n1);
  for (uint32_t p = 0; p < n1; p++) {
    for (int k = 1; k <= 3; k++) {
      fp->tw0[(2 * (k - 1)) * n1 + p] = wr[(k * p) % m];
      fp->tw0[(2 * (k - 1) + 1) * n1 + p] = wi[(k * p) % m];
    }
  }
  int st = 1;
  for (uint32_t n = m / 4, s = 4; n >= 4; n /= 4, s *= 4, st++) {
    uint32_t nn1 = n / 4, tstep = m / n;
    fp->tw[st] = (double *)aligned_alloc(64, sizeof(double) * 6 * (nn1 ? nn1 : 1));
    for (uint32_t p = 0; p < nn1; p++)
      for (int k = 1; k <= 3; k++) {
        fp->tw[st][6 * p + 2 * (k - 1)] = wr[(k * p * tstep) % m];
        fp->tw[st][6 * p + 2 * (k - 1) + 1] = wi[(k * p * tstep) % m];
      }
  }
  fp->stages = st;
}

#define ORC_CMUL(tr, ti, wr_, wi_, outr, outi)                   \
  do {                                                           \
    outr = _mm256_fmsub_pd(tr, wr_, _mm256_mul_pd(ti, wi_));     \
    outi = _mm256_fmadd_pd(tr, wi_, _mm256_mul_pd(ti, wr_));     \
  } while (0)

/* Forward DFT (e^{-2 pi i jk/m}) of (xr, xi); (yr, yi) is scratch.  Returns 0 if the result is in x, 1 if in y. */
__attribute__((target("avx2,fma"))) static int orc_cfft_avx2(double *xr, double *xi, double *yr, double *yi,
                                                              const orc_fastplan *fp) {
  const uint32_t m = fp->m, n1 = m / 4;
  /* ---- stage 0: s = 1; a_k = x[p + k n1]; y[4p + k] ---- */
  {
    const double *w = fp->tw0;
    for (uint32_t p = 0; p < n1; p += 4) {
      __m256d a0r = _mm256_loadu_pd(xr + p), a0i = _mm256_loadu_pd(xi + p);
      __m256d a1r = _mm256_loadu_pd(xr + p + n1), a1i = _mm256_loadu_pd(xi + p + n1);
      __m256d a2r = _mm256_loadu_pd(xr + p + 2 * n1), a2i = _mm256_loadu_pd(xi + p + 2 * n1);
      __m256d a3r = _mm256_loadu_pd(xr + p + 3 * n1), a3i = _mm256_loadu_pd(xi + p + 3 * n1);
      __m256d apcr = _mm256_add_pd(a0r, a2r), apci = _mm256_add_pd(a0i, a2i);
      __m256d amcr = _mm256_sub_pd(a0r, a2r), amci = _mm256_sub_pd(a0i, a2i);
      __m256d bpdr = _mm256_add_pd(a1r, a3r), bpdi = _mm256_add_pd(a1i, a3i);
      __m256d bmdr = _mm256_sub_pd(a1r, a3r), bmdi = _mm256_sub_pd(a1i, a3i);
      /* -j (b - d) = (bmdi, -bmdr) */
      __m256d t1r = _mm256_add_pd(amcr, bmdi), t1i = _mm256_sub_pd(amci, bmdr);
      __m256d t2r = _mm256_sub_pd(apcr, bpdr), t2i = _mm256_sub_pd(apci, bpdi);
      __m256d t3r = _mm256_sub_pd(amcr, bmdi), t3i = _mm256_add_pd(amci, bmdr);
      __m256d o0r = _mm256_add_pd(apcr, bpdr), o0i = _mm256_add_pd(apci, bpdi);
      __m256d o1r, o1i, o2r, o2i, o3r, o3i;
      ORC_CMUL(t1r, t1i, _mm256_loadu_pd(w + 0 * n1 + p), _mm256_loadu_pd(w + 1 * n1 + p), o1r, o1i);
      ORC_CMUL(t2r, t2i, _mm256_loadu_pd(w + 2 * n1 + p), _mm256_loadu_pd(w + 3 * n1 + p), o2r, o2i);
      ORC_CMUL(t3r, t3i, _mm256_loadu_pd(w + 4 * n1 + p), _mm256_loadu_pd(w + 5 * n1 + p), o3r, o3i);
      /* transpose (k, p) -> y[4 p + k] */
      __m256d u0 = _mm256_unpacklo_pd(o0r, o1r), u1 = _mm256_unpackhi_pd(o0r, o1r);
      __m256d u2 = _mm256_unpacklo_pd(o2r, o3r), u3 = _mm256_unpackhi_pd(o2r, o3r);
      _mm256_storeu_pd(yr + 4 * p + 0, _mm256_permute2f128_pd(u0, u2, 0x20));
      _mm256_storeu_pd(yr + 4 * p + 4, _mm256_permute2f128_pd(u1, u3, 0x20));
      _mm256_storeu_pd(yr + 4 * p + 8, _mm256_permute2f128_pd(u0, u2, 0x31));
      _mm256_storeu_pd(yr + 4 * p + 12, _mm256_permute2f128_pd(u1, u3, 0x31));
      u0 = _mm256_unpacklo_pd(o0i, o1i); u1 = _mm256_unpackhi_pd(o0i, o1i);
      u2 = _mm256_unpacklo_pd(o2i, o3i); u3 = _mm256_unpackhi_pd(o2i, o3i);
      _mm256_storeu_pd(yi + 4 * p + 0, _mm256_permute2f128_pd(u0, u2, 0x20));
      _mm256_storeu_pd(yi + 4 * p + 4, _mm256_permute2f128_pd(u1, u3, 0x20));
      _mm256_storeu_pd(yi + 4 * p + 8, _mm256_permute2f128_pd(u0, u2, 0x31));
      _mm256_storeu_pd(yi + 4 * p + 12, _mm256_permute2f128_pd(u1, u3, 0x31));
    }
  }
  /* ---- stages k >= 1: vectors run along the contiguous stride s >= 4 ---- */
  double *ar = yr, *ai = yi, *br = xr, *bi = xi;
  int in_y = 1, st = 1;
  for (uint32_t n = m / 4, s = 4; n >= 4; n /= 4, s *= 4, st++) {
    const uint32_t nn1 = n / 4;
    const double *tw = fp->tw[st];
    for (uint32_t p = 0; p < nn1; p++) {
      const __m256d w1r = _mm256_broadcast_sd(tw + 6 * p), w1i = _mm256_broadcast_sd(tw + 6 * p + 1);
      const __m256d w2r = _mm256_broadcast_sd(tw + 6 * p + 2), w2i = _mm256_broadcast_sd(tw + 6 * p + 3);
      const __m256d w3r = _mm256_broadcast_sd(tw + 6 * p + 4), w3i = _mm256_broadcast_sd(tw + 6 * p + 5);
      const double *a0r = ar + s * p, *a0i = ai + s * p;
      const size_t sn = (size_t)s * nn1;
      double *o0r = br + (size_t)s * 4 * p, *o0i = bi + (size_t)s * 4 * p;
      for (uint32_t q = 0; q < s; q += 4) {
        __m256d x0r = _mm256_loadu_pd(a0r + q), x0i = _mm256_loadu_pd(a0i + q);
        __m256d x1r = _mm256_loadu_pd(a0r + sn + q), x1i = _mm256_loadu_pd(a0i + sn + q);
        __m256d x2r = _mm256_loadu_pd(a0r + 2 * sn + q), x2i = _mm256_loadu_pd(a0i + 2 * sn + q);
        __m256d x3r = _mm256_loadu_pd(a0r + 3 * sn + q), x3i = _mm256_loadu_pd(a0i + 3 * sn + q);
        __m256d apcr = _mm256_add_pd(x0r, x2r), apci = _mm256_add_pd(x0i, x2i);
        __m256d amcr = _mm256_sub_pd(x0r, x2r), amci = _mm256_sub_pd(x0i, x2i);
        __m256d bpdr = _mm256_add_pd(x1r, x3r), bpdi = _mm256_add_pd(x1i, x3i);
        __m256d bmdr = _mm256_sub_pd(x1r, x3r), bmdi = _mm256_sub_pd(x1i, x3i);
        __m256d t1r = _mm256_add_pd(amcr, bmdi), t1i = _mm256_sub_pd(amci, bmdr);
        __m256d t2r = _mm256_sub_pd(apcr, bpdr), t2i = _mm256_sub_pd(apci, bpdi);
        __m256d t3r = _mm256_sub_pd(amcr, bmdi), t3i = _mm256_add_pd(amci, bmdr);
        _mm256_storeu_pd(o0r + q, _mm256_add_pd(apcr, bpdr));
        _mm256_storeu_pd(o0i + q, _mm256_add_pd(apci, bpdi));
        __m256d or_, oi_;
        if (p == 0) {  /* unit twiddles */
          _mm256_storeu_pd(o0r + s + q, t1r); _mm256_storeu_pd(o0i + s + q, t1i);
          _mm256_storeu_pd(o0r + 2 * s + q, t2r); _mm256_storeu_pd(o0i + 2 * s + q, t2i);
          _mm256_storeu_pd(o0r + 3 * s + q, t3r); _mm256_storeu_pd(o0i + 3 * s + q, t3i);
        } else {
          ORC_CMUL(t1r, t1i, w1r, w1i, or_, oi_);
          _mm256_storeu_pd(o0r + s + q, or_); _mm256_storeu_pd(o0i + s + q, oi_);
          ORC_CMUL(t2r, t2i, w2r, w2i, or_, oi_);
          _mm256_storeu_pd(o0r + 2 * s + q, or_); _mm256_storeu_pd(o0i + 2 * s + q, oi_);
          ORC_CMUL(t3r, t3i, w3r, w3i, or_, oi_);
          _mm256_storeu_pd(o0r + 3 * s + q, or_); _mm256_storeu_pd(o0i + 3 * s + q, oi_);
        }
      }
    }
    double *t;
    t = ar; ar = br; br = t; t = ai; ai = bi; bi = t;
    in_y ^= 1;
  }
  return in_y;
}

/* complex_untwist (simd/scalar.rs:26-35) on split arrays: out[j] = round(re(x[j] * n_inv * conj(tw[j]))), out[j + m] =
 * round(im(..)); f64::round = half away from zero: t = trunc(v), t += copysign(1, v) where |v - t| >= 0.5. */
__attribute__((target("avx2,fma"))) static void orc_untwist_round_avx2(const double *rr, const double *ri, const double *twr,
                                                                        const double *twi, double n_inv, double *out, uint32_t m) {
  const __m256d ninv = _mm256_set1_pd(n_inv), half = _mm256_set1_pd(0.5), one = _mm256_set1_pd(1.0);
  const __m256d signmask = _mm256_set1_pd(-0.0);
  for (uint32_t j = 0; j < m; j += 4) {
    const __m256d ar = _mm256_mul_pd(_mm256_loadu_pd(rr + j), ninv), ai = _mm256_mul_pd(_mm256_loadu_pd(ri + j), ninv);
    const __m256d wr = _mm256_loadu_pd(twr + j), wi = _mm256_loadu_pd(twi + j);
    __m256d v[2];
    v[0] = _mm256_fmadd_pd(ar, wr, _mm256_mul_pd(ai, wi));
    v[1] = _mm256_fmsub_pd(ai, wr, _mm256_mul_pd(ar, wi));
    for (int c = 0; c < 2; c++) {
      const __m256d t = _mm256_round_pd(v[c], _MM_FROUND_TO_ZERO | _MM_FROUND_NO_EXC);
      const __m256d d = _mm256_andnot_pd(signmask, _mm256_sub_pd(v[c], t));
      const __m256d bump = _mm256_and_pd(_mm256_cmp_pd(d, half, _CMP_GE_OQ), _mm256_or_pd(one, _mm256_and_pd(v[c], signmask)));
      _mm256_storeu_pd(out + j + c * m, _mm256_add_pd(t, bump));
    }
  }
}

/* vector_mod_pow2_q_f64 for q = 2^64 + the saturating `as i64` (simd/scalar.rs:75-119, torus.rs:177-186) on
 * integer-valued doubles: h = rint(a / 2^64) 2^64 by the 1.5 * 2^116 trick, lo = a - h exact with |lo| <= 2^63, then
 * lo = hi32 * 2^32 + rem with hi32 = floor(lo / 2^32) (fits an int32) and rem in [0, 2^32) read out of the mantissa of
 * rem + 2^52.  Returns the number of lanes with |lo| == 2^63 or non-finite input (the caller redoes those with the
 * scalar reference form; probability ~2^-53 per coefficient). */
__attribute__((target("avx2,fma"))) static int orc_mod_pow2_64_avx2(uint64_t *c, const double *a, size_t len) {
  const __m256d magic = _mm256_set1_pd(124615124604835863084731911901282304.0);  /* 1.5 * 2^116 */
  const __m256d p63 = _mm256_set1_pd(9223372036854775808.0), inv32 = _mm256_set1_pd(1.0 / 4294967296.0);
  const __m256d p32 = _mm256_set1_pd(4294967296.0), p52 = _mm256_set1_pd(4503599627370496.0);
  const __m256d signmask = _mm256_set1_pd(-0.0);
  const __m256i lomask = _mm256_set1_epi64x(0xFFFFFFFFll);
  int bad = 0;
  size_t j = 0;
  for (; j + 4 <= len; j += 4) {
    const __m256d x = _mm256_loadu_pd(a + j);
    const __m256d h = _mm256_sub_pd(_mm256_add_pd(x, magic), magic);
    const __m256d lo = _mm256_sub_pd(x, h);
    /* |lo| >= 2^63 (or NaN): not (|lo| < 2^63) */
    bad += _mm256_movemask_pd(_mm256_cmp_pd(_mm256_andnot_pd(signmask, lo), p63, _CMP_NLT_UQ)) != 0;
    const __m256d hi = _mm256_floor_pd(_mm256_mul_pd(lo, inv32));
    const __m256d rem = _mm256_fnmadd_pd(hi, p32, lo);
    const __m256i hi64 = _mm256_slli_epi64(_mm256_cvtepi32_epi64(_mm256_cvttpd_epi32(hi)), 32);
    const __m256i rem64 = _mm256_and_si256(_mm256_castpd_si256(_mm256_add_pd(rem, p52)), lomask);
    _mm256_storeu_si256((__m256i *)(c + j), _mm256_or_si256(hi64, rem64));
  }
  return bad + (j < len);
}

/* complex_mad (simd/scalar.rs:12-16; the reference dispatches an AVX-512 version, simd/x86_64/avx512.rs:15-79):
 * c += a * b on interleaved (re, im) arrays, two complex numbers per vector. */
__attribute__((target("avx2,fma"))) static void orc_complex_mad_avx2(double *c, const double *a, const double *b, uint32_t len) {
  for (uint32_t j = 0; j < 2 * len; j += 4) {
    const __m256d va = _mm256_loadu_pd(a + j), vb = _mm256_loadu_pd(b + j);
    const __m256d br = _mm256_movedup_pd(vb), bi = _mm256_permute_pd(vb, 0xF), asw = _mm256_permute_pd(va, 0x5);
    const __m256d prod = _mm256_fmaddsub_pd(va, br, _mm256_mul_pd(asw, bi));
    _mm256_storeu_pd(c + j, _mm256_add_pd(_mm256_loadu_pd(c + j), prod));
  }
}

/* i64 -> f64, round to nearest even (`as f64`, entities/polynomial.rs:264-268): hi32 * 2^32 (exact) + lo32 (exact), one
 * rounding in the add -- the same double cvtsi2sd produces. */
__attribute__((target("avx2,fma"))) static void orc_i64_to_f64_avx2(double *out, const uint64_t *in, uint32_t n) {
  const __m256d p52 = _mm256_set1_pd(4503599627370496.0), p32 = _mm256_set1_pd(4294967296.0);
  const __m256i lomask = _mm256_set1_epi64x(0xFFFFFFFFll), p52i = _mm256_castpd_si256(p52);
  const __m256i idx = _mm256_setr_epi32(1, 3, 5, 7, 0, 0, 0, 0);
  for (uint32_t j = 0; j < n; j += 4) {
    const __m256i x = _mm256_loadu_si256((const __m256i *)(in + j));
    const __m256d lo = _mm256_sub_pd(_mm256_castsi256_pd(_mm256_or_si256(_mm256_and_si256(x, lomask), p52i)), p52);
    const __m128i hi32 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(x, idx));
    const __m256d hi = _mm256_cvtepi32_pd(hi32);
    _mm256_storeu_pd(out + j, _mm256_fmadd_pd(hi, p32, lo));
  }
}

/* complex_twist (simd/scalar.rs:19-23) into split arrays: (x[j] + i x[j + m]) * tw[j] */
__attribute__((target("avx2,fma"))) static void orc_twist_avx2(const double *x, const double *twr, const double *twi, double *xr,
                                                                double *xi, uint32_t m) {
  for (uint32_t j = 0; j < m; j += 4) {
    const __m256d re = _mm256_loadu_pd(x + j), im = _mm256_loadu_pd(x + j + m);
    const __m256d wr = _mm256_loadu_pd(twr + j), wi = _mm256_loadu_pd(twi + j);
    _mm256_storeu_pd(xr + j, _mm256_fmsub_pd(re, wr, _mm256_mul_pd(im, wi)));
    _mm256_storeu_pd(xi + j, _mm256_fmadd_pd(re, wi, _mm256_mul_pd(im, wr)));
  }
}
/* split (re[], im[]) -> interleaved (re, im) pairs and back */
__attribute__((target("avx2"))) static void orc_interleave_avx2(const double *rr, const double *ri, double *out, uint32_t m) {
  for (uint32_t j = 0; j < m; j += 4) {
    const __m256d r = _mm256_loadu_pd(rr + j), i = _mm256_loadu_pd(ri + j);
    const __m256d lo = _mm256_unpacklo_pd(r, i), hi = _mm256_unpackhi_pd(r, i);  /* r0 i0 r2 i2 | r1 i1 r3 i3 */
    _mm256_storeu_pd(out + 2 * j, _mm256_permute2f128_pd(lo, hi, 0x20));
    _mm256_storeu_pd(out + 2 * j + 4, _mm256_permute2f128_pd(lo, hi, 0x31));
  }
}
__attribute__((target("avx2"))) static void orc_deinterleave_avx2(const double *in, double *rr, double *ri, uint32_t m) {
  for (uint32_t j = 0; j < m; j += 4) {
    const __m256d a = _mm256_loadu_pd(in + 2 * j), b = _mm256_loadu_pd(in + 2 * j + 4);  /* r0 i0 r1 i1 | r2 i2 r3 i3 */
    const __m256d lo = _mm256_permute2f128_pd(a, b, 0x20), hi = _mm256_permute2f128_pd(a, b, 0x31);  /* r0 i0 r2 i2 | r1 i1 r3 i3 */
    _mm256_storeu_pd(rr + j, _mm256_unpacklo_pd(lo, hi));
    _mm256_storeu_pd(ri + j, _mm256_unpackhi_pd(lo, hi));
  }
}

#endif
