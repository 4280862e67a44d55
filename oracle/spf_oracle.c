/*
 * spf_oracle.c -- CPU ORACLE (test infrastructure, see spf_oracle.h).  Plain C restatement
 * of the reference's circuit-bootstrapping hot path.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/).
 *
 * Not a copy: the reference is Rust over flat "dst" arrays with iterator adaptors; this is
 * index arithmetic over the same flat layouts.  The complex FFT butterflies are this file's
 * own radix-4 Stockham (the reference calls the un-vendored crate rustfft 6.3.0), so FFT-domain
 * values agree with the reference to f64 rounding, not bit-for-bit.
 *
 * Two builds of this file (oracle/Makefile):
 *   libspf_oracle.so       -ffp-contract=off, plain C: THE ORACLE (parity checker, keygen, decrypt).
 *   libspf_oracle_fast.so  -DORC_FAST -ffp-contract=fast + AVX2/FMA intrinsics (fft_avx2.h) for the FFT, the
 *                          f64 -> torus conversion and the complex MAD: the CPU BASELINE bench.py times, a stand-in
 *                          for the reference's +avx2,+fma build (/root/reference/.cargo/config.toml) with rustfft and
 *                          hand-vectorised MADs (math/simd/x86_64/avx512.rs:15-79).  Same algorithm, same call graph;
 *                          tests/test_oracle_fast.py checks it against the strict build.
 */
#include "spf_oracle.h"
#ifdef ORC_FAST
#include "fft_avx2.h"
#endif

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef orc_c64 c64;

#define ORC_MAX_LOGN 12 /* negacyclic/mod.rs:17 "assert!(log_n < 13)" */

/* ------------------------------------------------------------------------------------ */
/* parameters & sizes                                                                   */
/* ------------------------------------------------------------------------------------ */

/* parasol_runtime/src/params.rs:107-134 DEFAULT_128; sunscreen_tfhe/src/params.rs:219-222,258-264 */
void orc_default_128(orc_params *p) {
  p->lwe_n = 637;  p->lwe_std = 7.25e-5;
  p->glwe_k = 1;   p->glwe_n = 2048;  p->glwe_std = 7e-16;
  p->cbs  = (orc_radix){4, 4};
  p->pbs  = (orc_radix){16, 2};
  p->pfks = (orc_radix){17, 2};
  p->ks   = (orc_radix){2, 6};
  p->ss   = (orc_radix){3, 15};
  p->tr   = (orc_radix){7, 6};
}

static inline uint32_t ilog2u(uint32_t x) { uint32_t l = 0; while ((1u << (l + 1)) <= x) l++; return l; }

/* entities/lwe_ciphertext.rs:24-32 */
size_t orc_size_lwe(uint32_t n) { return (size_t)n + 1; }
/* entities/glwe_ciphertext.rs:38-41 */
size_t orc_size_glwe(const orc_params *p) { return (size_t)(p->glwe_k + 1) * p->glwe_n; }
/* entities/glev_ciphertext.rs:26-31 */
size_t orc_size_glev(const orc_params *p, orc_radix r) { return orc_size_glwe(p) * r.count; }
/* entities/ggsw_ciphertext.rs (size = glev * (k+1)) */
size_t orc_size_ggsw(const orc_params *p, orc_radix r) { return orc_size_glev(p, r) * (p->glwe_k + 1); }
/* entities/ggsw_ciphertext_fft.rs:26-28 */
size_t orc_size_ggsw_fft(const orc_params *p, orc_radix r) { return orc_size_ggsw(p, r) / 2; }
/* entities/bootstrap_key.rs:122-124 */
size_t orc_size_bsk_fft(const orc_params *p) { return orc_size_ggsw_fft(p, p->pbs) * p->lwe_n; }
/* entities/lwe_keyswitch_key.rs:27-36 */
size_t orc_size_ksk(const orc_params *p) {
  return (size_t)p->glwe_k * p->glwe_n * p->ks.count * orc_size_lwe(p->lwe_n);
}
/* entities/automorphism_key_fft.rs:25-27; glwe_keyswitch_key.rs:30-32 */
size_t orc_size_ak_fft(const orc_params *p) {
  return orc_size_glev(p, p->tr) / 2 * p->glwe_k * ilog2u(p->glwe_n);
}
/* entities/scheme_switch_key.rs:49-60 */
size_t orc_size_ssk_fft(const orc_params *p) {
  size_t tri = (size_t)p->glwe_k * (p->glwe_k + 1) / 2;
  return orc_size_glev(p, p->ss) / 2 * tri;
}

/* ------------------------------------------------------------------------------------ */
/* negacyclic FFT: math/fft/negacyclic/mod.rs                                           */
/* ------------------------------------------------------------------------------------ */

typedef struct {
  uint32_t n;       /* polynomial degree N      */
  uint32_t m;       /* complex length N/2       */
  c64 *twist;       /* e^{2 pi i j / 2N}, j<m   (mod.rs:58-66) */
  double *wr, *wi;  /* e^{-2 pi i j / m}, j<m   (forward complex DFT twiddles, split) */
#ifdef ORC_FAST
  int fast;         /* m is a power of 4 >= 16: the AVX2 transform applies */
  orc_fastplan fp;
  double *twr, *twi; /* twist, split */
#endif
} orc_plan;

static orc_plan g_plans[ORC_MAX_LOGN + 1];
static pthread_once_t g_plan_once = PTHREAD_ONCE_INIT;

static void build_plans(void) {
  for (uint32_t ln = 1; ln <= ORC_MAX_LOGN; ln++) {
    orc_plan *pl = &g_plans[ln];
    uint32_t n = 1u << ln, m = n / 2;
    pl->n = n; pl->m = m;
    pl->twist = (c64 *)malloc(sizeof(c64) * m);
    pl->wr = (double *)malloc(sizeof(double) * m);
    pl->wi = (double *)malloc(sizeof(double) * m);
    for (uint32_t j = 0; j < m; j++) {
      /* mod.rs:58-66: (two_pi * x / n_2).sin_cos() with n_2 = 2N */
      double ang = (2.0 * M_PI) * (double)j / (double)(2 * n);
      pl->twist[j].re = cos(ang); pl->twist[j].im = sin(ang);
      long double a2 = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)m;
      pl->wr[j] = (double)cosl(a2); pl->wi[j] = (double)sinl(a2);
    }
#ifdef ORC_FAST
    pl->fast = orc_fast_ok(m);
    if (pl->fast) {
      orc_fastplan_build(&pl->fp, m, pl->wr, pl->wi);
      pl->twr = (double *)aligned_alloc(64, sizeof(double) * m);
      pl->twi = (double *)aligned_alloc(64, sizeof(double) * m);
      for (uint32_t j = 0; j < m; j++) { pl->twr[j] = pl->twist[j].re; pl->twi[j] = pl->twist[j].im; }
    }
#endif
  }
}

static const orc_plan *get_plan(uint32_t n) {
  pthread_once(&g_plan_once, build_plans);
  uint32_t ln = ilog2u(n);
  return &g_plans[ln];
}

/* Complex DFT of length m (natural order in and out), unnormalised; sgn = -1 forward
 * (e^{-2 pi i jk/m}), +1 inverse.  Stands in for rustfft's Fft::process_with_scratch
 * (negacyclic/mod.rs:106,119).  Stockham autosort, radix-4 with a radix-2 tail, split
 * re/im arrays so the inner loops vectorise. */
static void cfft_split(double *restrict xr, double *restrict xi, double *restrict yr, double *restrict yi,
                       const orc_plan *pl, int inverse) {
  uint32_t m = pl->m;
  const double *wr = pl->wr, *wi = pl->wi;
  double sg = inverse ? -1.0 : 1.0; /* multiplies the twiddle's imaginary part */
  uint32_t n = m, s = 1;
  double *ar = xr, *ai = xi, *br = yr, *bi = yi;
  while (n >= 4) {
    uint32_t n1 = n / 4, tstep = m / n;
    for (uint32_t p = 0; p < n1; p++) {
      double w1r = wr[p * tstep], w1i = sg * wi[p * tstep];
      double w2r = wr[2 * p * tstep], w2i = sg * wi[2 * p * tstep];
      double w3r = wr[3 * p * tstep], w3i = sg * wi[3 * p * tstep];
      const double *restrict a0r = ar + s * p, *restrict a0i = ai + s * p;
      const double *restrict a1r = a0r + s * n1, *restrict a1i = a0i + s * n1;
      const double *restrict a2r = a1r + s * n1, *restrict a2i = a1i + s * n1;
      const double *restrict a3r = a2r + s * n1, *restrict a3i = a2i + s * n1;
      double *restrict o0r = br + s * 4 * p, *restrict o0i = bi + s * 4 * p;
      double *restrict o1r = o0r + s, *restrict o1i = o0i + s, *restrict o2r = o1r + s, *restrict o2i = o1i + s,
             *restrict o3r = o2r + s, *restrict o3i = o2i + s;
#pragma GCC ivdep
      for (uint32_t q = 0; q < s; q++) {
        double apcr = a0r[q] + a2r[q], apci = a0i[q] + a2i[q];
        double amcr = a0r[q] - a2r[q], amci = a0i[q] - a2i[q];
        double bpdr = a1r[q] + a3r[q], bpdi = a1i[q] + a3i[q];
        /* -j*(b-d) forward, +j*(b-d) inverse */
        double bmdr = a1r[q] - a3r[q], bmdi = a1i[q] - a3i[q];
        double jr = sg * bmdi, ji = -sg * bmdr;
        double t1r = amcr + jr, t1i = amci + ji;
        double t2r = apcr - bpdr, t2i = apci - bpdi;
        double t3r = amcr - jr, t3i = amci - ji;
        o0r[q] = apcr + bpdr;            o0i[q] = apci + bpdi;
        o1r[q] = t1r * w1r - t1i * w1i;  o1i[q] = t1r * w1i + t1i * w1r;
        o2r[q] = t2r * w2r - t2i * w2i;  o2i[q] = t2r * w2i + t2i * w2r;
        o3r[q] = t3r * w3r - t3i * w3i;  o3i[q] = t3r * w3i + t3i * w3r;
      }
    }
    n /= 4; s *= 4;
    double *t;
    t = ar; ar = br; br = t; t = ai; ai = bi; bi = t;
  }
  if (n == 2) {
    for (uint32_t q = 0; q < s; q++) {
      double a_r = ar[q], a_i = ai[q], b_r = ar[q + s], b_i = ai[q + s];
      br[q] = a_r + b_r; bi[q] = a_i + b_i;
      br[q + s] = a_r - b_r; bi[q + s] = a_i - b_i;
    }
    double *t;
    t = ar; ar = br; br = t; t = ai; ai = bi; bi = t;
  }
  if (ar != xr) { memcpy(xr, ar, sizeof(double) * m); memcpy(xi, ai, sizeof(double) * m); }
}

/* f64::round (half away from zero), exact: x - trunc(x) is representable. */
static inline double round_half_away(double x) {
  double t = trunc(x);
  if (fabs(x - t) >= 0.5) t += copysign(1.0, x);
  return t;
}

/* TwistedFft::forward (negacyclic/mod.rs:96-107) + simd::complex_twist (simd/scalar.rs:19-23) */
void orc_fft_forward(const double *x, c64 *out, uint32_t n) {
  const orc_plan *pl = get_plan(n);
  uint32_t m = pl->m;
  if (m == 0) return;
  double xr[2048] __attribute__((aligned(64))), xi[2048] __attribute__((aligned(64))), yr[2048] __attribute__((aligned(64))),
      yi[2048] __attribute__((aligned(64)));
#ifdef ORC_FAST
  if (pl->fast) {
    orc_twist_avx2(x, pl->twr, pl->twi, xr, xi, m);
    const int in_y = orc_cfft_avx2(xr, xi, yr, yi, &pl->fp);
    orc_interleave_avx2(in_y ? yr : xr, in_y ? yi : xi, (double *)out, m);
    return;
  }
#endif
  for (uint32_t j = 0; j < m; j++) {
    double re = x[j], im = x[j + m];
    c64 t = pl->twist[j];
    xr[j] = re * t.re - im * t.im;
    xi[j] = re * t.im + im * t.re;
  }
  cfft_split(xr, xi, yr, yi, pl, 0);
  for (uint32_t j = 0; j < m; j++) { out[j].re = xr[j]; out[j].im = xi[j]; }
}

/* TwistedFft::reverse (negacyclic/mod.rs:109-122) + complex_untwist (simd/scalar.rs:26-35):
 * unnormalised inverse DFT, * 1/len, * twist_inv, round() (half away from zero). */
void orc_fft_reverse(const c64 *in, double *out, uint32_t n) {
  const orc_plan *pl = get_plan(n);
  uint32_t m = pl->m;
  if (m == 0) return;
  double xr[2048] __attribute__((aligned(64))), xi[2048] __attribute__((aligned(64))), yr[2048] __attribute__((aligned(64))),
      yi[2048] __attribute__((aligned(64)));
  double n_inv = 1.0 / (double)m;
#ifdef ORC_FAST
  if (pl->fast) {
    orc_deinterleave_avx2((const double *)in, xr, xi, m);
    /* inverse = forward on swapped re / im (the result comes back swapped again, i.e. in place) */
    const int in_y = orc_cfft_avx2(xi, xr, yi, yr, &pl->fp);
    const double *rr = in_y ? yr : xr, *ri = in_y ? yi : xi;
    orc_untwist_round_avx2(rr, ri, pl->twr, pl->twi, n_inv, out, m);
    return;
  }
#endif
  for (uint32_t j = 0; j < m; j++) { xr[j] = in[j].re; xi[j] = in[j].im; }
  cfft_split(xr, xi, yr, yi, pl, 1);
  for (uint32_t j = 0; j < m; j++) {
    /* tmp = x * n_inv * twist_inv[j]; twist_inv = twist^-1 = conj(twist) (unit modulus) */
    double ar = xr[j] * n_inv, ai = xi[j] * n_inv;
    double tr = pl->twist[j].re, ti = -pl->twist[j].im;
    out[j] = round_half_away(ar * tr - ai * ti);
    out[j + m] = round_half_away(ar * ti + ai * tr);
  }
}

/* PolynomialRef::fft (entities/polynomial.rs:257-274): u64 -> i64 -> f64, then forward. */
void orc_poly_fft(const u64 *p, c64 *out, uint32_t n) {
  double stackx[4096] __attribute__((aligned(64)));
#ifdef ORC_FAST
  if (n % 4 == 0) orc_i64_to_f64_avx2(stackx, p, n);
  else
#endif
  for (uint32_t j = 0; j < n; j++) stackx[j] = (double)(i64)p[j];
  orc_fft_forward(stackx, out, n);
}

/* simd/scalar.rs:75-119 vector_mod_pow2_q_f64 with log2_q = 64, and torus.rs:177-186
 * FromF64 (`x as i64` is a saturating cast in Rust). */
void orc_mod_pow2_q_f64(u64 *c, const double *a, size_t len) {
#ifdef ORC_FAST
  if (len % 4 == 0 && orc_mod_pow2_64_avx2(c, a, len) == 0) return;  /* else: redo everything the reference way */
#endif
  const double q = 18446744073709551616.0;       /* 2^64 */
  const double q_div_2 = 9223372036854775808.0;  /* 2^63 */
  for (size_t j = 0; j < len; j++) {
    double v = a[j];
    v = fma(-trunc(v / q), q, v);
    if (v >= q_div_2) v -= q;
    else if (v <= -q_div_2) v += q;
    i64 s;
    if (v != v) s = 0;                                    /* NaN -> 0 (Rust `as`) */
    else if (v >= q_div_2) s = INT64_MAX;                 /* saturate */
    else if (v < -q_div_2) s = INT64_MIN;
    else s = (i64)v;
    c[j] = (u64)s;
  }
}

/* PolynomialFftRef::ifft (entities/polynomial_fft.rs:82-99) */
void orc_poly_ifft(const c64 *in, u64 *p, uint32_t n) {
  double stackx[4096];
  orc_fft_reverse(in, stackx, n);
  orc_mod_pow2_q_f64(p, stackx, n);
}

/* ------------------------------------------------------------------------------------ */
/* radix decomposition: math/radix.rs                                                   */
/* ------------------------------------------------------------------------------------ */

/* radix.rs:155-162 round() */
u64 orc_radix_round(u64 x, orc_radix r) {
  uint32_t shift = 64 - r.radix_log * r.count;
  u64 round_bit = (x >> (shift - 1)) & 1;
  return (x >> shift) + round_bit;
}

/* simd/scalar.rs:52-72 vector_next_decomp: signed digits in [-B/2, B/2), LSB first */
void orc_next_decomp(u64 *s, u64 *r, size_t len, uint32_t radix_log) {
  u64 mask = ((u64)1 << radix_log) - 1;
  for (size_t j = 0; j < len; j++) {
    u64 digit = s[j] & mask;
    s[j] >>= radix_log;
    u64 carry = digit >> (radix_log - 1);
    s[j] += carry;
    r[j] = digit - (carry << radix_log);
  }
}

/* ops/ciphertext/lwe_ciphertext_ops.rs:129-142 modulus_switch */
u64 orc_modulus_switch(u64 x, uint32_t log_chi, uint32_t log_v, uint32_t log_modulus) {
  u64 mask = ((u64)1 << log_modulus) - 1;
  x <<= log_chi;
  uint32_t shift = 64 - (log_modulus - log_v);
  u64 rnd = (x >> (shift - 1)) & 1;
  x >>= shift;
  return ((x + rnd) & mask) << log_v;
}

/* ops/polynomial/mod.rs:62-87 polynomial_pow_k: P(X) -> P(X^k) in Z[X]/(X^N+1) */
void orc_poly_pow_k(u64 *pk, const u64 *p, uint32_t n, uint32_t k) {
  for (uint32_t i = 0; i < n; i++) {
    u64 ik = (u64)i * k;
    uint32_t idx = (uint32_t)(ik % n);
    int neg = (int)((ik / n) & 1);
    pk[idx] = neg ? (u64)0 - p[i] : p[i];
  }
}

/* simd/scalar.rs:134-143 vector_shr_round */
void orc_shr_round(u64 *y, const u64 *x, size_t len, uint32_t n) {
  for (size_t j = 0; j < len; j++) y[j] = (x[j] >> n) + ((x[j] >> (n - 1)) & 1);
}

/* entities/polynomial.rs:171-248 mul_by_{positive,negative}_monomial_negacyclic */
void orc_poly_mul_monomial(u64 *p, uint32_t n, i64 degree) {
  u64 tmp[4096];
  int negative = degree < 0;
  u64 deg = (u64)(negative ? -degree : degree) % (2 * (u64)n);
  if (deg == 0) return;
  if (deg == n) { for (uint32_t j = 0; j < n; j++) p[j] = (u64)0 - p[j]; return; }
  uint32_t shift = (uint32_t)(deg % n);
  memcpy(tmp, p, sizeof(u64) * n);
  if (!negative) {
    /* rotate_right(shift); negate [0,degree) if degree<N else [shift,N) */
    memcpy(p + shift, tmp, sizeof(u64) * (n - shift));
    memcpy(p, tmp + (n - shift), sizeof(u64) * shift);
    uint32_t lo = deg < n ? 0 : shift, hi = deg < n ? (uint32_t)deg : n;
    for (uint32_t j = lo; j < hi; j++) p[j] = (u64)0 - p[j];
  } else {
    /* rotate_left(shift); negate [N-shift,N) if degree<N else [0,N-shift) */
    memcpy(p, tmp + shift, sizeof(u64) * (n - shift));
    memcpy(p + (n - shift), tmp, sizeof(u64) * shift);
    uint32_t lo = deg < n ? n - shift : 0, hi = deg < n ? n : n - shift;
    for (uint32_t j = lo; j < hi; j++) p[j] = (u64)0 - p[j];
  }
}

/* ops/bootstrapping/programmable_bootstrapping.rs:129-185 generate_lut.
 * table[f*p + x] = maps[f](x) for f<v, x < p = 2^plaintext_bits. */
void orc_generate_lut(u64 *c, const u64 *table, uint32_t v, uint32_t n, uint32_t plaintext_bits) {
  uint32_t p = 1u << plaintext_bits;
  uint32_t log_v = ilog2u(v);
  if ((1u << log_v) != v) log_v += 1;
  uint32_t ceil_v = 1u << log_v;
  uint32_t stride = n / p;
  uint32_t delta = 64 - plaintext_bits;
  for (uint32_t j = 0; j < p; j++)
    for (uint32_t k = 0; k < stride; k++) {
      uint32_t fn_id = k % ceil_v;
      u64 pi = fn_id < v ? table[(size_t)fn_id * p + j] : 0;
      c[j * stride + k] = pi << delta;
    }
  for (uint32_t j = 0; j < stride / 2; j++) c[j] = (u64)0 - c[j];
  /* rotate_left(stride/2) */
  u64 tmp[4096];
  memcpy(tmp, c, sizeof(u64) * n);
  for (uint32_t j = 0; j < n; j++) c[j] = tmp[(j + stride / 2) % n];
}

/* ------------------------------------------------------------------------------------ */
/* FFT-domain ops: ops/fft_ops.rs                                                       */
/* ------------------------------------------------------------------------------------ */

/* simd/scalar.rs:12-16 complex_mad: c += a*b */
static void complex_mad(c64 *c, const c64 *a, const c64 *b, uint32_t len) {
#ifdef ORC_FAST
  if (len % 2 == 0) { orc_complex_mad_avx2((double *)c, (const double *)a, (const double *)b, len); return; }
#endif
  for (uint32_t j = 0; j < len; j++) {
    c[j].re += a[j].re * b[j].re - a[j].im * b[j].im;
    c[j].im += a[j].re * b[j].im + a[j].im * b[j].re;
  }
}

/* fft_ops.rs:67-98 decomposed_polynomial_glev_mad: c(GLWE-FFT) += <Decomp(poly), GLEV-FFT>.
 * Digits come LSB first and pair with the GLEV's GLWEs in REVERSE order (:92). */
static void decomposed_poly_glev_mad(c64 *c, const u64 *poly, const c64 *glev_fft,
                                     const orc_params *p, orc_radix r) {
  uint32_t n = p->glwe_n, m = n / 2, k1 = p->glwe_k + 1;
  u64 state[4096], digit[4096];
  c64 dfft[2048];
  for (uint32_t j = 0; j < n; j++) state[j] = orc_radix_round(poly[j], r); /* radix.rs:89 */
  for (uint32_t t = 0; t < r.count; t++) {
    uint32_t level = r.count - 1 - t;
    orc_next_decomp(state, digit, n, r.radix_log);
    orc_poly_fft(digit, dfft, n);
    const c64 *glwe = glev_fft + (size_t)level * k1 * m;
    for (uint32_t q = 0; q < k1; q++) /* fft_ops.rs:107-124 glwe_polynomial_mad */
      complex_mad(c + (size_t)q * m, glwe + (size_t)q * m, dfft, m);
  }
}

/* fft_ops.rs:23-56 glwe_ggsw_mad: c_fft += GGSW [*] GLWE */
void orc_glwe_ggsw_mad(c64 *c_fft, const u64 *glwe, const c64 *ggsw_fft, const orc_params *p,
                       orc_radix r) {
  uint32_t n = p->glwe_n, m = n / 2, k1 = p->glwe_k + 1;
  size_t row_stride = (size_t)r.count * k1 * m;
  for (uint32_t i = 0; i < k1; i++) /* a polynomials then b (:37) against row i */
    decomposed_poly_glev_mad(c_fft, glwe + (size_t)i * n, ggsw_fft + i * row_stride, p, r);
}

/* GlweCiphertextFft::ifft: per polynomial ifft */
void orc_glwe_fft_ifft(const c64 *c_fft, u64 *glwe, const orc_params *p) {
  uint32_t n = p->glwe_n, m = n / 2, k1 = p->glwe_k + 1;
  for (uint32_t q = 0; q < k1; q++) orc_poly_ifft(c_fft + (size_t)q * m, glwe + (size_t)q * n, n);
}

/* fft_ops.rs:149-181 cmux: c = d0 + IFFT(GGSW [*] (d1 - d0)) (c is overwritten, :180) */
void orc_cmux(u64 *c, const u64 *d0, const u64 *d1, const c64 *ggsw_fft, const orc_params *p,
              orc_radix r) {
  size_t len = orc_size_glwe(p);
  u64 diff[2 * 4096], prod[2 * 4096];
  c64 prod_fft[2 * 2048];
  u64 *diffp = diff, *prodp = prod; c64 *pf = prod_fft;
  int heap = len > 2 * 4096;
  if (heap) { diffp = malloc(8 * len); prodp = malloc(8 * len); pf = malloc(16 * len / 2); }
  for (size_t j = 0; j < len; j++) diffp[j] = d1[j] - d0[j];
  memset(pf, 0, sizeof(c64) * len / 2);
  orc_glwe_ggsw_mad(pf, diffp, ggsw_fft, p, r);
  orc_glwe_fft_ifft(pf, prodp, p);
  for (size_t j = 0; j < len; j++) c[j] = prodp[j] + d0[j];
  if (heap) { free(diffp); free(prodp); free(pf); }
}

/* fft_ops.rs:203-220 glev_cmux */
void orc_glev_cmux(u64 *c, const u64 *d0, const u64 *d1, const c64 *ggsw_fft, const orc_params *p,
                   orc_radix glev_r, orc_radix ggsw_r) {
  size_t g = orc_size_glwe(p);
  for (uint32_t i = 0; i < glev_r.count; i++) orc_cmux(c + i * g, d0 + i * g, d1 + i * g, ggsw_fft, p, ggsw_r);
}

/* parasol_runtime/src/crypto/evaluation.rs:104-123 multiply_glwe_ggsw */
void orc_multiply_glwe_ggsw(u64 *out, const u64 *glwe, const c64 *ggsw_fft, const orc_params *p) {
  size_t len = orc_size_glwe(p);
  c64 *pf = calloc(len / 2, sizeof(c64));
  orc_glwe_ggsw_mad(pf, glwe, ggsw_fft, p, p->cbs);
  orc_glwe_fft_ifft(pf, out, p);
  free(pf);
}

/* fft_ops.rs:457-495 keyswitch_glwe_to_glwe (FFT version):
 * out = (0, b) - IFFT( sum_i <Decomp(a_i), KSK_row_i> ) */
void orc_keyswitch_glwe(u64 *out, const u64 *in, const c64 *ksk_fft, const orc_params *p, orc_radix r) {
  uint32_t n = p->glwe_n, m = n / 2, k = p->glwe_k, k1 = k + 1;
  size_t len = orc_size_glwe(p);
  c64 *sum = calloc(len / 2, sizeof(c64));
  u64 *sum_ifft = malloc(8 * len);
  size_t row_stride = (size_t)r.count * k1 * m;
  for (uint32_t i = 0; i < k; i++)
    decomposed_poly_glev_mad(sum, in + (size_t)i * n, ksk_fft + i * row_stride, p, r);
  orc_glwe_fft_ifft(sum, sum_ifft, p);
  for (size_t j = 0; j < (size_t)k * n; j++) out[j] = (u64)0 - sum_ifft[j];
  for (size_t j = (size_t)k * n; j < len; j++) out[j] = in[j] - sum_ifft[j];
  free(sum); free(sum_ifft);
}

/* ops/automorphisms/mod.rs:53-85 trace */
void orc_trace(u64 *out, const u64 *x, const c64 *ak_fft, const orc_params *p) {
  uint32_t n = p->glwe_n, k1 = p->glwe_k + 1, logn = ilog2u(n);
  size_t len = orc_size_glwe(p);
  size_t ksk_stride = orc_size_glev(p, p->tr) / 2 * p->glwe_k;
  u64 *glwe_k = malloc(8 * len), *ks = malloc(8 * len);
  memcpy(out, x, 8 * len);
  for (uint32_t i = 1; i <= logn; i++) {
    uint32_t kk = n / (1u << (i - 1)) + 1;
    for (uint32_t q = 0; q < k1; q++) orc_poly_pow_k(glwe_k + (size_t)q * n, out + (size_t)q * n, n, kk);
    orc_keyswitch_glwe(ks, glwe_k, ak_fft + (size_t)(i - 1) * ksk_stride, p, p->tr);
    for (size_t j = 0; j < len; j++) out[j] += ks[j];
  }
  free(glwe_k); free(ks);
}

/* entities/scheme_switch_key.rs:21-30 get_linear_index */
static size_t tri_index(uint32_t i, uint32_t j, uint32_t n) {
  uint32_t row = i <= j ? i : j, col = i <= j ? j : i;
  return (size_t)n * (n + 1) / 2 - (size_t)(n - row) * (n - row + 1) / 2 + col - row;
}

/* fft_ops.rs:225-279,403-442 scheme_switch_fft.  The reference accumulates into a fresh
 * (zeroed) output (circuit_processor/mod.rs:333,534); here the output is zeroed first. */
void orc_scheme_switch_fft(c64 *out, const u64 *glev, const c64 *ssk_fft, const orc_params *p) {
  uint32_t n = p->glwe_n, m = n / 2, k = p->glwe_k, k1 = k + 1, l = p->cbs.count;
  size_t g = orc_size_glwe(p), gf = g / 2;
  size_t ssk_glev = orc_size_glev(p, p->ss) / 2;
  memset(out, 0, sizeof(c64) * orc_size_ggsw_fft(p, p->cbs));
  for (uint32_t j = 0; j < k1; j++)
    for (uint32_t i = 0; i < l; i++) {
      c64 *y = out + ((size_t)j * l + i) * gf;
      const u64 *x = glev + (size_t)i * g;
      if (j == k) { /* :236-239 last row: plain FFT of x_i */
        for (uint32_t q = 0; q < k1; q++) orc_poly_fft(x + (size_t)q * n, y + (size_t)q * m, n);
        continue;
      }
      orc_poly_fft(x + (size_t)k * n, y + (size_t)j * m, n); /* :222-241 b -> a-slot j */
      for (uint32_t r = 0; r < k; r++)
        decomposed_poly_glev_mad(y, x + (size_t)r * n, ssk_fft + tri_index(j, r, k) * ssk_glev, p, p->ss);
    }
}

/* ------------------------------------------------------------------------------------ */
/* bootstrapping                                                                        */
/* ------------------------------------------------------------------------------------ */

/* ops/bootstrapping/programmable_bootstrapping.rs:342-410 generalized_programmable_bootstrap */
void orc_pbs_generalized(u64 *out, const u64 *lwe_in, const u64 *lut_glwe, const c64 *bsk_fft,
                         uint32_t log_chi, uint32_t log_v, const orc_params *p) {
  uint32_t n = p->glwe_n, k1 = p->glwe_k + 1, ln = p->lwe_n;
  uint32_t two_n = ilog2u(n) + 1;
  size_t len = orc_size_glwe(p);
  size_t ggsw = orc_size_ggsw_fft(p, p->pbs);
  u64 *ct = malloc(8 * (ln + 1)), *tmp = malloc(8 * len), *rot = malloc(8 * len);
  for (uint32_t i = 0; i <= ln; i++) ct[i] = orc_modulus_switch(lwe_in[i], log_chi, log_v, two_n); /* :378-379 */
  memcpy(out, lut_glwe, 8 * len);                                     /* V_0 * X^{-b} (:385-390) */
  for (uint32_t q = 0; q < k1; q++) orc_poly_mul_monomial(out + (size_t)q * n, n, -(i64)ct[ln]);
  for (uint32_t i = 0; i < ln; i++) {                                 /* :396-409 */
    memcpy(tmp, out, 8 * len);
    memcpy(rot, out, 8 * len);
    for (uint32_t q = 0; q < k1; q++) orc_poly_mul_monomial(rot + (size_t)q * n, n, (i64)ct[i]);
    orc_cmux(out, tmp, rot, bsk_fft + (size_t)i * ggsw, p, p->pbs);
  }
  free(ct); free(tmp); free(rot);
}

/* programmable_bootstrapping.rs:291-318 programmable_bootstrap_univariate */
void orc_pbs_univariate(u64 *lwe_out, const u64 *lwe_in, const u64 *lut_glwe, const c64 *bsk_fft,
                        const orc_params *p) {
  u64 *glwe = malloc(8 * orc_size_glwe(p));
  orc_pbs_generalized(glwe, lwe_in, lut_glwe, bsk_fft, 0, 0, p);
  orc_sample_extract(lwe_out, glwe, 0, p);
  free(glwe);
}

static uint32_t cbs_log_v(const orc_params *p) {
  uint32_t lv = ilog2u(p->cbs.count);
  if ((1u << lv) != p->cbs.count) lv += 1;
  return lv;
}

/* circuit_bootstrapping.rs:430-482 fill_multifunctional_cbs_decomposition_lut */
void orc_cbs_lut(u64 *lut_glwe, const orc_params *p) {
  uint32_t n = p->glwe_n, k = p->glwe_k;
  u64 levels[16];
  memset(lut_glwe, 0, 8 * orc_size_glwe(p));
  for (uint32_t i = 0; i < 16; i++) {
    uint32_t pb = p->cbs.radix_log * (i + 1) + 1;
    levels[i] = 0;
    if (pb < 64) levels[i] = ((((u64)1 << pb) - 1)) << (64 - pb); /* encode(-1, pb) */
  }
  uint32_t v = 1u << cbs_log_v(p);
  u64 *b = lut_glwe + (size_t)k * n;
  for (uint32_t i = 0; i < n; i++) { uint32_t f = i % v; b[i] = f < p->cbs.count ? levels[f] : 0; }
}

/* circuit_bootstrapping.rs:387-428 hi_noise_lwe_to_lo_noise_glwe; lwe_rotate homomorphisms/lwe.rs:9-20 */
void orc_cbs_pbs_stage(u64 *glwe_out, const u64 *lwe_in, const c64 *bsk_fft, const orc_params *p) {
  uint32_t ln = p->lwe_n;
  u64 *rot = malloc(8 * (ln + 1)), *lut = malloc(8 * orc_size_glwe(p));
  memcpy(rot, lwe_in, 8 * (ln + 1));
  rot[ln] += (u64)1 << 62; /* Torus::encode(1, PlaintextBits(2)) */
  orc_cbs_lut(lut, p);
  orc_pbs_generalized(glwe_out, rot, lut, bsk_fft, 0, cbs_log_v(p), p);
  free(rot); free(lut);
}

/* circuit_bootstrapping.rs:260-298 mod_switch_trace_and_rotate;
 * glwe_mod_switch_and_expand_pow_2 ops/ciphertext/glwe_ciphertext_ops.rs:268-281 */
void orc_cbs_trace_stage(u64 *glev_out, const u64 *glwe_in, const c64 *ak_fft, const orc_params *p) {
  uint32_t n = p->glwe_n, k = p->glwe_k, k1 = k + 1, shift = ilog2u(n);
  size_t len = orc_size_glwe(p);
  u64 *rotated = malloc(8 * len), *permuted = malloc(8 * len), *shifted = malloc(8 * len);
  memcpy(rotated, glwe_in, 8 * len);
  for (uint32_t i = 0; i < p->cbs.count; i++) {
    uint32_t pb = p->cbs.radix_log * (i + 1) + 1;
    rotated[(size_t)k * n + i] += (u64)1 << (64 - pb); /* :284-285 (cumulative on one buffer) */
    memcpy(permuted, rotated, 8 * len);
    for (uint32_t q = 0; q < k1; q++) orc_poly_mul_monomial(permuted + (size_t)q * n, n, -(i64)i);
    orc_shr_round(shifted, permuted, len, shift);
    orc_trace(glev_out + (size_t)i * len, shifted, ak_fft, p);
  }
  free(rotated); free(permuted); free(shifted);
}

/* circuit_bootstrapping.rs:342-384 circuit_bootstrap_via_trace_and_scheme_switch */
void orc_circuit_bootstrap(c64 *ggsw_out, const u64 *lwe_in, const c64 *bsk_fft, const c64 *ak_fft,
                           const c64 *ssk_fft, const orc_params *p) {
  u64 *glwe = malloc(8 * orc_size_glwe(p)), *glev = malloc(8 * orc_size_glev(p, p->cbs));
  orc_cbs_pbs_stage(glwe, lwe_in, bsk_fft, p);
  orc_cbs_trace_stage(glev, glwe, ak_fft, p);
  orc_scheme_switch_fft(ggsw_out, glev, ssk_fft, p);
  free(glwe); free(glev);
}

/* ops/keyswitch/lwe_keyswitch.rs:23-62 + lev_ciphertext_ops.rs:18-42 + lwe_ciphertext_ops.rs:48-66 */
void orc_keyswitch_lwe(u64 *out, const u64 *in, const u64 *ksk, const orc_params *p) {
  uint32_t n1 = p->glwe_k * p->glwe_n, n0 = p->lwe_n, l = p->ks.count;
  size_t lwe0 = (size_t)n0 + 1;
  u64 *sum = calloc(lwe0, 8);
  for (uint32_t i = 0; i < n1; i++) {
    u64 s = orc_radix_round(in[i], p->ks), d;
    for (uint32_t t = 0; t < l; t++) {
      orc_next_decomp(&s, &d, 1, p->ks.radix_log);
      const u64 *row = ksk + ((size_t)i * l + (l - 1 - t)) * lwe0;
      for (size_t j = 0; j < lwe0; j++) sum[j] += row[j] * d;
    }
  }
  for (uint32_t j = 0; j < n0; j++) out[j] = (u64)0 - sum[j];
  out[n0] = in[n1] - sum[n0];
  free(sum);
}

/* ops/ciphertext/glwe_ciphertext_ops.rs:31-76 sample_extract */
void orc_sample_extract(u64 *lwe, const u64 *glwe, uint32_t h, const orc_params *p) {
  uint32_t n = p->glwe_n, k = p->glwe_k;
  for (uint32_t i = 0; i < k; i++) {
    const u64 *a = glwe + (size_t)i * n;
    for (uint32_t j = 0; j <= h; j++) lwe[(size_t)i * n + j] = a[h - j];
    for (uint32_t j = h + 1; j < n; j++) lwe[(size_t)i * n + j] = (u64)0 - a[h + n - j];
  }
  lwe[(size_t)k * n] = glwe[(size_t)k * n + h];
}

/* evaluation.rs:53-55 xor = a + b (also glwe_ciphertext_ops.rs:79-99) */
void orc_glwe_add(u64 *c, const u64 *a, const u64 *b, const orc_params *p) {
  size_t len = orc_size_glwe(p);
  for (size_t j = 0; j < len; j++) c[j] = a[j] + b[j];
}
/* evaluation.rs:48-50 not = input + trivial_one; trivial one = b[0] = encode(1,1bit) */
void orc_glwe_not(u64 *c, const u64 *a, const orc_params *p) {
  size_t len = orc_size_glwe(p);
  memcpy(c, a, 8 * len);
  c[(size_t)p->glwe_k * p->glwe_n] += (u64)1 << 63;
}
/* evaluation.rs:58-65 mul_xn -> rotate_glwe_positive_monomial_negacyclic */
void orc_glwe_mul_xn(u64 *c, const u64 *a, uint32_t n, const orc_params *p) {
  size_t len = orc_size_glwe(p);
  memcpy(c, a, 8 * len);
  for (uint32_t q = 0; q <= p->glwe_k; q++) orc_poly_mul_monomial(c + (size_t)q * p->glwe_n, p->glwe_n, (i64)n);
}

/* ------------------------------------------------------------------------------------ */
/* host-thread batches (the reference's rayon task model)                               */
/* ------------------------------------------------------------------------------------ */

int orc_hw_threads(void) { long v = sysconf(_SC_NPROCESSORS_ONLN); return v > 0 ? (int)v : 1; }

typedef struct {
  int kind, tid, nthreads; size_t batch; const orc_params *p;
  void *out; const void *in0, *in1, *in2; const void *k0, *k1, *k2;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const orc_params *p = j->p;
  for (size_t b = j->tid; b < j->batch; b += j->nthreads) {
    switch (j->kind) {
    case 0:
      orc_circuit_bootstrap((c64 *)j->out + b * orc_size_ggsw_fft(p, p->cbs),
                            (const u64 *)j->in0 + b * orc_size_lwe(p->lwe_n), j->k0, j->k1, j->k2, p);
      break;
    case 1: {
      size_t g = orc_size_glwe(p);
      orc_cmux((u64 *)j->out + b * g, (const u64 *)j->in0 + b * g, (const u64 *)j->in1 + b * g,
               (const c64 *)j->in2 + b * orc_size_ggsw_fft(p, p->cbs), p, p->cbs);
    } break;
    case 3: /* pointer tables: one (out, low, high, selector) quadruple per item */
      orc_cmux(((u64 **)j->out)[b], ((const u64 *const *)j->in0)[b], ((const u64 *const *)j->in1)[b],
               ((const c64 *const *)j->in2)[b], p, p->cbs);
      break;
    case 2:
      orc_keyswitch_lwe((u64 *)j->out + b * orc_size_lwe(p->lwe_n),
                        (const u64 *)j->in0 + b * orc_size_lwe(p->glwe_k * p->glwe_n), j->k0, p);
      break;
    }
  }
  return NULL;
}

static void run_batch(batch_job proto, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  batch_job *jobs = malloc(sizeof(batch_job) * nthreads);
  for (int t = 0; t < nthreads; t++) { jobs[t] = proto; jobs[t].tid = t; jobs[t].nthreads = nthreads;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]); }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

void orc_circuit_bootstrap_batch(c64 *out, const u64 *lwe_in, size_t batch, const c64 *bsk, const c64 *ak,
                                 const c64 *ssk, const orc_params *p, int nthreads) {
  batch_job j = {0}; j.kind = 0; j.batch = batch; j.p = p; j.out = out; j.in0 = lwe_in;
  j.k0 = bsk; j.k1 = ak; j.k2 = ssk;
  run_batch(j, nthreads);
}
void orc_cmux_batch(u64 *c, const u64 *d0, const u64 *d1, const c64 *ggsw, size_t batch,
                    const orc_params *p, int nthreads) {
  batch_job j = {0}; j.kind = 1; j.batch = batch; j.p = p; j.out = c; j.in0 = d0; j.in1 = d1; j.in2 = ggsw;
  run_batch(j, nthreads);
}
/* The same, operands given by pointer tables (a dependency level of a MUX tree: the reference runs every
 * ready CMux as one single-threaded task on its rayon pool, circuit_processor/mod.rs:201-223). */
void orc_cmux_batch_ptrs(u64 *const *c, const u64 *const *d0, const u64 *const *d1, const c64 *const *ggsw,
                         size_t batch, const orc_params *p, int nthreads) {
  batch_job j = {0}; j.kind = 3; j.batch = batch; j.p = p; j.out = (void *)c; j.in0 = d0; j.in1 = d1; j.in2 = ggsw;
  run_batch(j, nthreads);
}
void orc_keyswitch_lwe_batch(u64 *out, const u64 *in, size_t batch, const u64 *ksk, const orc_params *p,
                             int nthreads) {
  batch_job j = {0}; j.kind = 2; j.batch = batch; j.p = p; j.out = out; j.in0 = in; j.k0 = ksk;
  run_batch(j, nthreads);
}

/* ------------------------------------------------------------------------------------ */
/* client side: RNG, keygen, encrypt, decrypt                                           */
/* ------------------------------------------------------------------------------------ */

/* The reference draws from thread_rng() (rand.rs:23,34,39) and cannot be seeded; the harness
 * uses its own xoshiro256** so that oracle, CPU baseline and GPU see identical keys/inputs. */
static u64 splitmix64(u64 *x) {
  u64 z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void orc_rng_seed(orc_rng *g, u64 seed) {
  for (int i = 0; i < 4; i++) g->s[i] = splitmix64(&seed);
  g->have_spare = 0; g->spare = 0.0;
}
static inline u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
u64 orc_rng_u64(orc_rng *g) {
  u64 *s = g->s, result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return result;
}
static double rng_unit(orc_rng *g) { return ((double)(orc_rng_u64(g) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rng_gauss(orc_rng *g) {
  if (g->have_spare) { g->have_spare = 0; return g->spare; }
  double u1 = rng_unit(g), u2 = rng_unit(g);
  double r = sqrt(-2.0 * log(u1)), th = 2.0 * M_PI * u2;
  g->spare = r * sin(th); g->have_spare = 1;
  return r * cos(th);
}
/* rand.rs:20-30 normal_torus: e = round(N(0,std) * 2^64) as i64 -> u64 */
u64 orc_rng_normal_torus(orc_rng *g, double std) {
  double e0 = rng_gauss(g) * std;
  double e = round(e0 * 18446744073709551616.0);
  i64 s;
  if (e >= 9223372036854775808.0) s = INT64_MAX; else if (e < -9223372036854775808.0) s = INT64_MIN; else s = (i64)e;
  return (u64)s;
}

/* keygen: binary secret keys (rand.rs:38-47; keys.rs SecretKey::generate) */
void orc_keygen_secret(orc_rng *g, u64 *lwe0_sk, u64 *glwe1_sk, const orc_params *p) {
  for (uint32_t i = 0; i < p->lwe_n; i++) lwe0_sk[i] = orc_rng_u64(g) & 1;
  for (uint32_t i = 0; i < p->glwe_k * p->glwe_n; i++) glwe1_sk[i] = orc_rng_u64(g) & 1;
}

/* ops/encryption/lwe_encryption.rs:36-61 encrypt_lwe_ciphertext: b = <a,s> + m + e */
void orc_encrypt_lwe(orc_rng *g, u64 *ct, const u64 *sk, uint32_t n, double std, u64 msg) {
  u64 b = 0;
  for (uint32_t i = 0; i < n; i++) { ct[i] = orc_rng_u64(g); b += ct[i] * sk[i]; }
  ct[n] = b + msg + orc_rng_normal_torus(g, std);
}
/* entities/lwe_secret_key.rs:92-107 decrypt_without_decode */
u64 orc_decrypt_lwe_raw(const u64 *ct, const u64 *sk, uint32_t n) {
  u64 dot = 0;
  for (uint32_t i = 0; i < n; i++) dot += ct[i] * sk[i];
  return ct[n] - dot;
}
/* math/torus.rs:293-300 Torus::decode */
u64 orc_decode(u64 t, uint32_t pb) {
  u64 round_bit = (t >> (64 - pb - 1)) & 1;
  u64 mask = ((u64)1 << pb) - 1;
  return ((t >> (64 - pb)) + round_bit) & mask;
}

/* c += a * s over Z_q[X]/(X^N+1) for an arbitrary small-integer s (polynomial.rs:114-197
 * polynomial_external_mad, used only by keygen/encrypt/decrypt).  Exploits sparsity of s. */
static void poly_external_mad(u64 *c, const u64 *a, const u64 *s, uint32_t n) {
  for (uint32_t j = 0; j < n; j++) {
    u64 sj = s[j];
    if (sj == 0) continue;
    /* a * sj * X^j */
    for (uint32_t i = 0; i < n - j; i++) c[i + j] += a[i] * sj;
    for (uint32_t i = n - j; i < n; i++) c[i + j - n] -= a[i] * sj;
  }
}

/* ops/encryption/glwe_encryption.rs:22-64 encrypt_glwe_ciphertext_secret_generic */
void orc_encrypt_glwe(orc_rng *g, u64 *ct, const u64 *msg, const u64 *sk, const orc_params *p) {
  uint32_t n = p->glwe_n, k = p->glwe_k;
  u64 *b = ct + (size_t)k * n;
  memset(b, 0, 8 * n);
  for (uint32_t i = 0; i < k; i++) {
    u64 *a = ct + (size_t)i * n;
    for (uint32_t j = 0; j < n; j++) a[j] = orc_rng_u64(g);
    poly_external_mad(b, a, sk + (size_t)i * n, n);
  }
  for (uint32_t j = 0; j < n; j++) b[j] += msg[j];
  if (p->glwe_std == 0.0) return;
  for (uint32_t j = 0; j < n; j++) b[j] += orc_rng_normal_torus(g, p->glwe_std);
}
/* glwe_encryption.rs:104-128 decrypt_glwe_ciphertext: msg = b - sum a_i s_i */
void orc_decrypt_glwe_raw(u64 *msg, const u64 *ct, const u64 *sk, const orc_params *p) {
  uint32_t n = p->glwe_n, k = p->glwe_k;
  u64 tmp[4096];
  memset(tmp, 0, 8 * n);
  for (uint32_t i = 0; i < k; i++) poly_external_mad(tmp, ct + (size_t)i * n, sk + (size_t)i * n, n);
  for (uint32_t j = 0; j < n; j++) msg[j] = ct[(size_t)k * n + j] - tmp[j];
}
/* ops/encryption/rlwe_encryption.rs:47-61 rlwe_generate_public_key: the RLWE public key is a secret-key encryption of the
 * zero polynomial (glwe.dim.size must be 1) */
void orc_rlwe_generate_public_key(orc_rng *g, u64 *pk, const u64 *glwe_sk, const orc_params *p) {
  u64 zero[4096];
  memset(zero, 0, 8 * p->glwe_n);
  orc_encrypt_glwe(g, pk, zero, glwe_sk, p);
}
/* rlwe_encryption.rs:140-146: u = binary_torus_polynomial, e0 / e1 = normal_torus_polynomial(glwe.std) (rand.rs:20-47) */
void orc_rlwe_sample_randomness(orc_rng *g, u64 *u, u64 *e0, u64 *e1, const orc_params *p) {
  for (uint32_t j = 0; j < p->glwe_n; j++) u[j] = orc_rng_u64(g) & 1;
  for (uint32_t j = 0; j < p->glwe_n; j++) e0[j] = p->glwe_std == 0.0 ? 0 : orc_rng_normal_torus(g, p->glwe_std);
  for (uint32_t j = 0; j < p->glwe_n; j++) e1[j] = p->glwe_std == 0.0 ? 0 : orc_rng_normal_torus(g, p->glwe_std);
}
/* rlwe_encryption.rs:125-160 rlwe_encrypt_public_impl for given randomness (the double-LWE trick):
 * ct = (p0 * u + e0, p1 * u + e1 + m), every product the exact negacyclic one of polynomial_external_mad */
void orc_rlwe_encrypt_public(u64 *ct, const u64 *encoded_msg, const u64 *pk, const u64 *u, const u64 *e0, const u64 *e1,
                             const orc_params *p) {
  uint32_t n = p->glwe_n;
  u64 *a = ct, *b = ct + n;
  memset(ct, 0, 16 * (size_t)n);                     /* ct.clear() */
  poly_external_mad(a, pk, u, n);                    /* a = p0 * u */
  for (uint32_t j = 0; j < n; j++) a[j] += e0[j];
  poly_external_mad(b, pk + n, u, n);                /* b = p1 * u */
  for (uint32_t j = 0; j < n; j++) b[j] += e1[j];
  for (uint32_t j = 0; j < n; j++) b[j] += encoded_msg[j];
}
/* ops/encryption/glev_encryption.rs:23-79: GLWE_j encrypts msg * q/B^{j+1} */
void orc_encrypt_glev(orc_rng *g, u64 *ct, const u64 *msg, const u64 *sk, const orc_params *p, orc_radix r) {
  uint32_t n = p->glwe_n;
  u64 scaled[4096];
  for (uint32_t j = 0; j < r.count; j++) {
    u64 f = (u64)1 << (64 - r.radix_log * (j + 1));
    for (uint32_t c = 0; c < n; c++) scaled[c] = msg[c] * f;
    orc_encrypt_glwe(g, ct + (size_t)j * orc_size_glwe(p), scaled, sk, p);
  }
}
/* ops/encryption/ggsw_encryption.rs:16-72: row i<k encrypts -(msg*s_i), row k encrypts msg */
void orc_encrypt_ggsw(orc_rng *g, u64 *ct, const u64 *msg, const u64 *sk, const orc_params *p, orc_radix r) {
  uint32_t n = p->glwe_n, k = p->glwe_k;
  u64 ms[4096];
  for (uint32_t i = 0; i <= k; i++) {
    if (i < k) {
      memset(ms, 0, 8 * n);
      poly_external_mad(ms, msg, sk + (size_t)i * n, n);
      for (uint32_t c = 0; c < n; c++) ms[c] = (u64)0 - ms[c];
    } else memcpy(ms, msg, 8 * n);
    orc_encrypt_glev(g, ct + (size_t)i * orc_size_glev(p, r), ms, sk, p, r);
  }
}
void orc_ggsw_fft(c64 *out, const u64 *ggsw, const orc_params *p, orc_radix r) {
  uint32_t n = p->glwe_n; size_t polys = orc_size_ggsw(p, r) / n;
  for (size_t q = 0; q < polys; q++) orc_poly_fft(ggsw + q * n, out + q * (n / 2), n);
}
void orc_ggsw_ifft(u64 *out, const c64 *ggsw_fft, const orc_params *p, orc_radix r) {
  uint32_t n = p->glwe_n; size_t polys = orc_size_ggsw(p, r) / n;
  for (size_t q = 0; q < polys; q++) orc_poly_ifft(ggsw_fft + q * (n / 2), out + q * n, n);
}

/* ---- compute key: parasol_runtime/src/crypto/keys.rs:214-280 ---- */
typedef struct {
  int tid, nthreads; u64 seed; const orc_params *p; const u64 *lwe0_sk, *glwe1_sk; c64 *bsk_fft;
} bsk_job;

/* programmable_bootstrapping.rs:34-58 generate_bootstrap_key (rayon par_iter over key bits) +
 * fft::fft_bootstrap_key.  Each key bit gets its own PRNG stream so threads are deterministic. */
static void *bsk_worker(void *arg) {
  bsk_job *j = (bsk_job *)arg;
  const orc_params *p = j->p;
  uint32_t n = p->glwe_n;
  size_t gs = orc_size_ggsw(p, p->pbs);
  u64 *ggsw = malloc(8 * gs), *msg = calloc(n, 8);
  for (uint32_t i = j->tid; i < p->lwe_n; i += j->nthreads) {
    orc_rng g; orc_rng_seed(&g, j->seed ^ (0x5851F42D4C957F2Dull * (i + 1)));
    msg[0] = j->lwe0_sk[i];
    orc_encrypt_ggsw(&g, ggsw, msg, j->glwe1_sk, p, p->pbs);
    orc_ggsw_fft(j->bsk_fft + (size_t)i * (gs / 2), ggsw, p, p->pbs);
  }
  free(ggsw); free(msg);
  return NULL;
}

void orc_keygen_compute(orc_rng *g, const u64 *lwe0_sk, const u64 *glwe1_sk, c64 *bsk_fft, u64 *ksk,
                        c64 *ssk_fft, c64 *ak_fft, const orc_params *p, int nthreads) {
  uint32_t n = p->glwe_n, k = p->glwe_k, m = n / 2, k1 = k + 1, n0 = p->lwe_n;
  /* bootstrapping key */
  if (nthreads < 1) nthreads = 1;
  {
    u64 seed = orc_rng_u64(g);
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    bsk_job *jobs = malloc(sizeof(bsk_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
      jobs[t] = (bsk_job){t, nthreads, seed, p, lwe0_sk, glwe1_sk, bsk_fft};
      pthread_create(&th[t], NULL, bsk_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
  }
  /* LWE keyswitch key: ops/keyswitch/lwe_keyswitch_key.rs:16-50; glwe_1 flat as LWE key
   * (entities/glwe_secret_key.rs:156-158; keys.rs:223-229) */
  for (uint32_t i = 0; i < k * n; i++)
    for (uint32_t j = 0; j < p->ks.count; j++) {
      u64 f = (u64)1 << (64 - p->ks.radix_log * (j + 1));
      orc_encrypt_lwe(g, ksk + ((size_t)i * p->ks.count + j) * (n0 + 1), lwe0_sk, n0, p->lwe_std, f * glwe1_sk[i]);
    }
  /* scheme switch key: ops/bootstrapping/scheme_switch.rs:22-62: GLEV_s(s_i*s_j), i<=j */
  {
    size_t glev = orc_size_glev(p, p->ss);
    u64 *ct = malloc(8 * glev), *prod = malloc(8 * n);
    size_t idx = 0;
    for (uint32_t i = 0; i < k; i++)
      for (uint32_t j = i; j < k; j++, idx++) {
        memset(prod, 0, 8 * n);
        poly_external_mad(prod, glwe1_sk + (size_t)i * n, glwe1_sk + (size_t)j * n, n);
        orc_encrypt_glev(g, ct, prod, glwe1_sk, p, p->ss);
        for (size_t q = 0; q < glev / n; q++) orc_poly_fft(ct + q * n, ssk_fft + idx * (glev / 2) + q * m, n);
      }
    free(ct); free(prod);
  }
  /* automorphism key: ops/automorphisms/mod.rs:18-44; glwe_keyswitch_key.rs:32-69 */
  {
    uint32_t logn = ilog2u(n);
    size_t glwe = orc_size_glwe(p);
    u64 *sk_k = malloc(8 * k * n), *ct = malloc(8 * glwe), *scaled = malloc(8 * n);
    c64 *dst = ak_fft;
    for (uint32_t r = 1; r <= logn; r++) {
      uint32_t kk = n / (1u << (r - 1)) + 1;
      for (uint32_t i = 0; i < k; i++) orc_poly_pow_k(sk_k + (size_t)i * n, glwe1_sk + (size_t)i * n, n, kk);
      for (uint32_t i = 0; i < k; i++)
        for (uint32_t j = 0; j < p->tr.count; j++) {
          u64 f = (u64)1 << (64 - p->tr.radix_log * (j + 1));
          for (uint32_t c = 0; c < n; c++) scaled[c] = sk_k[(size_t)i * n + c] * f;
          orc_encrypt_glwe(g, ct, scaled, glwe1_sk, p);
          for (uint32_t q = 0; q < k1; q++, dst += m) orc_poly_fft(ct + (size_t)q * n, dst, n);
        }
    }
    free(sk_k); free(ct); free(scaled);
  }
}

/* Seconds per forward negacyclic transform of degree n on this core (bench.py prints it next to the CPU baseline so the
 * FFT class of the baseline is visible: rustfft-class is 2-4 us at n = 2048). */
double orc_bench_fft_forward(uint32_t n, int iters) {
  double x[4096] __attribute__((aligned(64)));
  c64 f[2048];
  for (uint32_t j = 0; j < n; j++) x[j] = (double)((j * 7919u) % 65536u) - 32768.0;
  struct timespec t0, t1;
  orc_fft_forward(x, f, n);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < iters; i++) { x[0] = (double)i; orc_fft_forward(x, f, n); }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / (double)iters + 0.0 * f[1].re;
}
