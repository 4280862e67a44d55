/*
 * spf_oracle.h -- CPU ORACLE for the spf_b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (Sunscreen-tech/spf v0.9.0, CPU-only Rust)
 * TFHE algorithms for the circuit-bootstrapping path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call this.  The product
 * (libspf_b200.so) never does.
 *
 * Parity status: the integer / index arithmetic is pinned bit-exactly by the reference's own
 * known-answer tests (tests/test_oracle_kat.py lists every one with file:line).  The f64 FFT
 * butterflies of the reference live in the un-vendored crate rustfft 6.3.0 (Cargo.lock:2337),
 * so ciphertext-level parity for CMUX/PBS/CBS is "parity unpinned": there are no golden
 * ciphertexts, keys or seeds in the reference (RNG is thread_rng()).  Those ops are pinned the
 * way the reference's tests pin them -- by decryption.
 *
 * All citations are relative to /root/reference/.
 */
#ifndef SPF_ORACLE_H
#define SPF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } orc_c64;

/* RadixDecomposition (sunscreen_tfhe/src/params.rs) */
typedef struct { uint32_t radix_log; uint32_t count; } orc_radix;

/* parasol_runtime/src/params.rs:59-91 (Params) flattened. */
typedef struct {
  uint32_t lwe_n;    /* l0_params.dim           (637)    */
  double   lwe_std;  /* l0_params.std           (7.25e-5)*/
  uint32_t glwe_k;   /* l1_params.dim.size      (1)      */
  uint32_t glwe_n;   /* l1_params.dim.polynomial_degree (2048) */
  double   glwe_std; /* l1_params.std           (7e-16)  */
  orc_radix cbs, pbs, ks, pfks, ss, tr;
} orc_params;

void orc_default_128(orc_params *p); /* parasol_runtime/src/params.rs:107-134 */

/* ---- entity sizes in elements (the OverlaySize::size functions) ---- */
size_t orc_size_lwe(uint32_t n);                                   /* n+1 u64                  */
size_t orc_size_glwe(const orc_params *p);                         /* (k+1)N u64               */
size_t orc_size_glev(const orc_params *p, orc_radix r);            /* l(k+1)N u64              */
size_t orc_size_ggsw(const orc_params *p, orc_radix r);            /* (k+1) l (k+1) N  u64     */
size_t orc_size_ggsw_fft(const orc_params *p, orc_radix r);        /* (k+1) l (k+1) N/2 c64    */
size_t orc_size_bsk_fft(const orc_params *p);                      /* n * ggsw_fft(pbs)        */
size_t orc_size_ksk(const orc_params *p);                          /* kN * l_ks * (n+1) u64    */
size_t orc_size_ak_fft(const orc_params *p);                       /* log2N * k * l_tr (k+1) N/2 */
size_t orc_size_ssk_fft(const orc_params *p);                      /* tri(k) * l_ss (k+1) N/2  */

/* ---- math ---- */
void orc_fft_forward(const double *x, orc_c64 *out, uint32_t n);   /* TwistedFft::forward */
void orc_fft_reverse(const orc_c64 *in, double *out, uint32_t n);  /* TwistedFft::reverse (incl. round) */
void orc_poly_fft(const uint64_t *p, orc_c64 *out, uint32_t n);    /* PolynomialRef::fft  */
void orc_poly_ifft(const orc_c64 *in, uint64_t *p, uint32_t n);    /* PolynomialFftRef::ifft */
void orc_mod_pow2_q_f64(uint64_t *c, const double *a, size_t len); /* vector_mod_pow2_q_f64, q=2^64 */
uint64_t orc_radix_round(uint64_t x, orc_radix r);
/* one step of vector_next_decomp over len values: s (state) updated, r (digit) written */
void orc_next_decomp(uint64_t *s, uint64_t *r, size_t len, uint32_t radix_log);
uint64_t orc_modulus_switch(uint64_t x, uint32_t log_chi, uint32_t log_v, uint32_t log_modulus);
void orc_poly_pow_k(uint64_t *pk, const uint64_t *p, uint32_t n, uint32_t k);
void orc_shr_round(uint64_t *y, const uint64_t *x, size_t len, uint32_t n);
void orc_poly_mul_monomial(uint64_t *p, uint32_t n, int64_t degree); /* in place, negacyclic */
void orc_generate_lut(uint64_t *lut_poly, const uint64_t *table /*[v][p]*/, uint32_t v,
                      uint32_t n, uint32_t plaintext_bits);

/* ---- evaluation ops (sunscreen_tfhe::ops) ---- */
void orc_glwe_ggsw_mad(orc_c64 *c_fft, const uint64_t *glwe, const orc_c64 *ggsw_fft,
                       const orc_params *p, orc_radix r);
void orc_glwe_fft_ifft(const orc_c64 *c_fft, uint64_t *glwe, const orc_params *p);
void orc_cmux(uint64_t *c, const uint64_t *d0, const uint64_t *d1, const orc_c64 *ggsw_fft,
              const orc_params *p, orc_radix r);
void orc_glev_cmux(uint64_t *c, const uint64_t *d0, const uint64_t *d1, const orc_c64 *ggsw_fft,
                   const orc_params *p, orc_radix glev_r, orc_radix ggsw_r);
void orc_keyswitch_glwe(uint64_t *out, const uint64_t *in, const orc_c64 *ksk_fft,
                        const orc_params *p, orc_radix r);
void orc_trace(uint64_t *out, const uint64_t *x, const orc_c64 *ak_fft, const orc_params *p);
void orc_scheme_switch_fft(orc_c64 *ggsw_out, const uint64_t *glev, const orc_c64 *ssk_fft,
                           const orc_params *p);
void orc_pbs_generalized(uint64_t *glwe_out, const uint64_t *lwe_in, const uint64_t *lut_glwe,
                         const orc_c64 *bsk_fft, uint32_t log_chi, uint32_t log_v,
                         const orc_params *p);
void orc_pbs_univariate(uint64_t *lwe_out /*kN+1*/, const uint64_t *lwe_in, const uint64_t *lut_glwe,
                        const orc_c64 *bsk_fft, const orc_params *p);
void orc_cbs_lut(uint64_t *lut_glwe, const orc_params *p);
void orc_cbs_pbs_stage(uint64_t *glwe_out, const uint64_t *lwe_in, const orc_c64 *bsk_fft,
                       const orc_params *p);                 /* hi_noise_lwe_to_lo_noise_glwe */
void orc_cbs_trace_stage(uint64_t *glev_out, const uint64_t *glwe_in, const orc_c64 *ak_fft,
                         const orc_params *p);               /* mod_switch_trace_and_rotate */
void orc_circuit_bootstrap(orc_c64 *ggsw_out, const uint64_t *lwe_in, const orc_c64 *bsk_fft,
                           const orc_c64 *ak_fft, const orc_c64 *ssk_fft, const orc_params *p);
void orc_keyswitch_lwe(uint64_t *lwe0_out, const uint64_t *lwe1_in, const uint64_t *ksk,
                       const orc_params *p);
void orc_sample_extract(uint64_t *lwe_out, const uint64_t *glwe, uint32_t h, const orc_params *p);
void orc_glwe_add(uint64_t *c, const uint64_t *a, const uint64_t *b, const orc_params *p);
void orc_glwe_not(uint64_t *c, const uint64_t *a, const orc_params *p);
void orc_glwe_mul_xn(uint64_t *c, const uint64_t *a, uint32_t n, const orc_params *p);
void orc_multiply_glwe_ggsw(uint64_t *out, const uint64_t *glwe, const orc_c64 *ggsw_fft,
                            const orc_params *p);

/* batch CBS over nthreads host threads: the reference's parallelisation model (one
 * single-threaded op per task, tasks spread over the pool; circuit_processor/mod.rs:201-209) */
void orc_circuit_bootstrap_batch(orc_c64 *ggsw_out, const uint64_t *lwe_in, size_t batch,
                                 const orc_c64 *bsk_fft, const orc_c64 *ak_fft,
                                 const orc_c64 *ssk_fft, const orc_params *p, int nthreads);
void orc_cmux_batch(uint64_t *c, const uint64_t *d0, const uint64_t *d1, const orc_c64 *ggsw_fft,
                    size_t batch, const orc_params *p, int nthreads);
void orc_cmux_batch_ptrs(uint64_t *const *c, const uint64_t *const *d0, const uint64_t *const *d1,
                         const orc_c64 *const *ggsw_fft, size_t batch, const orc_params *p, int nthreads);
void orc_keyswitch_lwe_batch(uint64_t *lwe0_out, const uint64_t *lwe1_in, size_t batch,
                             const uint64_t *ksk, const orc_params *p, int nthreads);

/* ---- client side (keygen / encrypt / decrypt), seeded PRNG of the harness' own ---- */
typedef struct { uint64_t s[4]; int have_spare; double spare; } orc_rng;
void orc_rng_seed(orc_rng *g, uint64_t seed);
uint64_t orc_rng_u64(orc_rng *g);
uint64_t orc_rng_normal_torus(orc_rng *g, double std); /* rand.rs:20-30 */

void orc_keygen_secret(orc_rng *g, uint64_t *lwe0_sk /*n*/, uint64_t *glwe1_sk /*kN*/,
                       const orc_params *p);
/* ComputeKeyNonFft::generate + .fft (parasol_runtime/src/crypto/keys.rs:214-280) */
void orc_keygen_compute(orc_rng *g, const uint64_t *lwe0_sk, const uint64_t *glwe1_sk,
                        orc_c64 *bsk_fft, uint64_t *ksk, orc_c64 *ssk_fft, orc_c64 *ak_fft,
                        const orc_params *p, int nthreads);

void orc_encrypt_lwe(orc_rng *g, uint64_t *ct, const uint64_t *sk, uint32_t n, double std,
                     uint64_t torus_msg);
uint64_t orc_decrypt_lwe_raw(const uint64_t *ct, const uint64_t *sk, uint32_t n);
uint64_t orc_decode(uint64_t torus, uint32_t plaintext_bits);
void orc_encrypt_glwe(orc_rng *g, uint64_t *ct, const uint64_t *torus_msg, const uint64_t *sk,
                      const orc_params *p);
void orc_decrypt_glwe_raw(uint64_t *torus_msg, const uint64_t *ct, const uint64_t *sk,
                          const orc_params *p);
/* RLWE public-key encryption (ops/encryption/rlwe_encryption.rs:47-160), k = 1 */
void orc_rlwe_generate_public_key(orc_rng *g, uint64_t *pk /*2N*/, const uint64_t *glwe_sk, const orc_params *p);
void orc_rlwe_sample_randomness(orc_rng *g, uint64_t *u, uint64_t *e0, uint64_t *e1, const orc_params *p);
void orc_rlwe_encrypt_public(uint64_t *ct, const uint64_t *encoded_msg, const uint64_t *pk, const uint64_t *u,
                             const uint64_t *e0, const uint64_t *e1, const orc_params *p);
void orc_encrypt_glev(orc_rng *g, uint64_t *ct, const uint64_t *msg, const uint64_t *sk,
                      const orc_params *p, orc_radix r);
void orc_encrypt_ggsw(orc_rng *g, uint64_t *ct, const uint64_t *msg, const uint64_t *sk,
                      const orc_params *p, orc_radix r);
void orc_ggsw_fft(orc_c64 *out, const uint64_t *ggsw, const orc_params *p, orc_radix r);
void orc_ggsw_ifft(uint64_t *out, const orc_c64 *ggsw_fft, const orc_params *p, orc_radix r);

int orc_hw_threads(void);
double orc_bench_fft_forward(uint32_t n, int iters); /* seconds per forward transform on this core */

#ifdef __cplusplus
}
#endif
#endif
