// serial.inl -- the reference's serialized key / ciphertext layouts, host only.
// bincode 1.3.3 (Cargo.lock:244-245) with fixint little-endian encoding: every sunscreen_tfhe
// entity is a single `data` sequence (sunscreen_tfhe/src/dst.rs:31-33) = u64 length || elements;
// Torus<u64> is serde(transparent) (math/torus.rs:217-220), Complex<f64> is (re, im).
// ComputeKey = bs_key || ks_key || ss_key || auto_key (parasol_runtime/src/crypto/keys.rs:306-318).
// Deserialisation mirrors parasol_runtime/src/safe_bincode.rs:16-27: a byte limit of
// GetSize::get_size(params), trailing bytes allowed, then check_is_valid (length == OverlaySize).
// Included by capi.cu inside its anonymous-namespace helpers' scope (uses fail(), len_*()).

namespace {

uint64_t rd_u64le(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
  return v;
}
void wr_u64le(uint8_t* p, uint64_t v) {
  for (int i = 0; i < 8; i++) { p[i] = (uint8_t)(v & 0xFF); v >>= 8; }
}

struct SeqSpec {
  size_t elems;      // expected element count (OverlaySize::size)
  size_t elem_size;  // 8 (u64) or 16 (Complex<f64>)
  const char* name;
};

// Walks `n` consecutive sequences; off[i] receives the byte offset of sequence i's first element.
int parse_seqs(const uint8_t* buf, size_t len, size_t limit, const SeqSpec* specs, int n, size_t* off) {
  if (!buf || !off) return fail(nullptr, SPF_E_INVALID, "NULL buffer");
  size_t pos = 0;
  for (int i = 0; i < n; i++) {
    if (pos + 8 > len) return fail(nullptr, SPF_E_INVALID, std::string(specs[i].name) + ": truncated length field");
    if (pos + 8 > limit) return fail(nullptr, SPF_E_INVALID, std::string(specs[i].name) + ": byte limit exceeded");
    const uint64_t got = rd_u64le(buf + pos);
    pos += 8;
    if (got != (uint64_t)specs[i].elems)
      return fail(nullptr, SPF_E_INVALID, std::string(specs[i].name) + ": sequence length " + std::to_string(got) +
                                              " != expected " + std::to_string(specs[i].elems));
    const size_t bytes = specs[i].elems * specs[i].elem_size;
    if (bytes > len - pos) return fail(nullptr, SPF_E_INVALID, std::string(specs[i].name) + ": truncated data");
    if (pos + bytes > limit) return fail(nullptr, SPF_E_INVALID, std::string(specs[i].name) + ": byte limit exceeded");
    off[i] = pos;
    pos += bytes;
  }
  return 0;
}

void compute_key_specs(const spf_params* p, SeqSpec (&s)[4]) {
  s[0] = SeqSpec{spf_b200_len_bsk(p), 16, "bs_key"};
  s[1] = SeqSpec{spf_b200_len_ksk(p), 8, "ks_key"};
  s[2] = SeqSpec{spf_b200_len_ssk(p), 16, "ss_key"};
  s[3] = SeqSpec{spf_b200_len_ak(p), 16, "auto_key"};
}

int ct_elems(const spf_params* p, int kind, size_t* elems) {
  switch (kind) {
    case SPF_CT_LWE0: *elems = spf_b200_len_lwe_l0(p); return 0;
    case SPF_CT_LWE1: *elems = spf_b200_len_lwe_l1(p); return 0;
    case SPF_CT_GLWE1: *elems = spf_b200_len_glwe_l1(p); return 0;
    case SPF_CT_GLEV1: *elems = spf_b200_len_glev_l1(p); return 0;
    default: return fail(nullptr, SPF_E_INVALID, "unknown ciphertext kind (L1GgswCiphertext is not serialisable, encryption.rs:94-98)");
  }
}

}  // namespace

extern "C" {

size_t spf_b200_serialized_size_compute_key(const spf_params* p) {
  if (!p) return 0;
  SeqSpec s[4];
  compute_key_specs(p, s);
  size_t total = 0;
  for (const SeqSpec& q : s) total += 8 + q.elems * q.elem_size;
  return total;
}

// ComputeKey::get_size (keys.rs:326-349): every key counted as Complex<f64> plus 4 length fields.
size_t spf_b200_serialized_limit_compute_key(const spf_params* p) {
  if (!p) return 0;
  SeqSpec s[4];
  compute_key_specs(p, s);
  size_t total = 0;
  for (const SeqSpec& q : s) total += q.elems;
  return total * 16 + 4 * 8;
}

int spf_b200_parse_compute_key(const spf_params* params, const uint8_t* buf, size_t len, size_t offsets[4]) {
  if (int rc = validate_params(params)) return rc;
  SeqSpec s[4];
  compute_key_specs(params, s);
  return parse_seqs(buf, len, spf_b200_serialized_limit_compute_key(params), s, 4, offsets);
}

int spf_b200_write_compute_key(const spf_params* params, uint8_t* out, size_t cap, const double* bsk_fft,
                               const uint64_t* ksk, const double* ssk_fft, const double* ak_fft, size_t* written) {
  if (int rc = validate_params(params)) return rc;
  if (!out || !bsk_fft || !ksk || !ssk_fft || !ak_fft) return fail(nullptr, SPF_E_INVALID, "NULL buffer");
  const size_t need = spf_b200_serialized_size_compute_key(params);
  if (cap < need) return fail(nullptr, SPF_E_INVALID, "output buffer too small");
  SeqSpec s[4];
  compute_key_specs(params, s);
  const void* src[4] = {bsk_fft, ksk, ssk_fft, ak_fft};
  size_t pos = 0;
  for (int i = 0; i < 4; i++) {
    wr_u64le(out + pos, s[i].elems);
    pos += 8;
    memcpy(out + pos, src[i], s[i].elems * s[i].elem_size);  // x86-64 / aarch64 hosts are little-endian
    pos += s[i].elems * s[i].elem_size;
  }
  if (written) *written = pos;
  return 0;
}

int spf_b200_create_from_serialized(const spf_params* params, const uint8_t* buf, size_t len, int device,
                                    spf_b200_ctx** out) {
  size_t off[4];
  if (int rc = spf_b200_parse_compute_key(params, buf, len, off)) return rc;
  // cudaMemcpy has no alignment requirement on the host side: upload straight out of the buffer.
  return spf_b200_create(params, (const double*)(buf + off[0]), spf_b200_len_bsk(params),
                         (const uint64_t*)(buf + off[1]), spf_b200_len_ksk(params), (const double*)(buf + off[2]),
                         spf_b200_len_ssk(params), (const double*)(buf + off[3]), spf_b200_len_ak(params), device, out);
}

// GetSize for the ciphertext newtypes (encryption.rs:454-519): (size + 1) * 8 bytes, which for
// these single-sequence types is also the exact serialized size.
size_t spf_b200_serialized_size_ciphertext(const spf_params* p, int kind) {
  size_t e = 0;
  if (!p || ct_elems(p, kind, &e)) return 0;
  return (e + 1) * 8;
}

int spf_b200_parse_ciphertext(const spf_params* params, int kind, const uint8_t* buf, size_t len, size_t* offset) {
  if (!params) return fail(nullptr, SPF_E_INVALID, "params is NULL");
  size_t e = 0;
  if (int rc = ct_elems(params, kind, &e)) return rc;
  const SeqSpec s{e, 8, "ciphertext"};
  return parse_seqs(buf, len, (e + 1) * 8, &s, 1, offset);
}

int spf_b200_write_ciphertext(const spf_params* params, int kind, const uint64_t* data, uint8_t* out, size_t cap,
                              size_t* written) {
  if (!params || !data || !out) return fail(nullptr, SPF_E_INVALID, "NULL argument");
  size_t e = 0;
  if (int rc = ct_elems(params, kind, &e)) return rc;
  if (cap < (e + 1) * 8) return fail(nullptr, SPF_E_INVALID, "output buffer too small");
  wr_u64le(out, e);
  memcpy(out + 8, data, e * 8);
  if (written) *written = (e + 1) * 8;
  return 0;
}

}  // extern "C"
