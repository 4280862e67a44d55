// C++17 consumer of include/spf_b200.hpp: the host-only parts run here (MUX circuits, graph construction, planner,
// error mapping); the GPU parts (Evaluation, CircuitProcessor, CompiledGraph) are compiled and linked, and run only
// when the program is started with the argument "gpu" on a box that has one.
// Built and run by tests/test_abi.py::test_cpp_host_mirror.
#include <cstdio>
#include <cstring>
#include <string>

#include "spf_b200.hpp"

static int fail(int code, const char* what) {
  std::fprintf(stderr, "hpp_smoke: check %d failed: %s\n", code, what);
  return code;
}

// Evaluation::circuit_bootstrap + cmux + an 8-bit add graph, as examples/basic_add does -- needs a GPU and keys.
static void gpu_path(const spf_params& p, const std::vector<double>& bsk, const std::vector<std::uint64_t>& ksk,
                     const std::vector<double>& ssk, const std::vector<double>& ak, std::vector<std::uint64_t>& io) {
  spf::Evaluation ev(p, bsk, ksk, ssk, ak, 0);
  std::vector<double> ggsw;
  std::vector<std::uint64_t> lwe0(ev.len_lwe_l0(), 0), out;
  ev.circuit_bootstrap(ggsw, lwe0);
  std::vector<std::uint64_t> a(ev.len_glwe_l1(), 0), b(ev.len_glwe_l1(), 1);
  ev.cmux(out, ggsw, a, b);
  spf::FheCircuit c;
  const int x = c.input(SPF_OP_INPUT_GLWE1, io.data());
  c.output(SPF_OP_OUTPUT_GLWE1, c.add(SPF_OP_NOT, x), io.data() + ev.len_glwe_l1());
  spf::CircuitProcessor proc(ev);
  proc.run_graph_blocking(c);
  spf::CompiledGraph g = proc.compile(c);
  g.run();
  std::printf("gpu path ok: %d levels, %llu launches\n", g.levels(), (unsigned long long)g.launches());
}

int main(int argc, char** argv) {
  const spf_params p = spf::default_128();
  if (p.lwe_n != 637 || p.glwe_n != 2048) return fail(1, "default_128");

  // mux_circuits::add::ripple_carry_adder(8, 8, false) on plaintext bits: 2 + 7 (inputs interleaved a0 b0 a1 b1 ...)
  const spf::MuxCircuit adder = spf::MuxCircuit::ripple_carry_adder(8, 8);
  std::vector<int> bits;
  for (int i = 0; i < 8; i++) { bits.push_back((2 >> i) & 1); bits.push_back((7 >> i) & 1); }
  int sum = 0;
  const std::vector<int> out = adder.evaluate(bits);
  for (std::size_t i = 0; i < out.size(); i++) sum |= out[i] << i;
  if (sum != 9 || out.size() != 9) return fail(2, "adder");
  if (spf::MuxCircuit::unsigned_multiplier(8, 8).mux_gates() != 3228) return fail(3, "multiplier size");

  // the adder as an FheCircuit: selectors are stand-in GGSW constants, outputs go to caller buffers
  std::vector<std::uint64_t> bufs(9 * 4096, 0);
  spf::FheCircuit c;
  std::vector<int> sel;
  for (int i = 0; i < 16; i++) sel.push_back(c.add(SPF_OP_ONE_GGSW1));
  const std::vector<int> sums = spf::insert_mux_circuit(c, adder, sel);
  for (std::size_t i = 0; i < sums.size(); i++) c.output(SPF_OP_OUTPUT_GLWE1, sums[i], bufs.data() + 4096 * i);
  const auto [level, owner] = c.plan(p, 2);
  int deepest = 0;
  for (std::int32_t l : level) deepest = l > deepest ? l : deepest;
  if (deepest < 16 || level.size() != c.size()) return fail(4, "plan");

  // error behaviour: malformed graphs and bad sizes throw spf::Error with the C ABI's code and message
  try {
    spf::FheCircuit bad;
    static std::uint64_t lwe0[638];
    bad.add(SPF_OP_KEYSWITCH_L1_TO_L0, bad.input(SPF_OP_INPUT_LWE0, lwe0));
    bad.plan(p);
    return fail(5, "malformed graph accepted");
  } catch (const spf::Error& e) {
    if (e.code != SPF_E_GRAPH || !std::strstr(e.what(), "wrong ciphertext kind")) return fail(6, e.what());
  }
  try {
    spf::MuxCircuit::unsigned_multiplier(0, 4);
    return fail(7, "bad size accepted");
  } catch (const spf::Error& e) {
    if (e.code != SPF_E_INVALID) return fail(8, e.what());
  }

  // circuits/mul.rs mul_impl in C++: an 18 x 18 product = four 9 x 9 blocks + the 4-way reduction behind one bootstrap
  // level; its Boolean skeleton (conversions passed through) multiplies
  {
    spf::FheCircuit m;
    std::vector<int> ma, mb;
    for (int i = 0; i < 18; i++) ma.push_back(m.add(SPF_OP_ONE_GGSW1));
    for (int i = 0; i < 18; i++) mb.push_back(m.add(SPF_OP_ONE_GGSW1));
    const std::vector<int> prod = spf::append_uint_multiply(m, ma, mb);
    if (prod.size() != 36) return fail(9, "product width");
    const unsigned long long x = 0x2F1A7ull & 0x3FFFF, y = 0x31C59ull & 0x3FFFF;
    std::vector<int> val(m.size(), 0);
    std::size_t n_cbs = 0;
    for (std::size_t i = 0; i < m.size(); i++) {
      const spf_node& n = m.nodes()[i];
      switch (n.op) {
        case SPF_OP_ONE_GGSW1: {
          std::size_t k = 0;
          while (k < 18 && ma[k] != (int)i) k++;
          if (k < 18) { val[i] = (x >> k) & 1; break; }
          k = 0;
          while (mb[k] != (int)i) k++;
          val[i] = (y >> k) & 1;
        } break;
        case SPF_OP_ONE_GLWE1: val[i] = 1; break;
        case SPF_OP_ZERO_GLWE1: val[i] = 0; break;
        case SPF_OP_CMUX: val[i] = val[n.in[0]] ? val[n.in[2]] : val[n.in[1]]; break;
        case SPF_OP_CIRCUIT_BOOTSTRAP: n_cbs++; val[i] = val[n.in[0]]; break;
        default: val[i] = val[n.in[0]]; break;
      }
    }
    unsigned long long got = 0;
    for (std::size_t i = 0; i < prod.size(); i++) got |= (unsigned long long)val[prod[i]] << i;
    if (got != x * y || n_cbs != 4 * 18) return fail(10, "append_uint_multiply skeleton");
    m.plan(p, 4);  // kinds line up, no cycle
  }

  if (argc > 1 && std::string(argv[1]) == "gpu") {
    std::vector<double> bsk(2 * spf_b200_len_bsk(&p)), ssk(2 * spf_b200_len_ssk(&p)), ak(2 * spf_b200_len_ak(&p));
    std::vector<std::uint64_t> ksk(spf_b200_len_ksk(&p)), io(2 * 4096, 0);
    gpu_path(p, bsk, ksk, ssk, ak, io);  // all-zero keys: exercises the calls, not the cryptography
  }
  std::printf("hpp ok: 2 + 7 = %d, %zu graph nodes over %d levels\n", sum, c.size(), deepest + 1);
  return 0;
}
