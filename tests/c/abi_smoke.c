/* Plain-C consumer of the drop-in boundary (include/spf_b200.h): host-only entry points, no GPU needed.
 * Built and run by tests/test_abi.py::test_plain_c_consumer. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spf_b200.h"

int main(void) {
  spf_params p;
  spf_b200_default_128(&p);
  if (spf_b200_len_lwe_l0(&p) != 638 || spf_b200_len_ggsw_l1(&p) != 16384) return 1;

  /* MUX circuit of an 8 x 8 multiplier: the reference ships 3228 multiplexers for it */
  spf_mux_node *mux = NULL;
  size_t n_mux = 0, gates = 0;
  if (spf_b200_mux_circuit(SPF_MUX_UNSIGNED_MULTIPLIER, 8, 8, 0, &mux, &n_mux) != SPF_OK) return 2;
  for (size_t i = 0; i < n_mux; i++) gates += mux[i].op == SPF_MUX_MUX;
  spf_b200_mux_free(mux);
  if (gates != 3228) return 3;

  /* a malformed graph is rejected by the host-only planner with a message */
  static uint64_t lwe0[638];
  spf_node bad[2];
  memset(bad, 0, sizeof bad);
  bad[0].op = SPF_OP_INPUT_LWE0; bad[0].in[0] = bad[0].in[1] = bad[0].in[2] = -1; bad[0].io = lwe0;
  bad[1].op = SPF_OP_KEYSWITCH_L1_TO_L0; bad[1].in[0] = 0; bad[1].in[1] = bad[1].in[2] = -1;
  if (spf_b200_graph_plan(&p, bad, 2, 1, NULL, NULL) != SPF_E_GRAPH) return 4;
  if (!strstr(spf_b200_last_error(NULL), "wrong ciphertext kind")) return 5;

  /* a well-formed one: levels and owners */
  static uint64_t glwe[4096];
  spf_node ok[3];
  memset(ok, 0, sizeof ok);
  for (int i = 0; i < 3; i++) ok[i].in[0] = ok[i].in[1] = ok[i].in[2] = -1;
  ok[0].op = SPF_OP_INPUT_GLWE1; ok[0].io = glwe;
  ok[1].op = SPF_OP_SAMPLE_EXTRACT; ok[1].in[0] = 0; ok[1].arg = 5;
  ok[2].op = SPF_OP_KEYSWITCH_L1_TO_L0; ok[2].in[0] = 1;
  int32_t level[3], owner[3];
  if (spf_b200_graph_plan(&p, ok, 3, 2, level, owner) != SPF_OK) return 6;
  if (level[0] != 0 || level[1] != 1 || level[2] != 2 || owner[0] != -1 || owner[1] != owner[2] || owner[1] < 0) return 7;

  printf("c abi ok: %zu multiplexers, levels %d %d %d, owner %d\n", gates, level[0], level[1], level[2], owner[1]);
  return 0;
}
