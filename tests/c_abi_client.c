/* Plain-C client of libtramp_b200.so: the header must compile as C99 and the
 * entry points must link and reject bad arguments without a GPU.  Built and run
 * by tests/test_host_logic.py::test_c_abi_from_plain_c. */
#include <stdio.h>
#include <string.h>
#include "tramp_b200.h"
int main(void) {
  trb_factor f; memset(&f, 0, sizeof f); f.kind = 99;
  if (trb_version() != TRB_VERSION) return 1;
  if (trb_sizeof_factor() != sizeof(trb_factor)) return 2;
  if (trb_sizeof_sweep() != sizeof(trb_sweep)) return 3;
  if (trb_sizeof_se() != sizeof(trb_se)) return 4;
  int rc = trb_factor_posterior(&f, 1, 4, 4, (const double*)8, 0, (const double*)8, NULL, (double*)8, (double*)8, 0, NULL);
  if (rc != TRB_ERR_INVALID || !strstr(trb_last_error(), "unknown factor kind")) return 5;
  rc = trb_se_run(NULL, 0, 1, NULL);
  if (rc != TRB_ERR_INVALID) return 6;
  printf("ok %d\n", trb_device_sm_count());
  return 0;
}
