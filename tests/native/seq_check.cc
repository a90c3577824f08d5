// Host build of the closed-form jump-ahead (volumetric_mapping_b200/csrc/vmap_seq.cuh) next to
// the plain sequential loop it must reproduce; driven by tests/test_seq_closed_form.py.
#include "../../volumetric_mapping_b200/csrc/vmap_seq.cuh"

extern "C" {

uint64_t seq_advance_closed(double* t, double dt, double tau, uint64_t max_steps) {
  return vmapseq::seq_advance(t, dt, tau, max_steps);
}

uint64_t seq_advance_loop(double* t_io, double dt, double tau, uint64_t max_steps) {
  double t = *t_io;
  uint64_t n = 0;
  while (n < max_steps && t < tau) {
    t = vmapseq::seq_add(t, dt);
    n++;
  }
  *t_io = t;
  return n;
}

// bulk driver: returns the number of mismatching cases
uint64_t seq_check_many(const double* t0, const double* dt, const double* tau, const uint64_t* cap,
                        uint64_t n_cases, uint64_t* first_bad) {
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n_cases; i++) {
    double a = t0[i], b = t0[i];
    const uint64_t na = vmapseq::seq_advance(&a, dt[i], tau[i], cap[i]);
    const uint64_t nb = seq_advance_loop(&b, dt[i], tau[i], cap[i]);
    if (na != nb || vmapseq::seq_bits(a) != vmapseq::seq_bits(b)) {
      if (!bad && first_bad) *first_bad = i;
      bad++;
    }
  }
  return bad;
}
}
