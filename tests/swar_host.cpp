// Host build of the SWAR primitives in csrc/mpc_device.cuh for exhaustive CPU checks (tests/test_swar_host.py).
#include "mpc_device.cuh"
extern "C" {
unsigned t_row2_cost(unsigned w, unsigned* nz) { return mpcdev::row2_cost(w, nz); }
unsigned t_sub(unsigned a, unsigned b) { return mpcdev::sub_u8x4(a, b); }
unsigned t_add(unsigned a, unsigned b) { return mpcdev::add_u8x4(a, b); }
unsigned t_xc(unsigned r, unsigned keep) { return mpcdev::xor_planes_consecutive(r, keep); }
unsigned t_xf(unsigned r, unsigned keep) { return mpcdev::xor_planes_first(r, keep); }
unsigned t_zero_run_cost(unsigned long long z) { return mpcdev::zero_run_cost(z); }
unsigned t_lzr(unsigned long long z, unsigned rows) { return mpcdev::leading_zero_rows(z, rows); }
// all 65536 row values: out[v] = cost of row v placed in the low / high halfword next to `other`
void t_row_costs(unsigned other, unsigned* lo, unsigned* hi) {
  for (unsigned v = 0; v < 65536; v++) {
    unsigned nz;
    unsigned c = mpcdev::row2_cost(v | (other << 16), &nz);
    lo[v] = (c & 0xffff) | ((nz & 1) << 31);
    c = mpcdev::row2_cost((v << 16) | other, &nz);
    hi[v] = (c >> 16) | ((nz >> 16) << 31);
  }
}
}
