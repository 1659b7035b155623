// Host build of the SWAR primitives in csrc/mpc_device.cuh for exhaustive CPU checks (tests/test_swar_host.py).
#include <string.h>
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

// ---- secondary variants (csrc/mpc_variants.cuh) on the host ----
#include "mpc_variants.cuh"
extern "C" void t_variant_run(int alg, const unsigned char* lines, unsigned long long n, unsigned* sizes, unsigned long long* counts) {
  for (unsigned long long i = 0; i < n; i++) {
    uint32_t x[32];
    __builtin_memcpy(x, lines + i * 128, 128);
    if (alg == 1) {
      int st;
      sizes[i] = mpcvar::bdi_block(x, &st);
      counts[st]++;
    } else if (alg == 2) {
      uint64_t c8;
      sizes[i] = mpcvar::fpc_block(x, &c8);
      for (int p = 0; p < 8; p++) counts[p] += (c8 >> (8 * p)) & 0xff;
    } else {
      uint64_t p8;
      uint32_t w;
      sizes[i] = mpcvar::bpc_block(x, &p8, &w);
      for (int p = 0; p < 7; p++) counts[p] += (p8 >> (8 * p)) & 0xff;
      counts[7] += w;
    }
  }
}

// the same for lines of W words (32- / 64-byte lines)
template <int W>
static void variant_run_w(int alg, const unsigned char* lines, unsigned long long n, unsigned* sizes, unsigned long long* counts) {
  for (unsigned long long i = 0; i < n; i++) {
    uint32_t x[32] = {0};
    __builtin_memcpy(x, lines + i * 4 * W, 4 * W);
    if (alg == 1) {
      int st;
      sizes[i] = mpcvar::bdi_block<W>(x, &st);
      counts[st]++;
    } else if (alg == 2) {
      uint64_t c8;
      sizes[i] = mpcvar::fpc_block<W>(x, &c8);
      for (int p = 0; p < 8; p++) counts[p] += (c8 >> (8 * p)) & 0xff;
    } else {
      uint64_t p8;
      uint32_t w;
      sizes[i] = mpcvar::bpc_block<W>(x, &p8, &w);
      for (int p = 0; p < 7; p++) counts[p] += (p8 >> (8 * p)) & 0xff;
      counts[7] += w;
    }
  }
}
extern "C" void t_variant_run_l(int alg, const unsigned char* lines, unsigned long long n, unsigned L, unsigned* sizes, unsigned long long* counts) {
  if (L == 32) variant_run_w<8>(alg, lines, n, sizes, counts);
  else if (L == 64) variant_run_w<16>(alg, lines, n, sizes, counts);
  else variant_run_w<32>(alg, lines, n, sizes, counts);
}

// BDI closed-form range tests vs the written-down reduceSign rule (BDI.cpp:203-218)
extern "C" int t_bdi_fits(unsigned long long x, int D) {
  return D == 1 ? mpcvar::bdi_fits64<1>(x) : D == 2 ? mpcvar::bdi_fits64<2>(x) : mpcvar::bdi_fits64<4>(x);
}
extern "C" int t_bdi_fits_rule(unsigned long long x, int D) {
  const unsigned long long limit = D == 1 ? 0xffull : D == 2 ? 0xffffull : 0xffffffffull;
  return mpcvar::bdi_reduce_sign(x) <= limit;
}
extern "C" int t_bdi_delta32(unsigned base, unsigned v, int D) {
  return D == 1 ? mpcvar::bdi_delta_fits32<1>(base, v) : mpcvar::bdi_delta_fits32<2>(base, v);
}

// PATTERN: per-block selection + immediates of the selected layout (mpcvar::pattern_block) and the content hash
extern "C" void t_pattern_run(const unsigned char* lines, unsigned long long n, unsigned* sizes, int* sels, unsigned* imms,
                              unsigned long long* hashes) {
  for (unsigned long long i = 0; i < n; i++) {
    uint32_t x[32];
    memcpy(x, lines + i * 128, 128);
    sizes[i] = mpcvar::pattern_block(x, &sels[i], &imms[i]);
    hashes[i] = mpcvar::block_hash64(x);
  }
}
