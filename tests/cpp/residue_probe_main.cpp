// Host check of the residue arithmetic the config compiler generates (tools/specgen --residues): every generated residue word
// against the per-byte definition -- predicted byte = line[psrc] as it is / plus a constant / shifted (PredictorModule.cpp:37-173),
// residue = (line[xsrc] - predicted) mod 256 with the root byte first and raw (ResidueModule.cpp:12-41).
// Built by tests/test_specgen_residues.py: g++ -I csrc -DPROBE_FILE="..." residue_probe_main.cpp
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "mpc_device.cuh"

using mpcdev::add_u8x4;
using mpcdev::sub_u8x4;
// host forms of the device-only helpers of mpc_spec.cuh
static inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {  // __byte_perm: selector nibbles 0-3 = bytes of a, 4-7 = bytes of b
  const uint64_t ab = ((uint64_t)b << 32) | a;
  uint32_t d = 0;
  for (int q = 0; q < 4; q++) d |= (uint32_t)((ab >> (8 * ((sel >> (4 * q)) & 7u))) & 0xffu) << (8 * q);
  return d;
}
static inline uint32_t sub_u8x4_shared(uint32_t a, uint32_t b, uint32_t ah, uint32_t bh) {
  const uint32_t t = ah - bh + 0x80808080u;
  return t ^ ((a ^ ~b) & 0x80808080u);
}
template <class F> static inline uint32_t shiftmix(uint32_t p, F f) { return f(p); }

#include PROBE_FILE

static uint64_t rng_state = 0x243F6A8885A308D3ull;
static uint32_t rnd() {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 16);
}

int main() {
  const int L = kProbeL, W = L / 4;
  long long bad = 0, checked = 0;
  for (int m = 0; m < kProbeModules; m++) {
    for (int it = 0; it < 20000; it++) {
      uint32_t x[32] = {0}, r[32] = {0};
      uint8_t line[128];
      const int mode = it % 5;
      for (int i = 0; i < L; i++) {
        uint32_t v = rnd();
        if (mode == 1) v = (v & 1) ? 0xff : 0x00;            // borrows everywhere
        else if (mode == 2) v = (v & 1) ? 0x80 : 0x7f;       // the sign bit of every byte
        else if (mode == 3) v = 1u << (v & 7);               // single bits: shifted-out bits
        else if (mode == 4) v = 0xffu ^ (1u << (v & 7));
        line[i] = (uint8_t)v;
      }
      memcpy(x, line, (size_t)L);
      kProbeFn[m](x, r);
      uint8_t got[128];
      memcpy(got, r, (size_t)L);
      for (int j = 0; j < L; j++) {
        uint8_t want;
        if (j == 0) want = line[kProbeRoot[m]];
        else {
          const uint8_t b = line[kProbeP[m][j]];
          const int v = kProbeV[m][j];
          uint8_t pred = b;
          if (kProbeOp[m] == 1) pred = (uint8_t)(b + (uint8_t)v);
          else if (kProbeOp[m] == 2) pred = (v <= -8 || v >= 8) ? 0 : (v < 0 ? (uint8_t)(b >> -v) : (uint8_t)(b << v));
          want = (uint8_t)(line[kProbeX[m][j]] - pred);
        }
        checked++;
        if (got[j] != want) {
          if (bad < 10) fprintf(stderr, "module #%d byte %d: got %02x want %02x\n", m, j, got[j], want);
          bad++;
        }
      }
    }
  }
  // canonical layout of a plane-major scan: c[16 h + k], byte lane q = residue byte of the k-th column (scan order) of chunk 4 h + q
  for (int it = 0; it < 2000 && kProbeCanonWords > 0; it++) {
    uint32_t r[32] = {0}, c[32] = {0};
    uint8_t rb[128];
    for (int i = 0; i < L; i++) rb[i] = (uint8_t)rnd();
    memcpy(r, rb, (size_t)L);
    probe_canon(r, c);
    for (int i = 0; i < kProbeCanonWords; i++)
      for (int q = 0; q < 4; q++) {
        const int chunk = 4 * (i / 16) + q;
        if (chunk >= L / 16) continue;  // lanes past the line: whatever the gather leaves there is masked by the encoder
        const uint8_t want = rb[kProbeCols[16 * chunk + i % 16]], got = (uint8_t)(c[i] >> (8 * q));
        checked++;
        if (got != want) {
          if (bad < 10) fprintf(stderr, "canonical word %d lane %d: got %02x want %02x\n", i, q, got, want);
          bad++;
        }
      }
  }
  (void)W;
  printf("%lld %lld\n", checked, bad);
  return bad ? 1 : 0;
}
