// Per-block size models of the stateless secondary compressors (BDI, FPC, BPC) for 128-byte blocks held as 32
// little-endian words.  __host__ __device__ so that the same code is checked on the CPU (tests/test_swar_host.py)
// and runs one block per thread on the GPU (mpc_variants.cu).  Reference: src/compressor/{BDI,FPC,BPC}.cpp;
// observable quirks are kept.
#pragma once
#include <stdint.h>

#include "mpc_device.cuh"

namespace mpcvar {

MPC_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
MPC_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
MPC_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}

// ---- BDI ------------------------------------------------------------------------------------------------------
// BDI::reduceSign, BDI.cpp:203-218: a negative value keeps (index of its highest zero bit + 2) low bits; all ones
// comes back unchanged (and therefore never fits a delta).  Kept as the written-down rule; the kernels use the
// closed form below.
MPC_HD uint64_t bdi_reduce_sign(uint64_t x) {
  if ((x >> 63) == 0 || x == ~0ull) return x;
  const int hz = 63 - clz64(~x);  // highest zero bit, 0..62
  const int keep = hz + 2;        // 2..64
  return keep >= 64 ? x : (x & ((1ull << keep) - 1ull));
}

// reduceSign(x) <= 2^(8D) - 1 in closed form.  For x = -k < 0 the kept value is 2^keep - k with
// 2^(keep-1) <= 2^keep - k < 2^keep and keep = (bit length of k-1) + 1, so it fits exactly when keep <= 8D, i.e.
// 2 <= k <= 2^(8D-1) (k = 1, all ones, never fits); x >= 0 fits when x < 2^(8D).  As one range test on x + 2^(8D-1):
// x in [-2^(8D-1), 2^(8D)) and x != -1.  tests/test_swar_host.py checks it against bdi_reduce_sign on the boundaries.
template <int D>
MPC_HD bool bdi_fits64(uint64_t x) {
  constexpr uint64_t half = 1ull << (8 * D - 1);
  return (x + half) < 3ull * half && x != ~0ull;
}
// the same test for base - v with both < 2^32 (base sizes 4 and 2), in 32-bit arithmetic: the 64-bit difference is
// d (v <= base) or d - 2^32 (v > base) with d = (base - v) mod 2^32
template <int D>
MPC_HD bool bdi_delta_fits32(uint32_t base, uint32_t v) {
  constexpr uint32_t half = 1u << (8 * D - 1);
  const uint32_t d = base - v;
  return v <= base ? d <= 2u * half - 1u : (d + half) <= half - 2u;
}

// W = words per line (32 / 16 / 8 for lines of 128 / 64 / 32 bytes); a line sits in x[0..W).
template <int B>
MPC_HD uint64_t bdi_value(const uint32_t (&x)[32], int i) {  // little-endian chunk, zero-extended (BDI.cpp:127-153)
  if (B == 8) return (uint64_t)x[2 * i] | ((uint64_t)x[2 * i + 1] << 32);
  if (B == 4) return x[i];
  return (x[i >> 1] >> (16 * (i & 1))) & 0xffffu;
}

// size of a check that fits, whatever the number of immediates: n + 8 (B + (n - 1) D), n = L / B values
template <int B, int D, int W>
MPC_HD constexpr uint32_t bdi_fit_size() { return (uint32_t)(4 * W / B) + 8u * (uint32_t)(B - D + (4 * W / B) * D); }

// BDI::checkBDI, BDI.cpp:108-201: immediates (values that fit D bytes on their own), the first other value is the
// base, every later one must be within a D-byte delta of it.
template <int B, int D, int W = 32>
MPC_HD uint32_t bdi_check(const uint32_t (&x)[32], uint32_t* imm_out = nullptr) {
  constexpr int n = 4 * W / B;
  uint32_t imm = 0;
  bool not_all = false;
  if (B == 8) {
    // base = first non-immediate value: select chain from the back
    uint64_t base = 0;
    uint32_t imm_mask = 0;
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
      const uint64_t v = bdi_value<8>(x, i);
      const bool im = bdi_fits64<D>(v);
      imm_mask |= (im ? 1u : 0u) << i;
      base = im ? base : v;
    }
    imm = (uint32_t)popc32(imm_mask);
    // deltas against the base, four values at a time: the first value out of range decides the check, so the rest is
    // skipped (incompressible data -- most of a dump -- leaves after the first group)
#pragma unroll
    for (int g = 0; g < n; g += 4) {
      if (!not_all) {
#pragma unroll
        for (int i = g; i < g + 4; i++) {
          const uint64_t v = bdi_value<8>(x, i);
          // immediates are skipped; the base itself passes (difference 0)
          not_all |= !((imm_mask >> i) & 1u) && !bdi_fits64<D>(base - v);
        }
      }
    }
  } else {
    constexpr uint32_t limit = D == 1 ? 0xffu : 0xffffu;  // zero-extended values are never negative
    uint32_t base = 0;
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
      const uint32_t v = (uint32_t)bdi_value<B>(x, i);
      const bool im = v <= limit;
      imm += im ? 1u : 0u;
      base = im ? base : v;
    }
#pragma unroll
    for (int g = 0; g < n; g += 8) {
      if (!not_all) {
#pragma unroll
        for (int i = g; i < g + 8; i++) {
          const uint32_t v = (uint32_t)bdi_value<B>(x, i);
          not_all |= v > limit && !bdi_delta_fits32<D>(base, v);
        }
      }
    }
  }
  if (imm_out) *imm_out = imm;
  if (not_all) return (uint32_t)n + 8u * (imm * (uint32_t)D + ((uint32_t)n - imm) * (uint32_t)B);
  return (uint32_t)n + 8u * (imm * (uint32_t)D + ((uint32_t)B + ((uint32_t)n - imm - 1u) * (uint32_t)D));  // wraps when imm == n
}

// bdi_check with the delta size as a RUN-TIME value (base sizes 8 and 4): one copy of the code per base size instead of one per
// (base, delta) pair.  The general form of the six checks shrinks from six unrolled bodies to three, and the BDI / PATTERN kernels
// -- 7 000 / 9 500 SASS instructions, a quarter to a half of their stall samples waiting for instruction fetches -- run the delta
// sizes of a base size as iterations of one loop.  Same rule, same sizes (tests/test_swar_host.py, tests/test_variants.py).
template <int B, int W>
MPC_HD uint32_t bdi_check_rt(const uint32_t (&x)[32], uint32_t D, uint32_t* imm_out) {
  constexpr int n = 4 * W / B;
  uint32_t imm = 0;
  bool not_all = false;
  if (B == 8) {
    const uint64_t half = 1ull << (8u * D - 1u);
    const uint64_t span = 3ull * half;  // x in [-half, 2 half) <=> x + half < 3 half (bdi_fits64)
    uint64_t base = 0;
    uint32_t imm_mask = 0;
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
      const uint64_t v = bdi_value<8>(x, i);
      const bool im = (v + half) < span && v != ~0ull;
      imm_mask |= (im ? 1u : 0u) << i;
      base = im ? base : v;
    }
    imm = (uint32_t)popc32(imm_mask);
#pragma unroll
    for (int g = 0; g < n; g += 4) {
      if (!not_all) {
#pragma unroll
        for (int i = g; i < g + 4; i++) {
          const uint64_t d = base - bdi_value<8>(x, i);
          not_all |= !((imm_mask >> i) & 1u) && !((d + half) < span && d != ~0ull);
        }
      }
    }
  } else {
    const uint32_t limit = D == 1u ? 0xffu : 0xffffu;  // zero-extended values are never negative
    const uint32_t half = 1u << (8u * D - 1u);
    uint32_t base = 0;
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
      const uint32_t v = (uint32_t)bdi_value<B>(x, i);
      const bool im = v <= limit;
      imm += im ? 1u : 0u;
      base = im ? base : v;
    }
#pragma unroll
    for (int g = 0; g < n; g += 8) {
      if (!not_all) {
#pragma unroll
        for (int i = g; i < g + 8; i++) {
          const uint32_t v = (uint32_t)bdi_value<B>(x, i);
          const uint32_t d = base - v;
          const bool fits = v <= base ? d <= 2u * half - 1u : (d + half) <= half - 2u;  // bdi_delta_fits32
          not_all |= v > limit && !fits;
        }
      }
    }
  }
  if (imm_out) *imm_out = imm;
  if (not_all) return (uint32_t)n + 8u * (imm * D + ((uint32_t)n - imm) * (uint32_t)B);
  return (uint32_t)n + 8u * (imm * D + ((uint32_t)B + ((uint32_t)n - imm - 1u) * D));  // wraps when imm == n
}

// ---- the common case in one pass per base size ------------------------------------------------------------------------
// Most of a dump (floats, pointers, noise) holds NO immediate for any delta size of a base size: then the base is value 0,
// a check either fits -- every value within a D-byte delta of value 0 -- or costs more than the raw line (n + 8 L bits), and
// the three (two, one) delta sizes of the base size are decided together by the widest delta.  bdi_no_imm<B> proves "no
// immediate" with one test per value that is sufficient for every delta size; bdi_need<B> returns the smallest delta size
// (1, 2, 4) that fits all deltas against value 0, or 9.  Lines with immediates take bdi_check.
// FROM..TO = range of words looked at (the callers vote on the first few words before paying for the whole line)
template <int B, int FROM, int TO>
MPC_HD bool bdi_no_imm(const uint32_t (&x)[32]) {
  bool none = true;
  if (B == 8) {
    // an 8-byte value is an immediate for some D <= 4 only if its upper word is 0 or all ones (bdi_fits64<4>)
#pragma unroll
    for (int i = FROM / 2; i < TO / 2; i++) none = none && (x[2 * i + 1] + 1u) > 1u;
  } else if (B == 4) {
#pragma unroll
    for (int i = FROM; i < TO; i++) none = none && x[i] > 0xffffu;   // zero-extended: immediate iff v <= 0xff / 0xffff
  } else {
#pragma unroll
    for (int i = FROM; i < TO; i++) none = none && mpcdev::min_u16x2(x[i] & 0xff00ff00u, 0x00010001u) == 0x00010001u;  // both halfwords > 0xff
  }
  return none;
}

// Which form a base size takes is decided per WARP on the GPU: lanes that took different forms would run both, one after
// the other, and a finely mixed dump would pay for the sum.  `vote(b)` is "b holds for every lane that runs the checks"
// (__all_sync over those lanes in the kernels, the identity on the host).  The first vote looks at four words only, so that
// a mixed warp is sent to the general form for a handful of instructions.
struct BdiSelfVote {
  MPC_HDM bool operator()(bool b) const { return b; }
};
template <int B, int W, class Vote>
MPC_HD bool bdi_fast_form(const uint32_t (&x)[32], const Vote& vote) {
  if (!vote(bdi_no_imm<B, 0, 4>(x))) return false;
  return vote(bdi_no_imm<B, 4, W>(x));
}

template <int B, int W>
MPC_HD int bdi_need(const uint32_t (&x)[32]) {
  constexpr int n = 4 * W / B;
  if (B == 8) {
    const uint64_t base = bdi_value<8>(x, 0);
    bool ok4 = true;
#pragma unroll
    for (int g = 0; g < n; g += 4) {
      if (ok4) {
#pragma unroll
        for (int i = g; i < g + 4; i++) ok4 = ok4 && bdi_fits64<4>(base - bdi_value<8>(x, i));
      }
    }
    if (!ok4) return 9;
    bool ok2 = true, ok1 = true;
#pragma unroll
    for (int i = 1; i < n; i++) {
      const uint64_t d = base - bdi_value<8>(x, i);
      ok2 = ok2 && bdi_fits64<2>(d);
      ok1 = ok1 && bdi_fits64<1>(d);
    }
    return ok1 ? 1 : (ok2 ? 2 : 4);
  } else {
    const uint32_t base = (uint32_t)bdi_value<B>(x, 0);
    constexpr int DM = B == 4 ? 2 : 1;  // widest delta of this base size
    bool okw = true;
#pragma unroll
    for (int g = 0; g < n; g += 8) {
      if (okw) {
#pragma unroll
        for (int i = g; i < g + 8; i++) okw = okw && bdi_delta_fits32<DM>(base, (uint32_t)bdi_value<B>(x, i));
      }
    }
    if (!okw) return 9;
    if (B == 2) return 1;
    bool ok1 = true;
#pragma unroll
    for (int i = 1; i < n; i++) ok1 = ok1 && bdi_delta_fits32<1>(base, (uint32_t)bdi_value<B>(x, i));
    return ok1 ? 1 : 2;
  }
}

// The six checks in the reference's order (BDI.cpp:30-66, Pattern.cpp:20-60); a later check only replaces an earlier one when
// strictly smaller.  A check either fits -- then its size is the constant bdi_fit_size whatever the number of immediates --
// or costs more than that, so a check whose fitting size cannot beat the best so far is skipped.  Returns the best size
// (raw = 8 L when nothing beats it); *sel = index 0..5 of the winning check (unchanged when none), *imm = its immediates.
template <int W, class Vote>
MPC_HD uint32_t bdi_best_check(const uint32_t (&x)[32], int* sel, uint32_t* imm, const Vote& vote) {
  uint32_t best = 32u * W, cur, im = 0;
  // all votes up front: every lane that runs the checks is here (further down lanes skip checks that cannot win)
  const bool fast8 = bdi_fast_form<8, W>(x, vote), fast4 = bdi_fast_form<4, W>(x, vote), fast2 = bdi_fast_form<2, W>(x, vote);
  if (fast8) {
    const int need = bdi_need<8, W>(x);
    if (need == 1) { best = bdi_fit_size<8, 1, W>(); *sel = 0; *imm = 0; }
    else if (need == 2 && best > bdi_fit_size<8, 2, W>()) { best = bdi_fit_size<8, 2, W>(); *sel = 1; *imm = 0; }
    else if (need == 4 && best > bdi_fit_size<8, 4, W>()) { best = bdi_fit_size<8, 4, W>(); *sel = 2; *imm = 0; }
  } else {
#pragma unroll 1
    for (int d = 0; d < 3; d++) {  // delta sizes 1, 2, 4: one body (bdi_check_rt)
      const uint32_t D = 1u << d, n8 = (uint32_t)(W / 2);
      if (d == 0 || best > n8 + 8u * (8u - D + n8 * D)) {  // bdi_fit_size<8, D, W>
        cur = bdi_check_rt<8, W>(x, D, &im);
        if (best > cur) { best = cur; *sel = d; *imm = im; }
      }
    }
  }
  if (best > bdi_fit_size<4, 1, W>()) {
    if (fast4) {
      const int need = bdi_need<4, W>(x);
      if (need == 1) { best = bdi_fit_size<4, 1, W>(); *sel = 3; *imm = 0; }
      else if (need == 2 && best > bdi_fit_size<4, 2, W>()) { best = bdi_fit_size<4, 2, W>(); *sel = 4; *imm = 0; }
    } else {
#pragma unroll 1
      for (int d = 0; d < 2; d++) {  // delta sizes 1, 2
        const uint32_t D = 1u << d, n4 = (uint32_t)W;
        if (d == 0 || best > n4 + 8u * (4u - D + n4 * D)) {  // bdi_fit_size<4, D, W>
          cur = bdi_check_rt<4, W>(x, D, &im);
          if (best > cur) { best = cur; *sel = 3 + d; *imm = im; }
        }
      }
    }
  }
  if (best > bdi_fit_size<2, 1, W>()) {
    if (fast2) {
      if (bdi_need<2, W>(x) == 1) { best = bdi_fit_size<2, 1, W>(); *sel = 5; *imm = 0; }
    } else {
      cur = bdi_check<2, 1, W>(x, &im); if (best > cur) { best = cur; *sel = 5; *imm = im; }
    }
  }
  return best;
}

// BDI::CompressLine, BDI.cpp:6-74.  Returns bits incl. the 4 encoding bits; *state = BDIState (BDI.h:10-21).
// make_vote(runs_checks) is called by every lane of the warp and returns the Vote over the lanes that run the checks
template <int W, class MakeVote>
MPC_HD uint32_t bdi_block_with(const uint32_t (&x)[32], int* state, const MakeVote& make_vote) {
  uint32_t any = 0, rep = 0;
#pragma unroll
  for (int i = 0; i < W; i++) { any |= x[i]; rep |= x[i] ^ x[i & 1]; }
  uint32_t best = 32u * W;
  int sel = 8;
  const auto vote = make_vote(any != 0 && rep != 0);
  if (any == 0) { best = 8; sel = 0; }
  else if (rep == 0) { best = 64; sel = 1; }
  else {
    int k = -1;
    uint32_t imm = 0;
    best = bdi_best_check<W>(x, &k, &imm, vote);
    sel = (k < 0 || best == 32u * W) ? 8 : 2 + k;
  }
  *state = sel;
  return best + 4u;
}
struct BdiMakeSelfVote {
  MPC_HDM BdiSelfVote operator()(bool) const { return BdiSelfVote(); }
};
template <int W = 32>
MPC_HD uint32_t bdi_block(const uint32_t (&x)[32], int* state) { return bdi_block_with<W>(x, state, BdiMakeSelfVote()); }

// ---- PATTERN (analysis tool) --------------------------------------------------------------------------------------
// Pattern::CompressLine, Pattern.cpp:6-75: the six base-delta checks in order (checkPattern, Pattern.cpp:109-199, is
// checkBDI word for word), strict improvement wins, no zero / repeat shortcut.  *sel = PatternState 0..5, or 9
// (NotDefined) when nothing beats the raw size; *imm = immediates of the selected check (countPattern,
// Pattern.cpp:201-320: every immediate adds baseSize implicit bytes, every other value baseSize explicit bytes).
// Returns bits incl. the 4 encoding bits.
template <int W = 32, class Vote = BdiSelfVote>
MPC_HD uint32_t pattern_block(const uint32_t (&x)[32], int* sel_out, uint32_t* imm_out, const Vote& vote = Vote()) {
  int k = -1;
  uint32_t imm = 0;
  const uint32_t best = bdi_best_check<W>(x, &k, &imm, vote);
  const bool none = k < 0 || best == 32u * W;
  *sel_out = none ? 9 : k;
  *imm_out = none ? 0u : imm;
  return best + 4u;
}

// 64-bit content hash of a block (temporal-locality pass: equal blocks are found by sorting the hashes, and every
// match is confirmed on the 128 bytes themselves, so the hash only has to spread well)
template <int W = 32>
MPC_HD uint64_t block_hash64(const uint32_t (&x)[32]) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int i = 0; i < W / 2; i++) {
    const uint64_t v = (uint64_t)x[2 * i] | ((uint64_t)x[2 * i + 1] << 32);
    h = (h ^ v) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
  }
  return mpcdev::splitmix64(h);
}

// ---- FPC ------------------------------------------------------------------------------------------------------
// FPC::CompressLine, FPC.cpp:7-87.  counts8 packs the eight per-word prefix counters, 8 bits each (<= 32 per block).
// The reference's unbounded zero-run scan (FPC.cpp:26) is bounded at the block end here.
// Branch-free: every lane of a warp holds a different pattern on a mixed dump, and an if / else chain per word made the warp walk
// every arm (the compiler kept the chain as branches: 86 BRA and 56 BSSY / BSYNC pairs in the 128-byte-line kernel).  Here the
// tests are range tests -- a word is a sign-extended k-bit value iff (v + 2^(k-1)) < 2^k, both halfwords are sign-extended bytes iff
// ((h + 0x80) mod 2^16) < 0x100 in each halfword --, the pattern is the first test that holds (a select chain from the back) and its
// cost comes out of a byte table with one PRMT.  What the chain was good at -- words that are decided by its first arms on EVERY
// lane (zero pages, small integers) -- is kept by votes: a warp whose lines are all zero is done at once, and the three tests past "sign-extended
// halfword" are skipped for a word that no lane needs them for (one vote per word; one per four words or one per line measured slower).  `vote(b)` = b holds on every lane of the warp
// (the identity on the host).
template <int W = 32, class Vote = BdiSelfVote>
MPC_HD uint32_t fpc_block(const uint32_t (&x)[32], uint64_t* counts8, const Vote& vote = Vote()) {
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < W; i++) any |= x[i];
  if (vote(any == 0u)) {  // one zero run: 6 bits, W words of pattern 0
    *counts8 = (uint64_t)W;
    return 6u;
  }
  uint32_t size = 0;
  uint64_t cnt = 0;
  bool prev_zero = false;
#pragma unroll
  for (int i = 0; i < W; i++) {
    const uint32_t v = x[i];
    const bool z = v == 0;
    const bool s16 = (v + 0x8000u) < 0x10000u;  // sign-extended 16 bits (covers 8, 4 and zero)
    uint32_t p = 7u;
    if (!vote(s16)) {  // some lane's word is none of the first four patterns
      p = (v == (v & 0xffu) * 0x01010101u) ? 6u : p;                          // one byte, four times
      p = ((mpcdev::add_u16x2(v, 0x00800080u) & 0xff00ff00u) == 0u) ? 5u : p;  // two sign-extended bytes in the halfwords
      p = ((v & 0x0000ffffu) == 0u) ? 4u : p;                                  // low halfword zero
    }
    p = s16 ? 3u : p;
    p = ((v + 0x80u) < 0x100u) ? 2u : p;
    p = ((v + 0x8u) < 0x10u) ? 1u : p;
    p = z ? 0u : p;
    // cost of pattern p (FPC.cpp:16-83): 6 once per zero run, 7, 11, 19, 19, 19, 11, 35
#if defined(__CUDA_ARCH__)
    uint32_t c = __byte_perm(0x130b0706u, 0x230b1313u, p) & 0xffu;
#else
    const uint32_t kCost[8] = {6, 7, 11, 19, 19, 19, 11, 35};
    uint32_t c = kCost[p];
#endif
    c = (z && prev_zero) ? 0u : c;
    prev_zero = z;
    size += c;
    cnt += 1ull << (8 * p);
  }
  *counts8 = cnt;
  return size;
}

// ---- BPC ------------------------------------------------------------------------------------------------------
// 32x32 bit-matrix transpose in registers: afterwards bit r of a[c] is the former bit c of a[r].
MPC_HD void transpose32(uint32_t (&a)[32]) {
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
      const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
      a[k] ^= t << j;
      a[k + j] ^= t;
    }
  }
}

// BPC::CompressLine, BPC.cpp:20-87 + encodeFirst (always 7, BPC.cpp:89-101) + encodeDeltas (BPC.cpp:103-185).
// pat8 packs the 7 pattern counters (BPC.h:13-22), 8 bits each; *words = value added to TotalWords.
template <int W = 32>
MPC_HD uint32_t bpc_block(const uint32_t (&x)[32], uint64_t* pat8, uint32_t* words) {
  uint32_t d[32];
  uint32_t neg = 0;  // bit r = delta r negative = bit 32 of the 33-bit delta (words are zero-extended, BPC.cpp:41-45)
#pragma unroll
  for (int r = 0; r < W - 1; r++) {
    d[r] = x[r + 1] - x[r];
    neg |= (x[r + 1] < x[r] ? 1u : 0u) << r;
  }
#pragma unroll
  for (int r = W - 1; r < 32; r++) d[r] = 0;  // a line of W words has W - 1 deltas: the planes are W - 1 bits wide
  transpose32(d);  // d[c] = delta bit plane c (bit r = delta r)
  uint32_t length = 7, run = 0, nwords = 0;
  uint64_t pat = 0;
  uint32_t prev = neg;  // DBP[32]
#pragma unroll
  for (int i = 32; i >= 0; i--) {
    const uint32_t dbp = (i == 32) ? neg : d[i];
    const uint32_t dbx = (i == 32) ? neg : (dbp ^ prev);
    prev = dbp;
    if (dbx == 0) { run++; continue; }
    if (run) { length += (run == 1) ? 3u : 7u; pat += 1ull << 8; nwords += run; run = 0; }
    int p;
    if (dbp == 0) { length += 5; p = 2; }
    else if (dbx == 0x7fffffffu) { length += 5; p = 6; }
    else {
      const int ones = popc32(dbx);
      if (ones == 1) { length += 10; p = 3; }
      else if (ones == 2 && (dbx & (dbx >> 1))) { length += 10; p = 4; }
      else { length += 32; p = 0; }
    }
    pat += 1ull << (8 * p);
    nwords += 1;
  }
  if (run) { length += (run == 1) ? 3u : 7u; pat += 1ull << 8; nwords += run; }
  *pat8 = pat;
  *words = nwords;
  return length;
}

}  // namespace mpcvar
