// A compressor and a loader written against the REFERENCE interfaces only (src/compressor/Compressor.h:18-33: CompressLine;
// src/loader/Loader.h:77-82: GetCacheline / GetCachelineSize / GetNumLines / Reset).  They must compile unchanged against the
// mirror headers, and the batched entry points the GPU path added (CompressBatch, GetChunk) must fall back to the reference's
// own per-line loops (main.cpp:229-244).  Prints what the per-line driver of the reference would have seen.
#include <cstdio>

#include "compressor/Compressor.h"
#include "loader/Loader.h"

namespace {

class SumCompressor : public comp::Compressor {  // "compressed size" = sum of the line's bytes
 public:
  explicit SumCompressor(unsigned lineSize) { m_Stat = new comp::CompResult(lineSize); }
  unsigned CompressLine(std::vector<uint8_t>& dataLine) override {
    unsigned s = 0;
    for (uint8_t b : dataLine) s += b;
    m_Stat->OriginalSize += dataLine.size() * 8;
    m_Stat->CompressedSize += s;
    return s;
  }
};

class CountingLoader : public trace::Loader {  // line i is filled with the byte i; isEnd on the call that returns the last line
 public:
  CountingLoader(unsigned rows, unsigned lineSize) : trace::Loader("none"), m_Rows(rows), m_L(lineSize) {}
  trace::MemReq_t* GetCacheline(trace::MemReq_t* r) override {
    r->reqSize = m_L;
    r->data.assign(m_L, (uint8_t)m_Cur);
    m_Cur++;
    r->isEnd = m_Cur >= m_Rows;
    return r;
  }
  unsigned GetCachelineSize() override { return m_L; }
  unsigned long long GetNumLines() override { return m_Rows; }
  void Reset() override { m_Cur = 0; }

 private:
  unsigned m_Rows, m_L, m_Cur = 0;
};

}  // namespace

int main() {
  const unsigned L = 32, rows = 10;
  CountingLoader loader(rows, L);
  SumCompressor comp(L);
  std::vector<uint8_t> buf(4 * L);
  unsigned long long lines = 0;
  uint64_t n;
  while ((n = loader.GetChunk(buf.data(), 4)) != 0) {  // chunks of 4, 4, 1: the 10th row is dropped like LoaderNPY's last row
    comp.CompressBatch(buf.data(), n);
    lines += n;
  }
  printf("%llu %llu %llu\n", lines, (unsigned long long)comp.GetResult()->OriginalSize,
         (unsigned long long)comp.GetResult()->CompressedSize);
  return 0;
}
