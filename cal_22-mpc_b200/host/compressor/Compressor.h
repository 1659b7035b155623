// comp::Compressor -- the plugin interface of the reference (src/compressor/Compressor.h:18-33) plus the
// batched entry point the GPU path is driven through.
#ifndef MPCB_COMPRESSOR_H_
#define MPCB_COMPRESSOR_H_

#include <cstdint>
#include <string>
#include <vector>

#include "CompResult.h"

namespace comp {

class Compressor {
 public:
  virtual ~Compressor() {}
  std::string GetCompressorName() { return m_Stat->CompressorName; }
  // reference contract: compressed size of the line in bits incl. encoding bits; statistics updated
  virtual unsigned CompressLine(std::vector<uint8_t>& dataLine) = 0;
  // batched form: nLines consecutive lines of m_Stat->LineSize bytes; statistics updated, no per-line value.
  // The default body is the reference's own per-line loop (main.cpp:237-243), so a compressor written against the
  // reference interface (CompressLine only) compiles and runs unchanged; the GPU compressors override it.
  virtual void CompressBatch(const uint8_t* lines, uint64_t nLines) {
    const size_t L = m_Stat ? m_Stat->LineSize : 0;
    std::vector<uint8_t> line(L);
    for (uint64_t i = 0; i < nLines && L; i++) {
      line.assign(lines + i * L, lines + (i + 1) * L);
      CompressLine(line);
    }
  }
  // file-backed form: nLines consecutive lines starting at byte `offset` of the open descriptor fd.  false = not supported by
  // this compressor (the driver then hands the lines over with CompressBatch).
  virtual bool CompressFile(int fd, uint64_t offset, uint64_t nLines, bool directIo) {
    (void)fd; (void)offset; (void)nLines; (void)directIo;
    return false;
  }
  virtual CompResult* GetResult() { return m_Stat; }

 protected:
  CompResult* m_Stat = nullptr;
};

}  // namespace comp
#endif
