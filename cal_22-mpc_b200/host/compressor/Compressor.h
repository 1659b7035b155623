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
  // batched form: nLines consecutive lines of GetCachelineSize() bytes; statistics updated, no per-line value
  virtual void CompressBatch(const uint8_t* lines, uint64_t nLines) = 0;
  virtual CompResult* GetResult() { return m_Stat; }

 protected:
  CompResult* m_Stat = nullptr;
};

}  // namespace comp
#endif
