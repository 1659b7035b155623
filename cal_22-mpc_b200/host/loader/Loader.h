// trace::Loader -- host-side mirror of the reference loader interface (reference src/loader/Loader.h:62-88).
// Same pure virtuals and MemReq_t fields, so code written against the reference keeps compiling; the
// batched entry point GetChunk() is what the GPU path uses instead of one GetCacheline() call per block.
#ifndef MPCB_LOADER_H_
#define MPCB_LOADER_H_

#include <cstdint>
#include <string>
#include <vector>

#define WORD_SIZE uint8_t
#define ACCESS_GRAN 32

typedef uint64_t addr_t;

namespace trace {

enum rw_t { READ, WRITE, NA };

struct MemReq_t {  // Loader.h:24-60
  addr_t addr = 0;
  rw_t rw = NA;
  uint32_t reqSize = 0;
  std::vector<WORD_SIZE> data;
  bool isEnd = false;
  virtual ~MemReq_t() {}
  virtual void Reset() {
    addr = 0;
    rw = NA;
    reqSize = 0;
    data.clear();
    isEnd = false;
  }
};

class Loader {
 public:
  explicit Loader(const std::string& filePath) : m_FilePath(filePath) {}
  virtual ~Loader() {}
  // reference interface (Loader.h:77-82)
  virtual MemReq_t* GetCacheline(MemReq_t*) = 0;
  virtual unsigned GetCachelineSize() = 0;
  virtual unsigned long long GetNumLines() = 0;
  virtual void Reset() = 0;
  // batched form: copies up to maxLines of the lines compressLines would still see (main.cpp:237-243)
  // into dst and returns how many were written; 0 = end of input.
  virtual uint64_t GetChunk(uint8_t* dst, uint64_t maxLines) = 0;
  // zero-copy form for loaders that hold the lines contiguously: pointer to the remaining lines, count
  // in *nLines, and the cursor moves to the end.  nullptr if unsupported.
  virtual const uint8_t* GetAll(uint64_t* nLines) { *nLines = 0; return nullptr; }

 protected:
  const std::string m_FilePath;
};

}  // namespace trace
#endif
