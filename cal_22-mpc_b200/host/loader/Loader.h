// trace::Loader -- host-side mirror of the reference loader interface (reference src/loader/Loader.h:62-88).
// Same pure virtuals and MemReq_t fields, so code written against the reference keeps compiling; the
// batched entry point GetChunk() is what the GPU path uses instead of one GetCacheline() call per block.
#ifndef MPCB_LOADER_H_
#define MPCB_LOADER_H_

#include <algorithm>
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
  virtual ~Loader() { delete m_ChunkReq; }
  // the request object GetCacheline fills; the reference's driver picks the type per loader (main.cpp:210-215)
  virtual MemReq_t* NewMemReq() { return new MemReq_t; }
  // reference interface (Loader.h:77-82)
  virtual MemReq_t* GetCacheline(MemReq_t*) = 0;
  virtual unsigned GetCachelineSize() = 0;
  virtual unsigned long long GetNumLines() = 0;
  virtual void Reset() = 0;
  // batched form: copies up to maxLines of the lines compressLines would still see (main.cpp:237-243)
  // into dst and returns how many were written; 0 = end of input.
  // The default body is the reference's own loop over GetCacheline (main.cpp:229-243: a request flagged isEnd is
  // not compressed), so a loader written against the reference interface compiles and runs unchanged.
  virtual uint64_t GetChunk(uint8_t* dst, uint64_t maxLines) {
    const size_t L = GetCachelineSize();
    if (!m_ChunkReq) m_ChunkReq = NewMemReq();
    uint64_t n = 0;
    while (n < maxLines && !m_ChunkEnd) {
      m_ChunkReq->Reset();
      MemReq_t* r = GetCacheline(m_ChunkReq);
      if (!r || r->isEnd) { m_ChunkEnd = true; break; }
      if (r->data.size() != L) continue;  // not a line of the dump (the reference's driver filters these, main.cpp:222-224)
      std::copy(r->data.begin(), r->data.end(), dst + n * L);
      n++;
    }
    return n;
  }
  // zero-copy form for loaders that hold the lines contiguously: pointer to the remaining lines, count
  // in *nLines, and the cursor moves to the end.  nullptr if unsupported.
  virtual const uint8_t* GetAll(uint64_t* nLines) { *nLines = 0; return nullptr; }
  // file-backed form for loaders whose remaining lines are one contiguous byte range of a file: an open descriptor (owned
  // by the loader), the range's offset and the line count; the cursor moves to the end.  -1 if unsupported.  The compressor
  // reads the range itself, straight into its pinned staging buffers (comp::Compressor::CompressFile).
  virtual int GetFile(uint64_t* dataOffset, uint64_t* nLines, bool directIo = false) { (void)directIo; *dataOffset = 0; *nLines = 0; return -1; }

 protected:
  const std::string m_FilePath;
  bool m_ChunkEnd = false;  // default GetChunk: the reference loop has seen isEnd
  MemReq_t* m_ChunkReq = nullptr;
};

}  // namespace trace
#endif
