// trace::LoaderNPY -- .npy memory-dump reader (mirror of reference src/loader/LoaderNPY.{h,cpp}).
// The file (uint8, C order, shape [N, L], NPY format 1.0/2.0/3.0) is memory-mapped instead of being read
// into a vector, so dumps larger than host RAM stream through the chunked H2D path.
#ifndef MPCB_LOADERNPY_H_
#define MPCB_LOADERNPY_H_

#include "Loader.h"

namespace trace {

class LoaderNPY : public Loader {
 public:
  explicit LoaderNPY(const std::string& filePath);
  ~LoaderNPY() override;
  MemReq_t* GetCacheline(MemReq_t* memReq) override;  // LoaderNPY.cpp:14-34
  unsigned GetCachelineSize() override;               // LoaderNPY.cpp:36-40
  unsigned long long GetNumLines() override;          // LoaderNPY.cpp:42-46 (row count of the file)
  void Reset() override;                              // LoaderNPY.cpp:48-54
  uint64_t GetChunk(uint8_t* dst, uint64_t maxLines) override;
  const uint8_t* GetAll(uint64_t* nLines) override;
  int GetFile(uint64_t* dataOffset, uint64_t* nLines, bool directIo = false) override;
  const std::string& Error() const { return m_Error; }

 private:
  void unmap();
  const uint8_t* m_Map = nullptr;
  size_t m_MapBytes = 0;
  const uint8_t* m_Data = nullptr;
  int m_Fd = -1;  // descriptor handed to CompressFile (opened on demand)
  uint64_t m_Rows = 0, m_LineSize = 0, m_CurrentLine = 0;
  std::string m_Error;
};

}  // namespace trace
#endif
