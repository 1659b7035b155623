#include "LoaderGPGPU.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace trace {
namespace gpgpusim {

namespace {
template <class T>
T rd(const uint8_t* p) {
  T v;
  memcpy(&v, p, sizeof(T));
  return v;
}
}  // namespace

LoaderGPGPU::LoaderGPGPU(const std::string& filePath) : Loader(filePath) { Reset(); }
LoaderGPGPU::~LoaderGPGPU() { unmap(); }

void LoaderGPGPU::unmap() {
  if (m_Map) munmap(const_cast<uint8_t*>(m_Map), m_MapBytes);
  m_Map = nullptr;
  m_MapBytes = 0;
}

// LoaderGPGPU.cpp:73-113: rewind and validate the header (messages and exit code are the reference's)
void LoaderGPGPU::Reset() {
  if (!m_Map) {
    int fd = open(m_FilePath.c_str(), O_RDONLY);
    if (fd < 0) {
      printf("Failed to open a file. Check the path of the file.\n");  // LoaderGPGPU.cpp:84-88
      exit(1);
    }
    struct stat st;
    fstat(fd, &st);
    m_MapBytes = (size_t)st.st_size;
    if (m_MapBytes) {
      void* p = mmap(nullptr, m_MapBytes, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p == MAP_FAILED) { m_Error = "mmap failed: " + m_FilePath; m_MapBytes = 0; }
      else { m_Map = static_cast<const uint8_t*>(p); madvise(p, m_MapBytes, MADV_SEQUENTIAL); }
    }
    close(fd);
  }
  // one byte key count, then NUM_KEYS x (6 + 1) bytes; the reference also fails when the header runs into the end of file
  const size_t header = 1 + (size_t)NUM_KEYS * 7;
  if (m_MapBytes < 1 || m_Map[0] != NUM_KEYS || m_MapBytes < header) {
    printf("The header of the GPGPU-sim trace file is not valid.\n");  // LoaderGPGPU.cpp:91-95, 108-112
    exit(1);
  }
  m_First = header;
  m_Cursor = m_First;
}

bool LoaderGPGPU::recordAt(size_t off, uint32_t* reqType, uint32_t* reqSize) const {
  if (off + kRecordHeader > m_MapBytes) return false;
  *reqType = rd<uint32_t>(m_Map + off + 38);
  *reqSize = rd<uint32_t>(m_Map + off + 58);
  // a record whose payload runs past the end of the file raises eof in the reference and is dropped (LoaderGPGPU.cpp:46-52)
  return off + kRecordHeader + (size_t)*reqSize <= m_MapBytes;
}

MemReq_t* LoaderGPGPU::GetCacheline(MemReq_t* memReq) {
  MemReqGPU_t* r = static_cast<MemReqGPU_t*>(memReq);
  uint32_t type = 0, size = 0;
  if (!recordAt(m_Cursor, &type, &size)) {
    r->isEnd = true;
    m_Cursor = m_MapBytes;
    return memReq;
  }
  const uint8_t* p = m_Map + m_Cursor;
  r->kernelID = p[0];
  r->mfType = (fetchTypeGPU)p[1];
  r->cycle = rd<uint64_t>(p + 2);
  r->tpc = rd<uint32_t>(p + 10);
  r->sid = rd<uint32_t>(p + 14);
  r->wid = rd<uint32_t>(p + 18);
  r->pc = rd<uint32_t>(p + 22);
  r->instCnt = rd<uint32_t>(p + 26);
  r->addr = rd<uint64_t>(p + 30);
  r->reqType = (reqTypeGPU)type;
  r->row = rd<uint32_t>(p + 42);
  r->chip = rd<uint32_t>(p + 46);
  r->bank = rd<uint32_t>(p + 50);
  r->col = rd<uint32_t>(p + 54);
  r->reqSize = size;
  r->data.assign(p + kRecordHeader, p + kRecordHeader + size);
  r->isEnd = false;
  m_Cursor += kRecordHeader + size;
  return memReq;
}

unsigned LoaderGPGPU::GetCachelineSize() {
  uint32_t type = 0, size = 0;
  recordAt(m_First, &type, &size);  // an empty trace reads as size 0, like the reference's default-constructed request
  return size;
}

unsigned long long LoaderGPGPU::GetNumLines() {
  unsigned long long n = 0;
  size_t off = m_First;
  uint32_t type = 0, size = 0;
  while (recordAt(off, &type, &size)) {
    off += kRecordHeader + size;
    n++;
  }
  return n;
}

uint64_t LoaderGPGPU::GetChunk(uint8_t* dst, uint64_t maxLines) {
  const uint32_t line = GetCachelineSize();
  uint64_t n = 0;
  uint32_t type = 0, size = 0;
  while (n < maxLines && recordAt(m_Cursor, &type, &size)) {
    if (type == GLOBAL_ACC_R || type == GLOBAL_ACC_W) {
      if (size != line) {
        // the reference hands such a record to the compressor as it is; its compressors are sized once from the first
        // record (main.cpp:86), so what they do with another size is undefined -- refuse instead
        printf("GPGPU-sim trace: record at byte %zu has req_size %u, the trace's line size is %u; mixed-size traces are not supported.\n",
               m_Cursor, size, line);
        exit(1);
      }
      memcpy(dst + n * line, m_Map + m_Cursor + kRecordHeader, line);
      n++;
    }
    m_Cursor += kRecordHeader + size;
  }
  return n;
}

}  // namespace gpgpusim
}  // namespace trace
