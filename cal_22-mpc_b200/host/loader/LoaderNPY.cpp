#ifndef _GNU_SOURCE
#define _GNU_SOURCE  // O_DIRECT
#endif
#include "LoaderNPY.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>

namespace trace {

LoaderNPY::LoaderNPY(const std::string& filePath) : Loader(filePath) { Reset(); }
LoaderNPY::~LoaderNPY() { unmap(); }

void LoaderNPY::unmap() {
  if (m_Fd >= 0) close(m_Fd);
  m_Fd = -1;
  if (m_Map) munmap(const_cast<uint8_t*>(m_Map), m_MapBytes);
  m_Map = nullptr;
  m_Data = nullptr;
}

// NPY header: magic "\x93NUMPY", major, minor, header length (u16 for v1, u32 for v2/v3), python dict literal
void LoaderNPY::Reset() {
  unmap();
  m_Rows = m_LineSize = m_CurrentLine = 0;
  m_Error.clear();
  int fd = open(m_FilePath.c_str(), O_RDONLY);
  if (fd < 0) { m_Error = "cannot open " + m_FilePath; return; }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 10) { m_Error = "not an .npy file: " + m_FilePath; close(fd); return; }
  m_MapBytes = (size_t)st.st_size;
  void* p = mmap(nullptr, m_MapBytes, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { m_Error = "mmap failed: " + m_FilePath; m_MapBytes = 0; return; }
  m_Map = static_cast<const uint8_t*>(p);
  madvise(p, m_MapBytes, MADV_SEQUENTIAL);
  if (memcmp(m_Map, "\x93NUMPY", 6) != 0) { m_Error = "bad .npy magic: " + m_FilePath; return; }
  const unsigned major = m_Map[6];
  size_t hlen, hoff;
  if (major == 1) { hlen = m_Map[8] | (m_Map[9] << 8); hoff = 10; }
  else { hlen = m_Map[8] | (m_Map[9] << 8) | (m_Map[10] << 16) | ((size_t)m_Map[11] << 24); hoff = 12; }
  if (hoff + hlen > m_MapBytes) { m_Error = "truncated .npy header"; return; }
  std::string hdr(reinterpret_cast<const char*>(m_Map + hoff), hlen);
  if (hdr.find("'fortran_order': True") != std::string::npos) { m_Error = "fortran-order .npy is not a block dump"; return; }
  size_t d = hdr.find("'descr':");
  if (d != std::string::npos) {
    size_t q0 = hdr.find('\'', d + 8), q1 = hdr.find('\'', q0 + 1);
    std::string descr = hdr.substr(q0 + 1, q1 - q0 - 1);
    if (descr != "|u1" && descr != "u1" && descr != "|i1" && descr != "<u1") { m_Error = "dump dtype must be uint8, got " + descr; return; }
  }
  size_t s = hdr.find("'shape':");
  if (s == std::string::npos) { m_Error = "no shape in .npy header"; return; }
  size_t a = hdr.find('(', s), b = hdr.find(')', a);
  std::vector<uint64_t> shape;
  for (size_t i = a + 1; i < b;) {
    while (i < b && (hdr[i] < '0' || hdr[i] > '9')) i++;
    if (i >= b) break;
    uint64_t v = 0;
    while (i < b && hdr[i] >= '0' && hdr[i] <= '9') v = v * 10 + (uint64_t)(hdr[i++] - '0');
    shape.push_back(v);
  }
  if (shape.size() != 2) { m_Error = "dump must be a 2-D [lines, lineSize] array"; return; }
  if (hoff + hlen + shape[0] * shape[1] > m_MapBytes) { m_Error = "truncated .npy data"; return; }
  m_Rows = shape[0];
  m_LineSize = shape[1];
  m_Data = m_Map + hoff + hlen;
}

MemReq_t* LoaderNPY::GetCacheline(MemReq_t* memReq) {
  memReq->addr = 0;
  memReq->rw = NA;
  memReq->reqSize = (uint32_t)m_LineSize;
  if (m_CurrentLine < m_Rows)
    memReq->data.assign(m_Data + m_CurrentLine * m_LineSize, m_Data + (m_CurrentLine + 1) * m_LineSize);
  m_CurrentLine++;
  // isEnd is raised on the call that RETURNS the last row, and the driver breaks before compressing it
  // (LoaderNPY.cpp:28-32, main.cpp:239-242): an N-row file yields N-1 blocks.
  memReq->isEnd = (m_CurrentLine >= m_Rows);
  return memReq;
}

unsigned LoaderNPY::GetCachelineSize() { return (unsigned)m_LineSize; }
unsigned long long LoaderNPY::GetNumLines() { return m_Rows; }

uint64_t LoaderNPY::GetChunk(uint8_t* dst, uint64_t maxLines) {
  const uint64_t usable = m_Rows ? m_Rows - 1 : 0;  // the last row is never compressed (see GetCacheline)
  if (m_CurrentLine >= usable) return 0;
  uint64_t n = usable - m_CurrentLine;
  if (n > maxLines) n = maxLines;
  memcpy(dst, m_Data + m_CurrentLine * m_LineSize, n * m_LineSize);
  m_CurrentLine += n;
  return n;
}

const uint8_t* LoaderNPY::GetAll(uint64_t* nLines) {
  const uint64_t usable = m_Rows ? m_Rows - 1 : 0;
  if (m_CurrentLine >= usable) { *nLines = 0; return m_Data; }
  const uint8_t* p = m_Data + m_CurrentLine * m_LineSize;
  *nLines = usable - m_CurrentLine;
  m_CurrentLine = usable;
  return p;
}

int LoaderNPY::GetFile(uint64_t* dataOffset, uint64_t* nLines, bool directIo) {
  *dataOffset = 0;
  *nLines = 0;
  if (!m_Data) return -1;
  if (m_Fd >= 0) close(m_Fd);
  m_Fd = open(m_FilePath.c_str(), O_RDONLY | (directIo ? O_DIRECT : 0));
  if (m_Fd < 0) return -1;
  const uint64_t usable = m_Rows ? m_Rows - 1 : 0;  // the last row is never compressed (see GetCacheline)
  if (m_CurrentLine < usable) {
    *dataOffset = (uint64_t)(m_Data - m_Map) + m_CurrentLine * m_LineSize;
    *nLines = usable - m_CurrentLine;
    m_CurrentLine = usable;
  }
  return m_Fd;
}

}  // namespace trace
