// trace::gpgpusim::LoaderGPGPU -- reader of binary GPGPU-Sim memory traces (.log), mirror of the reference's
// src/loader/LoaderGPGPU.{h,cpp} (namespace gpgpusim, LoaderGPGPU.cpp:8-113).
//
// File layout (LoaderGPGPU.cpp:26-54, 82-113): one byte = number of keys (must be 17), 17 x (6-char key, 1-byte size),
// then records of 62 header bytes
//   kid(1) mf_type(1) cycle(8) tpc(4) sid(4) wid(4) pc(4) inst_cnt(4) mem_addr(8) req_type(4) row(4) chip(4) bank(4)
//   col(4) req_size(4)
// followed by req_size payload bytes.  The reference reads one record per GetCacheline call and the driver keeps only
// GLOBAL_ACC_R / GLOBAL_ACC_W records (main.cpp:222-224).  Here the file is memory-mapped; GetChunk() walks the records
// once and gathers the payloads of the kept records into a dense block array for the batched GPU path.
#ifndef MPCB_LOADERGPGPU_H_
#define MPCB_LOADERGPGPU_H_

#include "Loader.h"

#define NUM_KEYS 17

namespace trace {
namespace gpgpusim {

enum reqTypeGPU {  // LoaderGPGPU.h:17-28
  GLOBAL_ACC_R = 0, LOCAL_ACC_R = 1, CONST_ACC_R = 2, TEXTURE_ACC_R = 3, GLOBAL_ACC_W = 4, LOCAL_ACC_W = 5,
  L1_WRBK_ACC = 6, L2_WRBK_ACC = 7, INST_ACC_R = 8,
};
enum fetchTypeGPU { READ_REQUEST = 0, WRITE_REQUEST = 1, READ_REPLY = 2, WRITE_ACK = 3 };  // LoaderGPGPU.h:31-37

struct MemReqGPU_t : public MemReq_t {  // LoaderGPGPU.h:39-105
  uint8_t kernelID = 0;
  uint64_t cycle = 0;
  uint32_t tpc = 0, sid = 0, wid = 0, pc = 0, instCnt = 0;
  reqTypeGPU reqType = GLOBAL_ACC_R;
  fetchTypeGPU mfType = READ_REQUEST;
  uint32_t row = 0, chip = 0, bank = 0, col = 0;
  void Reset() override {
    MemReq_t::Reset();
    kernelID = 0;
    cycle = 0;
    tpc = sid = wid = pc = instCnt = 0;
    reqType = GLOBAL_ACC_R;
    mfType = READ_REQUEST;
    row = chip = bank = col = 0;
  }
};

class LoaderGPGPU : public Loader {
 public:
  static constexpr size_t kRecordHeader = 62;
  explicit LoaderGPGPU(const std::string& filePath);
  ~LoaderGPGPU() override;
  MemReq_t* GetCacheline(MemReq_t* memReq) override;  // LoaderGPGPU.cpp:26-54: every record, any request type
  unsigned GetCachelineSize() override;               // LoaderGPGPU.cpp:16-24: req_size of the first record
  unsigned long long GetNumLines() override;          // LoaderGPGPU.cpp:56-70: number of complete records
  void Reset() override;                              // LoaderGPGPU.cpp:73-79
  // payloads of the GLOBAL_ACC_R/W records (the lines compressLines keeps, main.cpp:216-227), densely packed
  uint64_t GetChunk(uint8_t* dst, uint64_t maxLines) override;
  MemReq_t* NewMemReq() override { return new MemReqGPU_t; }
  const std::string& Error() const { return m_Error; }

 private:
  void unmap();
  bool recordAt(size_t off, uint32_t* reqType, uint32_t* reqSize) const;  // false: no complete record at off
  const uint8_t* m_Map = nullptr;
  size_t m_MapBytes = 0, m_First = 0, m_Cursor = 0;
  std::string m_Error;
};

}  // namespace gpgpusim
}  // namespace trace
#endif
