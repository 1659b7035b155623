// comp::CompResult -- totals and base CSV report (mirror of reference src/compressor/CompResult.h:24-86).
#ifndef MPCB_COMPRESULT_H_
#define MPCB_COMPRESULT_H_

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "../loader/Loader.h"
#include "../utils.h"

#define BYTE (8)
#define COMPSIZELIMIT ((ACCESS_GRAN * BYTE) + 32)  // CompResult.h:19: 288 histogram columns are printed

namespace comp {

struct CompResult {
  explicit CompResult(unsigned lineSize) : LineSize(lineSize), OriginalSize(0), CompressedSize(0), CompRatio(0) {}
  virtual ~CompResult() {}

  virtual void Print(std::string workloadName = "", std::string filePath = "") {
    std::ofstream file;
    std::ostream* os = &std::cout;
    if (filePath != "") {
      if (!isFileExists(filePath)) {
        file.open(filePath);
        if (!file.is_open()) {
          std::cout << "File is not open: \"" << filePath << "\"" << std::endl;
          exit(1);
        }
        file << "workload,original_size,compressed_size,compression_ratio," << std::endl;
        file.close();
      }
      file.open(filePath, std::ios_base::app);
      os = &file;
    }
    *os << workloadName << "," << OriginalSize << "," << CompressedSize << "," << formatDouble(CompRatio) << "," << std::endl;
  }
  virtual void PrintDetail(std::string workloadName = "", std::string filePath = "") {}

  std::string CompressorName;
  const unsigned LineSize;
  uint64_t OriginalSize;
  uint64_t CompressedSize;
  double CompRatio;
};

}  // namespace comp
#endif
