// TEST INFRASTRUCTURE ONLY -- stand-in for libnpy's npy.hpp as used by the reference's
// LoaderNPY.cpp:51: LoadArrayFromNumpy(path, shape, fortran_order, data) for uint8 arrays,
// NPY format v1/v2/v3 headers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace npy {
template <typename Scalar, typename ShapeT>
inline void LoadArrayFromNumpy(const std::string& path, std::vector<ShapeT>& shape, bool& fortran_order,
                               std::vector<Scalar>& data) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("npy: cannot open " + path);
  char magic[6];
  f.read(magic, 6);
  if (std::string(magic, 6) != std::string("\x93NUMPY", 6)) throw std::runtime_error("npy: bad magic");
  unsigned char ver[2];
  f.read(reinterpret_cast<char*>(ver), 2);
  uint32_t hlen = 0;
  if (ver[0] == 1) {
    unsigned char b[2];
    f.read(reinterpret_cast<char*>(b), 2);
    hlen = b[0] | (b[1] << 8);
  } else {
    unsigned char b[4];
    f.read(reinterpret_cast<char*>(b), 4);
    hlen = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
  }
  std::string hdr(hlen, '\0');
  f.read(&hdr[0], hlen);
  fortran_order = hdr.find("'fortran_order': True") != std::string::npos;
  shape.clear();
  size_t p = hdr.find("'shape':");
  if (p == std::string::npos) throw std::runtime_error("npy: no shape");
  p = hdr.find('(', p);
  size_t q = hdr.find(')', p);
  std::string dims = hdr.substr(p + 1, q - p - 1);
  size_t i = 0;
  while (i < dims.size()) {
    while (i < dims.size() && (dims[i] < '0' || dims[i] > '9')) i++;
    if (i >= dims.size()) break;
    unsigned long long v = 0;
    while (i < dims.size() && dims[i] >= '0' && dims[i] <= '9') v = v * 10 + (dims[i++] - '0');
    shape.push_back((ShapeT)v);
  }
  size_t n = 1;
  for (auto d : shape) n *= (size_t)d;
  data.resize(n);
  f.read(reinterpret_cast<char*>(data.data()), n * sizeof(Scalar));
}
}  // namespace npy
