#pragma once
#include <vector>
#include "numeric.h"

namespace jxlb {

struct HostNumericTables {
  NumericTables tables;             // pointers into the pools below (host addresses)
  std::vector<float> dequant_pool;
  std::vector<float> llf_pool;
  uint32_t llf_off[6];
};

const HostNumericTables& GetHostNumericTables();

}  // namespace jxlb
