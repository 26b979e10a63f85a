// Natural (zig-zag) coefficient orders for the 13 coefficient-order ids, built once on the host and uploaded to HBM.
#include <mutex>
#include <vector>

#include "frame_parser.h"
#include "vardct_sections.h"

namespace jxlb {

const NaturalOrders& NaturalOrderPoolHost() {
  static NaturalOrders nat;
  static std::vector<uint16_t> pool;
  static std::once_flag once;
  std::call_once(once, [] {
    uint32_t off = 0;
    for (uint32_t o = 0; o < kNumOrders; ++o) {
      uint32_t s = OrderRepresentative(o);
      std::vector<uint32_t> ord;
      NaturalCoeffOrder(StrategyCellsX(s), StrategyCellsY(s), &ord);
      nat.offset[o] = off;
      nat.size[o] = (uint32_t) ord.size();
      for (uint32_t v : ord) pool.push_back((uint16_t) v);
      off += (uint32_t) ord.size();
    }
    nat.pool = pool.data();
  });
  return nat;
}

}  // namespace jxlb
