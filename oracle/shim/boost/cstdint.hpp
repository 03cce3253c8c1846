// Minimal stand-in for <boost/cstdint.hpp> (test infrastructure; see oracle/README.md).
// Only the fixed-width integer names the reference uses (misc/defaults.hpp:18-22).
#ifndef UPB_SHIM_BOOST_CSTDINT_HPP
#define UPB_SHIM_BOOST_CSTDINT_HPP
#include <stdint.h>
namespace boost {
  using ::int8_t;  using ::int16_t;  using ::int32_t;  using ::int64_t;
  using ::uint8_t; using ::uint16_t; using ::uint32_t; using ::uint64_t;
}
#endif
