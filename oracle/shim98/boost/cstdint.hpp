// C++98 stand-in for <boost/cstdint.hpp>, used when compiling the reference's hot-path
// translation units (misc/kernel.cpp, data.cpp, peakcall.cpp) in its own -ansi mode.
#ifndef UPB_SHIM98_BOOST_CSTDINT_HPP
#define UPB_SHIM98_BOOST_CSTDINT_HPP
#include <stdint.h>
namespace boost {
  using ::int8_t;  using ::int16_t;  using ::int32_t;  using ::int64_t;
  using ::uint8_t; using ::uint16_t; using ::uint32_t; using ::uint64_t;
}
#endif
