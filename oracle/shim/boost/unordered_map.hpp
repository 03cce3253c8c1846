// Minimal stand-in for <boost/unordered_map.hpp> (misc/data.hpp:79 uses it as a name->index table).
#ifndef UPB_SHIM_BOOST_UNORDERED_MAP_HPP
#define UPB_SHIM_BOOST_UNORDERED_MAP_HPP
#include <unordered_map>
namespace boost { using std::unordered_map; }
#endif
