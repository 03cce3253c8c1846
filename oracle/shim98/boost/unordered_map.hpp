// C++98 stand-in for <boost/unordered_map.hpp>: only ContigTable (misc/data.cpp:196-261) uses it,
// and only from inside data.cpp, through a pointer member.
#ifndef UPB_SHIM98_BOOST_UNORDERED_MAP_HPP
#define UPB_SHIM98_BOOST_UNORDERED_MAP_HPP
#include <tr1/unordered_map>
namespace boost { using std::tr1::unordered_map; }
#endif
