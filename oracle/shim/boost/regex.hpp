// Minimal stand-in for <boost/regex.hpp>: the reference only needs regex, cmatch and
// regex_search on C strings (misc/format.hpp:22-34, misc/format.cpp). The patterns are
// plain ECMAScript, which std::regex accepts unchanged.
#ifndef UPB_SHIM_BOOST_REGEX_HPP
#define UPB_SHIM_BOOST_REGEX_HPP
#include <regex>
namespace boost {
  using std::regex;
  using std::cmatch;
  using std::smatch;
  using std::sub_match;
  using std::regex_search;
  using std::regex_match;
}
#endif
