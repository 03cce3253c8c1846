// C++98 stand-in for <boost/regex.hpp>: the hot-path translation units only see the
// `const regex` globals declared in misc/format.hpp:22-34 and never match with them.
#ifndef UPB_SHIM98_BOOST_REGEX_HPP
#define UPB_SHIM98_BOOST_REGEX_HPP
namespace boost {
  struct regex { regex (const char*) {} };
}
#endif
