// Minimal stand-in for <boost/algorithm/string.hpp>: starts_with, ends_with and
// split(is_any_of) with empty fields kept (token_compress_off, Boost's default).
#ifndef UPB_SHIM_BOOST_ALGORITHM_STRING_HPP
#define UPB_SHIM_BOOST_ALGORITHM_STRING_HPP
#include <string>
#include <vector>
namespace boost {
  inline bool starts_with (const std::string& s, const std::string& p) {
    return s.size() >= p.size() && s.compare(0, p.size(), p) == 0;
  }
  inline bool ends_with (const std::string& s, const std::string& p) {
    return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0;
  }
  struct is_any_of_pred {
    std::string chars;
    bool operator() (char c) const { return chars.find(c) != std::string::npos; }
  };
  inline is_any_of_pred is_any_of (const std::string& chars) { is_any_of_pred p; p.chars = chars; return p; }
  template <class Pred>
  inline std::vector<std::string>& split (std::vector<std::string>& out, const std::string& s, Pred pred) {
    out.clear();
    std::string cur;
    for (size_t i = 0; i < s.size(); ++i) {
      if (pred(s[i])) { out.push_back(cur); cur.clear(); } else cur += s[i];
    }
    out.push_back(cur);
    return out;
  }
}
#endif
