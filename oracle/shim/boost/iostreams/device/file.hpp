// Part of the Boost.Iostreams stand-in: everything lives in filtering_stream.hpp.
#include <boost/iostreams/filtering_stream.hpp>
