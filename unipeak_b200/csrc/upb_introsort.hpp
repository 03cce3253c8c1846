// upb_introsort.hpp — the leading part of what std::sort would produce, without sorting the rest.
//
// src/strand_shift.cpp:198 sorts ALL candidate regions by Region::sum() with std::sort (unstable: the order of ties is whatever
// libstdc++'s introsort makes of the emission-ordered vector) and then looks only at the first ~1000 that qualify.
// To return the same regions in the same order, the prefix has to come out of the same algorithm.  This runs libstdc++'s own
// introsort steps (std::__unguarded_partition_pivot, std::__partial_sort, std::__insertion_sort, std::__unguarded_insertion_sort
// — the pieces std::sort is made of, bits/stl_algo.h) but does not descend into partitions that lie entirely beyond `bound`:
// a partition step never moves an element across the cut, so what happens right of the cut cannot change anything left of it.
// Result: [first, first + sortedEnd) is exactly the first sortedEnd elements of std::sort(first, last, comp), sortedEnd >= bound.
// 107 K regions: 5 ms for the full sort, 0.2 ms for a 2 K prefix.
#pragma once

#include <algorithm>
#include <cstddef>

namespace upb {

template <class It, class Cmp>
void introsort_prefix_loop (It first, It last, long depth, It bound, It* sortedEnd, Cmp comp) {
  while (last - first > 16) {                                         // std::_S_threshold
    if (depth == 0) { std::__partial_sort(first, last, last, comp); break; }
    --depth;
    It cut = std::__unguarded_partition_pivot(first, last, comp);
    if (cut < bound) introsort_prefix_loop(cut, last, depth, bound, sortedEnd, comp);
    else if (cut < *sortedEnd) *sortedEnd = cut;                      // everything from `cut` on is left as it is
    last = cut;
  }
  if (last >= bound && last < *sortedEnd) *sortedEnd = last;
}

// returns the number of leading elements that are in their std::sort places (>= min(bound, n))
template <class It, class Compare>
size_t sort_prefix_like_std_sort (It first, It last, size_t bound, Compare compare) {
  const size_t n = (size_t)(last - first);
  if (n == 0) return 0;
  if (bound > n) bound = n;
  auto comp = __gnu_cxx::__ops::__iter_comp_iter(compare);
  It end = last;
  introsort_prefix_loop(first, last, (long)std::__lg(last - first) * 2, first + bound, &end, comp);
  if (end - first > 16) {                                             // std::__final_insertion_sort, cut short at `end`
    std::__insertion_sort(first, first + 16, comp);
    std::__unguarded_insertion_sort(first + 16, end, comp);
  } else {
    std::__insertion_sort(first, end, comp);
  }
  return (size_t)(end - first);
}

}  // namespace upb
