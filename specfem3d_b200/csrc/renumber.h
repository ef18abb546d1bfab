// renumber.h -- the library's own numbering of the global GLL points.
//
// The Fortran solver numbers the points by first touch in element order (get_global.f90 / Appendix D of SURVEY.md): the 64
// points an interior element introduces are contiguous, but the 125 points it READS are spread over ~25 runs of 4-5
// points, so a 32-lane gather or RED.ADD of the force kernels touches ~8-10 32-byte sectors and fills them to 62 %
// (profiles/r01i_force_po2_ncu_summary.txt).  Global point ids never leave the library except as indices into arrays the
// library copies itself, so it renumbers them once, in prepare_constants_device:
//
//   walk the elements in (predicted) launch order -- per domain: inner elements, then outer ones; regular before irregular;
//   ascending ispec; groups of five, exactly like group_store_build -- and inside a group visit the points in the lane
//   order of the row-map kernels, k-plane by k-plane, row by row, ALONG xi THROUGH THE FIVE ELEMENTS; number on first touch.
//
// For xi-neighbouring elements a row of a group is then ~21 consecutive points, a k-plane of its new points 80 and the whole
// group 320 (1.3 KB per component): a warp's gather / RED touches 3-4 full sectors (tools/micro/gather_red_layouts.cu:
// gather + RED 1.33 -> 0.69 ms per 1M elements together with the row lane map).  For any other mesh it is just another
// valid numbering.  Where host arrays are indexed by global point (inverse mass matrices, wavefield transfers, interface
// point lists) the permutation is applied at the boundary; results are identical up to the order of the atomic sums.
#pragma once
#include <vector>

namespace sb200 {

// h_ibool: [nspec][125], 1-based.  Returns perm[old 0-based id] = new 0-based id (a bijection on [0, nglob)).
inline std::vector<int> compute_point_numbering(const int* h_ibool, int nspec, int nglob, const int* ispec_is_inner,
                                                const int* ispec_is_elastic, const int* ispec_is_acoustic,
                                                const int* irregular_element_number) {
  std::vector<int> perm((size_t)nglob, -1);
  int next = 0;
  std::vector<int> list;
  list.reserve(nspec);
  for (int domain = 0; domain < 3; ++domain)      // 0 elastic, 1 acoustic, 2 anything else
    for (int inner = 1; inner >= 0; --inner)      // inner elements are the bulk: they get the unbroken rows
      for (int irr = 0; irr < 2; ++irr) {         // group_store_build: regular groups first, then irregular ones
        list.clear();
        for (int e = 0; e < nspec; ++e) {
          const int dom = ispec_is_elastic[e] ? 0 : (ispec_is_acoustic[e] ? 1 : 2);
          if (dom != domain || (ispec_is_inner[e] != 0) != (inner != 0) || (irregular_element_number[e] > 0) != (irr != 0)) continue;
          list.push_back(e);
        }
        for (size_t g0 = 0; g0 < list.size(); g0 += 5) {
          const int nel = (int)(list.size() - g0 < 5 ? list.size() - g0 : 5);
          for (int k = 0; k < 5; ++k)
            for (int j = 0; j < 5; ++j)
              for (int el = 0; el < nel; ++el) {
                const int* ib = h_ibool + (size_t)list[g0 + el] * 125 + k * 25 + j * 5;
                for (int i = 0; i < 5; ++i) {
                  const int old = ib[i] - 1;
                  if (perm[old] < 0) perm[old] = next++;
                }
              }
        }
      }
  for (int p = 0; p < nglob; ++p)  // points no element refers to (none in a valid database)
    if (perm[p] < 0) perm[p] = next++;
  return perm;
}

}  // namespace sb200
