// kernels_elastic_col.cuh -- register-blocked elastic force kernel (placeholder: forwards to the per-point kernel).
#pragma once
#include "kernels_elastic.cuh"
namespace sb200 {
inline void launch_elastic_col(DeviceMesh* mp, const ElasticArgs& a) {
  k_elastic_forces_pt<false, false><<<a.nb_elements, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
}
}  // namespace sb200
