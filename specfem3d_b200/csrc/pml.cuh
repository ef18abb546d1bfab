// pml.cuh -- C-PML elastic elements.  Not implemented yet: every entry fails loudly.
#pragma once
#include "device_mesh.h"
namespace sb200 {
struct PML {};
inline void pml_destroy(DeviceMesh*) {}
inline void pml_snapshot_displ(DeviceMesh*, int, realw, realw, realw) { exit_on_error("C-PML is not implemented yet"); }
inline void pml_compute_forces(DeviceMesh*, int, int, realw) { exit_on_error("C-PML is not implemented yet"); }
}  // namespace sb200
extern "C" void SB200_FC(prepare_fields_elastic_pml)(long*, int*, int*, int*, int*, realw*, realw*, realw*, realw*, realw*, realw*) {
  sb200::exit_on_error("prepare_fields_elastic_pml: C-PML is not implemented yet");
}
extern "C" void SB200_FC(pml_impose_boundary_condition_elastic_cuda)(long*) {
  sb200::exit_on_error("pml_impose_boundary_condition_elastic_cuda: C-PML is not implemented yet");
}
extern "C" void SB200_FC(transfer_pml_displ_from_device)(int*, realw*, realw*, long*) {
  sb200::exit_on_error("transfer_pml_displ_from_device: C-PML is not implemented yet");
}
extern "C" void SB200_FC(transfer_pml_displ_to_device)(int*, realw*, realw*, long*) {
  sb200::exit_on_error("transfer_pml_displ_to_device: C-PML is not implemented yet");
}
