// host_loop.cu -- the native host side above the C ABI (include/specfem3d_b200_host.h): iterate_time of a forward run.
//
// Host code only (it sits in a .cu file so that the one build recipe of csrc/ picks it up): every device action goes through
// the Fortran-callable entry points, by reference, in the order the Fortran solver calls them:
//   iterate_time                                        src/specfem3D/iterate_time.F90:219-330
//   update_displ_Newmark (GPU branch)                   src/specfem3D/update_displacement_scheme.f90:281-312
//   compute_forces_acoustic_calling (GPU branches)      src/specfem3D/compute_forces_acoustic_calling_routine.f90:28-330
//   compute_forces_viscoelastic_calling (GPU branches)  src/specfem3D/compute_forces_viscoelastic_calling_routine.F90:30-435
//   assemble_MPI_vector_send_cuda / _write_cuda         src/specfem3D/assemble_MPI_vector.f90:640-830
//   write_seismograms (GPU branch)                      src/specfem3D/write_seismograms.f90:28-200
//   check_stability                                     src/specfem3D/check_stability.f90:85-120
// specfem3d_b200/time_loop.py is the same loop in Python; this one removes the interpreter from the step (configs[0] as four
// handles on one GPU is bound by it) and is what a C / C++ application would call.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "specfem3d_b200_host.h"

namespace {

typedef specfem_b200_host_rank Rank;
typedef specfem_b200_host_run Run;

[[noreturn]] void host_error(const char* msg) {
  fprintf(stderr, "ERROR (specfem3d_b200 host loop): %s\n", msg);
  exit(1);
}

Rank* rank_by_id(Run* run, int id) {
  for (int r = 0; r < run->nranks; ++r)
    if (run->ranks[r].myrank == id) return &run->ranks[r];
  return nullptr;
}

// The "MPI" between the ranks of this process: the send buffer of interface n of rank r lands in the receive buffer of the
// interface of the neighbour that points back to r (what isend / irecv pairs do in assemble_MPI_vector_send_cuda).
void exchange_on_host(Run* run, int ncomp) {
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    const size_t mx = (size_t)r.max_nibool_interfaces_ext_mesh;
    for (int n = 0; n < r.num_interfaces_ext_mesh; ++n) {
      Rank* nb = rank_by_id(run, r.my_neighbors_ext_mesh[n]);
      if (!nb) host_error("host_exchange: a neighbour rank is not driven by this process");
      int m = -1;
      for (int q = 0; q < nb->num_interfaces_ext_mesh; ++q)
        if (nb->my_neighbors_ext_mesh[q] == r.myrank) m = q;
      if (m < 0) host_error("host_exchange: the neighbour has no interface pointing back");
      const size_t cnt = (size_t)r.nibool_interfaces_ext_mesh[n];
      if ((int)cnt != nb->nibool_interfaces_ext_mesh[m]) host_error("host_exchange: interface sizes differ on the two sides");
      const size_t nbmx = (size_t)nb->max_nibool_interfaces_ext_mesh;
      if (ncomp == 3)
        memcpy(nb->buffer_recv_vector + 3 * nbmx * m, r.buffer_send_vector + 3 * mx * n, 3 * cnt * sizeof(realw));
      else
        memcpy(nb->buffer_recv_scalar + nbmx * m, r.buffer_send_scalar + mx * n, cnt * sizeof(realw));
    }
  }
}

void acoustic_step(Run* run, int it) {
  int fwd = 1, sifs = run->STACEY_INSTEAD_OF_FREE_SURFACE;
  const bool multi = run->NPROC > 1;
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (r.acoustic) acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &fwd);
  }
  std::vector<int> run_number(run->NSOURCES > 0 ? run->NSOURCES : 1, 0);  // run_number_of_the_source(:) = 0
  for (int iphase = 1; iphase <= 2; ++iphase) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (!r.acoustic) continue;
      compute_forces_acoustic_cuda_(r.Mesh_pointer, &iphase, &r.nspec_outer_acoustic, &r.nspec_inner_acoustic, &fwd);
      if (iphase == 1) {
        if (run->STACEY_ABSORBING_CONDITIONS && run->UNDO_ATTENUATION_AND_OR_PML) {
          compute_stacey_acoustic_undoatt_cuda_(r.Mesh_pointer, &iphase, &fwd);  // no record kept
        } else if (run->STACEY_ABSORBING_CONDITIONS) {
          compute_stacey_acoustic_cuda_(r.Mesh_pointer, &iphase, r.b_absorb_potential, &fwd);
          if (run->SAVE_FORWARD && r.absorb_potential_history) {  // write_abs(IOABS_AC, b_absorb_potential, b_reclen_potential, it)
            const size_t rec = (size_t)25 * r.num_abs_boundary_faces;
            memcpy(r.absorb_potential_history + rec * (it - 1), r.b_absorb_potential, rec * sizeof(realw));
          }
        }
        if (run->ELASTIC_SIMULATION) compute_coupling_ac_el_cuda_(r.Mesh_pointer, &iphase, &r.num_coupling_ac_el_faces, &fwd);
        if (run->NSOURCES > 0) {
          int ns = run->NSOURCES;
          compute_add_sources_ac_cuda_(r.Mesh_pointer, &ns, const_cast<double*>(run->stf_acoustic + (size_t)(it - 1) * ns), run_number.data());
        }
      }
    }
    if (multi) {
      if (iphase == 1) {
        for (int ir = 0; ir < run->nranks; ++ir) {
          Rank& r = run->ranks[ir];
          if (r.acoustic) transfer_boun_pot_from_device_(r.Mesh_pointer, r.buffer_send_scalar, &fwd);
        }
        if (run->host_exchange) exchange_on_host(run, 1);
      } else {
        for (int ir = 0; ir < run->nranks; ++ir) {
          Rank& r = run->ranks[ir];
          if (r.acoustic) transfer_asmbl_pot_to_device_(r.Mesh_pointer, r.buffer_recv_scalar, &fwd);
        }
      }
    }
  }
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (!r.acoustic) continue;
    kernel_3_acoustic_cuda_(r.Mesh_pointer, &run->deltatover2, &run->deltatover2, &fwd);
    acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &fwd);
  }
}

void elastic_step(Run* run, int it) {
  int fwd = 1, att = run->ATTENUATION ? 1 : 0;
  const bool multi = run->NPROC > 1;
  for (int iphase = 1; iphase <= 2; ++iphase) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (!r.elastic) continue;
      // COMPUTE_AND_STORE_STRAIN = ATTENUATION in a forward run (read_mesh_databases.F90)
      compute_forces_viscoelastic_cuda_(r.Mesh_pointer, &iphase, &run->deltat, &r.nspec_outer_elastic, &r.nspec_inner_elastic, &att, &att, &fwd);
      if (iphase == 1) {
        if (run->STACEY_ABSORBING_CONDITIONS && run->UNDO_ATTENUATION_AND_OR_PML) {
          compute_stacey_viscoelastic_undoatt_cuda_(r.Mesh_pointer, &iphase, &fwd);  // no record kept: the kernel run re-computes it
        } else if (run->STACEY_ABSORBING_CONDITIONS) {
          compute_stacey_viscoelastic_cuda_(r.Mesh_pointer, &iphase, r.b_absorb_field, &fwd);
          if (run->SAVE_FORWARD && r.absorb_field_history) {  // write_abs(IOABS, b_absorb_field, b_reclen_field, it)
            const size_t rec = (size_t)3 * 25 * r.num_abs_boundary_faces;
            memcpy(r.absorb_field_history + rec * (it - 1), r.b_absorb_field, rec * sizeof(realw));
          }
        }
        if (run->ACOUSTIC_SIMULATION && r.num_coupling_ac_el_faces > 0)
          compute_coupling_el_ac_cuda_(r.Mesh_pointer, &iphase, &r.num_coupling_ac_el_faces, &fwd);
        if (run->NSOURCES > 0) {
          int ns = run->NSOURCES;
          compute_add_sources_el_cuda_(r.Mesh_pointer, const_cast<double*>(run->stf_elastic + (size_t)(it - 1) * ns), &ns);
        }
        if (multi) transfer_boundary_from_device_a_(r.Mesh_pointer, &r.nspec_outer_elastic);
      }
    }
    if (multi && iphase == 2) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic) sync_copy_from_device_(r.Mesh_pointer, &iphase, r.buffer_send_vector);
      }
      if (run->host_exchange) exchange_on_host(run, 3);
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic)
          transfer_boundary_to_device_a_(r.Mesh_pointer, r.buffer_recv_vector, &r.num_interfaces_ext_mesh, &r.max_nibool_interfaces_ext_mesh);
      }
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic)
          transfer_asmbl_accel_to_device_(r.Mesh_pointer, r.buffer_recv_vector, &r.num_interfaces_ext_mesh, &r.max_nibool_interfaces_ext_mesh,
                                          r.nibool_interfaces_ext_mesh, r.ibool_interfaces_ext_mesh, &fwd);
      }
    }
  }
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (!r.elastic) continue;
    if (run->PML_CONDITIONS) pml_impose_boundary_condition_elastic_cuda_(r.Mesh_pointer);
    int ocean = run->APPROXIMATE_OCEAN_LOAD ? 1 : 0;
    kernel_3_a_cuda_(r.Mesh_pointer, &run->deltatover2, &run->deltatover2, &ocean, &fwd);
    if (ocean) {  // compute_forces_viscoelastic_calling_routine.F90:380-424
      compute_coupling_ocean_cuda_(r.Mesh_pointer, &fwd);
      kernel_3_b_cuda_(r.Mesh_pointer, &run->deltatover2, &run->deltatover2, &fwd);
    }
  }
}

// ---- kernel runs (SIMULATION_TYPE 3): specfem3d_b200/kernel_run.py is the same loop in Python -------------------------------
// the halo of the reconstructed wavefield between the ranks of this process (b_buffer_*)
void exchange_on_host_b(Run* run, int ncomp) {
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    const size_t mx = (size_t)r.max_nibool_interfaces_ext_mesh;
    for (int n = 0; n < r.num_interfaces_ext_mesh; ++n) {
      Rank* nb = rank_by_id(run, r.my_neighbors_ext_mesh[n]);
      if (!nb) host_error("host_exchange: a neighbour rank is not driven by this process");
      int m = -1;
      for (int q = 0; q < nb->num_interfaces_ext_mesh; ++q)
        if (nb->my_neighbors_ext_mesh[q] == r.myrank) m = q;
      if (m < 0) host_error("host_exchange: the neighbour has no interface pointing back");
      const size_t cnt = (size_t)r.nibool_interfaces_ext_mesh[n], nbmx = (size_t)nb->max_nibool_interfaces_ext_mesh;
      if (ncomp == 3)
        memcpy(nb->b_buffer_recv_vector + 3 * nbmx * m, r.b_buffer_send_vector + 3 * mx * n, 3 * cnt * sizeof(realw));
      else
        memcpy(nb->b_buffer_recv_scalar + nbmx * m, r.b_buffer_send_scalar + mx * n, cnt * sizeof(realw));
    }
  }
}

// foa = 0: compute_forces_viscoelastic_GPU_calling (both wavefields); 1 / 3: the separate routines of a coupled run
void elastic_step_kl(Run* run, int it, int foa) {
  const bool multi = run->NPROC > 1, fwd = foa != 3, bwd = foa != 1;
  int one = 1, three = 3, zero = 0, strain = 1, att = 0, NSTEP = run->NSTEP;
  realw dt_forces = foa == 3 ? run->b_deltat : run->deltat;
  for (int iphase = 1; iphase <= 2; ++iphase) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (r.elastic)
        compute_forces_viscoelastic_cuda_(r.Mesh_pointer, &iphase, &dt_forces, &r.nspec_outer_elastic, &r.nspec_inner_elastic, &strain, &att, &foa);
    }
    if (multi && iphase == 2 && fwd) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic) sync_copy_from_device_(r.Mesh_pointer, &iphase, r.buffer_send_vector);
      }
      if (run->host_exchange) exchange_on_host(run, 3);
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic)
          transfer_boundary_to_device_a_(r.Mesh_pointer, r.buffer_recv_vector, &r.num_interfaces_ext_mesh, &r.max_nibool_interfaces_ext_mesh);
      }
    }
    if (iphase == 1) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (!r.elastic) continue;
        if (run->STACEY_ABSORBING_CONDITIONS) {
          if (bwd && r.num_abs_boundary_faces > 0) {  // read_abs(IOABS, b_absorb_field, b_reclen_field, NSTEP - it + 1)
            const size_t rec = (size_t)3 * 25 * r.num_abs_boundary_faces;
            memcpy(r.b_absorb_field, r.absorb_field_history + rec * (NSTEP - it), rec * sizeof(realw));
          }
          compute_stacey_viscoelastic_cuda_(r.Mesh_pointer, &iphase, r.b_absorb_field, &foa);
        }
        if (run->ACOUSTIC_SIMULATION && r.num_coupling_ac_el_faces > 0)
          compute_coupling_el_ac_cuda_(r.Mesh_pointer, &iphase, &r.num_coupling_ac_el_faces, &foa);
        if (fwd && r.nrec_local > 0 && it < NSTEP) {
          int itf = it;
          add_sources_el_sim_type_2_or_3_(r.Mesh_pointer, r.source_adjoint, &run->nrec, &r.nrec_local, &NSTEP, &itf);
        }
        if (bwd && run->NSOURCES > 0) {
          int ns = run->NSOURCES;
          compute_add_sources_el_s3_cuda_(r.Mesh_pointer, const_cast<double*>(run->stf_elastic + (size_t)(NSTEP - it) * ns), &ns);
        }
        if (multi) {
          if (fwd) transfer_boundary_from_device_a_(r.Mesh_pointer, &r.nspec_outer_elastic);
          if (bwd) transfer_boun_accel_from_device_(r.Mesh_pointer, r.b_buffer_send_vector, &three);
        }
      }
      if (multi && bwd && run->host_exchange) exchange_on_host_b(run, 3);
    } else if (multi) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (!r.elastic) continue;
        if (fwd)
          transfer_asmbl_accel_to_device_(r.Mesh_pointer, r.buffer_recv_vector, &r.num_interfaces_ext_mesh, &r.max_nibool_interfaces_ext_mesh,
                                          r.nibool_interfaces_ext_mesh, r.ibool_interfaces_ext_mesh, &one);
        if (bwd)
          transfer_asmbl_accel_to_device_(r.Mesh_pointer, r.b_buffer_recv_vector, &r.num_interfaces_ext_mesh, &r.max_nibool_interfaces_ext_mesh,
                                          r.nibool_interfaces_ext_mesh, r.ibool_interfaces_ext_mesh, &three);
      }
    }
  }
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (!r.elastic) continue;
    if (fwd) kernel_3_a_cuda_(r.Mesh_pointer, &run->deltatover2, &run->b_deltatover2, &zero, &one);
    if (bwd) kernel_3_a_cuda_(r.Mesh_pointer, &run->deltatover2, &run->b_deltatover2, &zero, &three);
  }
}

void acoustic_step_kl(Run* run, int it, int foa) {
  const bool multi = run->NPROC > 1, fwd = foa != 3, bwd = foa != 1;
  int one = 1, three = 3, sifs = run->STACEY_INSTEAD_OF_FREE_SURFACE, NSTEP = run->NSTEP;
  std::vector<int> run_number(run->NSOURCES > 0 ? run->NSOURCES : 1, 0);
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (!r.acoustic) continue;
    if (fwd) acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &one);
    if (bwd) acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &three);
  }
  for (int iphase = 1; iphase <= 2; ++iphase) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (!r.acoustic) continue;
      compute_forces_acoustic_cuda_(r.Mesh_pointer, &iphase, &r.nspec_outer_acoustic, &r.nspec_inner_acoustic, &foa);
      if (iphase != 1) continue;
      if (run->STACEY_ABSORBING_CONDITIONS) {
        if (bwd && r.num_abs_boundary_faces > 0) {
          const size_t rec = (size_t)25 * r.num_abs_boundary_faces;
          memcpy(r.b_absorb_potential, r.absorb_potential_history + rec * (NSTEP - it), rec * sizeof(realw));
        }
        compute_stacey_acoustic_cuda_(r.Mesh_pointer, &iphase, r.b_absorb_potential, &foa);
      }
      if (run->ELASTIC_SIMULATION) {
        if (fwd) compute_coupling_ac_el_cuda_(r.Mesh_pointer, &iphase, &r.num_coupling_ac_el_faces, &one);
        if (bwd) compute_coupling_ac_el_cuda_(r.Mesh_pointer, &iphase, &r.num_coupling_ac_el_faces, &three);
      }
      if (fwd && r.nrec_local > 0 && it < NSTEP) {
        int itf = it;
        add_sources_ac_sim_2_or_3_cuda_(r.Mesh_pointer, r.source_adjoint, &run->nrec, &r.nrec_local, &NSTEP, &itf);
      }
      if (bwd && run->NSOURCES > 0) {
        int ns = run->NSOURCES;
        compute_add_sources_ac_s3_cuda_(r.Mesh_pointer, &ns, const_cast<double*>(run->stf_acoustic + (size_t)(NSTEP - it) * ns), run_number.data());
      }
    }
    if (!multi) continue;
    if (iphase == 1) {
      if (fwd) {
        for (int ir = 0; ir < run->nranks; ++ir) {
          Rank& r = run->ranks[ir];
          if (r.acoustic) transfer_boun_pot_from_device_(r.Mesh_pointer, r.buffer_send_scalar, &one);
        }
        if (run->host_exchange) exchange_on_host(run, 1);
      }
      if (bwd) {
        for (int ir = 0; ir < run->nranks; ++ir) {
          Rank& r = run->ranks[ir];
          if (r.acoustic) transfer_boun_pot_from_device_(r.Mesh_pointer, r.b_buffer_send_scalar, &three);
        }
        if (run->host_exchange) exchange_on_host_b(run, 1);
      }
    } else {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (!r.acoustic) continue;
        if (fwd) transfer_asmbl_pot_to_device_(r.Mesh_pointer, r.buffer_recv_scalar, &one);
        if (bwd) transfer_asmbl_pot_to_device_(r.Mesh_pointer, r.b_buffer_recv_scalar, &three);
      }
    }
  }
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (!r.acoustic) continue;
    kernel_3_acoustic_cuda_(r.Mesh_pointer, &run->deltatover2, &run->b_deltatover2, &foa);
    if (fwd) acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &one);
    if (bwd) acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &three);
  }
}

// ---- undo-attenuation kernel runs: specfem3d_b200/undo_att.py is the same loop in Python ----------------------------------
// the MPI part of one phase for the accel (foa 1) or b_accel (foa 3) field
void exchange_elastic_undoatt(Run* run, int iphase, int foa) {
  if (run->NPROC == 1) return;
  int three = 3;
  if (iphase == 1) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (!r.elastic) continue;
      if (foa == 1) transfer_boundary_from_device_a_(r.Mesh_pointer, &r.nspec_outer_elastic);
      else transfer_boun_accel_from_device_(r.Mesh_pointer, r.b_buffer_send_vector, &three);
    }
    if (foa == 3 && run->host_exchange) exchange_on_host_b(run, 3);
  } else {
    if (foa == 1) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic) sync_copy_from_device_(r.Mesh_pointer, &iphase, r.buffer_send_vector);
      }
      if (run->host_exchange) exchange_on_host(run, 3);
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic)
          transfer_boundary_to_device_a_(r.Mesh_pointer, r.buffer_recv_vector, &r.num_interfaces_ext_mesh, &r.max_nibool_interfaces_ext_mesh);
      }
    }
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (r.elastic)
        transfer_asmbl_accel_to_device_(r.Mesh_pointer, foa == 3 ? r.b_buffer_recv_vector : r.buffer_recv_vector, &r.num_interfaces_ext_mesh,
                                        &r.max_nibool_interfaces_ext_mesh, r.nibool_interfaces_ext_mesh, r.ibool_interfaces_ext_mesh, &foa);
    }
  }
}

// foa = 3: compute_forces_acoustic_backward_calling on the re-computed potentials (it = it_tmp); 1: the adjoint potentials
void acoustic_one_field_undoatt(Run* run, int it, int foa) {
  int sifs = run->STACEY_INSTEAD_OF_FREE_SURFACE, NSTEP = run->NSTEP;
  std::vector<int> run_number(run->NSOURCES > 0 ? run->NSOURCES : 1, 0);
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (r.acoustic) acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &foa);
  }
  for (int iphase = 1; iphase <= 2; ++iphase) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (!r.acoustic) continue;
      compute_forces_acoustic_cuda_(r.Mesh_pointer, &iphase, &r.nspec_outer_acoustic, &r.nspec_inner_acoustic, &foa);
      if (iphase != 1) continue;
      if (run->STACEY_ABSORBING_CONDITIONS) compute_stacey_acoustic_undoatt_cuda_(r.Mesh_pointer, &iphase, &foa);
      if (foa == 1 && r.nrec_local > 0 && it < NSTEP) {
        int itf = it;
        add_sources_ac_sim_2_or_3_cuda_(r.Mesh_pointer, r.source_adjoint, &run->nrec, &r.nrec_local, &NSTEP, &itf);
      }
      if (foa == 3 && run->NSOURCES > 0) {
        int ns = run->NSOURCES;
        compute_add_sources_ac_s3_cuda_(r.Mesh_pointer, &ns, const_cast<double*>(run->stf_acoustic + (size_t)(NSTEP - it) * ns), run_number.data());
      }
    }
    if (run->NPROC > 1) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (!r.acoustic) continue;
        if (iphase == 1) transfer_boun_pot_from_device_(r.Mesh_pointer, foa == 3 ? r.b_buffer_send_scalar : r.buffer_send_scalar, &foa);
        else transfer_asmbl_pot_to_device_(r.Mesh_pointer, foa == 3 ? r.b_buffer_recv_scalar : r.buffer_recv_scalar, &foa);
      }
      if (iphase == 1 && run->host_exchange) {
        if (foa == 3) exchange_on_host_b(run, 1);
        else exchange_on_host(run, 1);
      }
    }
  }
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (!r.acoustic) continue;
    kernel_3_acoustic_cuda_(r.Mesh_pointer, &run->deltatover2, &run->b_deltatover2, &foa);
    acoustic_enforce_free_surf_cuda_(r.Mesh_pointer, &sifs, &foa);
  }
}

// foa = 3: compute_forces_viscoelastic_backward_calling (it = it_tmp); 1: compute_forces_viscoelastic_calling on the adjoint wavefield
void elastic_one_field_undoatt(Run* run, int it, int foa) {
  int strain = 1, att = run->ATTENUATION ? 1 : 0, zero = 0, NSTEP = run->NSTEP;
  realw dt_forces = foa == 3 ? run->b_deltat : run->deltat;
  for (int iphase = 1; iphase <= 2; ++iphase) {
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (!r.elastic) continue;
      compute_forces_viscoelastic_cuda_(r.Mesh_pointer, &iphase, &dt_forces, &r.nspec_outer_elastic, &r.nspec_inner_elastic, &strain, &att, &foa);
      if (iphase != 1) continue;
      if (run->STACEY_ABSORBING_CONDITIONS) compute_stacey_viscoelastic_undoatt_cuda_(r.Mesh_pointer, &iphase, &foa);
      if (foa == 1 && r.nrec_local > 0 && it < NSTEP) {
        int itf = it;
        add_sources_el_sim_type_2_or_3_(r.Mesh_pointer, r.source_adjoint, &run->nrec, &r.nrec_local, &NSTEP, &itf);
      }
      if (run->NSOURCES > 0) {  // foa 1 too: compute_add_sources_viscoelastic_GPU adds the forward source into b_accel (unused here)
        int ns = run->NSOURCES;
        compute_add_sources_el_s3_cuda_(r.Mesh_pointer, const_cast<double*>(run->stf_elastic + (size_t)(NSTEP - it) * ns), &ns);
      }
    }
    exchange_elastic_undoatt(run, iphase, foa);
  }
  for (int ir = 0; ir < run->nranks; ++ir) {
    Rank& r = run->ranks[ir];
    if (r.elastic) kernel_3_a_cuda_(r.Mesh_pointer, &run->deltatover2, &run->b_deltatover2, &zero, &foa);
  }
}

}  // namespace

extern "C" void SB200_FC(specfem_b200_iterate_time_undoatt)(specfem_b200_host_run* run) {
  if (!run || run->nranks <= 0 || !run->ranks) host_error("specfem_b200_iterate_time_undoatt: no ranks");
  if ((run->ELASTIC_SIMULATION && run->ACOUSTIC_SIMULATION) || run->PML_CONDITIONS || run->APPROXIMATE_HESS_KL)
    host_error("specfem_b200_iterate_time_undoatt: elastic or acoustic domains (not coupled), no C-PML, no approximate Hessian");
  const int NSTEP = run->NSTEP;
  const int nt_dump = run->NT_DUMP_ATTENUATION < NSTEP ? run->NT_DUMP_ATTENUATION : NSTEP;
  if (nt_dump <= 0) host_error("specfem_b200_iterate_time_undoatt: NT_DUMP_ATTENUATION must be positive");
  const int nsubset = (NSTEP + nt_dump - 1) / nt_dump;
  const int every = run->NTSTEP_BETWEEN_OUTPUT_SAMPLE > 0 ? run->NTSTEP_BETWEEN_OUTPUT_SAMPLE : 1;
  int one = 1, three = 3, zero = 0, it_global = 0;
  std::vector<std::vector<realw>> buf_displ(run->nranks), buf_pot(run->nranks), buf_pdd(run->nranks);
  for (int iteration_on_subset = 1; iteration_on_subset <= nsubset; ++iteration_on_subset) {
    // read_forward_arrays_undoatt: frame NSUBSET_ITERATIONS - iteration_on_subset + 1 -> b_ fields
    const int frame = nsubset - iteration_on_subset;
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      int n = r.NGLOB_AB, n3 = 3 * r.NGLOB_AB;
      if (r.acoustic)
        transfer_b_fields_ac_to_device_(&n, r.frame_potential[frame], r.frame_potential_dot[frame], r.frame_potential_dot_dot[frame], r.Mesh_pointer);
      if (!r.elastic) continue;
      transfer_b_fields_to_device_(&n3, r.frame_displ[frame], r.frame_veloc[frame], r.frame_accel[frame], r.Mesh_pointer);
      if (run->ATTENUATION) {
        int size_R = 3 * 125 * r.NSPEC_AB;  // N_SLS * NGLL3 * NSPEC
        realw** R = r.frame_R + 6 * (size_t)frame;
        transfer_b_rmemory_to_device_(r.Mesh_pointer, R[0], R[1], R[2], R[3], R[4], R[5], &size_R);
      }
    }
    for (int ir = 0; ir < run->nranks; ++ir) {  // compute_strain_att_backward
      Rank& r = run->ranks[ir];
      if (r.elastic) compute_strain_cuda_(r.Mesh_pointer, &three);
    }
    const int it_temp = it_global;
    std::vector<int> seismo_temp(run->nranks);
    for (int ir = 0; ir < run->nranks; ++ir) seismo_temp[ir] = run->ranks[ir].seismo_current;
    // forward subsets are full except the LAST one, which the kernel run takes first
    const int it_subset_end = iteration_on_subset == 1 ? NSTEP - (nsubset - 1) * nt_dump : nt_dump;
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (r.elastic) buf_displ[ir].assign((size_t)it_subset_end * 3 * r.NGLOB_AB, 0.f);
      if (r.acoustic) {
        buf_pot[ir].assign((size_t)it_subset_end * r.NGLOB_AB, 0.f);
        buf_pdd[ir].assign((size_t)it_subset_end * r.NGLOB_AB, 0.f);
      }
    }
    // ---- the forward wavefield over this subset, forward in time ----
    for (int its = 1; its <= it_subset_end; ++its) {
      it_global += 1;
      for (int ir = 0; ir < run->nranks; ++ir) {  // update_displ_Newmark_backward: fluid first
        Rank& r = run->ranks[ir];
        if (r.acoustic) update_displacement_ac_cuda_(r.Mesh_pointer, &run->b_deltat, &run->b_deltatover2, &run->b_deltatsqover2, &three);
      }
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic) update_displacement_cuda_(r.Mesh_pointer, &run->b_deltat, &run->b_deltatover2, &run->b_deltatsqover2, &three);
      }
      const int it_tmp = NSTEP - (nsubset - iteration_on_subset) * nt_dump - its + 1;
      if (run->ACOUSTIC_SIMULATION) acoustic_one_field_undoatt(run, it_tmp, 3);
      if (run->ELASTIC_SIMULATION) elastic_one_field_undoatt(run, it_tmp, 3);
      for (int ir = 0; ir < run->nranks; ++ir) {  // b_displ_elastic_buffer(:,:,it_of_this_subset) = b_displ
        Rank& r = run->ranks[ir];
        int n = r.NGLOB_AB, n3 = 3 * r.NGLOB_AB;
        if (r.acoustic) {
          transfer_b_potential_ac_from_device_(&n, buf_pot[ir].data() + (size_t)(its - 1) * n, r.Mesh_pointer);
          transfer_b_potential_dot_dot_ac_from_device_(&n, buf_pdd[ir].data() + (size_t)(its - 1) * n, r.Mesh_pointer);
        }
        if (r.elastic) transfer_b_displ_from_device_(&n3, buf_displ[ir].data() + (size_t)(its - 1) * n3, r.Mesh_pointer);
      }
    }
    it_global = it_temp;
    for (int ir = 0; ir < run->nranks; ++ir) run->ranks[ir].seismo_current = seismo_temp[ir];
    for (int ir = 0; ir < run->nranks; ++ir) {  // compute_strain_att
      Rank& r = run->ranks[ir];
      if (r.elastic) compute_strain_cuda_(r.Mesh_pointer, &one);
    }
    // ---- the adjoint wavefield over the same steps; buffered frames in reverse order ----
    for (int its = 1; its <= it_subset_end; ++its) {
      it_global += 1;
      const int it = it_global;
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        int n = r.NGLOB_AB, n3 = 3 * r.NGLOB_AB;
        if (r.acoustic) {
          transfer_b_potential_ac_to_device_(&n, buf_pot[ir].data() + (size_t)(it_subset_end - its) * n, r.Mesh_pointer);
          transfer_b_potential_dot_dot_ac_to_device_(&n, buf_pdd[ir].data() + (size_t)(it_subset_end - its) * n, r.Mesh_pointer);
        }
        if (r.elastic) transfer_b_displ_to_device_(&n3, buf_displ[ir].data() + (size_t)(it_subset_end - its) * n3, r.Mesh_pointer);
      }
      for (int ir = 0; ir < run->nranks; ++ir) {  // update_displ_Newmark
        Rank& r = run->ranks[ir];
        if (r.acoustic) update_displacement_ac_cuda_(r.Mesh_pointer, &run->deltat, &run->deltatover2, &run->deltatsqover2, &one);
      }
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.elastic) update_displacement_cuda_(r.Mesh_pointer, &run->deltat, &run->deltatover2, &run->deltatsqover2, &one);
      }
      if (run->ACOUSTIC_SIMULATION) acoustic_one_field_undoatt(run, it, 1);
      if (run->ELASTIC_SIMULATION) elastic_one_field_undoatt(run, it, 1);
      if (run->has_receivers && (it - 1) % every == 0) {
        for (int ir = 0; ir < run->nranks; ++ir) {
          Rank& r = run->ranks[ir];
          r.seismo_current += 1;
          if (r.nrec_local == 0) continue;
          int itf = it, nl = run->nlength_seismogram;
          compute_seismograms_cuda_(r.Mesh_pointer, r.seismograms_d, r.seismograms_v, r.seismograms_a, r.seismograms_p, &r.seismo_current, &nl,
                                    &itf, &run->NSTEP, &run->ACOUSTIC_SIMULATION, &run->ELASTIC_SIMULATION, &zero);
        }
      }
      for (int ir = 0; ir < run->nranks; ++ir) {  // compute_kernels: the library re-computes the strain of b_displ
        Rank& r = run->ranks[ir];
        if (r.elastic) compute_kernels_elastic_cuda_(r.Mesh_pointer, &run->deltat);
      }
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        if (r.acoustic) compute_kernels_acoustic_cuda_(r.Mesh_pointer, &run->deltat);
      }
    }
  }
  if (it_global != NSTEP) host_error("specfem_b200_iterate_time_undoatt: the subsets do not add up to NSTEP");
}

extern "C" void SB200_FC(specfem_b200_iterate_time_kernel_run)(specfem_b200_host_run* run, int* it_begin, int* it_end) {
  if (!run || run->nranks <= 0 || !run->ranks) host_error("specfem_b200_iterate_time_kernel_run: no ranks");
  if (*it_begin < 1 || *it_end > run->NSTEP) host_error("specfem_b200_iterate_time_kernel_run: time steps outside 1 .. NSTEP");
  if (run->ATTENUATION || run->PML_CONDITIONS) host_error("specfem_b200_iterate_time_kernel_run: no attenuation / C-PML (iterate_time_undoatt)");
  int one = 1, three = 3, zero = 0;
  const int every = run->NTSTEP_BETWEEN_OUTPUT_SAMPLE > 0 ? run->NTSTEP_BETWEEN_OUTPUT_SAMPLE : 1;
  for (int it = *it_begin; it <= *it_end; ++it) {
    for (int ir = 0; ir < run->nranks; ++ir) {  // update_displ_Newmark
      Rank& r = run->ranks[ir];
      if (r.acoustic) update_displacement_ac_cuda_(r.Mesh_pointer, &run->deltat, &run->deltatover2, &run->deltatsqover2, &one);
      if (r.elastic) update_displacement_cuda_(r.Mesh_pointer, &run->deltat, &run->deltatover2, &run->deltatsqover2, &one);
    }
    for (int ir = 0; ir < run->nranks; ++ir) {  // update_displ_Newmark_backward
      Rank& r = run->ranks[ir];
      if (r.acoustic) update_displacement_ac_cuda_(r.Mesh_pointer, &run->b_deltat, &run->b_deltatover2, &run->b_deltatsqover2, &three);
      if (r.elastic) update_displacement_cuda_(r.Mesh_pointer, &run->b_deltat, &run->b_deltatover2, &run->b_deltatsqover2, &three);
    }
    if (run->ELASTIC_SIMULATION && run->ACOUSTIC_SIMULATION) {
      // coupled: adjoint wavefield elastic first, then acoustic; reconstructed wavefield acoustic first, then elastic
      elastic_step_kl(run, it, 1);
      acoustic_step_kl(run, it, 1);
      acoustic_step_kl(run, it, 3);
      elastic_step_kl(run, it, 3);
    } else {
      if (run->ACOUSTIC_SIMULATION) acoustic_step_kl(run, it, 0);
      if (run->ELASTIC_SIMULATION) elastic_step_kl(run, it, 0);
    }
    if (it == 1) {  // read_forward_arrays: the last frame of the forward run replaces the b_ fields after the first step
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        int n = r.NGLOB_AB, n3 = 3 * r.NGLOB_AB;
        if (r.acoustic) transfer_b_fields_ac_to_device_(&n, r.last_potential, r.last_potential_dot, r.last_potential_dot_dot, r.Mesh_pointer);
        if (r.elastic) transfer_b_fields_to_device_(&n3, r.last_displ, r.last_veloc, r.last_accel, r.Mesh_pointer);
      }
    }
    if (run->has_receivers && (it - 1) % every == 0) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        r.seismo_current += 1;
        if (r.nrec_local == 0) continue;
        int itf = it, nl = run->nlength_seismogram;
        compute_seismograms_cuda_(r.Mesh_pointer, r.seismograms_d, r.seismograms_v, r.seismograms_a, r.seismograms_p, &r.seismo_current, &nl,
                                  &itf, &run->NSTEP, &run->ACOUSTIC_SIMULATION, &run->ELASTIC_SIMULATION, &zero);
      }
    }
    for (int ir = 0; ir < run->nranks; ++ir) {  // compute_kernels
      Rank& r = run->ranks[ir];
      if (r.elastic) compute_kernels_elastic_cuda_(r.Mesh_pointer, &run->deltat);
      if (r.acoustic) compute_kernels_acoustic_cuda_(r.Mesh_pointer, &run->deltat);
      if (run->APPROXIMATE_HESS_KL) compute_kernels_hess_cuda_(r.Mesh_pointer, &run->deltat, &r.elastic, &r.acoustic);
    }
  }
}

extern "C" void SB200_FC(specfem_b200_iterate_time)(specfem_b200_host_run* run, int* it_begin, int* it_end) {
  if (!run || run->nranks <= 0 || !run->ranks) host_error("specfem_b200_iterate_time: no ranks");
  if (*it_begin < 1 || *it_end > run->NSTEP) host_error("specfem_b200_iterate_time: time steps outside 1 .. NSTEP");
  if (run->NSOURCES > 0 && ((run->ELASTIC_SIMULATION && !run->stf_elastic) || (run->ACOUSTIC_SIMULATION && !run->stf_acoustic)))
    host_error("specfem_b200_iterate_time: source time function table missing");
  int fwd = 1;
  const int every = run->NTSTEP_BETWEEN_OUTPUT_SAMPLE > 0 ? run->NTSTEP_BETWEEN_OUTPUT_SAMPLE : 1;
  for (int it = *it_begin; it <= *it_end; ++it) {
    // update_displ_Newmark: acoustic first, then elastic
    for (int ir = 0; ir < run->nranks; ++ir) {
      Rank& r = run->ranks[ir];
      if (r.acoustic) update_displacement_ac_cuda_(r.Mesh_pointer, &run->deltat, &run->deltatover2, &run->deltatsqover2, &fwd);
      if (r.elastic) update_displacement_cuda_(r.Mesh_pointer, &run->deltat, &run->deltatover2, &run->deltatsqover2, &fwd);
    }
    if (run->ACOUSTIC_SIMULATION) acoustic_step(run, it);
    if (run->ELASTIC_SIMULATION) elastic_step(run, it);
    if (run->has_receivers && (it - 1) % every == 0) {
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        r.seismo_current += 1;
        if (r.nrec_local == 0) continue;
        int itf = it, nl = run->nlength_seismogram;
        const bool flush = run->seismo_flush_every > 0 && (it - *it_begin + 1) % run->seismo_flush_every == 0;
        int it_last = flush ? it : run->NSTEP;
        compute_seismograms_cuda_(r.Mesh_pointer, r.seismograms_d, r.seismograms_v, r.seismograms_a, r.seismograms_p, &r.seismo_current, &nl,
                                  &itf, &it_last, &run->ACOUSTIC_SIMULATION, &run->ELASTIC_SIMULATION,
                                  &run->USE_TRICK_FOR_BETTER_PRESSURE);
      }
    }
    if (run->norm_every > 0 && run->norms && (it - *it_begin + 1) % run->norm_every == 0) {  // check_stability
      realw mx = 0.f;
      for (int ir = 0; ir < run->nranks; ++ir) {
        Rank& r = run->ranks[ir];
        realw v = 0.f;
        int one = 1;
        if (r.elastic) get_norm_elastic_from_device_(&v, r.Mesh_pointer, &one);
        else if (r.acoustic) get_norm_acoustic_from_device_(&v, r.Mesh_pointer, &one);
        mx = v > mx ? v : mx;
        if (!std::isfinite(v)) host_error("the wavefield became non-finite (check_stability): the simulation blew up");
      }
      run->norms[(it - *it_begin + 1) / run->norm_every - 1] = mx;
    }
  }
}
