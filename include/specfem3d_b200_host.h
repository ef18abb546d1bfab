/* specfem3d_b200_host.h -- a native (C++) host side above the C ABI of specfem3d_b200.h.
 *
 * The reference's host is compiled Fortran: iterate_time (src/specfem3D/iterate_time.F90:219-330) calls the device entry
 * points once per time step, in the order of update_displ_Newmark (update_displacement_scheme.f90:281-312),
 * compute_forces_acoustic_calling (compute_forces_acoustic_calling_routine.f90:28-330), compute_forces_viscoelastic_calling
 * (compute_forces_viscoelastic_calling_routine.F90:30-435) and write_seismograms (write_seismograms.f90:28-200).  No Fortran
 * compiler exists in this image, so the same control flow is restated here in C++ (csrc/host_loop.cu): it is a CALLER of the ABI
 * -- nothing in it touches the device except through the entry points of specfem3d_b200.h -- and stands where xspecfem3D's
 * time loop would stand.  specfem3d_b200/time_loop.py is the same loop in Python (kernel runs, SAVE_FORWARD and the tests'
 * call-by-call checks use that one); tests/test_gpu_parity.py::test_native_host_loop_equals_python_loop runs both.
 *
 * Plain C: pointers and sizes only.  All arrays are the caller's (Fortran layouts, as in specfem3d_b200.h).
 */
#ifndef SPECFEM3D_B200_HOST_H
#define SPECFEM3D_B200_HOST_H

#include "specfem3d_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one MPI rank of the solver = one Mesh_pointer handle (the ranks of a run may share a GPU or sit on several) */
typedef struct {
  long* Mesh_pointer;
  int myrank;
  int elastic, acoustic;                 /* the rank owns elements of the domain (ELASTIC / ACOUSTIC of the rank) */
  int nspec_outer_elastic, nspec_inner_elastic, nspec_outer_acoustic, nspec_inner_acoustic;
  int num_coupling_ac_el_faces;
  /* halo (assemble_MPI_vector.f90 / assemble_MPI_scalar.f90): interface tables and the rank's send / receive buffers */
  int num_interfaces_ext_mesh, max_nibool_interfaces_ext_mesh;
  const int* nibool_interfaces_ext_mesh; /* (num_interfaces) */
  const int* ibool_interfaces_ext_mesh;  /* (max_nibool, num_interfaces) */
  const int* my_neighbors_ext_mesh;      /* (num_interfaces) neighbour rank of every interface */
  realw *buffer_send_vector, *buffer_recv_vector; /* (NDIM, max_nibool, num_interfaces) */
  realw *buffer_send_scalar, *buffer_recv_scalar; /* (max_nibool, num_interfaces) */
  /* Stacey record buffers: one record each, (NDIM, NGLLSQUARE, num_abs_boundary_faces) / (NGLLSQUARE, num_abs_boundary_faces) */
  realw *b_absorb_field, *b_absorb_potential;
  /* seismograms (NDIM, nrec_local, nlength_seismogram) and the running sample counter of the rank */
  int nrec_local;
  realw *seismograms_d, *seismograms_v, *seismograms_a, *seismograms_p;
  int seismo_current;
  /* SAVE_FORWARD (run->SAVE_FORWARD != 0): what the Fortran solver writes to proc*_absorb_field.bin / proc*_absorb_potential.bin
     with write_abs(.., it) (compute_stacey_viscoelastic.F90:190-200, compute_stacey_acoustic.f90): record it - 1 of these
     caller-owned arrays, NSTEP records of 3 * 25 * num_abs_boundary_faces / 25 * num_abs_boundary_faces floats; NULL = not kept */
  int num_abs_boundary_faces;
  realw *absorb_field_history, *absorb_potential_history;
  /* kernel runs (SIMULATION_TYPE 3, specfem_b200_iterate_time_kernel_run): the *_history arrays above are READ (read_abs at
     record NSTEP - it); buffers of the reconstructed wavefield's halo; the adjoint sources of the rank's receivers
     source_adjoint(NDIM, nadj_rec_local, NSTEP) in forward time order; the last frame of the forward run
     (proc*_save_forward_arrays.bin, read_forward_arrays: it replaces the b_ fields after the first step) */
  int NGLOB_AB;
  realw *b_buffer_send_vector, *b_buffer_recv_vector, *b_buffer_send_scalar, *b_buffer_recv_scalar;
  realw* source_adjoint;
  realw *last_displ, *last_veloc, *last_accel, *last_potential, *last_potential_dot, *last_potential_dot_dot;
  /* undo-attenuation kernel runs (specfem_b200_iterate_time_undoatt): what proc*_save_frame_at*.bin holds at the start of every
     forward subset s = 0 .. NSUBSET_ITERATIONS - 1 -- frame_displ[s], frame_veloc[s], frame_accel[s] (NDIM, NGLOB), with ATTENUATION
     frame_R[6 s + {0..5}] = R_xx, R_yy, R_xy, R_xz, R_yz, R_trace (N_SLS, 125, NSPEC); fluid ranks: the three potentials */
  int NSPEC_AB;
  realw **frame_displ, **frame_veloc, **frame_accel, **frame_R;
  realw **frame_potential, **frame_potential_dot, **frame_potential_dot_dot;
} specfem_b200_host_rank;

typedef struct {
  int nranks;                            /* ranks driven by this process */
  specfem_b200_host_rank* ranks;
  int NPROC;                             /* ranks of the whole run (> 1: the halo entry points are called) */
  realw deltat, deltatover2, deltatsqover2;
  int NSTEP;
  int ELASTIC_SIMULATION, ACOUSTIC_SIMULATION, ATTENUATION, STACEY_ABSORBING_CONDITIONS, STACEY_INSTEAD_OF_FREE_SURFACE,
      PML_CONDITIONS, APPROXIMATE_OCEAN_LOAD;
  /* sources: stf_pre_compute of EVERY step, (NSOURCES, NSTEP) doubles, elastic and acoustic tables
     (compute_add_sources_viscoelastic.F90: get_stf_viscoelastic; compute_add_sources_acoustic.f90) */
  int NSOURCES;                          /* 0 = no source */
  const double *stf_elastic, *stf_acoustic;
  /* seismograms */
  int has_receivers, NTSTEP_BETWEEN_OUTPUT_SAMPLE, nlength_seismogram;
  /* pressure seismograms from -potential (the caller's stf_acoustic table then holds the SECOND DERIVATIVE of the source time
     function, compute_add_sources_acoustic.f90:736-750); handed to compute_seismograms_cuda like write_seismograms.f90:80 does */
  int USE_TRICK_FOR_BETTER_PRESSURE;
  /* 0: the seismograms come back when the device buffer is full or at NSTEP, like in the reference (write_seismograms.f90:
     NTSTEP_BETWEEN_OUTPUT_SEISMOS); n > 0: additionally after every n-th step of the call (compute_seismograms_cuda is told
     that the step is the last one, so it hands back -- blocking -- everything recorded so far): real-time monitoring */
  int seismo_flush_every;
  /* SIMULATION_TYPE 1 with SAVE_FORWARD: the absorbed Stacey tractions / fluxes of every step are kept (rank: *_history) */
  int SAVE_FORWARD;
  /* kernel runs: b_deltat = -deltat and its halves (prepare_timerun.F90:425-430), number of receivers of the whole run, and
     APPROXIMATE_HESS_KL */
  realw b_deltat, b_deltatover2, b_deltatsqover2;
  int nrec, APPROXIMATE_HESS_KL;
  /* undo-attenuation runs: length of a forward subset (NT_DUMP_ATTENUATION of the Par_file, clipped to NSTEP) */
  int NT_DUMP_ATTENUATION;
  /* forward runs of UNDO_ATTENUATION_AND_OR_PML keep no Stacey record: specfem_b200_iterate_time calls
     compute_stacey_*_undoatt_cuda instead (compute_stacey_viscoelastic.F90:453-457, compute_stacey_acoustic.f90:369-372) */
  int UNDO_ATTENUATION_AND_OR_PML;
  /* halo transport between the ranks of THIS process: 1 = copy send -> receive buffers on the host (what MPI does between
     processes), 0 = none (single rank, or the library moves the halo itself after specfem_b200_halo_p2p_connect) */
  int host_exchange;
  /* check_stability: after every norm_every-th step of the call (0 = never) the maximum displacement norm (potential_dot_dot
     norm of purely acoustic ranks) over the ranks goes to norms[0], norms[1], ...; a non-finite value stops the run like
     check_stability.f90:120 does */
  int norm_every;
  realw* norms;
} specfem_b200_host_run;

/* Time steps it_begin .. it_end (1-based, inclusive) of a forward run (SIMULATION_TYPE 1, with or without SAVE_FORWARD).
   Replaces the loop body of iterate_time.F90:219-330 in GPU_MODE. */
void SB200_FC(specfem_b200_iterate_time)(specfem_b200_host_run* run, int* it_begin, int* it_end);

/* The same for a kernel run (SIMULATION_TYPE 3; elastic, acoustic or coupled domains, no attenuation / C-PML): adjoint wavefield
   in the forward arrays driven by the adjoint sources, forward wavefield reconstructed backwards in time in the b_ arrays from
   the last frame and the Stacey records, kernels accumulated every step.  Replaces iterate_time.F90:233-300 with
   compute_forces_viscoelastic_GPU_calling / compute_forces_acoustic_GPU_calling (FORWARD_OR_ADJOINT 0) or, in coupled runs, the
   separate forward / backward routines (compute_forces_viscoelastic_calling_routine.F90:61-67). */
void SB200_FC(specfem_b200_iterate_time_kernel_run)(specfem_b200_host_run* run, int* it_begin, int* it_end);

/* The whole kernel run of UNDO_ATTENUATION_AND_OR_PML (iterate_time_undoatt.F90:63-470; elastic / viscoelastic OR acoustic domains,
   no C-PML, no approximate Hessian): the forward subsets are taken in reverse order; for each one the saved frame is restored into the
   b_ fields, the forward wavefield is re-computed FORWARD in time (b_deltat = deltat) and b_displ (fluid: b_potential,
   b_potential_dot_dot) of every step is buffered on the host, then the adjoint wavefield advances over the same steps while the
   buffered frames are handed back in reverse order and the kernels accumulate.  Always all NSTEP steps. */
void SB200_FC(specfem_b200_iterate_time_undoatt)(specfem_b200_host_run* run);

#ifdef __cplusplus
}
#endif
#endif
