!> fen_gpu_mod -- ISO_C_BINDING layer between FEN's Fortran drivers and libfen_gpu.so.
!>
!> SOURCE ONLY: the build image has no Fortran compiler (gfortran / flang / nvfortran / mpif90 are
!> all absent), so this file is not compiled or run by the test-suite; tests/test_fortran_shim.py checks it
!> textually against include/fen_gpu.h (every function bound, argument counts, VALUE attributes, the
!> bind(C) types member by member, the enum values).  It is kept mechanical: part 1
!> declares the bind(C) interfaces of include/fen_gpu.h one to one, part 2 wraps them in procedures
!> that carry the reference's own names and argument lists, so a driver program switches over by
!> changing its `use` lines (INTEGRATION.md).  The same call sequences are exercised through the
!> C ABI by fen_b200/api.py in tests/.
!>
!> Reference interfaces mirrored (paths relative to the FEN repository):
!>   solver_mod        init_solver(comp_grid)                        src/solver.f90:34
!>                     advance_solution(comp_grid, step, dt)         src/solver.f90:13-19,28
!>                     print_solver_status(log_id, step, time, dt)   src/solver.f90:21-25,29
!>                     destroy_solver()                              src/solver.f90:333
!>   navier_stokes_mod navier_stokes_solver(comp_grid, step, dt)     src/navier_stokes.f90:50
!>                     set_timestep(comp_grid, dt, U)                src/navier_stokes.f90:623
!>                     module scalars density, viscosity, g, CFL ... src/navier_stokes.f90:18-45
!>   poisson_mod       init_poisson_solver(phi), solve_poisson(phi), destroy_poisson_solver(phi)
!>                                                                   src/poisson.f90:51-52,57,1456
!>   halo_mod          update_halos(f, G, l)                         src/halo.f90:12
!>   scalar / vector   update_ghost_nodes                            src/scalar.f90:223, src/vector.f90:82
!>   -DMF builds (2-D): volume_of_fluid_mod  allocate_vof_fields, get_vof_from_distance, get_h_from_vof,
!>                     advect_vof(v, dt), check_vof_integral         src/volume_of_fluid.f90:54,676,228,434,722
!>                     multiphase_mod  update_material_properties    src/multiphase.f90:121
!>                     init_solver with the MF wiring                src/solver.f90:87-98
module fen_gpu_mod

    use, intrinsic :: iso_c_binding
    use precision_mod, only : dp
    use grid_mod     , only : grid
    use scalar_mod   , only : scalar
    use vector_mod   , only : vector

    implicit none

    ! enum fen_field (include/fen_gpu.h)
    integer(c_int), parameter :: FEN_P = 0, FEN_PHI = 1, FEN_RHO = 2, FEN_MU = 3
    integer(c_int), parameter :: FEN_VX = 4, FEN_VY = 5, FEN_VZ = 6
    integer(c_int), parameter :: FEN_DVOX = 10, FEN_DVOY = 11, FEN_DVOZ = 12
    integer(c_int), parameter :: FEN_SX = 16, FEN_SY = 17, FEN_SZ = 18
    integer(c_int), parameter :: FEN_VOF = 19, FEN_H = 20, FEN_D = 21, FEN_CURV = 22
    integer(c_int), parameter :: FEN_NORMX = 23, FEN_NORMY = 24, FEN_LX = 26, FEN_LY = 27
    integer(c_int), parameter :: FEN_PHAT = 29, FEN_PO = 30

    type, bind(C) :: fen_grid_desc
        integer(c_int) :: nx, ny, nz
        integer(c_int) :: ndim
        real(c_double) :: delta
        integer(c_int) :: bc(6)
        integer(c_int) :: rank, nranks
        integer(c_int) :: device
    end type fen_grid_desc

    type, bind(C) :: fen_ns_params
        real(c_double) :: density, viscosity
        real(c_double) :: g(3)
        real(c_double) :: CFL
        real(c_double) :: dt_o
        real(c_double) :: dt_visc, dt_conv
        integer(c_int) :: constant_CFL
    end type fen_ns_params

    type, bind(C) :: fen_mf_params
        real(c_double) :: rho_0, rho_1, mu_0, mu_1
        real(c_double) :: sigma
        real(c_double) :: beta
        real(c_double) :: cut
        integer(c_int) :: quadratic
        integer(c_int) :: x_first
        real(c_double) :: dt_surf
        real(c_double) :: rhomin, irhomin
    end type fen_mf_params

    !---------------------------------------------------------------------------------------------
    ! Part 1: the C ABI, one interface per function of include/fen_gpu.h
    !---------------------------------------------------------------------------------------------
    interface
        function fen_gpu_last_error() bind(C, name='fen_gpu_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        function fen_gpu_create(desc, ctx) bind(C, name='fen_gpu_create') result(ierr)
            import :: c_int, c_ptr, fen_grid_desc
            type(fen_grid_desc), intent(in) :: desc
            type(c_ptr), intent(out) :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_destroy(ctx) bind(C, name='fen_gpu_destroy') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_comm_handle_bytes() bind(C, name='fen_gpu_comm_handle_bytes') result(n)
            import :: c_int
            integer(c_int) :: n
        end function
        function fen_gpu_comm_export(ctx, handle) bind(C, name='fen_gpu_comm_export') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, handle
            integer(c_int) :: ierr
        end function
        function fen_gpu_comm_connect(ctx, all_handles) bind(C, name='fen_gpu_comm_connect') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, all_handles
            integer(c_int) :: ierr
        end function
        function fen_gpu_push(ctx, field, host, gl) bind(C, name='fen_gpu_push') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, host
            integer(c_int), value :: field, gl
            integer(c_int) :: ierr
        end function
        function fen_gpu_pull(ctx, field, host, gl) bind(C, name='fen_gpu_pull') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, host
            integer(c_int), value :: field, gl
            integer(c_int) :: ierr
        end function
        function fen_gpu_scalar_allocate(ctx, gl, loc, field) bind(C, name='fen_gpu_scalar_allocate') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: gl, loc
            integer(c_int), intent(out) :: field
            integer(c_int) :: ierr
        end function
        function fen_gpu_scalar_destroy(ctx, field) bind(C, name='fen_gpu_scalar_destroy') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            integer(c_int) :: ierr
        end function
        function fen_gpu_set_bc_type(ctx, field, face, bctype) bind(C, name='fen_gpu_set_bc_type') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: field, face, bctype
            integer(c_int) :: ierr
        end function
        function fen_gpu_set_bc_plane(ctx, field, face, plane, uniform) bind(C, name='fen_gpu_set_bc_plane') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, plane
            integer(c_int), value :: field, face, uniform
            integer(c_int) :: ierr
        end function
        function fen_gpu_update_ghost_nodes(ctx, field, ncomp) bind(C, name='fen_gpu_update_ghost_nodes') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: field, ncomp
            integer(c_int) :: ierr
        end function
        function fen_gpu_update_halos(ctx, field) bind(C, name='fen_gpu_update_halos') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            integer(c_int) :: ierr
        end function
        function fen_gpu_init_poisson_solver(ctx) bind(C, name='fen_gpu_init_poisson_solver') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_solve_poisson(ctx, field) bind(C, name='fen_gpu_solve_poisson') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            integer(c_int) :: ierr
        end function
        function fen_gpu_destroy_poisson_solver(ctx) bind(C, name='fen_gpu_destroy_poisson_solver') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_init_solver(ctx) bind(C, name='fen_gpu_init_solver') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_destroy_solver(ctx) bind(C, name='fen_gpu_destroy_solver') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_get_params(ctx, p) bind(C, name='fen_gpu_get_params') result(ierr)
            import :: c_int, c_ptr, fen_ns_params
            type(c_ptr), value :: ctx
            type(fen_ns_params), intent(out) :: p
            integer(c_int) :: ierr
        end function
        function fen_gpu_set_params(ctx, p) bind(C, name='fen_gpu_set_params') result(ierr)
            import :: c_int, c_ptr, fen_ns_params
            type(c_ptr), value :: ctx
            type(fen_ns_params), intent(in) :: p
            integer(c_int) :: ierr
        end function
        function fen_gpu_set_timestep(ctx, U, dt) bind(C, name='fen_gpu_set_timestep') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), value :: U
            real(c_double), intent(out) :: dt
            integer(c_int) :: ierr
        end function
        function fen_gpu_navier_stokes_solver(ctx, step, dt) bind(C, name='fen_gpu_navier_stokes_solver') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: step
            real(c_double), intent(inout) :: dt
            integer(c_int) :: ierr
        end function
        function fen_gpu_get_status(ctx, maxdiv, maxCFL) bind(C, name='fen_gpu_get_status') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: maxdiv, maxCFL
            integer(c_int) :: ierr
        end function
        ! ---- two-phase path (-DMF) ----
        function fen_gpu_mf_get_params(ctx, p) bind(C, name='fen_gpu_mf_get_params') result(ierr)
            import :: c_int, c_ptr, fen_mf_params
            type(c_ptr), value :: ctx
            type(fen_mf_params), intent(out) :: p
            integer(c_int) :: ierr
        end function
        function fen_gpu_mf_set_params(ctx, p) bind(C, name='fen_gpu_mf_set_params') result(ierr)
            import :: c_int, c_ptr, fen_mf_params
            type(c_ptr), value :: ctx
            type(fen_mf_params), intent(in) :: p
            integer(c_int) :: ierr
        end function
        function fen_gpu_allocate_vof_fields(ctx) bind(C, name='fen_gpu_allocate_vof_fields') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_get_vof_from_distance(ctx, fn, user, x0, y0) &
                bind(C, name='fen_gpu_get_vof_from_distance') result(ierr)
            import :: c_int, c_ptr, c_funptr, c_double
            type(c_ptr), value :: ctx, user
            type(c_funptr), value :: fn
            real(c_double), value :: x0, y0
            integer(c_int) :: ierr
        end function
        function fen_gpu_get_h_from_vof(ctx) bind(C, name='fen_gpu_get_h_from_vof') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_advect_vof(ctx, vector_x, dt) bind(C, name='fen_gpu_advect_vof') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: vector_x
            real(c_double), value :: dt
            integer(c_int) :: ierr
        end function
        function fen_gpu_check_vof_integral(ctx, i1, i2) bind(C, name='fen_gpu_check_vof_integral') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: i1, i2
            integer(c_int) :: ierr
        end function
        function fen_gpu_update_material_properties(ctx) bind(C, name='fen_gpu_update_material_properties') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_init_solver_mf(ctx, fn, user, x0, y0) bind(C, name='fen_gpu_init_solver_mf') result(ierr)
            import :: c_int, c_ptr, c_funptr, c_double
            type(c_ptr), value :: ctx, user
            type(c_funptr), value :: fn
            real(c_double), value :: x0, y0
            integer(c_int) :: ierr
        end function
        function fen_gpu_destroy_vof(ctx) bind(C, name='fen_gpu_destroy_vof') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        ! ---- the rest of the header: containers, field operators, stage-level calls, IO, hook, measurement ----
        function fen_gpu_version() bind(C, name='fen_gpu_version') result(v)
            import :: c_int
            integer(c_int) :: v
        end function
        function fen_gpu_synchronize(ctx) bind(C, name='fen_gpu_synchronize') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_local_bounds(ctx, lo, hi) bind(C, name='fen_gpu_local_bounds') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), intent(out) :: lo(3), hi(3)
            integer(c_int) :: ierr
        end function
        function fen_gpu_pull_async(ctx, field, host, gl) bind(C, name='fen_gpu_pull_async') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx, host
            integer(c_int), value :: field, gl
            integer(c_int) :: ierr
        end function
        function fen_gpu_pull_wait(ctx) bind(C, name='fen_gpu_pull_wait') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_set_to_value(ctx, field, val) bind(C, name='fen_gpu_set_to_value') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            real(c_double), value :: val
            integer(c_int) :: ierr
        end function
        function fen_gpu_get_bc_type(ctx, field, face, bctype) bind(C, name='fen_gpu_get_bc_type') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: field, face
            integer(c_int), intent(out) :: bctype
            integer(c_int) :: ierr
        end function
        function fen_gpu_max_value(ctx, field, val) bind(C, name='fen_gpu_max_value') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            real(c_double), intent(out) :: val
            integer(c_int) :: ierr
        end function
        function fen_gpu_integral(ctx, field, val) bind(C, name='fen_gpu_integral') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            real(c_double), intent(out) :: val
            integer(c_int) :: ierr
        end function
        function fen_gpu_gradient(ctx, scalar_in, vector_out_x) bind(C, name='fen_gpu_gradient') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: scalar_in, vector_out_x
            integer(c_int) :: ierr
        end function
        function fen_gpu_divergence(ctx, vector_in_x, scalar_out) bind(C, name='fen_gpu_divergence') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: vector_in_x, scalar_out
            integer(c_int) :: ierr
        end function
        function fen_gpu_laplacian(ctx, vector_in_x, vector_out_x) bind(C, name='fen_gpu_laplacian') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: vector_in_x, vector_out_x
            integer(c_int) :: ierr
        end function
        function fen_gpu_center_to_face(ctx, scalar_in, vector_out_x) bind(C, name='fen_gpu_center_to_face') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: scalar_in, vector_out_x
            integer(c_int) :: ierr
        end function
        function fen_gpu_laplacian_scalar(ctx, scalar_in, scalar_out) bind(C, name='fen_gpu_laplacian_scalar') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: scalar_in, scalar_out
            integer(c_int) :: ierr
        end function
        function fen_gpu_face_to_center(ctx, scalar_face, scalar_center, dir) bind(C, name='fen_gpu_face_to_center') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: scalar_face, scalar_center, dir
            integer(c_int) :: ierr
        end function
        function fen_gpu_curl(ctx, vector_in_x, vector_out_x) bind(C, name='fen_gpu_curl') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: vector_in_x, vector_out_x
            integer(c_int) :: ierr
        end function
        function fen_gpu_poisson_variant(ctx) bind(C, name='fen_gpu_poisson_variant') result(msg)
            import :: c_ptr
            type(c_ptr), value :: ctx
            type(c_ptr) :: msg
        end function
        function fen_gpu_status_line(ctx, step, time, dt, buf, buflen) bind(C, name='fen_gpu_status_line') result(ierr)
            import :: c_int, c_ptr, c_double, c_char
            type(c_ptr), value :: ctx
            integer(c_int), value :: step, buflen
            real(c_double), value :: time, dt
            character(kind=c_char), intent(out) :: buf(*)
            integer(c_int) :: ierr
        end function
        function fen_gpu_add_advection(ctx, rhs_vector_x) bind(C, name='fen_gpu_add_advection') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: rhs_vector_x
            integer(c_int) :: ierr
        end function
        function fen_gpu_compute_explicit_terms(ctx, rhs_vector_x) bind(C, name='fen_gpu_compute_explicit_terms') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: rhs_vector_x
            integer(c_int) :: ierr
        end function
        function fen_gpu_predicted_velocity_field(ctx, dt) bind(C, name='fen_gpu_predicted_velocity_field') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), value :: dt
            integer(c_int) :: ierr
        end function
        function fen_gpu_correct_velocity_field(ctx, dt) bind(C, name='fen_gpu_correct_velocity_field') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), value :: dt
            integer(c_int) :: ierr
        end function
        function fen_gpu_update_pressure(ctx) bind(C, name='fen_gpu_update_pressure') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int) :: ierr
        end function
        function fen_gpu_checks(ctx, dt) bind(C, name='fen_gpu_checks') result(ierr)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), value :: dt
            integer(c_int) :: ierr
        end function
        function fen_gpu_scalar_write(ctx, field, filename) bind(C, name='fen_gpu_scalar_write') result(ierr)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            character(kind=c_char), intent(in) :: filename(*)
            integer(c_int) :: ierr
        end function
        function fen_gpu_scalar_read(ctx, field, filename) bind(C, name='fen_gpu_scalar_read') result(ierr)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            integer(c_int), value :: field
            character(kind=c_char), intent(in) :: filename(*)
            integer(c_int) :: ierr
        end function
        function fen_gpu_save_state(ctx, filename) bind(C, name='fen_gpu_save_state') result(ierr)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            character(kind=c_char), intent(in) :: filename(*)
            integer(c_int) :: ierr
        end function
        function fen_gpu_load_state(ctx, filename) bind(C, name='fen_gpu_load_state') result(ierr)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            character(kind=c_char), intent(in) :: filename(*)
            integer(c_int) :: ierr
        end function
        function fen_gpu_save_fields(ctx, step, dir) bind(C, name='fen_gpu_save_fields') result(ierr)
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: ctx
            integer(c_int), value :: step
            character(kind=c_char), intent(in) :: dir(*)
            integer(c_int) :: ierr
        end function
        function fen_gpu_set_forcing_hook(ctx, fn, user) bind(C, name='fen_gpu_set_forcing_hook') result(ierr)
            import :: c_int, c_ptr, c_funptr
            type(c_ptr), value :: ctx, user
            type(c_funptr), value :: fn
            integer(c_int) :: ierr
        end function
        function fen_gpu_profile_enable(ctx, on) bind(C, name='fen_gpu_profile_enable') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: on
            integer(c_int) :: ierr
        end function
        function fen_gpu_profile_read(ctx, max_entries, names, ms, launches, n_out) &
                bind(C, name='fen_gpu_profile_read') result(ierr)
            import :: c_int, c_ptr, c_double, c_char
            type(c_ptr), value :: ctx
            integer(c_int), value :: max_entries
            character(kind=c_char), intent(out) :: names(32, *)
            real(c_double), intent(out) :: ms(*)
            integer(c_int), intent(out) :: launches(*)
            integer(c_int), intent(out) :: n_out
            integer(c_int) :: ierr
        end function
        function fen_gpu_launch_count(ctx) bind(C, name='fen_gpu_launch_count') result(n)
            import :: c_long_long, c_ptr
            type(c_ptr), value :: ctx
            integer(c_long_long) :: n
        end function
        function fen_gpu_stream(ctx) bind(C, name='fen_gpu_stream') result(s)
            import :: c_ptr
            type(c_ptr), value :: ctx
            type(c_ptr) :: s
        end function
    end interface

    !---------------------------------------------------------------------------------------------
    ! Part 2: the reference's names.  Module state mirrors navier_stokes_mod's public scalars; the
    ! fields p, v, ... stay the reference's own module-global host arrays (navier_stokes.f90:36-38).
    !---------------------------------------------------------------------------------------------
    type(c_ptr), save :: ctx = c_null_ptr        !< one solver instance per process, as in the reference
    integer, save     :: gpu_ndim = 3

    ! same procedure-pointer interface as solver_mod (src/solver.f90:13-29)
    abstract interface
        subroutine advance_solver(comp_grid, step, dt)
            import :: dp, grid
            type(grid), intent(in   ) :: comp_grid
            integer   , intent(in   ) :: step
            real(dp)  , intent(inout) :: dt
        end subroutine advance_solver
    end interface
    procedure(advance_solver), pointer :: advance_solution => Null()

contains

    !> prints like IO.f90:12 print_error_message and continues, as the reference does
    subroutine gpu_check(ierr, where)
        use IO_mod, only : print_error_message
        integer(c_int), intent(in) :: ierr
        character(*)  , intent(in) :: where
        character(kind=c_char), pointer :: cmsg(:)
        character(len=512) :: msg
        integer :: n
        if (ierr == 0) return
        call c_f_pointer(fen_gpu_last_error(), cmsg, [512])
        msg = ' '
        do n = 1, 512
            if (cmsg(n) == c_null_char) exit
            msg(n:n) = cmsg(n)
        end do
        call print_error_message('ERROR: '//where//': '//trim(msg))
        ! the one hard stop of the reference: unsupported Poisson BC combination (poisson.f90:91-95)
        if (ierr == 3 .and. index(where, 'init') > 0) stop
    end subroutine gpu_check

    integer(c_int) function bc_code(s)
        character(*), intent(in) :: s
        select case (trim(s))
        case ('Periodic'); bc_code = 0
        case ('Wall');     bc_code = 1
        case ('Inflow');   bc_code = 2
        case ('Outflow');  bc_code = 3
        case default;      bc_code = -1
        end select
    end function bc_code

    !> grid%setup has already run (src/grid.f90:67): hand its result to the device library.
    !> G%prow must be 1 (slabs); G%rank / G%nranks select the z slab and the GPU.
    subroutine gpu_attach_grid(G, device)
        type(grid), intent(in) :: G
        integer   , intent(in), optional :: device
        type(fen_grid_desc) :: d
        integer :: n
        d%nx = G%Nx; d%ny = G%Ny; d%nz = G%Nz
#if DIM==3
        d%ndim = 3
#else
        d%ndim = 2
#endif
        gpu_ndim = d%ndim
        d%delta = G%delta                                   ! = Lx/float(Nx), grid.f90:140
        d%bc = 0
        do n = 1, 2*d%ndim
            d%bc(n) = bc_code(G%boundary_conditions(n))
        end do
        d%rank = G%rank; d%nranks = G%nranks
        d%device = -1
        if (present(device)) d%device = device
        call gpu_check(fen_gpu_create(d, ctx), 'grid setup')
    end subroutine gpu_attach_grid

    !> multi-GPU wiring (replaces decomp_2d_init's communicator set-up, grid.f90:125): all-gather of
    !> the exported handles with the MPI the driver already has.
    subroutine gpu_connect(comm)
        use mpi
        integer, intent(in) :: comm
        integer :: nb, np, ierr
        character(kind=c_char), allocatable, target :: mine(:), everyone(:)
        call mpi_comm_size(comm, np, ierr)
        if (np == 1) return
        nb = fen_gpu_comm_handle_bytes()
        allocate(mine(nb), everyone(nb*np))
        call gpu_check(fen_gpu_comm_export(ctx, c_loc(mine)), 'comm export')
        call mpi_allgather(mine, nb, mpi_byte, everyone, nb, mpi_byte, comm, ierr)
        call gpu_check(fen_gpu_comm_connect(ctx, c_loc(everyone)), 'comm connect')
    end subroutine gpu_connect

    !> host -> device / device -> host for one reference scalar (the explicit transfer points)
    subroutine gpu_push(s, field)
        type(scalar), intent(in), target :: s
        integer(c_int), intent(in) :: field
        call gpu_check(fen_gpu_push(ctx, field, c_loc(s%f), int(s%gl, c_int)), 'push')
    end subroutine gpu_push

    subroutine gpu_pull(s, field)
        type(scalar), intent(inout), target :: s
        integer(c_int), intent(in) :: field
        call gpu_check(fen_gpu_pull(ctx, field, c_loc(s%f), int(s%gl, c_int)), 'pull')
    end subroutine gpu_pull

    !> bc%type_<face> and bc%<face> planes of a reference scalar -> device (call after the driver has
    !> set them, e.g. v%x%bc%top = U in lid_driven.f90:59)
    subroutine gpu_push_bc(s, field)
        type(scalar), intent(in), target :: s
        integer(c_int), intent(in) :: field
        call gpu_check(fen_gpu_set_bc_type(ctx, field, 0_c_int, int(s%bc%type_left  , c_int)), 'bc')
        call gpu_check(fen_gpu_set_bc_type(ctx, field, 1_c_int, int(s%bc%type_right , c_int)), 'bc')
        call gpu_check(fen_gpu_set_bc_type(ctx, field, 2_c_int, int(s%bc%type_bottom, c_int)), 'bc')
        call gpu_check(fen_gpu_set_bc_type(ctx, field, 3_c_int, int(s%bc%type_top   , c_int)), 'bc')
        call gpu_check(fen_gpu_set_bc_plane(ctx, field, 0_c_int, c_loc(s%bc%left)  , 0_c_int), 'bc')
        call gpu_check(fen_gpu_set_bc_plane(ctx, field, 1_c_int, c_loc(s%bc%right) , 0_c_int), 'bc')
        call gpu_check(fen_gpu_set_bc_plane(ctx, field, 2_c_int, c_loc(s%bc%bottom), 0_c_int), 'bc')
        call gpu_check(fen_gpu_set_bc_plane(ctx, field, 3_c_int, c_loc(s%bc%top)   , 0_c_int), 'bc')
#if DIM==3
        call gpu_check(fen_gpu_set_bc_type(ctx, field, 4_c_int, int(s%bc%type_front, c_int)), 'bc')
        call gpu_check(fen_gpu_set_bc_type(ctx, field, 5_c_int, int(s%bc%type_back , c_int)), 'bc')
        call gpu_check(fen_gpu_set_bc_plane(ctx, field, 4_c_int, c_loc(s%bc%front), 0_c_int), 'bc')
        call gpu_check(fen_gpu_set_bc_plane(ctx, field, 5_c_int, c_loc(s%bc%back) , 0_c_int), 'bc')
#endif
    end subroutine gpu_push_bc

    !> module scalars of navier_stokes_mod -> device (density, viscosity, g, CFL, constant_CFL, dt_o)
    subroutine gpu_push_params()
        use navier_stokes_mod, only : density, viscosity, g, CFL, constant_CFL, dt_o, dt_visc, dt_conv
        type(fen_ns_params) :: p
        p%density = density; p%viscosity = viscosity
        p%g = 0.0_dp
        p%g(1:gpu_ndim) = g(1:gpu_ndim)
        p%CFL = CFL; p%dt_o = dt_o; p%dt_visc = dt_visc; p%dt_conv = dt_conv
        p%constant_CFL = merge(1_c_int, 0_c_int, constant_CFL)
        call gpu_check(fen_gpu_set_params(ctx, p), 'set params')
    end subroutine gpu_push_params

    !=============================================================================================
    !> init_solver(comp_grid), src/solver.f90:34.  The host fields are still allocated by the
    !> reference routine (drivers poke them); the device side gets the same fields, BC wiring table
    !> and Poisson variant.
    subroutine init_solver(comp_grid)
        use navier_stokes_mod, only : allocate_navier_stokes_fields
#ifdef MF
        use volume_of_fluid_mod, only : allocate_vof_fields
        use multiphase_mod     , only : allocate_multiphase_fields
#endif
        type(grid), intent(in) :: comp_grid
        call allocate_navier_stokes_fields(comp_grid)            ! host mirrors, navier_stokes.f90:752
        if (.not. c_associated(ctx)) call gpu_attach_grid(comp_grid)
        call gpu_push_params()
#ifdef MF
        ! solver.f90:83-98: host mirrors of the VoF / multiphase fields, then the device does the rest
        call allocate_vof_fields(comp_grid)
        call allocate_multiphase_fields(comp_grid)
        call gpu_push_mf_params()
        call gpu_check(fen_gpu_init_solver_mf(ctx, c_funloc(gpu_distance), c_null_ptr, &
                                              comp_grid%origin(1), comp_grid%origin(2)), 'init_solver')
        call gpu_pull_mf_params()                                ! rhomin, irhomin
#else
        call gpu_check(fen_gpu_init_solver(ctx), 'init_solver')  ! device fields + init_poisson_solver
#endif
        advance_solution => navier_stokes_solver
    end subroutine init_solver

#ifdef MF
    !> C-callable trampoline for the reference's `distance` procedure pointer (volume_of_fluid.f90:40-46)
    function gpu_distance(user, x, y) bind(C) result(d)
        use volume_of_fluid_mod, only : distance
        type(c_ptr), value :: user
        real(c_double), value :: x, y
        real(c_double) :: d
        d = distance(x, y)
    end function gpu_distance

    !> module variables of multiphase_mod / volume_of_fluid_mod -> device
    subroutine gpu_push_mf_params()
        use multiphase_mod     , only : rho_0, rho_1, mu_0, mu_1, sigma
        use volume_of_fluid_mod, only : beta, cut, quadratic, x_first
        type(fen_mf_params) :: p
        call gpu_check(fen_gpu_mf_get_params(ctx, p), 'mf params')
        p%rho_0 = rho_0; p%rho_1 = rho_1; p%mu_0 = mu_0; p%mu_1 = mu_1; p%sigma = sigma
        p%beta = beta; p%cut = cut
        p%quadratic = merge(1_c_int, 0_c_int, quadratic)
        p%x_first = merge(1_c_int, 0_c_int, x_first)
        call gpu_check(fen_gpu_mf_set_params(ctx, p), 'mf params')
    end subroutine gpu_push_mf_params

    subroutine gpu_pull_mf_params()
        use multiphase_mod     , only : rhomin, irhomin
        use navier_stokes_mod  , only : dt_surf
        use volume_of_fluid_mod, only : x_first
        type(fen_mf_params) :: p
        call gpu_check(fen_gpu_mf_get_params(ctx, p), 'mf params')
        rhomin = p%rhomin; irhomin = p%irhomin; dt_surf = p%dt_surf
        x_first = p%x_first /= 0
    end subroutine gpu_pull_mf_params

    !> advect_vof(v, dt), volume_of_fluid.f90:434, for drivers that advect with their own velocity field
    !> (test/small_test/volume_of_fluid/reversed/reversed.f90:63): vfield = device id of v%x
    subroutine advect_vof(vfield, dt)
        integer(c_int), intent(in) :: vfield
        real(dp)      , intent(in) :: dt
        call gpu_check(fen_gpu_advect_vof(ctx, vfield, dt), 'advect_vof')
    end subroutine advect_vof

    subroutine get_h_from_vof()
        call gpu_check(fen_gpu_get_h_from_vof(ctx), 'get_h_from_vof')
    end subroutine get_h_from_vof

    subroutine check_vof_integral(int_phase_1, int_phase_2)
        real(dp), intent(out) :: int_phase_1, int_phase_2
        call gpu_check(fen_gpu_check_vof_integral(ctx, int_phase_1, int_phase_2), 'check_vof_integral')
    end subroutine check_vof_integral

    !> pull vof (and rho, mu for output) into the reference's host arrays
    subroutine gpu_pull_vof()
        use volume_of_fluid_mod, only : vof
        use navier_stokes_mod  , only : rho, mu
        call gpu_pull(vof, FEN_VOF)
        call gpu_pull(rho, FEN_RHO)
        call gpu_pull(mu, FEN_MU)
    end subroutine gpu_pull_vof
#endif

    !> push the initial condition and BC planes the driver wrote after init_solver
    subroutine gpu_push_state()
        use navier_stokes_mod, only : p, v
        call gpu_push_params()
        call gpu_push_bc(p, FEN_P);     call gpu_push(p, FEN_P)
        call gpu_push_bc(v%x, FEN_VX);  call gpu_push(v%x, FEN_VX)
        call gpu_push_bc(v%y, FEN_VY);  call gpu_push(v%y, FEN_VY)
#if DIM==3
        call gpu_push_bc(v%z, FEN_VZ);  call gpu_push(v%z, FEN_VZ)
#endif
    end subroutine gpu_push_state

    !> pull p and v back into the reference's host arrays (before output / post-processing)
    subroutine gpu_pull_state()
        use navier_stokes_mod, only : p, v
        call gpu_pull(p, FEN_P)
        call gpu_pull(v%x, FEN_VX)
        call gpu_pull(v%y, FEN_VY)
#if DIM==3
        call gpu_pull(v%z, FEN_VZ)
#endif
    end subroutine gpu_pull_state

    !> set_timestep(comp_grid, dt, U), src/navier_stokes.f90:623 (also sets dt_o = dt, :664)
    subroutine set_timestep(comp_grid, dt, U)
        use navier_stokes_mod, only : dt_o
        type(grid), intent(in   ) :: comp_grid
        real(dp)  , intent(  out) :: dt
        real(dp)  , intent(in   ) :: U
        call gpu_push_params()
        call gpu_check(fen_gpu_set_timestep(ctx, U, dt), 'set_timestep')
        dt_o = dt
    end subroutine set_timestep

    !> navier_stokes_solver(comp_grid, step, dt), src/navier_stokes.f90:50 == advance_solution
    subroutine navier_stokes_solver(comp_grid, step, dt)
        use navier_stokes_mod, only : maxdiv, maxCFL
        type(grid), intent(in   ) :: comp_grid
        integer   , intent(in   ) :: step
        real(dp)  , intent(inout) :: dt
        call gpu_check(fen_gpu_navier_stokes_solver(ctx, int(step, c_int), dt), 'navier_stokes_solver')
        call gpu_check(fen_gpu_get_status(ctx, maxdiv, maxCFL), 'checks')
    end subroutine navier_stokes_solver

    !> destroy_solver, src/solver.f90:333
    subroutine destroy_solver()
        call gpu_check(fen_gpu_destroy_solver(ctx), 'destroy_solver')
        call gpu_check(fen_gpu_destroy(ctx), 'destroy')
        ctx = c_null_ptr
    end subroutine destroy_solver

    !=============================================================================================
    ! poisson_mod (src/poisson.f90:51-52): phi is the reference's host scalar; the solve runs on the
    ! device copy of FEN_PHI.
    subroutine init_poisson_solver(phi)
        type(scalar), intent(in) :: phi
        if (.not. c_associated(ctx)) call gpu_attach_grid(phi%G)
        call gpu_check(fen_gpu_init_poisson_solver(ctx), 'init_poisson_solver')
    end subroutine init_poisson_solver

    subroutine solve_poisson(phi)
        type(scalar), intent(inout) :: phi
        call gpu_push(phi, FEN_PHI)
        call gpu_check(fen_gpu_solve_poisson(ctx, FEN_PHI), 'solve_poisson')
        call gpu_pull(phi, FEN_PHI)
    end subroutine solve_poisson

    subroutine destroy_poisson_solver(phi)
        type(scalar), intent(in) :: phi
        call gpu_check(fen_gpu_destroy_poisson_solver(ctx), 'destroy_poisson_solver')
    end subroutine destroy_poisson_solver

    !=============================================================================================
    !> update_halos(f, G, l), src/halo.f90:12 -- same explicit-shape dummy as the reference.
    !> Stand-alone use (tests/fields): round trip through a scratch device field.
    subroutine update_halos(f, G, l)
        integer   , intent(in   ) :: l
        type(grid), intent(in   ) :: G
        real(dp)  , intent(inout), target :: f(G%lo(1)-l:G%hi(1)+l, G%lo(2)-l:G%hi(2)+l, G%lo(3)-l:G%hi(3)+l)
        integer(c_int) :: id
        call gpu_check(fen_gpu_scalar_allocate(ctx, int(l, c_int), 0_c_int, id), 'update_halos')
        call gpu_check(fen_gpu_push(ctx, id, c_loc(f), int(l, c_int)), 'update_halos')
        call gpu_check(fen_gpu_update_halos(ctx, id), 'update_halos')
        call gpu_check(fen_gpu_pull(ctx, id, c_loc(f), int(l, c_int)), 'update_halos')
        call gpu_check(fen_gpu_scalar_destroy(ctx, id), 'update_halos')
    end subroutine update_halos

    !=============================================================================================
    ! solver_mod output and restart (src/solver.f90:103-329): same files, written from the device fields
    !=============================================================================================
    !> C string from a Fortran one
    pure function cstr(s) result(c)
        character(*), intent(in) :: s
        character(kind=c_char) :: c(len_trim(s) + 1)
        integer :: n
        do n = 1, len_trim(s)
            c(n) = s(n:n)
        end do
        c(len_trim(s) + 1) = c_null_char
    end function cstr

    !> save_state(step), src/solver.f90:160: data/state_<step7>.raw
    subroutine save_state(step)
        integer, intent(in) :: step
        character(len=7) :: sn
        write(sn, '(I0.7)') step
        call gpu_check(fen_gpu_save_state(ctx, cstr('data/state_'//sn//'.raw')), 'save_state')
    end subroutine save_state

    !> load_state(step), src/solver.f90:244 (also refreshes the ghost nodes, :283-297)
    subroutine load_state(step)
        integer, intent(in) :: step
        character(len=7) :: sn
        write(sn, '(I0.7)') step
        call gpu_check(fen_gpu_load_state(ctx, cstr('data/state_'//sn//'.raw')), 'load_state')
    end subroutine load_state

    !> save_fields(step), src/solver.f90:103: data/vx_<step7>.raw, vy_, [vz_], p_, [vof_]
    subroutine save_fields(step)
        integer, intent(in) :: step
        call gpu_check(fen_gpu_save_fields(ctx, int(step, c_int), cstr('data')), 'save_fields')
    end subroutine save_fields

    !> scalar%write / scalar%read of a device field (src/scalar.f90:428, :400)
    subroutine gpu_scalar_write(field, filename)
        integer(c_int), intent(in) :: field
        character(*)  , intent(in) :: filename
        call gpu_check(fen_gpu_scalar_write(ctx, field, cstr(filename)), 'scalar%write')
    end subroutine gpu_scalar_write

    subroutine gpu_scalar_read(field, filename)
        integer(c_int), intent(in) :: field
        character(*)  , intent(in) :: filename
        call gpu_check(fen_gpu_scalar_read(ctx, field, cstr(filename)), 'scalar%read')
    end subroutine gpu_scalar_read

    !> print_solver_status(log_id, step, time, dt), src/navier_stokes.f90:734 (format :746)
    subroutine print_solver_status(log_id, step, time, dt)
        use global_mod, only : myrank
        integer , intent(in) :: log_id, step
        real(dp), intent(in) :: time, dt
        character(kind=c_char) :: buf(256)
        character(len=255) :: line
        integer :: n
        call gpu_check(fen_gpu_status_line(ctx, int(step, c_int), time, dt, buf, 256_c_int), 'print_solver_status')
        line = ' '
        do n = 1, 255
            if (buf(n) == c_null_char) exit
            line(n:n) = buf(n)
        end do
        if (myrank == 0) write(log_id, '(A)') trim(line)
    end subroutine print_solver_status

    !> host callback between the predictor and the Poisson solve: the call site of apply_ibm_forcing(v, dt)
    !> (src/navier_stokes.f90:106-108).  `fn` is a bind(C) function (user, step, dt) -> integer(c_int); it may
    !> gpu_pull(v%x, FEN_VX) ..., force the host arrays and gpu_push them back.
    subroutine gpu_set_forcing_hook(fn)
        type(c_funptr), intent(in) :: fn
        call gpu_check(fen_gpu_set_forcing_hook(ctx, fn, c_null_ptr), 'set_forcing_hook')
    end subroutine gpu_set_forcing_hook

    !> fields_mod operators on device fields (src/fields.f90:31,120,298,175); arguments are device field ids
    subroutine gpu_gradient(s, vx)
        integer(c_int), intent(in) :: s, vx
        call gpu_check(fen_gpu_gradient(ctx, s, vx), 'gradient')
    end subroutine gpu_gradient
    subroutine gpu_divergence(vx, s)
        integer(c_int), intent(in) :: vx, s
        call gpu_check(fen_gpu_divergence(ctx, vx, s), 'divergence')
    end subroutine gpu_divergence
    subroutine gpu_laplacian(vx, ox)
        integer(c_int), intent(in) :: vx, ox
        call gpu_check(fen_gpu_laplacian(ctx, vx, ox), 'laplacian')
    end subroutine gpu_laplacian
    subroutine gpu_center_to_face(s, vx)
        integer(c_int), intent(in) :: s, vx
        call gpu_check(fen_gpu_center_to_face(ctx, s, vx), 'center_to_face')
    end subroutine gpu_center_to_face
    !> laplacian_of_scalar (src/fields.f90:256), face_to_center (:210; face = 'x', 'y' or 'z'), curl (:347)
    subroutine gpu_laplacian_scalar(s, o)
        integer(c_int), intent(in) :: s, o
        call gpu_check(fen_gpu_laplacian_scalar(ctx, s, o), 'laplacian_of_scalar')
    end subroutine gpu_laplacian_scalar
    subroutine gpu_face_to_center(sf, sc, face)
        integer(c_int), intent(in) :: sf, sc
        character(len=1), intent(in) :: face
        call gpu_check(fen_gpu_face_to_center(ctx, sf, sc, int(index('xyz', face) - 1, c_int)), 'face_to_center')
    end subroutine gpu_face_to_center
    subroutine gpu_curl(vx, ox)
        integer(c_int), intent(in) :: vx, ox
        call gpu_check(fen_gpu_curl(ctx, vx, ox), 'curl')
    end subroutine gpu_curl

    !> scalar%max_value / scalar%integral of a device field, reduced over all ranks (src/scalar.f90:179, :201)
    real(dp) function gpu_max_value(field)
        integer(c_int), intent(in) :: field
        call gpu_check(fen_gpu_max_value(ctx, field, gpu_max_value), 'max_value')
    end function gpu_max_value
    real(dp) function gpu_integral(field)
        integer(c_int), intent(in) :: field
        call gpu_check(fen_gpu_integral(ctx, field, gpu_integral), 'integral')
    end function gpu_integral

    !> scalar%update_ghost_nodes on the device copy of a solver field (src/scalar.f90:223)
    subroutine update_ghost_nodes(field, ncomp)
        integer(c_int), intent(in) :: field
        integer       , intent(in) :: ncomp
        call gpu_check(fen_gpu_update_ghost_nodes(ctx, field, int(ncomp, c_int)), 'update_ghost_nodes')
    end subroutine update_ghost_nodes

end module fen_gpu_mod
