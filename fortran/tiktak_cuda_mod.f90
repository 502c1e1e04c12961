! tiktak_cuda_mod.f90 -- ISO_C_BINDING interface of libtiktak_cuda (include/tiktak_cuda.h) for the reference driver.
!
! The reference (TiktakGlobalSearch.f90) keeps its state machine, CLI, config.txt parser and .dat writers; the bodies
! of setupSobol (:825-838), solveAtSobolPoints (:405-471), chooseSobol (:840-899), LocalMinimizations (:473-544) and
! lastSearch (:635-681) call the functions below instead (INTEGRATION.md section 2 lists the replacement per procedure).
! Add this file before minimize.o in the reference Makefile's object list and link with -ltiktak_cuda -lcudart.
! No Fortran compiler exists in the build image: this source is reviewed, not compiled here;
! tests/test_abi_and_config.py checks that every NAME= below is declared in the header and exported by the library.
MODULE tiktak_cuda_mod
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  TYPE, BIND(C) :: tiktak_config            ! include/tiktak_cuda.h: struct tiktak_config
     INTEGER(C_INT32_T) :: nx, nm, maxeval, qr_ndraw, legitimate, maxpoints_plus1, search_type, lottery_points
     REAL(C_DOUBLE)     :: fvalmax
     TYPE(C_PTR)        :: range, init, bound          ! p_range(nx,2), p_init(nx), p_bound(nx,2): column-major as is
     INTEGER(C_INT32_T) :: objective_id, device, rank, world
     INTEGER(C_INT32_T) :: nm_itmax, nm_shrink_mode, sobol_quirk, text_roundtrip, round_size
     REAL(C_DOUBLE)     :: nm_ftol
     TYPE(C_PTR)        :: user
  END TYPE
  INTERFACE
     INTEGER(C_INT) FUNCTION tiktak_cuda_create(h, cfg) BIND(C, NAME='tiktak_cuda_create')
       IMPORT; TYPE(C_PTR), INTENT(OUT) :: h; TYPE(tiktak_config), INTENT(IN) :: cfg
     END FUNCTION
     SUBROUTINE tiktak_cuda_destroy(h) BIND(C, NAME='tiktak_cuda_destroy')
       IMPORT; TYPE(C_PTR), VALUE :: h
     END SUBROUTINE
     INTEGER(C_INT) FUNCTION tiktak_cuda_sobol_points(h, first, count, out) BIND(C, NAME='tiktak_cuda_sobol_points')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT64_T), VALUE :: first, count; REAL(C_DOUBLE) :: out(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_pretest(h, n_eval, n_legit) BIND(C, NAME='tiktak_cuda_pretest')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT64_T), INTENT(OUT) :: n_eval, n_legit
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_pretest_values(h, first, count, fval) BIND(C, NAME='tiktak_cuda_pretest_values')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT64_T), VALUE :: first, count; REAL(C_DOUBLE) :: fval(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_choose(h, x_starts, sobol_idx) BIND(C, NAME='tiktak_cuda_choose')
       IMPORT; TYPE(C_PTR), VALUE :: h; REAL(C_DOUBLE) :: x_starts(*); INTEGER(C_INT32_T) :: sobol_idx(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_local(h, alg, first_k, n_k, starts, results, evals) BIND(C, NAME='tiktak_cuda_local')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg, first_k, n_k
       REAL(C_DOUBLE) :: starts(*), results(*); INTEGER(C_INT32_T) :: evals(*)
     END FUNCTION
     ! the same round without host synchronisation: queue every round, read once (include/tiktak_cuda.h)
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_enqueue(h, alg, first_k, n_k, send) BIND(C, NAME='tiktak_cuda_local_enqueue')
       IMPORT; TYPE(C_PTR), VALUE :: h, send; INTEGER(C_INT32_T), VALUE :: alg, first_k, n_k
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_ingest(h, first_k, n_k, recv) BIND(C, NAME='tiktak_cuda_local_ingest')
       IMPORT; TYPE(C_PTR), VALUE :: h, recv; INTEGER(C_INT32_T), VALUE :: first_k, n_k
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_ingest_rounds(h, first_k, n_k, round_size, recv, accepted) &
          BIND(C, NAME='tiktak_cuda_local_ingest_rounds')
       IMPORT; TYPE(C_PTR), VALUE :: h, recv; INTEGER(C_INT32_T), VALUE :: first_k, n_k, round_size
       INTEGER(C_INT32_T), INTENT(OUT) :: accepted
     END FUNCTION
     ! enqueue + commit of restarts first_k.. as ONE call (what fortran/TiktakGlobalSearch.patch uses): `accepted` restarts were
     ! committed to the device-resident searchResults table; queue the rest again.  Collective when world > 1.
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_launch(h, alg, first_k, n_k, round_size, accepted) &
          BIND(C, NAME='tiktak_cuda_local_launch')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg, first_k, n_k, round_size
       INTEGER(C_INT32_T), INTENT(OUT) :: accepted
     END FUNCTION
     ! 0 = the search ended like the reference's; 2 = DFNLS refused the start
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_status(h, first_k, n_k, status) BIND(C, NAME='tiktak_cuda_local_status')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: first_k, n_k; INTEGER(C_INT32_T) :: status(*)
     END FUNCTION
     ! asynchronous instances: restarts first_k.. by `instances` blocks that take the next restart when they finish one (:546-600)
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_async(h, alg, first_k, n_k, instances, used) BIND(C, NAME='tiktak_cuda_local_async')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg, first_k, n_k, instances
       INTEGER(C_INT32_T), INTENT(OUT) :: used
     END FUNCTION
     ! which row each start of an asynchronous run was blended toward (-1: none) and the restart's place in the order of the file
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_async_info(h, first_k, n_k, base, order) BIND(C, NAME='tiktak_cuda_local_async_info')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: first_k, n_k; INTEGER(C_INT32_T) :: base(*), order(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_async_supported(h, alg, yes) BIND(C, NAME='tiktak_cuda_local_async_supported')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg; INTEGER(C_INT32_T), INTENT(OUT) :: yes
     END FUNCTION
     ! the tick of the virtual clock: 0 = pass of amoeba's loop (four-warp instances), 1 = objective evaluation (one-warp instances, DFNLS)
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_async_clock(h, alg, instances, clock) BIND(C, NAME='tiktak_cuda_local_async_clock')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg, instances; INTEGER(C_INT32_T), INTENT(OUT) :: clock
     END FUNCTION
     ! RESCUE_H calls of each restart's DFNLS search (minimize.f90:643-650)
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_nrescue(h, first_k, n_k, nrescue) BIND(C, NAME='tiktak_cuda_local_nrescue')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: first_k, n_k; INTEGER(C_INT32_T) :: nrescue(*)
     END FUNCTION
     ! SolveMissingSearch (:1104-1146) for one restart missing from searchResults.dat; start(nx), row(nx+4), evals
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_missing(h, alg, k, start, row, evals) BIND(C, NAME='tiktak_cuda_local_missing')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg, k
       REAL(C_DOUBLE) :: start(*), row(*); INTEGER(C_INT32_T), INTENT(OUT) :: evals
     END FUNCTION
     ! several GPUs in one process: handles(1:n) created with rank = 0..n-1, world = n, device = rank; one communicator for all
     INTEGER(C_INT) FUNCTION tiktak_cuda_comm_init_all(handles, n) BIND(C, NAME='tiktak_cuda_comm_init_all')
       IMPORT; TYPE(C_PTR) :: handles(*); INTEGER(C_INT32_T), VALUE :: n
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_capacity(h, alg, restarts) BIND(C, NAME='tiktak_cuda_local_capacity')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: alg; INTEGER(C_INT32_T), INTENT(OUT) :: restarts
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_local_fetch(h, first_k, n_k, starts, results, evals) BIND(C, NAME='tiktak_cuda_local_fetch')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT32_T), VALUE :: first_k, n_k
       REAL(C_DOUBLE) :: starts(*), results(*); INTEGER(C_INT32_T) :: evals(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_last_search(h, fval, x, iters) BIND(C, NAME='tiktak_cuda_last_search')
       IMPORT; TYPE(C_PTR), VALUE :: h; REAL(C_DOUBLE), INTENT(OUT) :: fval; REAL(C_DOUBLE) :: x(*)
       INTEGER(C_INT32_T), INTENT(OUT) :: iters
     END FUNCTION
     INTEGER(C_INT) FUNCTION tiktak_cuda_best(h, fbest, xbest, k) BIND(C, NAME='tiktak_cuda_best')
       IMPORT; TYPE(C_PTR), VALUE :: h; REAL(C_DOUBLE), INTENT(OUT) :: fbest; REAL(C_DOUBLE) :: xbest(*)
       INTEGER(C_INT32_T), INTENT(OUT) :: k
     END FUNCTION
     ! objFun at n points, x(n, nx) column-major (objective.f90:125-142): chooseSobol's p_init row, CLI option d
     INTEGER(C_INT) FUNCTION tiktak_cuda_objfun(h, x, n, fval) BIND(C, NAME='tiktak_cuda_objfun')
       IMPORT; TYPE(C_PTR), VALUE :: h; REAL(C_DOUBLE) :: x(*); INTEGER(C_INT64_T), VALUE :: n; REAL(C_DOUBLE) :: fval(*)
     END FUNCTION
     ! counts of the queued pre-test (read when first needed: lastSobol.dat / legitSobol.dat)
     INTEGER(C_INT) FUNCTION tiktak_cuda_pretest_result(h, n_eval, n_legit) BIND(C, NAME='tiktak_cuda_pretest_result')
       IMPORT; TYPE(C_PTR), VALUE :: h; INTEGER(C_INT64_T), INTENT(OUT) :: n_eval, n_legit
     END FUNCTION
     ! update options 2 / 3 (genericParams.f90:118-251): the rows of searchResults.dat, (n, nx+4) column-major, re-enter the device table
     INTEGER(C_INT) FUNCTION tiktak_cuda_push_results(h, rows, n) BIND(C, NAME='tiktak_cuda_push_results')
       IMPORT; TYPE(C_PTR), VALUE :: h; REAL(C_DOUBLE) :: rows(*); INTEGER(C_INT32_T), VALUE :: n
     END FUNCTION
     ! the -1 defaults of config.txt (genericParams.f90:307-372), applied in place
     INTEGER(C_INT) FUNCTION tiktak_cuda_config_defaults(nx, maxeval, qr_ndraw, legitimate, fvalmax, maxpoints, lottery_points) &
          BIND(C, NAME='tiktak_cuda_config_defaults')
       IMPORT; INTEGER(C_INT32_T), VALUE :: nx; INTEGER(C_INT32_T) :: maxeval, qr_ndraw, legitimate, maxpoints, lottery_points
       REAL(C_DOUBLE) :: fvalmax
     END FUNCTION
     ! messages: C strings (use C_F_POINTER / a loop up to the NUL to copy them into a CHARACTER variable for exitState)
     TYPE(C_PTR) FUNCTION tiktak_cuda_strerror(code) BIND(C, NAME='tiktak_cuda_strerror')
       IMPORT; INTEGER(C_INT), VALUE :: code
     END FUNCTION
     TYPE(C_PTR) FUNCTION tiktak_cuda_last_error() BIND(C, NAME='tiktak_cuda_last_error')
       IMPORT
     END FUNCTION
  END INTERFACE
END MODULE
