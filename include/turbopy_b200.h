/*
 * turbopy_b200 -- C ABI of the B200-native turboPy particle hot path.
 *
 * This header is the drop-in boundary: plain pointers, sizes and POD structs,
 * no torch / C++ types.  The reference (arichar6/turbopy) is pure Python with
 * no FFI of its own; each entry point below cites the reference method whose
 * arithmetic it replaces.  The Python ComputeTools in turbopy_b200/tools.py
 * bind these through ctypes (INTEGRATION.md shows the binding a turboPy
 * maintainer would add).
 *
 * Conventions
 *   - every function returns int: 0 = TP_OK, >0 = a cudaError_t, <0 = TP_E_*;
 *     tp_error_string() turns either into text.  Nothing here falls back to
 *     the CPU: without a CUDA device every compute entry point fails.
 *   - all device pointers are fp64 (double) unless stated otherwise; `stream`
 *     is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 *     asynchronous on that stream and may be captured into a CUDA graph.
 *   - a "(n,3) array" is described by (ptr, stride_n, stride_c) in ELEMENTS:
 *     element (i,c) lives at ptr[i*stride_n + c*stride_c].  Two layouts are
 *     accelerated: SoA (stride_n = 1, stride_c = ld >= n) and AoS, numpy's
 *     C-order (stride_n = 3, stride_c = 1).  Anything else -> TP_E_LAYOUT.
 *   - arithmetic: IEEE fp64, round-to-nearest, no FMA contraction, operation
 *     order of the reference; results are bit-identical to the reference's
 *     numpy on the same inputs ("Arithmetic contract").
 */
#ifndef TURBOPY_B200_H
#define TURBOPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TP_VERSION 100            /* 0.1.0 */

/* ---- error codes (negative; positive values are cudaError_t) ------------ */
#define TP_OK            0
#define TP_E_ARG        -1        /* bad argument (null pointer, n < 0, ...)    */
#define TP_E_LAYOUT     -2        /* strides are neither SoA nor AoS            */
#define TP_E_NODEVICE   -3        /* no CUDA device / not an sm_100 device      */
#define TP_E_GRID       -4        /* fewer than 2 grid nodes, dr <= 0, ...      */
#define TP_E_MODE       -5        /* unknown field mode / formula / axis        */
#define TP_E_GRAPH      -6        /* graph capture misuse                       */
#define TP_RECORDS_DOUBLES_PER_NODE 18   /* caller-owned record scratch of tp_gather_push_windows_f64 */
#define TP_E_NCCL       -100      /* libnccl not loadable; -100 - k = NCCL returned ncclResult_t k */

/* ---- how E / B are given to tp_boris_push_f64 ---------------------------- */
#define TP_FIELD_UNIFORM_HOST 0   /* ptr -> 3 doubles in HOST memory, read at call time   */
#define TP_FIELD_UNIFORM_DEV  1   /* ptr -> 3 doubles in DEVICE memory, read by the kernel */
#define TP_FIELD_PER_PARTICLE 2   /* ptr -> (n,3) DEVICE array with its own strides        */

/* ---- linear-gather formulas (they round differently) --------- */
#define TP_INTERP_A 1             /* Grid.create_interpolator          turbopy/core.py:711-755 */
#define TP_INTERP_B 2             /* interp1d, 1-D y  -> numpy.interp  turbopy/computetools.py:480 */
#define TP_INTERP_C 3             /* interp1d, N-D y  -> scipy _call_linear (same call site) */

/* ---- per-particle status bits written by the gather kernels -------------- */
#define TP_ST_BELOW      1        /* coordinate < r[0]   (interp1d: ValueError; core.py:726 assert) */
#define TP_ST_ABOVE      2        /* coordinate > r[ng-1](interp1d: ValueError; core.py:727 assert) */
#define TP_ST_BAD_WINDOW 4        /* formula A found 0 or >2 nodes in (x-dr, x+dr): core.py:729 assert */

typedef void* tp_stream_t;

/* (n,3) position and momentum arrays, updated IN PLACE (same storage the
 * caller published to other modules: turbopy/core.py:334-339,371-383). */
typedef struct tp_particles {
    double* pos;  int64_t pos_stride_n;  int64_t pos_stride_c;
    double* mom;  int64_t mom_stride_n;  int64_t mom_stride_c;
    int64_t n;
} tp_particles;

/* scalars of BorisPush.push(position, momentum, charge, mass, E, B)
 * (turbopy/computetools.py:407,427): dt = owner.clock.dt read at call time,
 * c2 = BorisPush.c2 (:405).  mass_sq is the caller's `mass**2` (:430) -- the
 * only way the push uses the mass -- evaluated by the host language so that it
 * carries the reference's own rounding of that power. */
typedef struct tp_push_consts {
    double charge;
    double mass_sq;
    double dt;
    double c2;
} tp_push_consts;

/* 1-D grid fields for the fused gather.  r: ng node coordinates (Grid.r,
 * turbopy/core.py:666-667), dr: Grid.dr (:625); r_first / r_last: host copies
 * of r[0] and r[ng-1] (Grid.r_min / r_max for a Grid), the bounds the kernels
 * check every coordinate against.  comp[0..2] = Ex,Ey,Ez and
 * comp[3..5] = Bx,By,Bz: DEVICE pointers to node 0 of each component, NULL
 * for a component that is identically zero; node j of component c is
 * comp[c][j*comp_stride[c]] (stride 1 for a (ng,) field, 3 for a column of a
 * Grid.generate_field(3) array, :675-698). */
typedef struct tp_grid_fields {
    const double* r;
    int64_t       ng;
    double        dr;
    double        r_first;
    double        r_last;
    const double* comp[6];
    int64_t       comp_stride[6];
} tp_grid_fields;

/* A grid whose node coordinates are EXACTLY r[i] = start + span*((double)i*step) for i < n-1 and
 * r[n-1] = start + span*1.0, each operation rounded once -- the arithmetic of turboPy's
 * Grid.r = r_min + (r_max - r_min)*np.linspace(0, 1, N) (turbopy/core.py:666-667,709).  The caller
 * asserts this (the Python layer checks it bit for bit against Grid.r on the host before using the
 * *_lin_* entry points); the kernels then rebuild r in registers instead of reading it from HBM. */
typedef struct tp_linspace {
    double  start;
    double  span;
    double  step;
    int64_t n;
} tp_linspace;

/* ---- library / device ----------------------------------------------------- */
int         tp_version(void);
const char* tp_error_string(int code);
/* sm_count, compute capability, L2 bytes, max opt-in shared memory per block */
int         tp_device_info(int* sm_count, int* cc_major, int* cc_minor,
                           int64_t* l2_bytes, int64_t* smem_optin_bytes);

/* ---- kernel 1a: BorisPush.push  (turbopy/computetools.py:407-441) -------- */
int tp_boris_push_f64(const tp_particles* p, const tp_push_consts* k,
                      const double* E, int e_mode, int64_t e_stride_n, int64_t e_stride_c,
                      const double* B, int b_mode, int64_t b_stride_n, int64_t b_stride_c,
                      tp_stream_t stream);
/* tp_boris_push_f64 with the energy / moment diagnostic of the step riding on it: out8 receives
 * {KE, Px, Py, Pz, sum|p|^2, n, 0, 0} of the momenta as they are BEFORE this push (what tp_reduce_moments_f64 called
 * just before it would deliver, to summation-order rounding, 1e-12); the particles are pushed exactly as by
 * tp_boris_push_f64.  With uniform E, B and SoA particles the TMA ring kernel accumulates the sums from the momenta it
 * has just loaded (no second pass over the momentum arrays); otherwise the standalone reduction is launched in front
 * of the push.  See tp_gather_push_moments_f64.  reduce_scratch: tp_reduce_scratch_doubles() doubles. */
int tp_boris_push_moments_f64(const tp_particles* p, const tp_push_consts* k,
                              const double* E, int e_mode, int64_t e_stride_n, int64_t e_stride_c,
                              const double* B, int b_mode, int64_t b_stride_n, int64_t b_stride_c,
                              double mass, double mass_sq, double c2, double* out8, double* reduce_scratch,
                              tp_stream_t stream);
/* nsteps consecutive pushes with the same E and B (the reference's time loop calling
 * BorisPush.push nsteps times with unchanged field arguments, e.g. a uniform static
 * E x B) in ONE pass over the particles: each particle is read once, stepped nsteps
 * times in registers and written once.  Bit-identical to nsteps calls of
 * tp_boris_push_f64; the traffic per push drops from 96 B to 96/nsteps B, so the
 * kernel leaves the HBM roofline and becomes fp64-bound. */
int tp_boris_push_steps_f64(const tp_particles* p, const tp_push_consts* k,
                            const double* E, int e_mode, int64_t e_stride_n, int64_t e_stride_c,
                            const double* B, int b_mode, int64_t b_stride_n, int64_t b_stride_c,
                            int nsteps, tp_stream_t stream);

/* ---- kernel 1b: fused linear gather + BorisPush.push ---------------------
 * E,B at each particle = linear interpolation of the grid fields at
 * pos[:,axis] (the gather of ChargedParticle.update,
 * tests/fixtures/particle_in_field/particle_in_field.py:61-63, vectorised over
 * particles), then the push above.  status: 4 device uint64 words
 * {count of flagged particles, lowest flagged index, OR of TP_ST_* bits, 0},
 * initialised by tp_status_reset; flagged particles are left untouched.
 * Node coordinates and fields are staged into shared memory by TMA bulk copies
 * when they fit (query with tp_gather_smem_bytes), else read through L2. */
int tp_gather_push_f64(const tp_particles* p, const tp_push_consts* k,
                       const tp_grid_fields* g, int axis, int formula,
                       uint64_t* status, tp_stream_t stream);
/* nsteps consecutive gather+push steps on UNCHANGED grid fields (a static field map) in one
 * pass over the particles: bit-identical to nsteps calls of tp_gather_push_f64, 96/nsteps
 * bytes of traffic per push.  A particle that leaves the grid during the pass is flagged
 * once and keeps the state it had when it left (what the remaining single-step calls would
 * have done: leave it untouched). */
int tp_gather_push_steps_f64(const tp_particles* p, const tp_push_consts* k, const tp_grid_fields* g,
                             int axis, int formula, int nsteps, uint64_t* status, tp_stream_t stream);
/* The same single step for grids TOO LARGE for shared memory, with a per-tile WINDOW of the node records staged
 * next to every 960-particle tile (one cp.async.bulk per tile) -- the fast path once the particles are ordered by
 * cell (tp_sort_by_cell), when a tile only touches a short run of consecutive cells.  tile_bounds: DEVICE array
 * of tp_gather_windows_len() int32 words, two per full tile {first cell, number of cells}; it is read by this
 * call to size the windows and REWRITTEN by it with the cells the particles moved into, i.e. the windows of the
 * next call.  It is a hint: particles whose cell is not covered (zeroed or stale bounds, particles in random
 * order) take their records from L2 like tp_gather_push_f64 -- identical results either way.  With a NULL
 * tile_bounds, or when tp_gather_windows_len() is 0, this IS tp_gather_push_f64.  records_scratch: where the
 * per-step record table is packed each call (formula B: cell records {r[j], r[j+1], y[j], slope_j} with
 * numpy.interp's per-cell slopes, 144-byte pitch; formulas A, C: node records {r, E, B, 0}, 80-byte pitch); caller-owned memory makes the call capturable in a
 * CUDA graph on any stream without a warm-up call (the library-owned scratch is allocated per stream). */
int tp_gather_push_windows_f64(const tp_particles* p, const tp_push_consts* k,
                               const tp_grid_fields* g, int axis, int formula,
                               int32_t* tile_bounds, int64_t tile_bounds_len,
                               double* records_scratch /* DEVICE, TP_RECORDS_DOUBLES_PER_NODE*ng doubles, or NULL: library-owned */,
                               uint64_t* status, tp_stream_t stream);
/* A gather+push step with the energy / moment diagnostic of the step riding on it: out8 receives
 * {KE, Px, Py, Pz, sum|p|^2, n, 0, 0} of the momenta as they are BEFORE this push -- what
 * tp_reduce_moments_f64(mom, ..., mass, mass_sq, c2, out8, reduce_scratch) called just before the push would
 * deliver, to summation-order rounding (1e-12 relative; both orders are deterministic) -- and the particles are
 * pushed exactly as by tp_gather_push_windows_f64 with the same arguments (tile_bounds / records_scratch may be
 * NULL: tp_gather_push_f64).  With formula B and SoA particles the TMA-pipelined kernels (grid staged in shared
 * memory, or per-tile windows on a large grid) accumulate the sums from the momenta they have just loaded, so the
 * diagnostic costs no second pass over the momentum arrays (24 B per particle saved on top of the push's 96 B); in
 * every other case the standalone reduction is launched in front of the push.  The reference has no energy diagnostic
 * (SURVEY 8a-8); the Simulation loop it rides on is diagnose() -> update() (turbopy/core.py:152-158).
 * reduce_scratch: tp_reduce_scratch_doubles() doubles. */
int tp_gather_push_moments_f64(const tp_particles* p, const tp_push_consts* k,
                                       const tp_grid_fields* g, int axis, int formula,
                                       int32_t* tile_bounds, int64_t tile_bounds_len, double* records_scratch,
                                       double mass, double mass_sq, double c2, double* out8, double* reduce_scratch,
                                       uint64_t* status, tp_stream_t stream);
int64_t tp_gather_windows_len(const tp_particles* p, const tp_grid_fields* g);
/* cells per tile a staged window can hold (formula-dependent record size); 0 when windows would not be used */
int64_t tp_gather_windows_capacity(const tp_particles* p, const tp_grid_fields* g, int formula);
/* prime tile_bounds from the particles' current coordinates (8 B read per particle); a zeroed array is valid too */
int tp_gather_windows_init(const tp_particles* p, const tp_grid_fields* g, int axis,
                           int32_t* tile_bounds, int64_t tile_bounds_len, tp_stream_t stream);
int tp_status_reset(uint64_t* status, tp_stream_t stream);
/* shared-memory bytes the staged variant would need; *staged = 1 if it fits */
int tp_gather_smem_bytes(const tp_grid_fields* g, int64_t* bytes, int* staged);

/* ---- gather only: Interpolators.interpolate1D(x, y)(x_new)
 * (turbopy/computetools.py:459-481) for ncomp fields at n points.
 * xq: n query points (stride xq_stride); y: component c, node j at
 * y[c*y_stride_c + j*y_stride_n]; out: component c, point i at
 * out[c*out_stride_c + i].  status as above (out-of-range points get NaN). */
int tp_interp1d_f64(const double* r, int64_t ng, double dr, double r_first, double r_last,
                    const double* y, int ncomp, int64_t y_stride_c, int64_t y_stride_n,
                    const double* xq, int64_t xq_stride, int64_t n,
                    double* out, int64_t out_stride_c, int formula,
                    uint64_t* status, tp_stream_t stream);
/* The same with the abscissas asserted to be a tp_linspace grid (r is still passed: the recovery path for
 * on-node / edge / out-of-range points reads it): the fast path rebuilds x[j], x[j+1] in registers instead of
 * loading them.  Used for formulas B and C on grids staged in shared memory; other cases run the generic kernel. */
int tp_interp1d_lin_f64(const double* r, int64_t ng, double dr, double r_first, double r_last,
                        const double* y, int ncomp, int64_t y_stride_c, int64_t y_stride_n,
                        const double* xq, int64_t xq_stride, int64_t n,
                        double* out, int64_t out_stride_c, int formula, const tp_linspace* grid,
                        uint64_t* status, tp_stream_t stream);

/* ---- kernel 2: FiniteDifference 1-D stencils ------------------------------
 * centered_difference (turbopy/computetools.py:104-120): d[i]=(y[i+1]-y[i-1])/two_dr,
 *   two_dr = 2*dr formed by the caller as the reference does; d[0]=d[n-1]=0.
 * upwind_left (:122-138): d[i]=(y[i]-y[i-1])/widths[i-1], d[0]=0 (widths: n-1).
 * del2_radial apply (:189-223, `del2_radial() @ f`): tridiagonal rows rebuilt in
 *   the kernel from r and dr with scipy's dia-matvec summation order. */
int tp_fd_centered_f64(const double* y, double* d, int64_t n, double two_dr, tp_stream_t stream);
int tp_fd_upwind_left_f64(const double* y, const double* widths, double* d, int64_t n,
                          tp_stream_t stream);
int tp_fd_del2_radial_apply_f64(const double* f, const double* r, double* out, int64_t n,
                                double g1 /* 1/(2.0*dr), :197 */, double g2 /* 1/(dr**2), :211 */,
                                tp_stream_t stream);
/* out = f + dtau * (del2_radial() @ f) in one pass (the explicit field update `f + dtau * (D @ f)`; numpy's three
 * roundings per point, no contraction); r: Grid.r, or NULL with `grid` describing it as a tp_linspace (or the other
 * way round).  out must not alias f (rows read their neighbours). */
int tp_fd_del2_radial_axpy_f64(const double* f, const double* r, const tp_linspace* grid, double* out, int64_t n,
                               double g1, double g2, double dtau, tp_stream_t stream);
/* generic banded operator apply (the dia_matrix objects the other builders
 * return, :140-187,225-381): out[i] = sum_k data[k][i+off[k]] * x[i+off[k]] in
 * the given diagonal order; data is (ndiag, n) row-major DEVICE memory,
 * offsets ndiag HOST ints (|off| <= 2, ndiag <= 5). */
int tp_fd_dia_apply_f64(const double* data, const int* offsets, int ndiag,
                        const double* x, double* out, int64_t n, tp_stream_t stream);
/* the same for operators whose diagonals are constant apart from <= 8 entries
 * within 4 columns of either boundary (ddx :140-155, ddr :246-263, del2
 * :225-244, BC_left_* :265-357, BC_right_extrap :359-381): coef[k] is the
 * constant of diagonal k (HOST), (ov_diag[o], ov_col[o], ov_val[o]) the
 * exceptions data[ov_diag][ov_col] = ov_val (HOST).  Reads only x: 16 B/point. */
int tp_fd_dia_const_apply_f64(const double* coef, const int* offsets, int ndiag,
                              const int* ov_diag, const int64_t* ov_col, const double* ov_val, int nov,
                              const double* x, double* out, int64_t n, tp_stream_t stream);

/* ---- PoissonSolver1DRadial.solve (turbopy/computetools.py:35-59): two fp64
 * prefix scans, I1 = cumsum(r*s*dr), I2 = cumsum(I1*dr/r with the axis fix),
 * out = I2 - I2[n-1].  mean_dr = np.mean(Grid.cell_widths) formed by the caller
 * (:49).  numpy's cumsum is sequential, a parallel scan is not: results agree to
 * rounding, not bit for bit; out[n-1] is exactly 0.  Two passes over the inputs (round 2): the tile sums of the
 * second scan are formed from per-tile aggregates of the first pass (csrc/scan_kernels.cu).
 * scratch: tp_poisson_scratch_doubles(n) doubles. */
/* The three r-dependent kernels on a tp_linspace grid: same results bit for bit, r (and cell_widths) never read:
 * upwind_left 24 -> 16 B/point, del2_radial 24 -> 16 B/point, the Poisson passes 40 -> 24 B/point. */
int tp_fd_upwind_left_lin_f64(const double* y, const tp_linspace* grid, double* d, int64_t n, tp_stream_t stream);
int tp_fd_del2_radial_apply_lin_f64(const double* f, const tp_linspace* grid, double* out, int64_t n,
                                    double g1, double g2, tp_stream_t stream);
int tp_poisson_radial_solve_lin_f64(const double* sources, const tp_linspace* grid, double mean_dr, double* out,
                                    int64_t n, double* scratch, tp_stream_t stream);
int64_t tp_poisson_scratch_doubles(int64_t n);
int tp_poisson_radial_solve_f64(const double* sources, const double* r, double mean_dr, double* out,
                                int64_t n, double* scratch, tp_stream_t stream);

/* ---- kernel 3: energy / moment reduction (new; no reference function) ----
 * out8 (device) = {KE, Px, Py, Pz, sum|p|^2, n, 0, 0} with
 * KE = sum |p|^2/(m1+mass), m1 = sqrt(mass_sq + |p|^2/c2) (the gamma*m of
 * turbopy/computetools.py:430-431).  Deterministic two-pass reduction:
 * scratch must hold tp_reduce_scratch_doubles() doubles. */
int64_t tp_reduce_scratch_doubles(void);
int tp_reduce_moments_f64(const double* mom, int64_t stride_n, int64_t stride_c, int64_t n,
                          double mass, double mass_sq, double c2,
                          double* out8, double* scratch, tp_stream_t stream);

/* ---- the path's only collective: SUM all-reduce of the diagnostic's <= 8 doubles --------------
 * One rank per GPU.  tp_nccl_unique_id on rank 0 -> ship the 128 bytes to every rank by any means (the Python
 * layer uses torch.distributed's store) -> tp_nccl_comm_init on every rank (collective) -> per diagnose step
 * tp_nccl_allreduce_sum_f64(comm, out8, 8, stream), IN PLACE on the reduction's output buffer and ON THE SAME
 * STREAM as tp_reduce_moments_f64, so reduction + all-reduce are stream-ordered without events and can be
 * captured in a CUDA graph together with the push steps.  libnccl.so.2 is resolved at run time (the copy the
 * process already has loaded, normally torch's); TP_E_NCCL if there is none. */
int tp_nccl_version(int* version);
int tp_nccl_unique_id(void* id128 /* 128 bytes out */);
int tp_nccl_comm_init(const void* id128, int world, int rank, void** comm);
int tp_nccl_allreduce_sum_f64(void* comm, double* buf, int64_t n, tp_stream_t stream);
int tp_nccl_comm_destroy(void* comm);

/* ---- end-to-end call on HOST buffers --------------------------------------
 * Same contract as tp_boris_push_f64 / tp_gather_push_f64 but pos/mom (and
 * per-particle E,B) are HOST arrays (numpy's (n,3) C-order or SoA): the call
 * streams them through the device in chunks (H2D, kernel, D2H overlapped on
 * internal streams) and returns when the host arrays hold the result.
 * Grid fields in tp_grid_fields are HOST pointers here and are uploaded by
 * the call.  status_host: 4 uint64 words, filled on return (may be NULL). */
int tp_boris_push_host_f64(const tp_particles* p_host, const tp_push_consts* k,
                           const double* E_host, int e_per_particle,
                           const double* B_host, int b_per_particle);
int tp_gather_push_host_f64(const tp_particles* p_host, const tp_push_consts* k,
                            const tp_grid_fields* g_host, int axis, int formula,
                            uint64_t* status_host);
/* bytes moved by the last host call (for bench.py's e2e accounting) */
/* Page-lock (cudaHostRegister) / unlock a caller-owned host range so that the host entry points above copy it by
 * asynchronous DMA.  Already page-locked memory is accepted as is.  The caller keeps the range alive and unpins it
 * before freeing it.  Returns 0, a cudaError_t (not fatal: unpinned arrays still work, slower) or TP_E_*. */
int tp_host_pin(void* ptr, int64_t bytes);
int tp_host_unpin(void* ptr);
int tp_host_last_transfer(int64_t* h2d_bytes, int64_t* d2h_bytes);
int tp_host_release(void);        /* free the internal staging buffers/streams */
/* What the host link of the current device sustains right now: pinned host <-> device cudaMemcpyAsync of `bytes`
 * per direction, `iters` times -- H2D alone, D2H alone, both at once (GB/s per direction).  The ceiling of the
 * host entry points above (48 B in + 48 B out per push); bench.py reports e2e against it. */
int tp_host_link_probe(int64_t bytes, int iters, double* h2d_gbs, double* d2h_gbs, double* duplex_gbs);

/* ---- on-device field update + clock for multi-step graphs ----
 * tp_plane_wave_f64: EMWave.update of the particle_in_field example
 *   (tests/fixtures/particle_in_field/particle_in_field.py:32-34):
 *   E[j*e_stride] = amplitude*cos(2*pi*(-omega*t + k*(r[j]-x0))), t = clock_dev[0].
 * tp_clock_advance_f64: SimulationClock.advance (turbopy/core.py:533-538) on a
 *   2-double device block {time, this_step}: this_step += 1; time = start + dt*this_step.
 * Both read/write the clock on the device, so a captured graph of
 * [plane_wave, gather_push, clock_advance] x K advances time on every replay. */
int tp_plane_wave_f64(const double* r, int64_t ng, double amplitude, double omega, double k, double x0,
                      const double* clock_dev, double* E, int64_t e_stride, tp_stream_t stream);
int tp_clock_advance_f64(double* clock_dev, double start_time, double dt, tp_stream_t stream);

/* ---- particle reordering by grid cell (additive; the reference never reorders) ----
 * Orders the particles by the cell they sit in along `axis`, so that the fused
 * gather on a grid too large for shared memory reads each warp's node records
 * from a few shared lines instead of one line per lane.  STABLE: the permutation
 * equals numpy.argsort(key, kind="stable") with
 *   key = min(trunc(max((x - r_first) * (1/dr), 0)), ncells-1) >> cells_per_bin_log2
 * (NaN -> key 0).  position and momentum are permuted in place (same storage);
 * perm_out (device, n int64, optional) receives new[i] = old[perm_out[i]] so that
 * callers can reorder their own per-particle arrays with tp_permute_f64.
 * scratch: tp_sort_scratch_bytes() bytes of device memory.  n < 2^32. */
int64_t tp_sort_scratch_bytes(int64_t n, int64_t ncells, int cells_per_bin_log2);
int tp_sort_by_cell_f64(const tp_particles* p, int axis, double r_first, double dr, int64_t ncells,
                        int cells_per_bin_log2, void* scratch, int64_t* perm_out, tp_stream_t stream);
/* Local re-sort of an ALREADY nearly cell-ordered ensemble: the same stable order by cell, applied independently
 * inside blocks of tp_resort_block_particles() (1920) consecutive particles -- block 0 = [0, block_offset), block b =
 * [block_offset + (b-1)*1920, block_offset + b*1920) -- in one read and one write of the particle state (the cost
 * of a push step; the global sort above gathers every component through a random permutation and costs ~18 of
 * them).  Alternate block_offset = 0 / 960 between calls so that particles migrate across block boundaries.  Per
 * block the result equals numpy.argsort(min(key - min(key), 2^20-1), kind="stable").  SoA layout only
 * (stride_n = 1, even stride_c, 16-byte aligned arrays), else TP_E_LAYOUT; block_offset even, < 1920. */
int tp_resort_blocks_f64(const tp_particles* p, int axis, double r_first, double dr, int64_t ncells,
                         int64_t block_offset, tp_stream_t stream);
int64_t tp_resort_block_particles(void);
/* array[i*stride_n] <- array[perm[i]*stride_n] for i < n; tmp: n doubles of device scratch */
int tp_permute_f64(double* array, int64_t stride_n, int64_t n, const int64_t* perm, double* tmp, tp_stream_t stream);

/* ---- CUDA-graph capture of a sequence of the calls above ------------------ */
int tp_graph_begin(tp_stream_t stream);
int tp_graph_end(tp_stream_t stream, void** graph_exec);
int tp_graph_launch(void* graph_exec, tp_stream_t stream);
int tp_graph_destroy(void* graph_exec);

/* ---- launch accounting and self tests -------------------------------------- */
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t tp_launch_count(void);
/* device self-test of the shared-reciprocal division ("Arithmetic
 * contract") against IEEE a/b on n pseudo-random pairs, including zeros,
 * infinities and operands outside the fast path's domain: *mismatches = pairs
 * the fast path accepted but got wrong (must be 0), *fast_taken = pairs it
 * accepted (may be NULL) */
int tp_selftest_division(int64_t n, uint64_t seed, int64_t* mismatches, int64_t* fast_taken,
                         tp_stream_t stream);
/* same for the branch-free in-range reciprocal / square root of the fast path
 * against IEEE 1.0/b and sqrt(a) */
int tp_selftest_roots(int64_t n, uint64_t seed, int64_t* mismatches, int64_t* fast_taken,
                      tp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TURBOPY_B200_H */
