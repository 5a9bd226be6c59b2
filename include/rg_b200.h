/*
 * rg_b200.h -- C ABI of the B200-native renormalization-group hot path (librg_b200.so).
 *
 * The reference (cchandre/RG) has no FFI of its own: its boundary is the Python API of the `RG`
 * class (RG.py:44-307) and the RG_modules drivers.  This header is what the Python host
 * (rg_b200/rg.py, rg_b200/rg_modules.py) binds through ctypes; every entry point cites the
 * reference interface it replaces.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *  - every call returns int: 0 = ok, <0 = argument error, >0 = CUDA error (cudaError_t);
 *    rg_last_error() returns a thread-local message for the last non-zero return.
 *  - no exceptions or callbacks cross the ABI; the caller owns every buffer it passes.
 *  - array arguments may be HOST or DEVICE pointers (detected with cudaPointerGetAttributes);
 *    host buffers are copied with cudaMemcpyAsync on the plan's stream into a grow-only device arena the
 *    plan owns (no cudaMalloc/cudaFree per call in the steady state; pinned host buffers make the copies
 *    truly asynchronous).  Device-resident inputs written on another stream must be synchronised by the
 *    caller before the call; every call returns after its own stream has drained.
 *  - Hamiltonian coefficient arrays use the reference layout: f[j][nu_1]...[nu_n], C-contiguous
 *    float64, nu_i in wrapped order 0..L,-L..-1 (RG.py:61,66-67); a batch is B such arrays
 *    back to back.  vecjl = (J+1)*(2L+1)^dim, M = (2L+1)^dim.
 *  - a plan owns its device memory and is bound to one device; a plan is not thread-safe,
 *    distinct plans are.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef RG_B200_H
#define RG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_CT_ONESTEP 0  /* RG.expm_onestep  RG.py:173-192 */
#define RG_CT_ADAPT 1    /* RG.expm_adapt    RG.py:194-205 */
#define RG_CT_MULTIPLY 2 /* RG.expm_multiply RG.py:152-171 */

#define RG_NORM_SUM 0 /* RG.py:93 */
#define RG_NORM_MAX 1 /* RG.py:94 */
#define RG_NORM_EUCLIDEAN 2 /* RG.py:95 */
#define RG_NORM_ANALYTIC 3 /* RG.py:96 */

#define RG_WHICH_NORM 0     /* RG.norm(f)                     RG.py:92-97  */
#define RG_WHICH_NORM_INT 1 /* RG.norm_int(h)                 RG.py:244-247 */
#define RG_WHICH_NORM_IMINUS 2 /* norm of f restricted to I^- RG.py:217-218 */

/* internal safety codes written to err[] where the reference would spin forever (quirk Q6) */
#define RG_ERR_TABLE_EXHAUSTED (-99) /* converge needed more rg_map iterates than the tables hold */
#define RG_ERR_WORK_BUDGET (-98)     /* a Hamiltonian exceeded max_liouville liouville evaluations */
#define RG_ERR_SERIES_CAP 1          /* Lie series hit the hard cap: reported like the TolMax abort */

typedef struct rg_plan rg_plan;     /* one per (device, parameter set): RG.__init__ RG.py:51-100 */
typedef struct rg_tables rg_tables; /* per-iterate frequency tables for one starting Omega, RG.py:210-213 */

/* Everything RG.__init__ (RG.py:51-100) precomputes on the host, handed over as tables. */
typedef struct {
    int32_t dim, L, J;
    int32_t norm_choice;  /* RG_NORM_*  (NormChoice, RG.py:92-97) */
    double norm_analytic; /* NormAnalytic */
    double tol_min, tol_max; /* TolMin, TolMax (RG.py:219,223,281) */
    int32_t max_lie;      /* MaxLie (RG.py:219,225) */
    int32_t ct;           /* RG_CT_*  (CanonicalTransformation, RG.py:220) */
    double min_step, abs_tol, rel_tol; /* expm_adapt controls (RG.py:196,201) */
    double eigenvalue;    /* self.Eigenvalue (RG.py:69) */
    const double* omega0_nu; /* [M]      self.omega0_nu (RG.py:71), wrapped order, host */
    const uint8_t* iminus;   /* [vecjl]  self.Iminus    (RG.py:86), host */
    const int32_t* reidx;    /* [M]      gather table of RG.py:87-91,214-215: flat wrapped mode index of
                                         sign(lambda)*N*nu, or -1 where the mask is false, host */
    const double* theta;     /* [n_theta] theta_m, m = 1.. (RG.py:98-100), host */
    int32_t n_theta;
    int32_t device;          /* CUDA device ordinal */
    int32_t max_series;      /* hard cap on Lie-series length (0 = default 100000); the reference has none */
    int64_t max_liouville;   /* work budget per Hamiltonian in liouville evaluations (0 = default 2 000 000) */
} rg_params;

/* per-point work counters (also the roofline numerator: liouville evaluations actually executed) */
typedef struct {
    int64_t liouville; /* RG.liouville evaluations (RG.py:115-117) */
    int64_t lie_steps; /* canonical transformations (RG.py:220) */
    int64_t rg_maps;   /* RG.rg_map calls (RG.py:207-228) */
    int64_t conv_fma_executed; /* warp-lane FMAs issued by the convolution loops (incl. predicated lanes) */
    int64_t conv_fma_useful;   /* FMAs on structurally non-zero, in-range operand pairs */
    int64_t clk_conv;          /* SM clocks inside the bracket convolution, summed over Hamiltonians */
    int64_t clk_epilogue;      /* SM clocks inside the series epilogue (update + norms), summed */
    int64_t clk_total;         /* SM clocks from ticket to result, summed over Hamiltonians */
    int64_t clk_conv_busy;     /* of clk_conv: clocks warp 0 spent on its own pieces (the rest is barrier wait) */
    float kernel_ms;           /* device time of the last batched call (CUDA events on the plan's stream) */
    int32_t launches;          /* kernels launched by the last call */
    int64_t liouville_full;    /* of `liouville`: evaluations on the full (not half-space) schedule, e.g. the 1-norm estimator's */
} rg_stats;

const char* rg_last_error(void);
int rg_device_count(int* n);

int rg_plan_create(const rg_params* params, rg_plan** out);
int rg_plan_destroy(rg_plan* plan);
/* geometry of the compiled schedule (for DESIGN.md / roofline arithmetic / tests) */
int rg_plan_info(const rg_plan* plan, int32_t* threads_per_cta, int32_t* smem_bytes, int32_t* ctas_per_sm,
                 int32_t* n_support, int64_t* fma_useful_per_liouville, int64_t* fma_executed_per_liouville,
                 int32_t* resident_ctas /* Hamiltonians in flight at once: SMs x CTAs per SM */);

/* nrm_tab[k] = ||N^T Omega_k||, Omega_nu_tab[k][M] = Omega_k . nu (wrapped order), k = 0..n_tab,
   computed by the host with the reference's own calls (RG.py:210-213). */
int rg_tables_create(rg_plan* plan, int32_t n_tab, const double* nrm_tab, const double* Omega_nu_tab, rg_tables** out);
int rg_tables_destroy(rg_tables* t);

/* RG_modules.point batched (RG_modules.py:121-125): generate_1Hamiltonian(amps, symmetric=True)
   (RG.py:255-264) followed by converge (RG.py:278-290), one Hamiltonian per CTA, fully on device.
   amps [n][nK] float64; K_flat[nK] flat wrapped index (j*M + mode) of each K entry.
   conv/count/err [n] int32: int(converge(h)), h.count, h.error.  stats may be NULL. */
int rg_region(rg_plan* plan, const rg_tables* tab, int32_t n, const double* amps, const int32_t* K_flat, int32_t nK,
              int32_t* conv, int32_t* count, int32_t* err, rg_stats* stats);

/* rg_region that also reports the work: liouville_per_point (int64 [n], may be NULL) receives the liouville
   evaluations each point needed (the roofline numerator, and the input of load-balance studies). */
int rg_region_work(rg_plan* plan, const rg_tables* tab, int32_t n, const double* amps, const int32_t* K_flat, int32_t nK,
                   int32_t* conv, int32_t* count, int32_t* err, int64_t* liouville_per_point, rg_stats* stats);

/* RG.converge batched (RG.py:278-290).  f [n][vecjl]; k0[n] = iterate index of each input's Omega in
   `tab` (may be NULL = all 0); err_io[n] carries h.error in and out (untouched on the converged branch). */
int rg_converge(rg_plan* plan, const rg_tables* tab, int32_t n, const double* f, const int32_t* k0,
                int32_t* conv, int32_t* count, int32_t* err_io, rg_stats* stats);

/* RG.rg_map applied n_iter times (RG.py:207-228).  f_out [n][vecjl]; err[n] = error of the last map. */
int rg_map(rg_plan* plan, const rg_tables* tab, int32_t n, const double* f, const int32_t* k0, int32_t n_iter,
           double* f_out, int32_t* err, rg_stats* stats);

/* RG.expm_onestep / expm_adapt / expm_multiply (RG.py:152-205) for caller-supplied y = (y0,y1,y2,y3).
   Omega_nu [M] is the RG object's current self.Omega_nu (hidden state, RG.py:213).  err_io carries
   h.error in (used by expm_adapt's accept branch, quirk Q4) and out. */
int rg_expm(rg_plan* plan, int32_t ct, int32_t n, const double* f, const double* y0, const double* y1,
            const double* y2, const double* y3, const double* Omega_nu, double step, double* f_out,
            int32_t* err_io, rg_stats* stats);

/* RG.liouville (RG.py:115-117): out = y0*dA f + conv(y1, Omega_nu f) - conv(y2, dA f). */
int rg_liouville(rg_plan* plan, int32_t n, const double* y0, const double* y1, const double* y2, const double* f,
                 const double* Omega_nu, double* out);

/* RG.approach (RG.py:292-307) for n bracketing pairs at once, the whole bisection on the device: while
   norm_int(f0 - f1) > reldist * norm(f0): mid = (f0 + f1) * 0.5; RG.converge(mid) replaces f0 (converged) or f1
   (diverged) and its count / error.  f0, f1 [n][vecjl] and count0/1, err0/1 [n] are in/out (the caller passes the
   values after RG.approach's `h.error = 0`, RG.py:294-295, and resets err1 afterwards, RG.py:306); levels[n]
   receives the number of midpoints evaluated.  All pairs share the frequency vector of `tab`.  Every pair
   advances at its own pace (one CTA or one cluster per pair). */
int rg_approach(rg_plan* plan, const rg_tables* tab, int32_t n, double* f0, double* f1, double reldist,
                int32_t* count0, int32_t* count1, int32_t* err0, int32_t* err1, int32_t* levels, rg_stats* stats);

/* RG.conv_product (RG.py:102-106): truncated linear convolution over (j, nu). */
int rg_conv_product(rg_plan* plan, int32_t n, const double* f1, const double* f2, double* out);

/* RG.generate_y (RG.py:108-113): y0[n], y1,y2,y3 [n][vecjl]. */
int rg_generate_y(rg_plan* plan, int32_t n, const double* f, const double* Omega_nu, double* y0, double* y1,
                  double* y2, double* y3);

/* RG.norm / RG.norm_int / I^- norm (RG.py:92-97,244-247,217-218): out[n]. */
int rg_norm(rg_plan* plan, int32_t n, const double* f, int32_t which, double* out);

/* RG.sym (RG.py:249-253). */
int rg_sym(rg_plan* plan, int32_t n, const double* f, double* out);

/* N-matrix re-indexing + rescaling, the head of RG.rg_map (RG.py:210-216). */
int rg_reindex_rescale(rg_plan* plan, const rg_tables* tab, int32_t n, const double* f, const int32_t* k0, double* out);

/* RG.compute_m_s(y, n0, tol, ell) (RG.py:130-150): (m*, s) per Hamiltonian for caller-supplied y; n0 and ell enter the
   test of RG.py:135 (the reference's defaults are n0 = 1, ell = 2; `tol` is unused there too).  norm1 (optional)
   receives the 1-norm estimate of L_y the selection started from. */
int rg_compute_m_s(rg_plan* plan, int32_t n, const double* y0, const double* y1, const double* y2,
                   const double* Omega_nu, double n0, double ell, int32_t* m_star, int32_t* scale, double* norm1);

/* RG.pnorm(y, p) (RG.py:119-123): onenormest((L_y)^p)^(1/p) per Hamiltonian, out[n]; the block 1-norm estimator is
   scipy.sparse.linalg.onenormest restated (Higham-Tisseur Alg. 2.4, t = 2, itmax = 5) with the reference's
   rmatvec = -matvec. */
int rg_pnorm(rg_plan* plan, int32_t n, const double* y0, const double* y1, const double* y2,
             const double* Omega_nu, int32_t p, double* out);

/* Host-only introspection of the convolution schedule (no device needed): used by tests to replay the
   schedule in numpy.  mode: 0 = I^- support, all output modes; 1 = dense (caller-supplied y); 2 = I^-
   support, half space nu <= 0 only (exactly symmetric operands).  Arrays are malloc'ed; free with
   rg_debug_free.  A schedule has n_stages (1 or 2) stages, each a partition of the output modes into boxes
   (stage 0 stores, stage 1 adds).  pieces: [n_stages][n_warps*max_pieces][4] = {box, begin, end, rank}, box = -1
   unused (max_pieces = piece slots per warp: 4, more when a warp has to own many boxes).
   entries: [n][2] list words: word 0 = the lanes of the piece's box whose nu - nu1 is on the grid (bit l = lane l: the
   range test of the truncated convolution is the planner's, the kernels only test their bit); word 1 = centred index of
   the source mode nu1 | its record << 16.
   cost[4] = {useful FMAs, issued FMAs (32 lanes x list cost), max per-warp list cost, total list cost}. */
int rg_debug_schedule(int32_t dim, int32_t L, int32_t J, const uint8_t* iminus_wrapped, int32_t mode,
                      int32_t* n_warps, int32_t* n_boxes, int32_t** lane_out /*[NB*32] centred idx*/,
                      int32_t** pieces, int32_t* n_entries, uint32_t** entries /*[n*2]*/,
                      int32_t* n_support, int32_t** supp_rec /*[M]*/, int64_t* cost, int32_t* n_stages,
                      int32_t* max_pieces);
/* Run ends of every piece slot of that schedule: runs[slot][0] = end of the piece's run of complete Taylor products,
   runs[slot][1] = end of its run truncated to LIM_B levels of y (the rest, to the piece's end, is truncated to LIM_A);
   product_rule = 1 if the plan is costed for the product-rule form of the bracket (BIG kernels).  malloc'ed. */
int rg_debug_piece_runs(int32_t dim, int32_t L, int32_t J, const uint8_t* iminus_wrapped, int32_t mode,
                        int32_t** runs, int32_t* n_slots, int32_t* product_rule);
void rg_debug_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
