/*
 * pinn_b200.h -- C ABI of the B200-native PINN residual/loss engine.
 *
 * This is the drop-in boundary for NeuralPDE.jl's PhysicsInformedNN hot path
 * (boundary B1).  The reference has no FFI of its own:
 * its extension point is the object `discretize` returns,
 *     OptimizationProblem(OptimizationFunction(full_loss_function, AutoZygote()), flat_init_params)
 *     (reference src/discretize.jl:776-780),
 * i.e. `f(theta, p)::Real` plus its reverse-mode gradient.  A Julia shim (see
 * INTEGRATION.md) lowers a `PINNRepresentation` to `pinn_problem_desc` once at
 * `discretize` time and then calls the entry points below through
 * `@ccall libpinn_b200.pinn_loss_grad(...)` with `CuPtr`s on CUDA.jl's task-local
 * stream.  All signatures are plain C: pointers, sizes, no torch / C++ types.
 *
 * Reference functions each entry point replaces (file:line in /root/reference):
 *   pinn_create            <- symbolic_discretize's closure construction
 *                             (src/discretize.jl:413-651), Phi (src/pinn_types.jl:79-90)
 *   pinn_set_points[_host] <- train-set placement in get_loss_function
 *                             (src/training_strategies.jl:215-221 Grid, :271-282 Stochastic)
 *   pinn_loss_grad[_host]  <- full_loss_function(theta,p) (src/discretize.jl:567-598)
 *                             + Zygote gradient (src/discretize.jl:778)
 *                             + numeric_derivative (src/pinn_types.jl:445-482; replaced
 *                               by exact forward-mode taps)
 *   pinn_term_residual     <- datafree_{pde,bc}_loss_functions[i](points, theta)
 *                             (src/pinn_types.jl:414-440; probe used by
 *                             test/Forward/forward__ode.jl:134-137)
 *   pinn_comm_init         <- (no counterpart: the reference is single-process)
 *
 * Conventions
 *   - Every function returns 0 on success, nonzero on error; the message is
 *     available from pinn_last_error() (thread-local).  No C++ exception crosses
 *     the ABI.
 *   - theta / grad use the reference's flat ComponentArray layout: per network,
 *     per Dense layer: weight (out x in, column-major) then bias (out); networks in
 *     depvar order; then the `p` block when param_estim=true
 *     (src/discretize.jl:451-465, src/pinn_types.jl:337-351).
 *   - Point sets are d x N column-major (one point = d contiguous scalars), the
 *     reference's train-set layout (src/discretize.jl:226-238).
 *   - Ownership: the caller owns theta / grad / output buffers and device point
 *     buffers passed to pinn_set_points (aliased, not copied, while the handle
 *     lives).  The engine owns its workspaces, host-staged copies and the NCCL
 *     communicator.
 *   - Threading: a handle is not thread-safe; one handle per GPU rank.  All
 *     device work is enqueued on the stream passed in; no hidden device
 *     synchronisation except in the *_host variants.
 */
#ifndef PINN_B200_H
#define PINN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_ABI_VERSION 2

/* hard limits (validated by pinn_create) */
#define PINN_MAX_LAYERS 16   /* Dense layers per network */
#define PINN_MAX_IN 8        /* network input dimension */
#define PINN_MAX_CH 10       /* propagated channels per (term, network): 1 + #first + #second + #third */
#define PINN_MAX_NETS 8
#define PINN_MAX_TAPS 32     /* taps per term */
#define PINN_MAX_INSTR 192   /* residual-program length per term */
#define PINN_MAX_TERMS 32
#define PINN_MAX_PARAMS 16   /* length of the theta.p block */
#define PINN_MAX_DIM 8       /* rows of a term's point matrix */

typedef struct pinn_engine* pinn_handle;

/* scalar type of theta / points / outputs (theta's eltype rules, src/eltype_matching.jl) */
enum { PINN_F32 = 0, PINN_F64 = 1 };

/* arithmetic mode of the layer contractions.
 * Shapes the tcgen05 modes accept (anything else: pinn_create fails with a message, never a silent fallback):
 *   1-output networks, linear last layer, >= 2 Dense layers, PINN_F32, <= 6 taps per term, and either
 *     - every hidden width in {16, 32, 48, 64} and <= 5 propagated channels per (term, network)      [both modes], or
 *     - every hidden width in {64, 128}, at least one hidden->hidden layer; terms that need more than 4 channels
 *       are evaluated in several passes over the same weights (pure second derivatives only)   [PINN_MODE_TC_BF16]. */
enum {
  PINN_MODE_FFMA = 0,      /* CUDA-core FMA in the scalar type (parity mode, fp32 / fp64)   */
  PINN_MODE_TC_BF16 = 1,   /* tcgen05, bf16 operands, fp32 accumulate                        */
  PINN_MODE_TC_SPLIT = 2   /* tcgen05, split-bf16 x2 (3 MMAs per product), fp32 accumulate   */
};

/* activations (Lux Dense: act.(W*x .+ b)) */
enum {
  PINN_ACT_IDENTITY = 0,
  PINN_ACT_TANH = 1,
  PINN_ACT_SIGMOID = 2,
  PINN_ACT_SIN = 3,
  PINN_ACT_SOFTPLUS = 4,
  PINN_ACT_SWISH = 5
};

/* residual-program opcodes.  The program is in SSA form: instruction i defines
 * value i; `a` / `b` name earlier values (or an index for the LOAD ops); the last
 * instruction's value is the residual r = lhs - rhs of the equation, i.e. what the
 * reference's generated function returns per point (src/symbolic_utilities.jl:360-370). */
enum {
  PINN_OP_CONST = 0,  /* imm                                                     */
  PINN_OP_COORD = 1,  /* row `a` of the term's point matrix (cord[[a],:])        */
  PINN_OP_TAP = 2,    /* tap `a` of the term (u(...) or a pure partial of it)    */
  PINN_OP_PARAM = 3,  /* theta.p[a]  (param_estim) -- differentiated             */
  PINN_OP_ADD = 4,
  PINN_OP_SUB = 5,
  PINN_OP_MUL = 6,
  PINN_OP_DIV = 7,
  PINN_OP_NEG = 8,
  PINN_OP_POW = 9,    /* v[a] ^ v[b]                                             */
  PINN_OP_POWI = 10,  /* v[a] ^ (int)imm                                         */
  PINN_OP_SIN = 11,
  PINN_OP_COS = 12,
  PINN_OP_EXP = 13,
  PINN_OP_LOG = 14,
  PINN_OP_TANH = 15,
  PINN_OP_SQRT = 16,
  PINN_OP_ABS = 17
};

typedef struct {
  int32_t op;
  int32_t a;
  int32_t b;
  int32_t _pad;
  double imm;
} pinn_instr;

/* one Dense MLP (Lux.Chain of Dense layers), one per dependent variable */
typedef struct {
  int32_t n_layers;      /* number of Dense layers                         */
  const int32_t* dims;   /* n_layers+1 entries: in, hidden..., out         */
  const int32_t* acts;   /* n_layers entries: PINN_ACT_* of each layer      */
  int64_t theta_offset;  /* start of this network's block inside theta      */
} pinn_net_desc;

/* a tap = u_k or one pure partial derivative of it, as produced by
 * _transform_expression (src/symbolic_utilities.jl:150-201).  dir[] index the
 * network's own input vector (dict_interior_indvars), not the point rows. */
typedef struct {
  int32_t net;     /* network (depvar) index                               */
  int32_t out;     /* output component of the network (0 for 1-output nets) */
  int32_t order;   /* 0, 1, 2 (any pair of directions) or 3 (pure: d^3/dx_i^3, PINN_MODE_FFMA) */
  int32_t dir[4];  /* derivative directions, `order` entries used            */
} pinn_tap_desc;

enum { PINN_REDUCE_MEAN = 0, /* mean(abs2, r)            training_strategies.jl:220 */
       PINN_REDUCE_WSUM = 1  /* scale * sum(w .* abs2(r)) fixed-node quadrature      */ };

typedef struct {
  int32_t dim;                 /* rows of the point matrix                              */
  int32_t n_taps;
  const pinn_tap_desc* taps;
  const int32_t* net_rows;     /* [n_nets][PINN_MAX_IN]: point row feeding input j of
                                  network k (cord_k = vcat(...), discretize.jl:111-116);
                                  ignored for networks the term does not tap            */
  int32_t n_instr;
  const pinn_instr* prog;
  int32_t reduction;           /* PINN_REDUCE_*                                         */
  double scale;                /* PINN_REDUCE_WSUM: multiplies the weighted sum (1/area) */
} pinn_term_desc;

typedef struct {
  int32_t abi_version;         /* PINN_ABI_VERSION                                      */
  int32_t dtype;               /* PINN_F32 / PINN_F64                                   */
  int32_t mode;                /* PINN_MODE_*                                           */
  int32_t device;              /* CUDA device ordinal                                   */
  int32_t n_nets;
  const pinn_net_desc* nets;
  int32_t n_terms;             /* PDE terms then BC terms, in full_loss_function order   */
  const pinn_term_desc* terms;
  int32_t n_params;            /* length of theta.p (0 unless param_estim)               */
  int64_t param_offset;        /* start of theta.p inside theta                          */
  int64_t n_theta;             /* total length of theta                                  */
} pinn_problem_desc;

/* ---- lifecycle ---------------------------------------------------------------- */
int pinn_create(const pinn_problem_desc* desc, pinn_handle* out);
int pinn_destroy(pinn_handle h);
const char* pinn_last_error(void);
int pinn_abi_version(void);

/* ---- point sets ---------------------------------------------------------------- */
/* Alias a device-resident d x n point matrix (and optional n quadrature weights).   */
int pinn_set_points(pinn_handle h, int32_t term, const void* dev_pts, int64_t n,
                    const void* dev_weights /* nullable */);
/* Copy a host d x n matrix into engine-owned device memory (Stochastic resampling:
 * host rand + upload every call, training_strategies.jl:277-281).  Async on `stream`. */
int pinn_set_points_host(pinn_handle h, int32_t term, const void* host_pts, int64_t n,
                         const void* host_weights /* nullable */, void* stream);
/* Number of points the mean is taken over when the term is sharded across ranks
 * (defaults to the local n). */
int pinn_set_global_count(pinn_handle h, int32_t term, int64_t n_global);

/* ---- the hot path --------------------------------------------------------------- */
/* loss and gradient of  sum_k w[k] * L_k(theta)  (discretize.jl:582-588).
 *   dev_theta       [n_theta]  device
 *   host_weights    [n_terms]  host doubles (adaptive-loss weights; NULL = all 1)
 *   dev_grad        [n_theta]  device, nullable (loss only)
 *   dev_term_losses [n_terms]  device: unweighted L_k (for logging / adaptive weights)
 *   dev_total       [1]        device
 * With a communicator attached, grad / losses are summed over ranks (one allreduce). */
int pinn_loss_grad(pinn_handle h, const void* dev_theta, const double* host_weights,
                   void* dev_grad, void* dev_term_losses, void* dev_total, void* stream);

/* Same, through HOST buffers: copies theta in, runs, copies grad / losses / total out
 * and synchronises.  This is the end-to-end call a CPU-resident optimizer makes. */
int pinn_loss_grad_host(pinn_handle h, const void* host_theta, const double* host_weights,
                        void* host_grad, void* host_term_losses, void* host_total);

/* residual vector r[n] of one term at its current point set (parity probe) */
int pinn_term_residual(pinn_handle h, int32_t term, const void* dev_theta, void* dev_r,
                       void* stream);
int pinn_term_residual_host(pinn_handle h, int32_t term, const void* host_theta, void* host_r);

/* ---- device-side StochasticTraining sampler -----------------------------------------
 * The reference draws `lb .+ (ub .- lb) .* rand(T, d, n)` on the host and uploads it on EVERY loss evaluation
 * (generate_random_points src/training_strategies.jl:242-245, get_loss_function :271-282).  pinn_set_sampler registers
 * a term's box (one [lb, ub] per point row) and point count and draws the first sample into engine-owned memory;
 * pinn_resample draws the next sample of every registered term (one small Philox4x32-10 kernel per term, counter =
 * (point, row group, draw), key = seed + term), so a training iteration has no host round trip.  pinn_adam_iterate
 * resamples before every step when samplers are registered.  pinn_get_points_host copies a term's current points out
 * (callbacks, tests).  The random stream necessarily differs from Julia's default RNG; the distribution is the same. */
int pinn_set_sampler(pinn_handle h, int32_t term, int64_t n, const double* host_lb, const double* host_ub, uint64_t seed,
                     void* stream);
/* QuasiRandomTraining on the device: kind = PINN_SAMPLER_LHS draws a Latin hypercube sample per call -- the reference's
 * default `sampling_alg = LatinHypercubeSample()` (src/training_strategies.jl:285-334; QuasiMonteCarlo.sample on the host
 * + upload per call, :365-389).  Each row's n strata hold exactly one point per draw: stratum index = a keyed Feistel
 * permutation of the point index, position inside the stratum uniform (Philox).  pinn_set_sampler == kind UNIFORM. */
enum { PINN_SAMPLER_UNIFORM = 0, PINN_SAMPLER_LHS = 1 };
int pinn_set_sampler_ex(pinn_handle h, int32_t term, int32_t kind, int64_t n, const double* host_lb, const double* host_ub,
                        uint64_t seed, void* stream);
int pinn_resample(pinn_handle h, void* stream);
int pinn_get_points_host(pinn_handle h, int32_t term, void* host_pts);

/* max |dL_term/dtheta_i| and mean |dL_term/dtheta_i| of ONE term's unweighted loss (the other terms enter with
 * weight 0): what GradientScaleAdaptiveLoss needs per reweighting (reference src/adaptive_losses.jl:115-123 calls
 * Zygote.gradient(pde_loss_function, theta) per term and takes maximum(abs, .) / mean(abs, .) on the host).
 * One fused evaluation + one reduction launch; only the two scalars cross to the host.  Synchronises the stream. */
int pinn_term_grad_stats(pinn_handle h, int32_t term, const void* dev_theta, double* host_max_abs,
                         double* host_mean_abs, void* stream);
int pinn_term_grad_stats_host(pinn_handle h, int32_t term, const void* host_theta, double* host_max_abs,
                              double* host_mean_abs);

/* ---- device-resident optimizer loop ------------------------------------ */
/* Adam (Optimisers.Adam semantics: m, v, bias-corrected step) applied in the TAIL of the fused kernel, right after
 * the in-kernel gradient reduction (and, at nranks > 1, the peer-memory sum): a training iteration is ONE launch
 * (+ one sampler launch per sampled term, + the weight-pack launch on the 128-wide path) with no host round trip.
 * theta, m, v, the step counter and the sampler draw counter live in engine-owned device memory, so
 * pinn_adam_iterate captures its n_steps iterations once into a CUDA graph and replays it while the arguments
 * stay the same (PINN_B200_NO_GRAPH=1 disables the capture).  Point sets: fixed (Grid / fixed-node quadrature /
 * non-resampled) or drawn by the device-side samplers.  Multi-GPU: every rank applies the identical update to its
 * replica (needs the peer-memory path, see pinn_comm_info).  The reference's per-iteration host loop
 * (Optimization.solve + Zygote + Optimisers.Adam) is what this replaces. */
int pinn_adam_begin(pinn_handle h, const void* host_theta0, double lr, double beta1, double beta2, double eps);
/* run n_steps iterations; host_total (nullable) receives the loss of the LAST evaluated theta,
 * host_term_losses (nullable) its per-term losses.  Synchronises at the end. */
int pinn_adam_iterate(pinn_handle h, int32_t n_steps, const double* host_weights, void* host_total,
                      void* host_term_losses);
int pinn_adam_theta(pinn_handle h, void* host_theta_out);

/* ---- multi-GPU -------------------------------------------------------------------- */
/* Attach an NCCL communicator built from a 128-byte ncclUniqueId that the caller
 * distributed (rank 0 obtains it from pinn_comm_unique_id). */
int pinn_comm_unique_id(void* out_128_bytes);
int pinn_comm_init(pinn_handle h, const void* unique_id_128_bytes, int32_t rank, int32_t nranks);
/* pinn_comm_init also maps every rank's symmetric gradient buffer into this process (CUDA IPC over NVLink peer
 * access, handles exchanged through the communicator).  When that succeeds on all ranks the sum over ranks runs INSIDE
 * the fused kernel: each CTA publishes its slice of the reduced gradient, signals per-slice flags in the peers'
 * memory and adds the peers' slices in rank order -- no ncclAllReduce, no extra launch, identical bits on every
 * rank.  Otherwise (GPUs without peer access, ranks that are threads of one process, PINN_B200_NO_P2P=1) the step is
 * fused kernel -> ncclAllReduce -> unpack.  pinn_comm_info reports which: *fused_p2p = 1 / 0, and returns the reason
 * for a fallback ("" when none).  All ranks must issue the same sequence of pinn_loss_grad / pinn_adam_iterate calls
 * (SPMD), and synchronise with each other before destroying their handles. */
const char* pinn_comm_info(pinn_handle h, int32_t* fused_p2p);

/* ---- introspection ------------------------------------------------------------------ */
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t pinn_launch_count(pinn_handle h);
/* device time (ms) of the main fused kernel in the most recent pinn_loss_grad, measured
 * with CUDA events on the launching stream; requires pinn_set_timing(h, 1). Blocks. */
int pinn_set_timing(pinn_handle h, int32_t enable);
double pinn_last_kernel_ms(pinn_handle h);
/* bytes of device workspace owned by the handle */
int64_t pinn_workspace_bytes(pinn_handle h);
/* algorithmic FLOPs of one pinn_loss_grad at the current point sets:
 * 6 * sum_terms N * sum_nets C * S   */
double pinn_flops_per_eval(pinn_handle h);

#ifdef __cplusplus
}
#endif
#endif /* PINN_B200_H */
