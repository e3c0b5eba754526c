/* qkan_b200.h - C ABI of the B200-native batched QKANLayer.forward.
 *
 * The reference (javiergonzalez10upf/QKAN_Implementation) is pure Python and has no FFI;
 * each entry point below names the reference interface it stands in for
 * (paths relative to QKAN_Steps_original/).  Plain pointers and sizes only - no torch
 * types cross this boundary.  Every function returns QKAN_OK (0) or a negative
 * qkan_status; nothing throws.  qkan_last_error() gives the message of the last failure
 * on the calling thread.  See INTEGRATION.md for the ctypes binding a maintainer of the
 * reference would add.
 */
#ifndef QKAN_B200_H
#define QKAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QKAN_OK = 0,
    QKAN_ERR_BAD_SHAPE = -1,     /* N, K < 1, D < 0, B < 0, null pointers                      */
    QKAN_ERR_UNSUPPORTED = -2,   /* no kernel for this (N, K, D, dtype, mode): D > 31, ...     */
    QKAN_ERR_WEIGHT_RANGE = -3,  /* some |w| > 1 (MulStep.py:36-37 raises ValueError)          */
    QKAN_ERR_CUDA = -4,          /* CUDA runtime / launch failure, message in qkan_last_error  */
    QKAN_ERR_NO_WEIGHTS = -5     /* forward before set_weights                                  */
} qkan_status;

/* amplitude representation of the simulated statevector */
#define QKAN_COMPLEX128 0   /* default: complex double                                          */
#define QKAN_COMPLEX64 1    /* complex float (inputs stay float64)                               */
#define QKAN_REAL64 2       /* real double: every gate of the circuit is real, imag == 0 exactly */

/* term polynomial */
#define QKAN_MODE_COMPAT 0  /* reference semantics: T_D(x) for every LCU term (MulStep.py:20,59) */
#define QKAN_MODE_PAPER 1   /* T_d(x) for term d                                                 */

/* state preparation */
#define QKAN_PREP_ANALYTIC 1 /* H^(x)(m+l)|0> written in closed form (default)                   */
#define QKAN_PREP_GATES 0    /* start from |0...0>, run the initial Hadamards as passes          */

typedef struct qkan_layer qkan_layer;   /* one QKANLayer(N, K, max_degree) on one device */

/* QKANLayer.__init__ (QKANLayer.py:13-28).  device = CUDA ordinal. */
int qkan_layer_create(qkan_layer** out, int N, int K, int max_degree, int dtype, int mode, int prep, int device);
void qkan_layer_destroy(qkan_layer* layer);

/* MulStep.set_weights for all degrees at once (MulStep.py:24-39; called from
 * QKANLayer.forward, QKANLayer.py:124-125).  W is float64 [max_degree+1, N*K] row-major;
 * on_device != 0: W is a device pointer on the layer's device, else a host pointer.
 * validate != 0: returns QKAN_ERR_WEIGHT_RANGE if any |w| > 1 (synchronises the stream).
 * Builds the device rotation tables used by every later forward. */
int qkan_layer_set_weights(qkan_layer* layer, const double* W, int on_device, int validate, void* cuda_stream);

/* QKANLayer.forward (QKANLayer.py:77-135) over a batch, device pointers.
 *   x    float64 [B, N] row-major                      (values outside [-1,1] are clipped,
 *                                                        ChebyshevStep.py:52, and counted)
 *   out  float64 [B, K]
 *   amps optional (may be NULL): post-selected amplitudes [B, K], complex double for
 *        QKAN_COMPLEX128 / QKAN_REAL64, complex float for QKAN_COMPLEX64.
 * Asynchronous on cuda_stream (a cudaStream_t, NULL = default stream). */
int qkan_layer_forward(qkan_layer* layer, const double* x, int64_t B, double* out, void* amps, void* cuda_stream);

/* Forward fused with the output gather of a multi-GPU run: rows [row_offset, row_offset + B)
 * of the full result are stored by the kernel itself into EVERY buffer of out_ptrs - out_ptrs[0] is the
 * local [B_total, K] float64 buffer, the others are the same buffer of the NVLink peers mapped into this
 * process (e.g. torch symmetric memory) - so the transfer overlaps the arithmetic store by store and no
 * separate collective runs.  The caller synchronises the ranks afterwards (barrier).  Block engine only. */
int qkan_layer_forward_peers(qkan_layer* layer, const double* x, int64_t B, void* const* out_ptrs, int n_ptrs,
                             int64_t row_offset, void* cuda_stream);

/* The same through an NVLink multicast (NVLS) mapping of the result buffer: mc_out is the multicast address of
 * the [B_total, K] buffer (e.g. torch symmetric memory's multicast_ptr); the kernel issues ONE multimem.st per
 * result and the NVSwitch replicates it into every rank's memory, the local one included. */
int qkan_layer_forward_multicast(qkan_layer* layer, const double* x, int64_t B, void* mc_out, int64_t row_offset,
                                 void* cuda_stream);

/* Same call with HOST buffers (QKANLayer.forward with NumPy arrays, QKANLayer.py:77).  The batch is cut in chunks of about
 * 8 MiB of traffic (at most 16, never less than one wave of the forward kernel) that overlap on three streams.  Pinned
 * (page-locked, device-mapped) buffers, the default: the DMA engine copies x chunk by chunk while the kernel of the previous
 * chunk stores its results straight into the host buffer (no result staging copy); the pipeline of a call is replayed as a
 * CUDA graph while the caller keeps passing the same buffers.  QKAN_HOST_PATH = zero_copy (the kernel also reads x from host
 * memory itself, one launch, no chunks) | staged (H2D copy, kernel, D2H copy; always for pageable buffers) | copy_out selects
 * the other combinations.  Synchronous: returns when `out` (and `amps`) are complete. */
int qkan_layer_forward_host(qkan_layer* layer, const double* x, int64_t B, double* out, void* amps);

/* The chunk schedule qkan_layer_forward_host uses for a batch of B samples of an N -> K layer whose forward kernel works in
 * tiles of `tile_samples` samples and fills the GPU with `wave_samples` samples: cuts[0] = 0 < cuts[1] < ... < cuts[n] = B
 * (cuts holds max_chunks + 1 entries), returns n >= 1 or a negative error code.  About 8 MiB of traffic per chunk, at most 16
 * chunks, never less than one wave, boundaries on tile multiples.  Pure host arithmetic (no GPU needed). */
int qkan_plan_host_chunks(int64_t B, int N, int K, int64_t tile_samples, int64_t wave_samples, int64_t* cuts, int max_chunks);

/* Number of x entries seen outside [-1-1e-8, 1+1e-8] since the last call (the reference
 * prints them, ChebyshevStep.py:46-49).  Synchronises the device; resets the counter. */
int qkan_layer_out_of_range(qkan_layer* layer, uint64_t* count);

/* Diagonals of QKANLayer.get_intermediate_matrices (QKANLayer.py:52-66) for a batch, device
 * pointers, any output may be NULL:  cheb [B, N*K] = diag of create_dilated_chebyshev
 * (ChebyshevStep.py:55-65), weighted [B, D+1, N*K] = diag of get_weighted_polynomial_matrix
 * (MulStep.py:41-72), lcu [B, N*K] = diag of get_combined_matrix (LCUStep.py:18-37). */
int qkan_layer_diagonals(qkan_layer* layer, const double* x, int64_t B, double* cheb, double* weighted,
                         double* lcu, void* cuda_stream);

/* The same three stages read out of the SIMULATED circuit (QKANLayer.py:52-66): the post-selected (f_x, f_w) = (0, 0)
 * block amplitudes after the CHEB sequence (cheb), after the SELECT rotation of every degree term (weighted) and after
 * the degree sum (lcu), produced by the evolution functions the forward kernels run - scaled rotations + SELECT for
 * compat mode with 1 <= D <= 16, plain (cos, sin) rotations otherwise - instead of the closed form cos(D arccos x). */
int qkan_layer_stage_snapshots(qkan_layer* layer, const double* x, int64_t B, double* cheb, double* weighted,
                               double* lcu, void* cuda_stream);

/* Description of the kernel the layer resolved to, and its work per sample. */
typedef struct {
    int engine;                 /* 0 = block engine (prep = analytic, default); 1 = staged full-statevector
                                   engine (prep = gates)                                                  */
    int n_a, n_b, l, qubits;    /* register sizes ceil(log2 N), ceil(log2 K), ceil(log2 (D+1)); total     */
    /* block engine */
    int blocks;                 /* N*K*(D+1) live four-amplitude (f_x, f_w) blocks per sample             */
    int unroll;                 /* U: blocks (4U amplitudes) per lane in registers                        */
    int samples_per_lane;       /* SU: samples a lane evolves at a time (they share every table entry)    */
    int lanes_per_sample, lanes_per_row, rows_in_parallel, passes, row_steps;
    /* staged engine */
    int tile_qubits, tile_na, tile_nb, local_qubits, stages, sectors_total, sectors_run;
    /* launch */
    int threads_per_cta, min_ctas_per_sm, samples_per_cta;
    int grid, smem_bytes;       /* of the most recent launch (0 before the first)                         */
    /* work per sample */
    int passes_survey, passes_exec;
    int scaled_rotations;       /* form of the CHEB passes (block engine, compat mode, 1 <= D <= 16):
                                   2: sin-weighted basis (u, s v): entry (c, 1 - c^2), no square root (D <= 8);
                                   1: scaled form Ry = gamma * M(t), one FMA per real output, gamma^D and the
                                      quarter turns deferred to the last pass;
                                   0: plain (cos, sin) rotations (generic kernels)                            */
    int input_window;           /* > 0: window kernel (wide input rows) - rotation entries are built per row step
                                   from this many inputs instead of once per sample from all N                 */
    int element_owner;          /* 1: element-owner kernel (wide input rows) - the lanes of a row own its input elements: each
                                   loads x[n], runs CHEB in registers and applies SELECT to the K (D+1) blocks the element
                                   feeds; no shared memory, samples_per_lane samples share every SELECT entry        */
    int degree_factored;        /* 1: a-major kernels - blocks that share a CHEB evolution share its arithmetic: the D + 1
                                   degree copies of an (a, b) block (the state is (block) (x) |+>_deg until SELECT) and
                                   the blocks that read the same input element x[(a + N b) / K] (repeated entries of
                                   the multiplexor's angle table).  CHEB runs once per input element, SELECT per block:
                                   8 D N + 4 N K (D+1) FP instructions per sample instead of N K (D+1)(8 D + 4)   */
    int cheb_elements;          /* CHEB evaluations per sample: N (window kernel: the sum of the row steps' windows;
                                   direct kernel: K)                                                            */
    int direct_rows;            /* 1: direct kernel - every output row reads one input element (K a multiple of N): the
                                   lane of the row evaluates it in registers, no shared memory / barriers / tiles  */
    double flops_survey;        /* 6 * 2^qubits * ((D+1) + m + 2l + n_a)                     */
    double flops_per_block_basis; /* round-1 accounting: every (a, b, d) block evolved on its own through the whole
                                   sequence, (16 D + 4) flops each in the scaled form - reported so that throughput can
                                   be compared on the old basis; equals flops_exec when degree_factored == 0      */
    double flops_exec;          /* arithmetic the kernel executes (DFMA = 2, DMUL = DADD = 1)             */
    double fp_inst_exec;        /* FP64 (FP32 for complex64) lane-instructions behind flops_exec          */
    double layout_efficiency;   /* live block slots / issued block slots (block engine)                   */
    double io_bytes;            /* 8 N + 8 K                                                              */
} qkan_kernel_info;
int qkan_layer_info(qkan_layer* layer, qkan_kernel_info* info);

/* One-shot convenience in the shape SURVEY.md 8(b) proposes: device pointers, builds (and
 * caches per thread) a layer, sets weights without validation, runs forward. */
int qkan_forward(const void* x, const void* w, void* out, int64_t B, int N, int K, int D,
                 int dtype, int mode, void* amps, void* cuda_stream);

/* Gate-list statevector simulation (what the reference's unit tests do with Qiskit Aer's
 * unitary_simulator: MulStep.py:115-166, LCUStep.py:69-107, SUMStep.py:40-78).  Device pointers.
 *   gates   int[n_gates][3] = {kind, q0, q1}; kind 0 H(q0), 1 RY(q0, params[g]), 2 CX(control q0, target q1),
 *           3 SWAP(q0, q1), 4 X(q0), 5 Z(q0); qubit 0 = least significant bit of the amplitude index
 *   basis   int64[n_states]: initial basis state of each run
 *   state_out complex128 [n_states, 2^n_qubits]: evolved states (row j = column `basis[j]` of the unitary) */
int qkan_simulate_circuit(const int* gates, const double* params, int n_gates, int n_qubits,
                          const long long* basis, int64_t n_states, void* state_out, void* cuda_stream);

/* ---- DegreeOptimizer.evaluate_degree (original_degree_optimizer/DegreeOptimizer.py:122-158).
 * Chebyshev features T_k(clip(x, -1, 1)), k = 0..D (ChebyshevStep.py:32-53) of x [n, F] (row-major, device),
 * columns in the reference's np.hstack order: column k F + f; P = F (D+1).
 *
 * qkan_cheb_gram: G [(P+1), (P+1)] (device, row-major) = A^T A of the augmented matrix A = [X_D | y]: its
 *   leading F (d+1) block is X_d^T X_d for every d <= D, column P holds X_D^T y, and G[P][P] = y^T y.  Fused
 *   feature generation + FP64 tensor-core (DMMA) SYRK; partial tiles are summed in a fixed order, so the result is
 *   deterministic.  `workspace` (device) must hold qkan_cheb_gram_workspace() bytes.  0 <= D <= 16.
 *   The F all-ones T_0 columns (DegreeOptimizer.py:96-119: transforms[0] = ones) are computed once; G is the full matrix.
 * qkan_cheb_residuals: explicit residuals r_d = y - X_d c_d of all D+1 fits in one pass.  coef [D+1][P] (device;
 *   zero where a column's degree exceeds d), w [n] sample weights or NULL, ybar = mean(y).  Per CTA c (count from
 *   qkan_cheb_residuals_ctas) partial sums, to be added up by the caller in CTA order:
 *     sums [ctas][D+1][2] = {sum r_d^2, sum w r_d^2};   tail [ctas][4] = {sum (y - ybar)^2, sum w y^2, sum w, sum y};
 *     xtr  [ctas][D+1][P] = X_D^T r_d (NULL to skip; used for one step of iterative refinement).
 * qkan_cheb_features: out [D+1][n][F] = the reference's transforms[d] (DegreeOptimizer.py:96-119). */
int qkan_cheb_gram_workspace(int64_t n, int F, int D, int64_t* bytes, int* slices);
int qkan_cheb_gram(const double* x, const double* y, int64_t n, int F, int D, double* G, void* workspace,
                   int64_t workspace_bytes, void* cuda_stream);
int qkan_cheb_residuals_ctas(int* ctas);
int qkan_cheb_residuals(const double* x, const double* y, const double* w, int64_t n, int F, int D, const double* coef,
                        double ybar, double* sums, double* tail, double* xtr, void* cuda_stream);
int qkan_cheb_features(const double* x, int64_t n, int F, int D, double* out, void* cuda_stream);

/* Measured peaks used as roofline denominators: dependent-free FMA chains on every SM.
 * fp64 != 0: DFMA, else FFMA.  Returns TFLOP/s (2 flops per FMA). */
int qkan_measure_fma_peak(int device, int fp64, double* tflops);
/* FP64 tensor-core peak: independent mma.sync.m8n8k4.f64 (DMMA) chains on every SM.  TFLOP/s. */
int qkan_measure_dmma_peak(int device, double* tflops);

const char* qkan_last_error(void);
int qkan_set_last_error(const char* msg);   /* internal: lets the other translation units report */
void qkan_version(int* major, int* minor, int* patch);

#ifdef __cplusplus
}
#endif
#endif /* QKAN_B200_H */
