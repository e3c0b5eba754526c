// qkan_capi.cu - C ABI (include/qkan_b200.h): layer handle, weight tables, kernel selection,
// device and host-buffer forward, FMA peak microbenchmark.
#include "qkan_kernel.cuh"
#include "qkan_block.cuh"
#include "qkan_amajor.cuh"
#include "qkan_instances.h"
#include "../../include/qkan_b200.h"

#include <cstdio>
#include <nvtx3/nvToolsExt.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace qkan;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(QKAN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                  \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);       \
    } while (0)

const std::vector<KernelInfo>& registry() {
    static std::vector<KernelInfo> reg;
    static std::once_flag once;
    std::call_once(once, [] { qkan_register_all(reg); });
    return reg;
}

const std::vector<BlockKernelInfo>& block_registry() {
    static std::vector<BlockKernelInfo> reg;
    static std::once_flag once;
    std::call_once(once, [] { qkan_register_all_block(reg); });
    return reg;
}

int clog2(int n) {
    int r = 0;
    while ((1 << r) < n) ++r;
    return r;
}

constexpr int MAX_CHUNKS = 64;

// NVTX range around the host side of an entry point (visible in Nsight Systems / ncu --nvtx; header-only NVTX 3: no cost
// when no tool is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// every entry point runs on the layer's device and leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;               // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev)                                            \
    DeviceGuard _guard(dev);                                      \
    if (_guard.err != cudaSuccess) return cuda_fail(_guard.err, "cudaSetDevice")

__global__ void qkan_check_weights_kernel(const double* W, long long n, unsigned long long* bad) {
    unsigned cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (!(fabs(W[i]) <= 1.0)) ++cnt;
    if (cnt) atomicAdd(bad, (unsigned long long)cnt);
}

}  // namespace

struct qkan_layer {
    int N, K, D, NA, NB, L;
    int dtype, mode, prep, device, sm_count;
    int engine = 0;                           // 0 = block engine, 1 = staged full-statevector engine
    const KernelInfo* kern = nullptr;         // staged
    const BlockKernelInfo* bkern = nullptr;   // block
    BlockLayout lay{};
    int window = 0;                           // window kernel: widest row-step input window (entries per cs row)
    void* wtab = nullptr;
    int* xidx = nullptr;
    unsigned long long* counters = nullptr;   // [0] out-of-range x, [1] |w| > 1
    double* W_dev = nullptr;
    bool weights_set = false;
    int last_grid = 0, last_smem = 0;
    // host-buffer path
    double* d_x = nullptr;
    double* d_out = nullptr;
    void* d_amps = nullptr;
    int64_t cap_x = 0, cap_out = 0, cap_amps = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {};
    cudaEvent_t ev_w = nullptr;               // recorded after the table build of set_weights: the host path's own
                                              // (non-blocking) streams wait on it, whatever stream the caller used
    // host-buffer path: the chunk pipeline (copies + kernels of one call) captured as a CUDA graph, replayed while the
    // caller keeps passing the same buffers (one launch instead of ~7 enqueue calls per chunk)
    struct HostKey { const void* x; void* out; void* amps; int64_t B; int in_direct, out_direct, nchunk; };
    HostKey last_key{}, graph_key{};
    cudaGraphExec_t graph = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join_in = nullptr, ev_join_out = nullptr;
};

static size_t amp_real_size(int dtype) { return dtype == QKAN_COMPLEX64 ? 4 : 8; }

extern "C" int qkan_layer_create(qkan_layer** out, int N, int K, int max_degree, int dtype, int mode, int prep,
                                 int device) {
    if (!out) return fail(QKAN_ERR_BAD_SHAPE, "null output handle");
    *out = nullptr;
    if (N < 1 || K < 1) return fail(QKAN_ERR_BAD_SHAPE, "N and K must be >= 1");
    if (max_degree < 0) return fail(QKAN_ERR_BAD_SHAPE, "Degree must be positive integer.");   // ChebyshevStep.py:14-15
    if (dtype < 0 || dtype > 2 || mode < 0 || mode > 1 || prep < 0 || prep > 1)
        return fail(QKAN_ERR_BAD_SHAPE, "bad dtype / mode / prep");
    const int NA = clog2(N), NB = clog2(K), L = clog2(max_degree + 1);
    const KernelInfo* best = nullptr;
    const BlockKernelInfo* bbest = nullptr;
    BlockLayout lay{};
    int window = 0;
    if (prep == QKAN_PREP_ANALYTIC) {
        // block engine: any N, K, D.  Pick the lane layout and the CTA size whose x tile fits.
        if ((long long)N * K >= (1 << 20) || max_degree >= 2048 || (mode == QKAN_MODE_PAPER && max_degree >= 128))
            return fail(QKAN_ERR_UNSUPPORTED, "block engine limits: N*K < 2^20, D < 2048 (D < 128 in paper mode)");
        const char* tune = getenv("QKAN_BLOCK_TUNE");          // tuning aid: "U:NT:MINB:SU" (0 = any)
        int fU = 0, fNT = 0, fMINB = 0, fSU = 0;
        if (tune) sscanf(tune, "%d:%d:%d:%d", &fU, &fNT, &fMINB, &fSU);
        const int asz = dtype == QKAN_COMPLEX128 ? 16 : 8;     // sizeof(amplitude): complex128 16, complex64 8, real64 8
        // ---- a-major scaled-rotation kernels (qkan_amajor.cuh): compat mode, 1 <= D <= 16
        const bool amajor_ok = mode == QKAN_MODE_COMPAT && max_degree >= TAN_MIN_DT && max_degree <= TAN_MAX_DT &&
                               fU <= 1 && !getenv("QKAN_BLOCK_NO_DT");
        // two samples per lane share every SELECT entry and the per-pass bookkeeping: pays for shallow sequences
        const int want_SU = fSU ? fSU : 2;                     // tile kernel
        const int want_SU_direct = fSU ? fSU : 4;              // direct kernel (profiles/r02e_tune_direct.jsonl)
        auto find_amajor = [&](int NT, int SU, bool window_kernel, bool direct_kernel = false, bool elem_kernel = false) -> const BlockKernelInfo* {
            for (const BlockKernelInfo& k : block_registry()) {
                if (!k.amajor || (k.window != 0) != window_kernel || (k.direct != 0) != direct_kernel || (k.elem != 0) != elem_kernel ||
                    k.amp != dtype || k.DT != max_degree || k.NT != NT) continue;
                if (!window_kernel && k.SU != SU) continue;
                if (fMINB ? (k.MINB != fMINB) : !k.is_default) continue;
                return &k;
            }
            return nullptr;
        };
        // resident warps per SM that the shared memory of one CTA allows (registers cap it at 32 / 24)
        auto warps_for = [](size_t smem, int NT) { return (int)((220 * 1024 / (smem + 1024)) * (size_t)(NT / 32)); };
        // rows that read a single input element: the direct kernel (no shared memory)
        const bool force_elem = getenv("QKAN_BLOCK_FORCE_ELEM") != nullptr;      // A/B aid
        if (amajor_ok && !getenv("QKAN_BLOCK_NO_DIRECT") && !force_elem) {
            const BlockLayout cand = plan_amajor_layout(N, K, 0);
            if (amajor_direct_ok(N, K, cand)) {
                const int NTs[2] = {256, 128};
                for (int ni = 0; ni < 2 && !bbest; ++ni) {
                    if (fNT && NTs[ni] != fNT) continue;
                    for (int SU = want_SU_direct; SU >= 1 && !bbest; SU >>= 1) {
                        const BlockKernelInfo* k = find_amajor(NTs[ni], SU, false, true);
                        if (k) { bbest = k; lay = cand; }
                    }
                }
            }
        }
        // Rows that read several inputs.  Pass 0: the element-owner kernel when its walk wastes little (rows much wider than
        // K: measured faster than the tile kernel on every such shape, profiles/r02j_tune_c4.jsonl), else the tile kernel
        // when shared memory leaves >= 24 resident warps; pass 1: whatever launches.
        for (int pass = 0; pass < 2 && amajor_ok && !bbest; ++pass) {
            if (!getenv("QKAN_BLOCK_NO_ELEM") && (!fNT || fNT == 256)) {
                const ElemLayout el = plan_elem_layout(N, K, 0);
                if (force_elem || el.efficiency >= (pass == 0 ? 0.8 : 0.0)) {
                    for (int SU = fSU ? fSU : 4; SU >= 1 && !bbest; SU >>= 1) {
                        const BlockKernelInfo* k = find_amajor(256, SU, false, false, true);
                        if (k) {
                            bbest = k;
                            lay.U = 1; lay.g_r_log2 = el.g_r_log2; lay.g_k_log2 = el.g_k_log2; lay.passes = el.passes; lay.brows = el.brows;
                            lay.efficiency = el.efficiency;
                        }
                    }
                }
            }
            const int NTs[2] = {256, 128};
            for (int ni = 0; ni < 2 && !bbest && !force_elem; ++ni) {
                const int NT = NTs[ni];
                if (fNT && NT != fNT) continue;
                for (int min_g = 0; min_g <= 5 && !bbest; ++min_g) {
                    const BlockLayout cand = plan_amajor_layout(N, K, min_g);
                    const int G = 1 << (cand.g_r_log2 + cand.g_k_log2);
                    if (G < (1 << min_g) || G > NT) continue;
                    const int SPC = NT / G;
                    const int row_bytes = amajor_row_amps(N + 1, G, asz) * asz;
                    for (int SU = want_SU; SU >= 1 && !bbest; --SU) {
                        const size_t smem = amajor_smem_bytes(N, SPC, row_bytes, SU, SU);     // smallest tile the launch can use
                        if (smem > AMAJOR_SMEM_CAP || (pass == 0 && warps_for(smem, NT) < 24)) continue;
                        const BlockKernelInfo* k = find_amajor(NT, SU, false);
                        if (k) { bbest = k; lay = cand; }
                    }
                }
            }
        }
        // ---- generic (cos, sin) kernels: paper mode, D = 0, D > 16, or rows too wide for the above
        const int NTs[4] = {256, 128, 64, 32};
        for (int ni = 0; ni < 4 && !bbest; ++ni) {
            const int NT = NTs[ni];
            if (fNT && NT != fNT) continue;
            for (int min_g = 0; min_g <= 5 && !bbest; ++min_g) {
                BlockLayout cand = plan_block_layout(N, K, max_degree, min_g, fU);
                int G = 1 << (cand.g_r_log2 + cand.g_k_log2);
                if (G < (1 << min_g) || G > NT) continue;
                auto cs_bytes_for = [&](int) { return (size_t)(NT / G) * (N + 1) * 2 * amp_real_size(dtype); };
                // wide input rows: shared memory limits the resident warps (< 24 per SM), so keep four blocks
                // per lane in flight instead of one (measured on N784 K10 D5: 3.7 -> 4.0 M samples/s)
                if (!fU && (220 * 1024 / (cs_bytes_for(cand.U) + 1024)) * (size_t)(NT / 32) < 24) {
                    const BlockLayout wide = plan_block_layout(N, K, max_degree, min_g, 4);
                    const int Gw = 1 << (wide.g_r_log2 + wide.g_k_log2);
                    if (Gw == G) cand = wide;
                }
                // the launch adds at most 16 KiB of raw-x staging (wider tiles read x directly) to one SPC-row cs tile
                if (cs_bytes_for(cand.U) > 72 * 1024) continue;
                for (const BlockKernelInfo& k : block_registry()) {
                    if (k.amajor || k.window || k.amp != dtype || k.mode != mode || k.U != cand.U || k.NT != NT || k.DT != 0) continue;
                    if (fMINB ? (k.MINB != fMINB) : !k.is_default) continue;
                    bbest = &k;
                    break;
                }
                if (bbest) lay = cand;
            }
        }
        if (getenv("QKAN_DEBUG_SELECT") && bbest)
            fprintf(stderr, "qkan select: N=%d K=%d D=%d dtype=%d -> elem=%d direct=%d amajor=%d U=%d SU=%d NT=%d MINB=%d DT=%d tan=%d window=%d (W=%d) g_r=%d g_k=%d passes=%d rows=%d\n",
                    N, K, max_degree, dtype, bbest->elem, bbest->direct, bbest->amajor, bbest->U, bbest->SU, bbest->NT, bbest->MINB, bbest->DT, bbest->tan, bbest->window, window,
                    lay.g_r_log2, lay.g_k_log2, lay.passes, lay.brows);
        if (!bbest) {
            char buf[160];
            snprintf(buf, sizeof buf, "no block kernel for N=%d K=%d D=%d dtype=%d mode=%d (input row too wide for shared memory?)",
                     N, K, max_degree, dtype, mode);
            return fail(QKAN_ERR_UNSUPPORTED, buf);
        }
    } else {
        int best_prio = 0;
        const char* venv = getenv("QKAN_VARIANT");          // tuning aid: prefer one variant of a tile
        const int want_variant = venv ? atoi(venv) : -1;
        for (const KernelInfo& k : registry()) {
            if (k.amp != dtype || k.mode != mode || k.prep != prep || k.L != L || k.NAT > NA || k.NBT > NB) continue;
            if (k.full && (k.NAT != NA || k.NBT != NB)) continue;
            const int prio = k.prio + ((k.variant == want_variant) ? 100 : 0);
            if (!best) { best = &k; best_prio = prio; continue; }
            const int a = k.NAT + k.NBT, b = best->NAT + best->NBT;
            if (prio > best_prio || (prio == best_prio && (a > b || (a == b && k.NBT > best->NBT)))) { best = &k; best_prio = prio; }
        }
        if (!best) {
            char buf[200];
            snprintf(buf, sizeof buf, "no staged (prep=gates) sm_100a kernel for N=%d K=%d D=%d (l=%d) dtype=%d mode=%d: "
                     "that engine covers D <= 31, compat mode", N, K, max_degree, L, dtype, mode);
            return fail(QKAN_ERR_UNSUPPORTED, buf);
        }
    }
    ON_DEVICE(device);
    qkan_layer* l = new qkan_layer();
    l->N = N; l->K = K; l->D = max_degree; l->NA = NA; l->NB = NB; l->L = L;
    l->dtype = dtype; l->mode = mode; l->prep = prep; l->device = device;
    l->kern = best;
    l->bkern = bbest;
    l->lay = lay;
    l->window = window;
    l->engine = (prep == QKAN_PREP_ANALYTIC) ? 0 : 1;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete l; return cuda_fail(e, "cudaGetDeviceProperties"); }
    l->sm_count = prop.multiProcessorCount;
    const size_t nab = (size_t)1 << (NA + NB);
    const size_t nblk = (size_t)N * K * (max_degree + 1);
    (void)nblk;
    if (l->engine == 0) {
        // + one pass of padding slots: the streaming kernel prefetches one pass ahead
        const size_t G = (size_t)1 << (l->lay.g_r_log2 + l->lay.g_k_log2);
        // (+ 8 more, never read: the window kernel's L1 prefetches run WINDOW_PREFETCH passes ahead)
        const size_t slots = ((size_t)l->lay.brows * l->lay.passes + 1 + 8) * l->lay.U * G;
        // a-major tables: D + 1 SELECT entries per (row step, pass, lane) step
        const size_t per_slot = l->bkern->elem ? (size_t)K * (max_degree + 1) : (l->bkern->amajor ? (size_t)(max_degree + 1) : 1);
        e = cudaMalloc(&l->wtab, slots * per_slot * 2 * amp_real_size(dtype));
        if (e == cudaSuccess) e = cudaMalloc(&l->xidx, slots * sizeof(int));
    } else {
        e = cudaMalloc(&l->wtab, (nab << L) * 2 * amp_real_size(dtype));
        if (e == cudaSuccess) e = cudaMalloc(&l->xidx, nab * sizeof(int));
    }
    if (e == cudaSuccess) e = cudaMalloc(&l->counters, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(l->counters, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&l->W_dev, (size_t)(max_degree + 1) * N * K * sizeof(double));
    if (e != cudaSuccess) { qkan_layer_destroy(l); return cuda_fail(e, "cudaMalloc(layer tables)"); }
    *out = l;
    return QKAN_OK;
}

extern "C" void qkan_layer_destroy(qkan_layer* l) {
    if (!l) return;
    DeviceGuard guard(l->device);
    cudaFree(l->wtab); cudaFree(l->xidx); cudaFree(l->counters); cudaFree(l->W_dev);
    cudaFree(l->d_x); cudaFree(l->d_out); cudaFree(l->d_amps);
    if (l->s_in) cudaStreamDestroy(l->s_in);
    if (l->s_k) cudaStreamDestroy(l->s_k);
    if (l->s_out) cudaStreamDestroy(l->s_out);
    for (int i = 0; i < MAX_CHUNKS; ++i) {
        if (l->ev_in[i]) cudaEventDestroy(l->ev_in[i]);
        if (l->ev_k[i]) cudaEventDestroy(l->ev_k[i]);
    }
    if (l->ev_w) cudaEventDestroy(l->ev_w);
    if (l->graph) cudaGraphExecDestroy(l->graph);
    if (l->ev_fork) cudaEventDestroy(l->ev_fork);
    if (l->ev_join_in) cudaEventDestroy(l->ev_join_in);
    if (l->ev_join_out) cudaEventDestroy(l->ev_join_out);
    delete l;
}

extern "C" int qkan_layer_set_weights(qkan_layer* l, const double* W, int on_device, int validate, void* cuda_stream) {
    NvtxRange nvtx_range("qkan_layer_set_weights");
    if (!l || !W) return fail(QKAN_ERR_BAD_SHAPE, "null layer or weights");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    ON_DEVICE(l->device);
    const size_t nw = (size_t)(l->D + 1) * l->N * l->K;
    if (validate) {
        // |w| <= 1 (MulStep.py:36-37) is checked BEFORE anything is overwritten: a rejected matrix leaves the previous
        // weights and tables in place, like the reference's set_weights, which raises before it assigns
        unsigned long long bad = 0;
        if (on_device) {
            CU(cudaMemsetAsync(l->counters + 1, 0, sizeof(unsigned long long), stream));
            qkan_check_weights_kernel<<<64, 256, 0, stream>>>(W, (long long)nw, l->counters + 1);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(&bad, l->counters + 1, sizeof bad, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
        } else {
            for (size_t i = 0; i < nw; ++i) bad += !(fabs(W[i]) <= 1.0);
        }
        if (bad) return fail(QKAN_ERR_WEIGHT_RANGE, "Weight magnitudes must be <= 1 for unitarity");   // MulStep.py:37
    }
    if (W != l->W_dev)
        CU(cudaMemcpyAsync(l->W_dev, W, nw * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           stream));
    const double* Wd = l->W_dev;
    CU(cudaMemsetAsync(l->counters + 1, 0, sizeof(unsigned long long), stream));
    if (l->engine == 0 && l->bkern->elem) {
        ElemLayout el{l->lay.g_r_log2, l->lay.g_k_log2, l->lay.passes, l->lay.brows, l->lay.efficiency};
        const long long steps = elem_steps(el);
        const unsigned nt = 128, nb = (unsigned)((steps + nt - 1) / nt);
        if (l->dtype == QKAN_COMPLEX64)
            qkan_prepare_elem_tables_kernel<float><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, el.passes, el.brows, el.g_r_log2,
                                                                          el.g_k_log2, steps, (CS<float>*)l->wtab, l->xidx, l->counters + 1);
        else
            qkan_prepare_elem_tables_kernel<double><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, el.passes, el.brows, el.g_r_log2,
                                                                           el.g_k_log2, steps, (CS<double>*)l->wtab, l->xidx, l->counters + 1);
    } else if (l->engine == 0 && l->bkern->amajor) {
        const long long steps = amajor_steps(l->lay);
        const unsigned nt = 128, nb = (unsigned)((steps + nt - 1) / nt);
        if (l->dtype == QKAN_COMPLEX64)
            qkan_prepare_amajor_tables_kernel<float><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, l->lay.passes, l->lay.brows,
                                                                            l->lay.g_r_log2, l->lay.g_k_log2, l->bkern->amp_bytes, 0, steps,
                                                                            (CS<float>*)l->wtab, l->xidx, l->counters + 1);
        else
            qkan_prepare_amajor_tables_kernel<double><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, l->lay.passes, l->lay.brows,
                                                                             l->lay.g_r_log2, l->lay.g_k_log2, l->bkern->amp_bytes, 0, steps,
                                                                             (CS<double>*)l->wtab, l->xidx, l->counters + 1);
    } else if (l->engine == 0) {
        const long long G = 1ll << (l->lay.g_r_log2 + l->lay.g_k_log2);
        // one extra row step of slots decodes to b >= K, i.e. padding: covers the prefetch overrun
        const long long slots = ((long long)l->lay.brows * l->lay.passes + 1) * l->lay.U * G;
        if ((long long)l->N * l->K * (l->D + 1) > slots) return fail(QKAN_ERR_BAD_SHAPE, "internal: slot table smaller than W");
        const unsigned nt = 128, nb = (unsigned)((slots + nt - 1) / nt);
        if (l->dtype == QKAN_COMPLEX64)
            qkan_prepare_block_tables_kernel<float><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, l->lay.U, l->lay.passes,
                                                                           l->lay.g_r_log2, l->lay.g_k_log2, l->mode,
                                                                           8, 0, slots,
                                                                           (CS<float>*)l->wtab, l->xidx, l->counters + 1);
        else
            qkan_prepare_block_tables_kernel<double><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, l->lay.U, l->lay.passes,
                                                                            l->lay.g_r_log2, l->lay.g_k_log2, l->mode,
                                                                            16, 0, slots,
                                                                            (CS<double>*)l->wtab, l->xidx, l->counters + 1);
    } else {
    const unsigned nab = 1u << (l->NA + l->NB);
    const unsigned nt = 128, nb = (nab + nt - 1) / nt;
    if (l->dtype == QKAN_COMPLEX64)
        qkan_prepare_tables_kernel<float><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, l->NA, l->NB, l->L,
                                                                 (CS<float>*)l->wtab, l->xidx, l->counters + 1);
    else
        qkan_prepare_tables_kernel<double><<<nb, nt, 0, stream>>>(Wd, l->N, l->K, l->D, l->NA, l->NB, l->L,
                                                                  (CS<double>*)l->wtab, l->xidx, l->counters + 1);
    }
    CU(cudaGetLastError());
    if (!l->ev_w) CU(cudaEventCreateWithFlags(&l->ev_w, cudaEventDisableTiming));
    CU(cudaEventRecord(l->ev_w, stream));
    l->weights_set = true;
    return QKAN_OK;
}

static int launch_on(qkan_layer* l, const double* x, int64_t B, double* out, void* amps, cudaStream_t stream,
                     void* const* peer_outs = nullptr, int n_peers = 0, int64_t row0 = 0, void* mc_out = nullptr) {
    if (l->engine == 0) {
        BlockParams p;
        p.x = x; p.cstab = l->wtab; p.xotab = l->xidx; p.amps = amps; p.oor = l->counters;
        for (int q = 0; q < 8; ++q) p.outs[q] = nullptr;
        p.outs[0] = out; p.n_out = 1; p.row0 = row0; p.mc_out = (double*)mc_out;
        if (peer_outs) {
            p.n_out = n_peers;
            for (int q = 0; q < n_peers; ++q) p.outs[q] = (double*)peer_outs[q];
        }
        p.B = B; p.N = l->N; p.K = l->K; p.D = l->D;
        p.g_r_log2 = l->lay.g_r_log2; p.g_k_log2 = l->lay.g_k_log2;
        p.passes = l->lay.passes; p.brows = l->lay.brows;
        p.sub = 1; p.tma_ok = 0; p.direct_x = 0; p.window = 0;
        p.plane_bytes = 0;
        p.plain = (p.n_out == 1 && !p.mc_out && !amps) ? 1 : 0;
        for (int q = 0; q < 8; ++q) p.init[q] = 0.0;
        p.init[0] = 1.0;                                   // PREPARE'd block state (1, 0, 0, 0), un-normalised
        p.out_scale = 1.0 / ((double)l->N * (double)(l->D + 1));
        p.amp_scale = pow(2.0, -0.5 * (double)(l->NA + l->NB + 2 * l->L + l->NA));
        cudaError_t e = l->bkern->launch(p, 1 << (l->lay.g_r_log2 + l->lay.g_k_log2), l->sm_count, stream, &l->last_grid,
                                         &l->last_smem);
        if (e != cudaSuccess) return cuda_fail(e, "qkan_block_kernel launch");
        return QKAN_OK;
    }
    LaunchParams p;
    p.x = x; p.wtab = l->wtab; p.xidx = l->xidx; p.out = out; p.amps = amps; p.oor = l->counters;
    p.B = B; p.N = l->N; p.K = l->K; p.D = l->D; p.NA = l->NA; p.NB = l->NB;
    p.out_scale = 1.0 / ((double)l->N * (double)(l->D + 1));
    p.amp_scale = pow(2.0, -0.5 * (double)(l->NA + l->NB + 2 * l->L + l->NA));
    p.tma_ok = 0;
    cudaError_t e = l->kern->launch(p, l->sm_count, stream, &l->last_grid, &l->last_smem);
    if (e != cudaSuccess) return cuda_fail(e, "qkan_forward_kernel launch");
    return QKAN_OK;
}

extern "C" int qkan_layer_forward(qkan_layer* l, const double* x, int64_t B, double* out, void* amps, void* cuda_stream) {
    NvtxRange nvtx_range("qkan_layer_forward");
    if (!l) return fail(QKAN_ERR_BAD_SHAPE, "null layer");
    if (B < 0) return fail(QKAN_ERR_BAD_SHAPE, "negative batch");
    if (!l->weights_set) return fail(QKAN_ERR_NO_WEIGHTS, "forward called before set_weights");
    if (B == 0) return QKAN_OK;
    if (!x || !out) return fail(QKAN_ERR_BAD_SHAPE, "null x / out");
    ON_DEVICE(l->device);
    return launch_on(l, x, B, out, amps, (cudaStream_t)cuda_stream);
}

static int ensure_host_path(qkan_layer* l, int64_t B, bool want_amps) {
    if (!l->s_in) {
        CU(cudaStreamCreateWithFlags(&l->s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&l->s_k, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&l->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            CU(cudaEventCreateWithFlags(&l->ev_in[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&l->ev_k[i], cudaEventDisableTiming));
        }
    }
    if (l->cap_x < B) {
        cudaFree(l->d_x); l->d_x = nullptr; l->cap_x = 0;
        CU(cudaMalloc(&l->d_x, (size_t)B * l->N * sizeof(double)));
        l->cap_x = B;
    }
    if (l->cap_out < B) {
        cudaFree(l->d_out); l->d_out = nullptr; l->cap_out = 0;
        CU(cudaMalloc(&l->d_out, (size_t)B * l->K * sizeof(double)));
        l->cap_out = B;
    }
    if (want_amps && l->cap_amps < B) {
        cudaFree(l->d_amps); l->d_amps = nullptr; l->cap_amps = 0;
        CU(cudaMalloc(&l->d_amps, (size_t)B * l->K * 2 * amp_real_size(l->dtype)));
        l->cap_amps = B;
    }
    return QKAN_OK;
}

extern "C" int qkan_layer_forward_peers(qkan_layer* l, const double* x, int64_t B, void* const* out_ptrs, int n_ptrs,
                                        int64_t row_offset, void* cuda_stream) {
    NvtxRange nvtx_range("qkan_layer_forward_peers");
    if (!l) return fail(QKAN_ERR_BAD_SHAPE, "null layer");
    if (B < 0 || row_offset < 0) return fail(QKAN_ERR_BAD_SHAPE, "negative batch / row offset");
    if (!out_ptrs || n_ptrs < 1 || n_ptrs > 8) return fail(QKAN_ERR_BAD_SHAPE, "need 1..8 result buffers");
    for (int q = 0; q < n_ptrs; ++q) if (!out_ptrs[q]) return fail(QKAN_ERR_BAD_SHAPE, "null result buffer");
    if (l->engine != 0) return fail(QKAN_ERR_UNSUPPORTED, "the fused output gather is implemented by the block engine (prep = analytic)");
    if (!l->weights_set) return fail(QKAN_ERR_NO_WEIGHTS, "forward called before set_weights");
    if (B == 0) return QKAN_OK;
    if (!x) return fail(QKAN_ERR_BAD_SHAPE, "null x");
    ON_DEVICE(l->device);
    return launch_on(l, x, B, (double*)out_ptrs[0], nullptr, (cudaStream_t)cuda_stream, out_ptrs, n_ptrs, row_offset);
}

extern "C" int qkan_layer_forward_multicast(qkan_layer* l, const double* x, int64_t B, void* mc_out, int64_t row_offset,
                                            void* cuda_stream) {
    NvtxRange nvtx_range("qkan_layer_forward_multicast");
    if (!l) return fail(QKAN_ERR_BAD_SHAPE, "null layer");
    if (B < 0 || row_offset < 0) return fail(QKAN_ERR_BAD_SHAPE, "negative batch / row offset");
    if (!mc_out) return fail(QKAN_ERR_BAD_SHAPE, "null multicast pointer");
    if (l->engine != 0) return fail(QKAN_ERR_UNSUPPORTED, "the fused output gather is implemented by the block engine (prep = analytic)");
    if (!l->weights_set) return fail(QKAN_ERR_NO_WEIGHTS, "forward called before set_weights");
    if (B == 0) return QKAN_OK;
    if (!x) return fail(QKAN_ERR_BAD_SHAPE, "null x");
    ON_DEVICE(l->device);
    return launch_on(l, x, B, nullptr, nullptr, (cudaStream_t)cuda_stream, nullptr, 0, row_offset, mc_out);
}

// Chunk schedule of qkan_layer_forward_host: cuts[0] = 0 < cuts[1] < ... < cuts[n] = B, returns n (<= max_chunks <= 64).
// Enough chunks to overlap copy-in / compute / copy-out: about 8 MiB of traffic per chunk but at most 16 chunks (each DMA copy
// carries ~10 us of fixed cost: N4 K4, 1M samples: 8 chunks 0.84 ms, 16 +9 %, 64 +60 %; N16 K16, 256 MB: 8 or 16 chunks 3.13 ms,
// 32 chunks 3.26 ms), and never less than one full wave of the forward kernel (a chunk's kernel takes the time of one CTA chunk
// however few CTAs it has: N784 K10 D5 cut into 64 chunks of 1 562 samples ran 17.2 ms against 11.8 ms with 8); every boundary a
// multiple of the CTA tile `tile_samples`.  Making the first and the last chunk smaller (QKAN_HOST_EDGE_DIV > 1; nothing overlaps
// the first copy-in and the last result write) or other boundaries (QKAN_HOST_CUTS) measured no better:
// profiles/r02v_e2e_ab.txt, r02I_e2e_chunk_boundaries.txt.  Pure host arithmetic: exported so that it is tested without a GPU.
extern "C" int qkan_plan_host_chunks(int64_t B, int N, int K, int64_t tile_samples, int64_t wave_samples, int64_t* cuts, int max_chunks) {
    if (B < 1 || N < 1 || K < 1 || tile_samples < 1 || wave_samples < 1 || !cuts || max_chunks < 1)
        return fail(QKAN_ERR_BAD_SHAPE, "qkan_plan_host_chunks: bad arguments");
    if (max_chunks > MAX_CHUNKS) max_chunks = MAX_CHUNKS;
    const int64_t spi = tile_samples, wave = wave_samples;
    int64_t per = ((int64_t)8 << 20) / ((int64_t)(N + K) * 8);
    if (per < 1) per = 1;
    bool forced = false;
    if (const char* e = getenv("QKAN_HOST_CHUNKS")) {        // tuning aid: this many uniform-size chunks
        const int nc = atoi(e);
        if (nc >= 1) { per = (B + nc - 1) / nc; forced = true; }
    }
    if (!forced) {
        if (per * 16 < B) per = (B + 15) / 16;
        if (per < wave) per = wave;
    }
    if (per * max_chunks < B) per = (B + max_chunks - 1) / max_chunks;
    per = (per + spi - 1) / spi * spi;
    int edge_div = 1;
    if (const char* e = getenv("QKAN_HOST_EDGE_DIV")) edge_div = atoi(e);
    int nchunk = 0;
    cuts[0] = 0;
    {
        int64_t edge = edge_div > 1 ? (per / edge_div + spi - 1) / spi * spi : 0;
        if (edge < spi || 2 * edge + per > B || max_chunks < 3) edge = 0;       // too small a batch for edges
        const int64_t mid = B - 2 * edge;
        const int room = max_chunks - (edge ? 2 : 0);
        int nmid = forced ? (int)((mid + per - 1) / per)      // a forced count: rounded up
                          : (int)(mid / per);                 // else rounded down: no chunk below `per` (one wave / ~8 MiB)
        if (nmid < 1) nmid = 1;
        if (nmid > room) nmid = room;
        const int64_t mper = ((mid + nmid - 1) / nmid + spi - 1) / spi * spi;
        if (edge) cuts[++nchunk] = edge;
        for (int i = 0; i < nmid; ++i) {
            int64_t hi = cuts[nchunk] + mper;
            if (hi > B - edge || i == nmid - 1) hi = B - edge;
            if (hi > cuts[nchunk]) cuts[++nchunk] = hi;
        }
        if (edge) cuts[++nchunk] = B;
    }
    if (const char* e = getenv("QKAN_HOST_CUTS")) {          // tuning aid: explicit chunk boundaries (samples, increasing)
        nchunk = 0;
        const char* q = e;
        while (*q && nchunk < max_chunks - 1) {
            char* end = nullptr;
            const long long v = strtoll(q, &end, 10);
            if (end == q) break;
            if (v > cuts[nchunk] && v < B) cuts[++nchunk] = v;
            q = *end ? end + 1 : end;
        }
        cuts[++nchunk] = B;
    }
    return nchunk;
}

extern "C" int qkan_layer_forward_host(qkan_layer* l, const double* x, int64_t B, double* out, void* amps) {
    NvtxRange nvtx_range("qkan_layer_forward_host");
    if (!l) return fail(QKAN_ERR_BAD_SHAPE, "null layer");
    if (B < 0) return fail(QKAN_ERR_BAD_SHAPE, "negative batch");
    if (!l->weights_set) return fail(QKAN_ERR_NO_WEIGHTS, "forward called before set_weights");
    if (B == 0) return QKAN_OK;
    if (!x || !out) return fail(QKAN_ERR_BAD_SHAPE, "null x / out");
    ON_DEVICE(l->device);
    // Pinned (page-locked, device-mapped) host buffers can be used by the kernel directly: it streams x from host memory
    // itself and / or stores every result straight into the host buffer, so that side needs no staging copy and crosses
    // PCIe concurrently with the arithmetic.  Each side is either "direct" (kernel access) or "copy" (chunked
    // cudaMemcpyAsync on its own stream, overlapped with the kernels of the neighbouring chunks):
    //   QKAN_HOST_PATH = copy_in (copy, direct; default for pinned buffers) | zero_copy (direct, direct) | staged (copy,
    //   copy; always for pageable buffers) | copy_out (direct, copy)
    // Measured on B200 / PCIe 5 (profiles/r02g_e2e_ab.txt, r02f_pcie_probe_n1.txt): the DMA engines move 55 GB/s one way and
    // 47 GB/s each way when both directions run; SM-issued reads reach 50 GB/s alone but reads + writes together only
    // 38 GB/s each way.  DMA reads + SM writes is the best combination on the balanced N4 K4 shape (0.855 ms against 0.873
    // zero-copy and 0.90 staged per 1M samples) and within 2 % of the best on the input-heavy N784 K10 shape (12.0 ms
    // against 14.9 zero-copy).
    auto mapped = [](const void* ptr, void** dev) -> bool {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
        if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
        *dev = at.devicePointer;
        return true;
    };
    void *dx = nullptr, *dout = nullptr, *damps = nullptr;
    const bool x_mapped = mapped(x, &dx);
    const bool out_mapped = mapped(out, &dout) && (!amps || mapped(amps, &damps));
    bool in_direct = false, out_direct = out_mapped;
    if (const char* m = getenv("QKAN_HOST_PATH")) {
        if (strcmp(m, "staged") == 0) { in_direct = false; out_direct = false; }
        else if (strcmp(m, "zero_copy") == 0) { in_direct = x_mapped; }
        else if (strcmp(m, "copy_out") == 0) { in_direct = x_mapped; out_direct = false; }
    }
    int rc = ensure_host_path(l, (in_direct && out_direct) ? 0 : B, amps != nullptr && !out_direct);
    if (rc) return rc;
    if (l->ev_w) CU(cudaStreamWaitEvent(l->s_k, l->ev_w, 0));
    if (in_direct && out_direct) {                           // one launch on the host pointers
        rc = launch_on(l, (const double*)dx, B, (double*)dout, damps, l->s_k);
        if (rc) return rc;
        CU(cudaStreamSynchronize(l->s_k));
        return QKAN_OK;
    }
    // the chunk schedule (qkan_plan_host_chunks below: about 8 MiB of traffic per chunk, at most 16, never less than one wave)
    const int64_t spi = l->engine == 0 ? (l->bkern->NT >> (l->lay.g_r_log2 + l->lay.g_k_log2)) : l->kern->spi;
    const int64_t wave = l->engine == 0 ? (int64_t)l->sm_count * 2 * spi * (l->bkern->SU > 0 ? l->bkern->SU : 1) : spi;
    int64_t cut[MAX_CHUNKS + 1];
    const int nchunk = qkan_plan_host_chunks(B, l->N, l->K, spi, wave, cut, MAX_CHUNKS);
    if (nchunk < 1) return nchunk < 0 ? nchunk : fail(QKAN_ERR_BAD_SHAPE, "internal: empty chunk schedule");
    const size_t asz = 2 * amp_real_size(l->dtype);
    auto enqueue = [&]() -> int {
        for (int i = 0; i < nchunk; ++i) {
            const int64_t off = cut[i], n = cut[i + 1] - cut[i];
            const double* xin = (const double*)dx + off * l->N;
            if (!in_direct) {
                CU(cudaMemcpyAsync(l->d_x + off * l->N, x + off * l->N, (size_t)n * l->N * sizeof(double),
                                   cudaMemcpyHostToDevice, l->s_in));
                CU(cudaEventRecord(l->ev_in[i], l->s_in));
                CU(cudaStreamWaitEvent(l->s_k, l->ev_in[i], 0));
                xin = l->d_x + off * l->N;
            }
            double* yout = out_direct ? (double*)dout + off * l->K : l->d_out + off * l->K;
            void* aout = !amps ? nullptr : (out_direct ? (void*)((char*)damps + (size_t)off * l->K * asz)
                                                       : (void*)((char*)l->d_amps + (size_t)off * l->K * asz));
            int rcl = launch_on(l, xin, n, yout, aout, l->s_k);
            if (rcl) return rcl;
            if (!out_direct) {
                CU(cudaEventRecord(l->ev_k[i], l->s_k));
                CU(cudaStreamWaitEvent(l->s_out, l->ev_k[i], 0));
                CU(cudaMemcpyAsync(out + off * l->K, l->d_out + off * l->K, (size_t)n * l->K * sizeof(double),
                                   cudaMemcpyDeviceToHost, l->s_out));
                if (amps)
                    CU(cudaMemcpyAsync((char*)amps + (size_t)off * l->K * asz, (char*)l->d_amps + (size_t)off * l->K * asz,
                                       (size_t)n * l->K * asz, cudaMemcpyDeviceToHost, l->s_out));
            }
        }
        return QKAN_OK;
    };
    // CUDA graph of the pipeline: built on the second consecutive call with the same buffers, replayed afterwards.
    // Pinned buffers only (copies from pageable memory are staged by the driver and cannot be captured usefully).
    unsigned long long cut_hash = (unsigned long long)nchunk;
    for (int i = 1; i <= nchunk; ++i) cut_hash = cut_hash * 1000003ull + (unsigned long long)cut[i];
    const qkan_layer::HostKey key{x, out, amps, B, in_direct ? 1 : 0, out_direct ? 1 : 0, (int)(cut_hash & 0x7fffffffull)};
    auto same = [](const qkan_layer::HostKey& a, const qkan_layer::HostKey& b) {
        return a.x == b.x && a.out == b.out && a.amps == b.amps && a.B == b.B && a.in_direct == b.in_direct &&
               a.out_direct == b.out_direct && a.nchunk == b.nchunk;
    };
    const bool graph_ok = x_mapped && out_mapped && !getenv("QKAN_HOST_NO_GRAPH");
    if (graph_ok && l->graph && same(key, l->graph_key)) {
        CU(cudaGraphLaunch(l->graph, l->s_k));
        CU(cudaStreamSynchronize(l->s_k));
        return QKAN_OK;
    }
    if (graph_ok && same(key, l->last_key)) {
        if (l->graph) { cudaGraphExecDestroy(l->graph); l->graph = nullptr; }
        if (!l->ev_fork) {
            CU(cudaEventCreateWithFlags(&l->ev_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&l->ev_join_in, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&l->ev_join_out, cudaEventDisableTiming));
        }
        CU(cudaStreamSynchronize(l->s_k));
        CU(cudaStreamBeginCapture(l->s_k, cudaStreamCaptureModeThreadLocal));
        // fork the copy streams into the capture, enqueue the pipeline, join them back
        cudaError_t ce = cudaEventRecord(l->ev_fork, l->s_k);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(l->s_in, l->ev_fork, 0);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(l->s_out, l->ev_fork, 0);
        int rcq = ce == cudaSuccess ? enqueue() : QKAN_ERR_CUDA;
        if (ce == cudaSuccess) ce = cudaEventRecord(l->ev_join_in, l->s_in);
        if (ce == cudaSuccess) ce = cudaEventRecord(l->ev_join_out, l->s_out);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(l->s_k, l->ev_join_in, 0);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(l->s_k, l->ev_join_out, 0);
        cudaGraph_t g = nullptr;
        cudaError_t ee = cudaStreamEndCapture(l->s_k, &g);
        if (rcq == QKAN_OK && ce == cudaSuccess && ee == cudaSuccess && g) {
            ee = cudaGraphInstantiate(&l->graph, g, 0);
            cudaGraphDestroy(g);
            if (ee == cudaSuccess) {
                l->graph_key = key;
                CU(cudaGraphLaunch(l->graph, l->s_k));
                CU(cudaStreamSynchronize(l->s_k));
                return QKAN_OK;
            }
            l->graph = nullptr;
        } else if (g) {
            cudaGraphDestroy(g);
        }
        cudaGetLastError();                                  // capture failed: fall through to the plain pipeline
    }
    l->last_key = key;
    rc = enqueue();
    if (rc) return rc;
    CU(cudaStreamSynchronize(l->s_k));
    if (!out_direct) CU(cudaStreamSynchronize(l->s_out));
    return QKAN_OK;
}

extern "C" int qkan_layer_out_of_range(qkan_layer* l, uint64_t* count) {
    if (!l || !count) return fail(QKAN_ERR_BAD_SHAPE, "null argument");
    ON_DEVICE(l->device);
    CU(cudaDeviceSynchronize());
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, l->counters, sizeof v, cudaMemcpyDeviceToHost));
    CU(cudaMemset(l->counters, 0, sizeof v));
    *count = v;
    return QKAN_OK;
}

extern "C" int qkan_layer_info(qkan_layer* l, qkan_kernel_info* info) {
    if (!l || !info) return fail(QKAN_ERR_BAD_SHAPE, "null argument");
    memset(info, 0, sizeof *info);
    info->engine = l->engine;
    info->n_a = l->NA; info->n_b = l->NB; info->l = l->L;
    info->qubits = l->L + 2 + l->NA + l->NB;
    info->grid = l->last_grid; info->smem_bytes = l->last_smem;
    info->passes_survey = (l->D + 1) + (l->NA + l->NB) + 2 * l->L + l->NA;
    info->flops_survey = 6.0 * (double)(1ll << info->qubits) * info->passes_survey;
    info->io_bytes = 8.0 * l->N + 8.0 * l->K;
    const double cf = (l->dtype == QKAN_REAL64) ? 0.5 : 1.0;      // real-only representation: half the arithmetic
    if (l->engine == 0) {
        const BlockKernelInfo& k = *l->bkern;
        const int G = 1 << (l->lay.g_r_log2 + l->lay.g_k_log2);
        info->blocks = l->N * l->K * (l->D + 1);
        info->unroll = k.U;
        info->samples_per_lane = k.SU;
        info->lanes_per_sample = G;
        info->lanes_per_row = 1 << l->lay.g_r_log2;
        info->rows_in_parallel = 1 << l->lay.g_k_log2;
        info->passes = l->lay.passes;
        info->row_steps = l->lay.brows;
        info->threads_per_cta = k.NT;
        info->min_ctas_per_sm = k.MINB;
        info->samples_per_cta = k.NT / G;
        info->passes_exec = l->D + 1;
        // per live block (see evolve_blocks): D-1 full CHEB passes (16 instr / 24 flops), the last CHEB pass
        // and the SELECT rotation pruned to the light cone of the read-out (8 + 4 instr / 12 + 6 flops), one
        // complex add (2 / 2).  DFMA = 2 flops, DMUL = DADD = 1.
        const double Dd = (double)l->D;
        info->flops_exec = cf * (double)info->blocks * (l->D > 0 ? 24.0 * Dd - 4.0 : 8.0);
        info->fp_inst_exec = cf * (double)info->blocks * (l->D > 0 ? 16.0 * Dd - 2.0 : 6.0);
        info->scaled_rotations = k.tan;
        info->input_window = l->window;
        info->flops_per_block_basis = info->flops_exec;
        if (k.amajor) {
            // a-major scaled-rotation kernels (qkan_amajor.cuh).  Per input element the pre-pass runs the CHEB sequence
            // once - D-1 full passes of 8 FMA and the pruned last pass alpha u + beta v (4 MUL + 4 FMA); the window
            // kernel evaluates an element once per row step whose window holds it.  Per (a, b, d) block: the SELECT
            // rotation fused with the read-out sum (4 FMA).  (The conversion of x into the rotation entry, one
            // reciprocal square root per element, is not counted.)
            double elems = (double)l->N;
            if (k.direct) elems = (double)l->K;               // every row evaluates its own element
            info->direct_rows = k.direct;
            info->element_owner = k.elem;
            if (k.elem) {                                     // every row evaluates the elements it reads
                elems = 0.0;
                for (long long b = 0; b < l->K; ++b) {
                    long long f, la;
                    elem_row_range(l->N, l->K, b, &f, &la);
                    elems += (double)(la - f + 1);
                }
            }
            if (l->window) {
                elems = 0.0;
                for (int bi = 0; bi < l->lay.brows; ++bi) {
                    int lo, len;
                    block_window(l->N, l->K, l->lay.g_k_log2, bi, &lo, &len);
                    elems += len;
                }
            }
            info->degree_factored = 1;
            info->cheb_elements = (int)elems;
            if (cheb_uses_sw_form(l->D)) {
                // sin-weighted basis: s^2 (1 FMA), D - 1 full passes (4 MUL + 8 FMA), the pruned pass (4 FMA)
                info->scaled_rotations = 2;
                info->flops_exec = cf * (elems * (20.0 * (Dd - 1.0) + 8.0 + 2.0) + 8.0 * (double)info->blocks);
                info->fp_inst_exec = cf * (elems * (12.0 * (Dd - 1.0) + 4.0 + 1.0) + 4.0 * (double)info->blocks);
            } else {
                info->flops_exec = cf * (elems * (16.0 * (Dd - 1.0) + 12.0) + 8.0 * (double)info->blocks);
                info->fp_inst_exec = cf * (elems * 8.0 * Dd + 4.0 * (double)info->blocks);
            }
            info->flops_per_block_basis = cf * (double)info->blocks * (16.0 * Dd + 4.0);
        }
        info->layout_efficiency = l->lay.efficiency;
    } else {
        const KernelInfo& k = *l->kern;
        info->tile_na = k.NAT; info->tile_nb = k.NBT;
        info->tile_qubits = k.L + 2 + k.NAT + k.NBT;
        info->local_qubits = k.T;
        info->threads_per_cta = k.NT; info->samples_per_cta = k.spi; info->stages = k.stages;
        info->sectors_total = 1 << (l->NA - k.NAT + l->NB - k.NBT);
        const int na_run = (l->N + (1 << k.NAT) - 1) >> k.NAT, nb_run = (l->K + (1 << k.NBT) - 1) >> k.NBT;
        info->sectors_run = na_run * nb_run;
        const int hp = k.prep ? (k.L + k.NAT) : (2 * k.L + 2 * k.NAT + k.NBT);
        info->passes_exec = (l->D + 1) + hp;
        // upper bound: rotations 6 flops / 4 instructions per amplitude, un-normalised Hadamards 2 / 2; the last
        // stage's Hadamards are pruned by the compiler to the post-selected outputs, so the real count is lower
        const double amps_run = (double)(1ll << info->tile_qubits) * info->sectors_run;
        info->flops_exec = cf * amps_run * (6.0 * (l->D + 1) + 2.0 * hp);
        info->fp_inst_exec = cf * amps_run * (4.0 * (l->D + 1) + 2.0 * hp);
        info->layout_efficiency = 1.0;
    }
    return QKAN_OK;
}

// per-stage diagonals (debug / verbose path): one thread per (sample, i)
__global__ void qkan_diagonals_kernel(const double* x, const double* W, long long B, int N, int K, int D, int mode,
                                      double* cheb, double* weighted, double* lcu) {
    const long long NK = (long long)N * K;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * NK) return;
    const long long s = gid / NK;
    const int i = (int)(gid - s * NK);
    const double xc = clip_unit<double>(x[s * N + i / K]);          // ChebyshevStep.py:52,64
    const double th = acos(xc);
    const double cD = cos((double)D * th);                           // ChebyshevStep.py:30
    if (cheb) cheb[gid] = cD;
    double acc = 0.0;
    for (int d = 0; d <= D; ++d) {
        const double c = mode == 0 ? cD : cos((double)d * th);
        const double m = c * W[(long long)d * NK + i];               // MulStep.py:72
        if (weighted) weighted[(s * (D + 1) + d) * NK + i] = m;
        acc += m / (double)(D + 1);                                  // LCUStep.py:36
    }
    if (lcu) lcu[gid] = acc;
}

extern "C" int qkan_layer_diagonals(qkan_layer* l, const double* x, int64_t B, double* cheb, double* weighted,
                                    double* lcu, void* cuda_stream) {
    NvtxRange nvtx_range("qkan_layer_diagonals");
    if (!l) return fail(QKAN_ERR_BAD_SHAPE, "null layer");
    if (B < 0) return fail(QKAN_ERR_BAD_SHAPE, "negative batch");
    if (!l->weights_set) return fail(QKAN_ERR_NO_WEIGHTS, "diagonals called before set_weights");
    if (B == 0) return QKAN_OK;
    if (!x) return fail(QKAN_ERR_BAD_SHAPE, "null x");
    ON_DEVICE(l->device);
    const long long total = (long long)B * l->N * l->K;
    const unsigned nt = 256;
    const unsigned nb = (unsigned)((total + nt - 1) / nt);
    qkan_diagonals_kernel<<<nb, nt, 0, (cudaStream_t)cuda_stream>>>(x, l->W_dev, B, l->N, l->K, l->D, l->mode, cheb,
                                                                    weighted, lcu);
    CU(cudaGetLastError());
    return QKAN_OK;
}

// Stage snapshots of the SIMULATED circuit (debug / verbose path): the post-selected block amplitudes after the CHEB
// sequence, after SELECT and after the degree sum, read out of the same evolution functions the forward kernels run
// (cheb_element + the SELECT rotation of qkan_amajor.cuh for compat mode, 1 <= D <= 16; evolve_blocks of
// qkan_block.cuh otherwise) - not the closed form cos(D arccos x) of qkan_diagonals_kernel, which they are tested
// against.  One thread per (sample, i), i = a + N b (the reference's diagonal index, QKANLayer.py:131-132).
template <int DT>
__device__ __forceinline__ void snapshot_block(const Cplx<double> (&init)[4], double xc, const double* W, long long NK, int i, int D,
                                               int mode, double* cheb_out, double* wtd_out, long long wtd_stride, double* lcu_out) {
    typedef Cplx<double> A;
    double acc = 0.0;
    if constexpr (DT > 0) {
        A lo0, lo2;
        cheb_element<A, double, DT>(init, xc, lo0, lo2);      // f_x = 0 amplitudes of the block after CHEB
        if (cheb_out) *cheb_out = lo0.re;
        for (int d = 0; d <= D; ++d) {
            const double w = W[(long long)d * NK + i];
            const double sw = sqrt((1.0 - w) * (1.0 + w));
            const double z = fma(w, lo0.re, -(sw * lo2.re));  // (f_x, f_w) = (0, 0) amplitude after SELECT
            if (wtd_out) wtd_out[(long long)d * wtd_stride] = z;
            acc += z / (double)(D + 1);                       // LCUStep.py:36 (sequential, degree order)
        }
    } else {
        const double cx[1] = {xc}, sx[1] = {sqrt((1.0 - xc) * (1.0 + xc))};
        const double one[1] = {1.0}, zero[1] = {0.0};
        const int dmax[1] = {D};
        if (cheb_out) *cheb_out = evolve_blocks<A, double, 1, 1, 0>(init, cx, sx, one, zero, dmax, D).re;   // every application on
        for (int d = 0; d <= D; ++d) {
            const double w = W[(long long)d * NK + i];
            const double cw[1] = {w}, sw[1] = {sqrt((1.0 - w) * (1.0 + w))};
            const int deg[1] = {mode == 1 ? d : D};           // paper mode: term d gets d applications
            const double z = evolve_blocks<A, double, 1, 1, 0>(init, cx, sx, cw, sw, deg, D).re;
            if (wtd_out) wtd_out[(long long)d * wtd_stride] = z;
            acc += z / (double)(D + 1);
        }
    }
    if (lcu_out) *lcu_out = acc;
}

__global__ void qkan_stage_snapshot_kernel(const double* x, const double* W, long long B, int N, int K, int D, int mode,
                                           double* cheb, double* weighted, double* lcu) {
    const long long NK = (long long)N * K;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * NK) return;
    const long long s = gid / NK;
    const int i = (int)(gid - s * NK);
    const double xc = clip_unit<double>(x[s * N + i / K]);           // ChebyshevStep.py:52,64
    Cplx<double> init[4];
    for (int q = 0; q < 4; ++q) { init[q].re = q == 0 ? 1.0 : 0.0; init[q].im = 0.0; }   // PREPARE'd block state, un-normalised
    double* c = cheb ? cheb + gid : nullptr;
    double* w = weighted ? weighted + s * (D + 1) * NK + i : nullptr;
    double* l = lcu ? lcu + gid : nullptr;
    const bool tan = mode == 0 && D >= TAN_MIN_DT && D <= TAN_MAX_DT;
    switch (tan ? D : 0) {
#define SNAP_CASE(DT) case DT: snapshot_block<DT>(init, xc, W, NK, i, D, mode, c, w, NK, l); break;
        SNAP_CASE(1) SNAP_CASE(2) SNAP_CASE(3) SNAP_CASE(4) SNAP_CASE(5) SNAP_CASE(6) SNAP_CASE(7) SNAP_CASE(8)
        SNAP_CASE(9) SNAP_CASE(10) SNAP_CASE(11) SNAP_CASE(12) SNAP_CASE(13) SNAP_CASE(14) SNAP_CASE(15) SNAP_CASE(16)
#undef SNAP_CASE
        default: snapshot_block<0>(init, xc, W, NK, i, D, mode, c, w, NK, l); break;
    }
}

extern "C" int qkan_layer_stage_snapshots(qkan_layer* l, const double* x, int64_t B, double* cheb, double* weighted,
                                          double* lcu, void* cuda_stream) {
    NvtxRange nvtx_range("qkan_layer_stage_snapshots");
    if (!l) return fail(QKAN_ERR_BAD_SHAPE, "null layer");
    if (B < 0) return fail(QKAN_ERR_BAD_SHAPE, "negative batch");
    if (!l->weights_set) return fail(QKAN_ERR_NO_WEIGHTS, "stage snapshots called before set_weights");
    if (B == 0) return QKAN_OK;
    if (!x) return fail(QKAN_ERR_BAD_SHAPE, "null x");
    ON_DEVICE(l->device);
    const long long total = (long long)B * l->N * l->K;
    const unsigned nt = 128;
    const unsigned nb = (unsigned)((total + nt - 1) / nt);
    qkan_stage_snapshot_kernel<<<nb, nt, 0, (cudaStream_t)cuda_stream>>>(x, l->W_dev, B, l->N, l->K, l->D, l->mode, cheb,
                                                                         weighted, lcu);
    CU(cudaGetLastError());
    return QKAN_OK;
}

extern "C" int qkan_forward(const void* x, const void* w, void* out, int64_t B, int N, int K, int D, int dtype,
                            int mode, void* amps, void* cuda_stream) {
    thread_local qkan_layer* cached = nullptr;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (cached && (cached->N != N || cached->K != K || cached->D != D || cached->dtype != dtype ||
                   cached->mode != mode || cached->device != dev)) {
        qkan_layer_destroy(cached);
        cached = nullptr;
    }
    if (!cached) {
        int rc = qkan_layer_create(&cached, N, K, D, dtype, mode, QKAN_PREP_ANALYTIC, dev);
        if (rc) return rc;
    }
    int rc = qkan_layer_set_weights(cached, (const double*)w, 1, 0, cuda_stream);
    if (rc) return rc;
    return qkan_layer_forward(cached, (const double*)x, B, (double*)out, amps, cuda_stream);
}

// ------------------------------------------------------------- FMA peak
template <typename R> __global__ void __launch_bounds__(256) fma_peak_kernel(R* sink, int iters, R b, R c) {
    R a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (R)(threadIdx.x + i) * (R)1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = qk_fma(a[i], b, c);
        }
    }
    R s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == (R)123456.789) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int qkan_measure_fma_peak(int device, int fp64, double* tflops) {
    if (!tflops) return fail(QKAN_ERR_BAD_SHAPE, "null argument");
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8, nt = 256, iters = fp64 ? 8192 : 16384;
    void* sink = nullptr;
    CU(cudaMalloc(&sink, (size_t)grid * nt * 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0));
        if (fp64) fma_peak_kernel<double><<<grid, nt>>>((double*)sink, iters, 0.999999, 1e-6);
        else fma_peak_kernel<float><<<grid, nt>>>((float*)sink, iters, 0.9999f, 1e-4f);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8 * 4 * (double)iters * grid * nt;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return QKAN_OK;
}

extern "C" const char* qkan_last_error(void) { return g_last_error.c_str(); }
extern "C" int qkan_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; return 0; }
extern "C" void qkan_version(int* major, int* minor, int* patch) {
    if (major) *major = 0;
    if (minor) *minor = 1;
    if (patch) *patch = 0;
}
