// qkan_circuit.cu - generic gate-list statevector simulation, batched over initial basis states.
// Evaluates block-encoding circuits the way the reference's unit tests do with Qiskit Aer's unitary_simulator
// (MulStep.py:115-166, LCUStep.py:69-107, SUMStep.py:40-78): column j of the circuit's unitary = the state evolved
// from |j>.  One CTA per column.
//   * <= 13 qubits: the whole state (<= 128 KB) lives in shared memory for the whole gate list and is written out once;
//   * larger circuits: the state lives in the output buffer (2 MB per column at 17 qubits: L2 resident).
// Run fusion: a maximal run of consecutive gates that all TARGET the same qubit t - H / RY / X / Z on t and CX with
// target t (any control) - is applied in ONE pass over the state: a thread keeps the amplitude pair (bit t = 0 / 1) of a
// fixed value of the other qubits in registers and walks the run (a CX swaps the pair when the control bit of that
// value is set).  FABLE's oracle O_A is 2 * 4^n gates that all target the flag qubit (fable.py), so a block-encoding
// circuit is 3 n + 1 passes instead of 2 * 4^n + 3 n; that is what makes the 17-qubit LCU circuit of N16 K16
// (131 072 oracle gates) a sub-second job.
#include "../../include/qkan_b200.h"
#include <cuda_runtime.h>
#include <math.h>
#include <string>

namespace {
enum GateKind { G_H = 0, G_RY = 1, G_CX = 2, G_SWAP = 3, G_X = 4, G_Z = 5 };

__device__ __forceinline__ long long insert_zero(long long t, int q) {
    const long long low = t & ((1ll << q) - 1);
    return ((t >> q) << (q + 1)) | low;
}

// (cos, sin)(theta / 2) of every RY, once per call
__global__ void qkan_circuit_cs_kernel(const int* __restrict__ gates, const double* __restrict__ params, int n_gates, double2* cs) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gates) return;
    double c = 1.0, s = 0.0;
    if (gates[3 * g] == G_RY) sincos(0.5 * params[g], &s, &c);
    cs[g] = make_double2(c, s);
}

template <bool SMEM>
__global__ void __launch_bounds__(512) qkan_circuit_kernel(const int* __restrict__ gates, const double2* __restrict__ cs, int n_gates,
                                                           int n_qubits, const long long* __restrict__ basis, double2* __restrict__ state) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long S = 1ll << n_qubits;
    double2* out = state + (long long)blockIdx.x * S;
    double2* st = SMEM ? reinterpret_cast<double2*>(smem_raw) : out;
    const long long b0 = basis[blockIdx.x];
    for (long long i = threadIdx.x; i < S; i += blockDim.x) st[i] = make_double2(i == b0 ? 1.0 : 0.0, 0.0);
    const double r = 0.70710678118654752440;
    int g = 0;
    while (g < n_gates) {
        __syncthreads();
        const int kind = gates[3 * g], q0 = gates[3 * g + 1], q1 = gates[3 * g + 2];
        if (kind == G_SWAP) {
            for (long long i = threadIdx.x; i < S; i += blockDim.x) {
                if (((i >> q0) & 1) == 1 && ((i >> q1) & 1) == 0) {
                    const long long j = (i ^ (1ll << q0)) | (1ll << q1);
                    const double2 a = st[i]; st[i] = st[j]; st[j] = a;
                }
            }
            ++g;
            continue;
        }
        const int t = kind == G_CX ? q1 : q0;
        int e = g + 1;                                        // the run [g, e): same target qubit, no SWAP (uniform scan)
        while (e < n_gates) {
            const int k = gates[3 * e];
            if (k == G_SWAP || (k == G_CX ? gates[3 * e + 2] : gates[3 * e + 1]) != t) break;
            ++e;
        }
        // PP pairs per thread at a time: the gate record is read once for all of them
        constexpr int PP = 4;
        for (long long p = threadIdx.x; p < (S >> 1); p += (long long)PP * blockDim.x) {
            long long i0[PP];
            double2 a[PP], b[PP];
            bool live[PP];
#pragma unroll
            for (int u = 0; u < PP; ++u) {
                const long long pu = p + (long long)u * blockDim.x;
                live[u] = pu < (S >> 1);
                i0[u] = insert_zero(live[u] ? pu : 0, t);
                a[u] = st[i0[u]];
                b[u] = st[i0[u] | (1ll << t)];
            }
            for (int gg = g; gg < e; ++gg) {
                const int k = gates[3 * gg];
                if (k == G_RY) {
                    const double2 q = cs[gg];
#pragma unroll
                    for (int u = 0; u < PP; ++u) {
                        const double2 na = make_double2(q.x * a[u].x - q.y * b[u].x, q.x * a[u].y - q.y * b[u].y);
                        b[u] = make_double2(q.y * a[u].x + q.x * b[u].x, q.y * a[u].y + q.x * b[u].y);
                        a[u] = na;
                    }
                } else if (k == G_CX) {
                    const int c = gates[3 * gg + 1];
#pragma unroll
                    for (int u = 0; u < PP; ++u)
                        if ((i0[u] >> c) & 1) { const double2 tmp = a[u]; a[u] = b[u]; b[u] = tmp; }
                } else if (k == G_H) {
#pragma unroll
                    for (int u = 0; u < PP; ++u) {
                        const double2 na = make_double2((a[u].x + b[u].x) * r, (a[u].y + b[u].y) * r);
                        b[u] = make_double2((a[u].x - b[u].x) * r, (a[u].y - b[u].y) * r);
                        a[u] = na;
                    }
                } else if (k == G_X) {
#pragma unroll
                    for (int u = 0; u < PP; ++u) { const double2 tmp = a[u]; a[u] = b[u]; b[u] = tmp; }
                } else {                                      // Z
#pragma unroll
                    for (int u = 0; u < PP; ++u) b[u] = make_double2(-b[u].x, -b[u].y);
                }
            }
#pragma unroll
            for (int u = 0; u < PP; ++u)
                if (live[u]) {
                    st[i0[u]] = a[u];
                    st[i0[u] | (1ll << t)] = b[u];
                }
        }
        g = e;
    }
    if (SMEM) {
        __syncthreads();
        for (long long i = threadIdx.x; i < S; i += blockDim.x) out[i] = st[i];
    }
}
}  // namespace

extern "C" int qkan_set_last_error(const char* msg);   // defined in qkan_capi.cu

extern "C" int qkan_simulate_circuit(const int* gates, const double* params, int n_gates, int n_qubits,
                                     const long long* basis, int64_t n_states, void* state_out, void* cuda_stream) {
    if (!gates || !params || !basis || !state_out) { qkan_set_last_error("null argument"); return QKAN_ERR_BAD_SHAPE; }
    if (n_qubits < 1 || n_qubits > 28 || n_gates < 0 || n_states < 0) { qkan_set_last_error("bad circuit size"); return QKAN_ERR_BAD_SHAPE; }
    if (n_states == 0) return QKAN_OK;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    auto fail_cuda = [](cudaError_t e, const char* what) {
        qkan_set_last_error((std::string(what) + ": " + cudaGetErrorString(e)).c_str());
        return (int)QKAN_ERR_CUDA;
    };
    double2* cs = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&cs, sizeof(double2) * (size_t)(n_gates > 0 ? n_gates : 1), stream);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMallocAsync(circuit rotation table)");
    if (n_gates > 0) qkan_circuit_cs_kernel<<<(n_gates + 255) / 256, 256, 0, stream>>>(gates, params, n_gates, cs);
    const size_t state_bytes = sizeof(double2) << n_qubits;
    const long long pairs = 1ll << (n_qubits - 1);
    const int nt = pairs >= 512 ? 512 : (pairs >= 32 ? (int)pairs : 32);
    if (n_qubits <= 13) {                                     // the state stays in shared memory for the whole gate list
        e = cudaFuncSetAttribute(qkan_circuit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)state_bytes);
        if (e == cudaSuccess) {
            qkan_circuit_kernel<true><<<(unsigned)n_states, nt, state_bytes, stream>>>(gates, cs, n_gates, n_qubits, basis, (double2*)state_out);
            e = cudaGetLastError();
        }
    } else {
        qkan_circuit_kernel<false><<<(unsigned)n_states, nt, 0, stream>>>(gates, cs, n_gates, n_qubits, basis, (double2*)state_out);
        e = cudaGetLastError();
    }
    cudaFreeAsync(cs, stream);
    if (e != cudaSuccess) return fail_cuda(e, "qkan_circuit_kernel");
    return QKAN_OK;
}
