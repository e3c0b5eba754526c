// qkan_circuit.cu - generic gate-list statevector simulation, one CTA per initial basis state.
// Used to evaluate block-encoding circuits (the reference does this with Qiskit Aer's
// unitary_simulator in its unit tests: MulStep.py:115-166, LCUStep.py:69-107, SUMStep.py:40-78):
// column j of the circuit's unitary = the state evolved from |j>.  Not a hot path: the state lives
// in global memory (L2) and every gate is one pass.
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

__global__ void __launch_bounds__(256) qkan_circuit_kernel(const int* __restrict__ gates, const double* __restrict__ params,
                                                          int n_gates, int n_qubits, const long long* __restrict__ basis,
                                                          double2* __restrict__ state) {
    const long long S = 1ll << n_qubits;
    double2* st = state + (long long)blockIdx.x * S;
    const long long b0 = basis[blockIdx.x];
    for (long long i = threadIdx.x; i < S; i += blockDim.x) st[i] = make_double2(i == b0 ? 1.0 : 0.0, 0.0);
    const double r = 0.70710678118654752440;
    for (int g = 0; g < n_gates; ++g) {
        __syncthreads();
        const int kind = gates[3 * g], q0 = gates[3 * g + 1], q1 = gates[3 * g + 2];
        if (kind == G_H || kind == G_RY || kind == G_X || kind == G_Z) {
            double c = 0.0, s = 0.0;
            if (kind == G_RY) sincos(0.5 * params[g], &s, &c);
            for (long long t = threadIdx.x; t < (S >> 1); t += blockDim.x) {
                const long long i0 = insert_zero(t, q0), i1 = i0 | (1ll << q0);
                const double2 a = st[i0], b = st[i1];
                if (kind == G_H) {
                    st[i0] = make_double2((a.x + b.x) * r, (a.y + b.y) * r);
                    st[i1] = make_double2((a.x - b.x) * r, (a.y - b.y) * r);
                } else if (kind == G_RY) {
                    st[i0] = make_double2(c * a.x - s * b.x, c * a.y - s * b.y);
                    st[i1] = make_double2(s * a.x + c * b.x, s * a.y + c * b.y);
                } else if (kind == G_X) {
                    st[i0] = b; st[i1] = a;
                } else {
                    st[i1] = make_double2(-b.x, -b.y);
                }
            }
        } else if (kind == G_CX) {          // q0 = control, q1 = target
            for (long long t = threadIdx.x; t < (S >> 1); t += blockDim.x) {
                const long long i0 = insert_zero(t, q1), i1 = i0 | (1ll << q1);
                if ((i0 >> q0) & 1) { const double2 a = st[i0]; st[i0] = st[i1]; st[i1] = a; }
            }
        } else if (kind == G_SWAP) {
            for (long long i = threadIdx.x; i < S; i += blockDim.x) {
                if (((i >> q0) & 1) == 1 && ((i >> q1) & 1) == 0) {
                    const long long j = (i ^ (1ll << q0)) | (1ll << q1);
                    const double2 a = st[i]; st[i] = st[j]; st[j] = a;
                }
            }
        }
    }
}
}  // namespace

extern "C" int qkan_set_last_error(const char* msg);   // defined in qkan_capi.cu

extern "C" int qkan_simulate_circuit(const int* gates, const double* params, int n_gates, int n_qubits,
                                     const long long* basis, int64_t n_states, void* state_out, void* cuda_stream) {
    if (!gates || !params || !basis || !state_out) { qkan_set_last_error("null argument"); return QKAN_ERR_BAD_SHAPE; }
    if (n_qubits < 1 || n_qubits > 28 || n_gates < 0 || n_states < 0) { qkan_set_last_error("bad circuit size"); return QKAN_ERR_BAD_SHAPE; }
    if (n_states == 0) return QKAN_OK;
    qkan_circuit_kernel<<<(unsigned)n_states, 256, 0, (cudaStream_t)cuda_stream>>>(gates, params, n_gates, n_qubits, basis,
                                                                                  (double2*)state_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { qkan_set_last_error((std::string("qkan_circuit_kernel: ") + cudaGetErrorString(e)).c_str()); return QKAN_ERR_CUDA; }
    return QKAN_OK;
}
