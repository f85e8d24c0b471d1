// kernels.cuh -- launch interface of the CUDA interpreter (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace exo
{
struct KernelParams
{
    const unsigned char* program;  // device copy of Program::image
    int32_t program_bytes;
    int32_t dbl_offset;            // byte offset of the double pool in the image
    int32_t scratch_offset;        // byte offset of the per-thread scratch inside dynamic shared memory
    int32_t T;                     // timesteps (evaluation e is timestep e % T)
    const double* x;
    double* phi;
    double* jac;                   // may be NULL
    const double* generated;       // [T][n_generated][12] or NULL
    int64_t n_evals;
    // element strides (in doubles): addr = e * s_e + k * s_k  (Phi, X)   /  e * s_e + row * s_r + col * s_c  (J)
    int64_t x_se, x_sk;
    int64_t phi_se, phi_sk;
    int64_t j_se, j_sr, j_sc;
};

struct LaunchConfig
{
    int threads = 128;
    int blocks = 0;
    size_t smem_bytes = 0;
};

// Computes threads/blocks/shared memory for a program on the current device.
cudaError_t plan_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, LaunchConfig& cfg, KernelParams& params);
cudaError_t launch_eval(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream);

// Register-resident kernel for serial-chain programs (chain_kernel.cu)
bool chain_supported(const ProgramHeader& hdr);
cudaError_t plan_chain_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, LaunchConfig& cfg, KernelParams& params);
cudaError_t launch_chain(int chain_joints, const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream);

}  // namespace exo
