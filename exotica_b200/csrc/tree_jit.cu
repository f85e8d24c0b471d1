// tree_jit.cu -- run-time compilation (NVRTC, sm_100a) and launch of the per-robot kernels tree_codegen.cpp writes.
//
// libnvrtc is opened with dlopen and the driver's module API is resolved through cudaGetDriverEntryPoint, so the library
// links against neither; a box without NVRTC simply keeps tree programs on the interpreter (tree_jit_get reports why).
// Compiled modules are cached per (device, source text) for the life of the process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "kernels.cuh"

namespace exo
{
namespace
{
// the headers the generated source includes, embedded at build time (Makefile: *.inc)
const char kSrcDeviceMath[] =
#include "device_math.cuh.inc"
    ;
const char kSrcProgram[] =
#include "program.h.inc"
    ;
const char kSrcApi[] =
#include "exotica_b200.h.inc"
    ;
const char kSrcStdint[] =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long long uintptr_t; typedef long long intptr_t;\n";
const char kSrcStddef[] = "#pragma once\ntypedef unsigned long size_t;\n";
const char kSrcMathConstants[] =
    "#pragma once\n"
    "#define CUDART_PIO2 1.5707963267948966e+0\n"
    "#define CUDART_PI 3.1415926535897931e+0\n";

typedef int nvrtcResult;
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc
{
    void* lib = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    std::string err;
};

Nvrtc& nvrtc()
{
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names)
    {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.lib) break;
    }
    if (!n.lib)
    {
        n.err = "libnvrtc not found (dlopen)";
        return n;
    }
#define EXO_SYM(field, name)                                              \
    *reinterpret_cast<void**>(&n.field) = dlsym(n.lib, name);              \
    if (!n.field)                                                          \
    {                                                                      \
        n.err = std::string("libnvrtc lacks ") + name;                     \
        n.lib = nullptr;                                                   \
        return n;                                                          \
    }
    EXO_SYM(CreateProgram, "nvrtcCreateProgram")
    EXO_SYM(CompileProgram, "nvrtcCompileProgram")
    EXO_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    EXO_SYM(GetCUBIN, "nvrtcGetCUBIN")
    EXO_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    EXO_SYM(GetProgramLog, "nvrtcGetProgramLog")
    EXO_SYM(DestroyProgram, "nvrtcDestroyProgram")
    EXO_SYM(GetErrorString, "nvrtcGetErrorString")
#undef EXO_SYM
    return n;
}

struct Driver
{
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    bool ok = false;
};

Driver& driver()
{
    static Driver d;
    static bool tried = false;
    if (tried) return d;
    tried = true;
    auto get = [](const char* name) -> void* {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return p;
    };
    *reinterpret_cast<void**>(&d.ModuleLoadData) = get("cuModuleLoadData");
    *reinterpret_cast<void**>(&d.ModuleGetFunction) = get("cuModuleGetFunction");
    *reinterpret_cast<void**>(&d.FuncSetAttribute) = get("cuFuncSetAttribute");
    *reinterpret_cast<void**>(&d.LaunchKernel) = get("cuLaunchKernel");
    d.ok = d.ModuleLoadData && d.ModuleGetFunction && d.FuncSetAttribute && d.LaunchKernel;
    return d;
}

struct Compiled
{
    CUfunction fn = nullptr;
    size_t smem_set = 0;
};
std::mutex g_mu;
std::map<std::pair<int, std::string>, Compiled*> g_cache;
}  // namespace

namespace
{
bool compile_source(const std::string& source, std::vector<char>& cubin, std::string& err)
{
    Nvrtc& n = nvrtc();
    if (!n.lib)
    {
        err = n.err;
        return false;
    }
    const char* headers[] = {kSrcDeviceMath, kSrcProgram, kSrcApi, kSrcStdint, kSrcStddef, kSrcMathConstants};
    const char* names[] = {"device_math.cuh", "program.h", "../../include/exotica_b200.h", "stdint.h", "stddef.h", "math_constants.h"};
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.CreateProgram(&prog, source.c_str(), "exo_tree_kernel.cu", 6, headers, names);
    if (r != 0)
    {
        err = std::string("nvrtcCreateProgram: ") + n.GetErrorString(r);
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    r = n.CompileProgram(prog, 3, opts);
    if (r != 0)
    {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        if (log.size() > 4000) log.resize(4000);
        err = std::string("nvrtcCompileProgram: ") + n.GetErrorString(r) + "\n" + log;
        n.DestroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    r = cs ? n.GetCUBIN(prog, cubin.data()) : 1;
    n.DestroyProgram(&prog);
    if (r != 0)
    {
        err = "nvrtcGetCUBIN failed";
        return false;
    }
    if (const char* dump = getenv("EXO_TREE_DUMP"))  // keep the artefacts for cuobjdump / inspection
    {
        const std::string base = dump;
        if (FILE* f = fopen((base + ".cu").c_str(), "w"))
        {
            fwrite(source.data(), 1, source.size(), f);
            fclose(f);
        }
        if (FILE* f = fopen((base + ".cubin").c_str(), "wb"))
        {
            fwrite(cubin.data(), 1, cubin.size(), f);
            fclose(f);
        }
    }
    return true;
}

// module cache: compiled kernels by (device, source text)
Compiled* get_kernel(const std::string& source, const char* entry, std::string& err)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
    {
        cudaGetLastError();
        err = "no CUDA device";
        return nullptr;
    }
    if (cudaFree(0) != cudaSuccess)  // the primary context must exist before a module is loaded into it
    {
        cudaGetLastError();
        err = "CUDA context";
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_mu);
    const auto key = std::make_pair(dev, source);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    Driver& d = driver();
    if (!d.ok)
    {
        err = "driver module API not available";
        return nullptr;
    }
    std::vector<char> cubin;
    if (!compile_source(source, cubin, err)) return nullptr;
    CUmodule mod = nullptr;
    CUresult cr = d.ModuleLoadData(&mod, cubin.data());
    if (cr != CUDA_SUCCESS)
    {
        err = "cuModuleLoadData failed (" + std::to_string((int)cr) + ")";
        return nullptr;
    }
    Compiled* c = new Compiled();
    cr = d.ModuleGetFunction(&c->fn, mod, entry);
    if (cr != CUDA_SUCCESS)
    {
        err = "cuModuleGetFunction failed (" + std::to_string((int)cr) + ")";
        delete c;
        return nullptr;
    }
    g_cache[key] = c;
    return c;
}

cudaError_t launch_kernel(Compiled* c, size_t smem_bytes, int threads, int64_t n_evals, void* params, cudaStream_t stream)
{
    Driver& d = driver();
    if (!c || !d.ok) return cudaErrorInvalidDeviceFunction;
    if (smem_bytes > c->smem_set)
    {
        if (d.FuncSetAttribute(c->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_bytes) != CUDA_SUCCESS) return cudaErrorInvalidValue;
        c->smem_set = smem_bytes;
    }
    const int64_t blocks = (n_evals + threads - 1) / threads;
    if (blocks > 0x7fffffff) return cudaErrorInvalidConfiguration;
    void* args[] = {params};
    const CUresult r = d.LaunchKernel(c->fn, (unsigned)blocks, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem_bytes, (CUstream)stream, args, nullptr);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}
}  // namespace

bool tree_jit_compile(const TreeGenPlan& plan, std::vector<char>& cubin, std::string& err)
{
    return compile_source(plan.source, cubin, err);
}

void* tree_jit_get(const TreeGenPlan& plan, std::string& err)
{
    return get_kernel(plan.source, plan.kernel_name, err);
}

bool tree_costgrad_jit_compile(const TreeCostPlan& plan, std::vector<char>& cubin, std::string& err)
{
    return compile_source(plan.source, cubin, err);
}

void* tree_costgrad_jit_get(const TreeCostPlan& plan, std::string& err)
{
    return get_kernel(plan.source, "exo_tree_costgrad_kernel", err);
}

void* tree_park_jit_get(const TreeParkPlan& plan, std::string& err)
{
    return get_kernel(plan.source, "exo_tree_park_kernel", err);
}

cudaError_t tree_park_jit_launch(void* kernel, const TreeParkPlan& plan, const TreeParkParams& P, cudaStream_t stream)
{
    TreeParkParams params = P;
    return launch_kernel(static_cast<Compiled*>(kernel), 0, plan.threads, P.n_evals, &params, stream);
}

cudaError_t tree_costgrad_jit_launch(void* kernel, const TreeCostPlan& plan, const TreeCostParams& P, cudaStream_t stream)
{
    TreeCostParams params = P;
    return launch_kernel(static_cast<Compiled*>(kernel), plan.smem_bytes, plan.warps_per_cta * 32, P.n_evals, &params, stream);
}

bool tree_jit_bind(const TreeGenPlan& plan, TreeGenParams& P, const double* x, double* phi, double* jac, int64_t n_evals, int64_t x_se, int64_t x_sk,
                   int64_t phi_se, int64_t phi_sk, int64_t j_se, int64_t j_sr, int64_t j_sc)
{
    // evaluations contiguous in every buffer, 16-byte aligned row pitches
    if (n_evals < 1 || n_evals > 0x7fffffff || x_se != 1 || phi_se != 1 || (jac && j_se != 1)) return false;
    if (phi_sk < n_evals || (phi_sk & 1)) return false;
    if (jac && (j_sc < n_evals || (j_sc & 1) || (j_sr & 1))) return false;
    if (jac && plan.needs_flat_jacobian && j_sr != (int64_t)plan.n_ctrl * j_sc) return false;
    if ((int)plan.maps.size() > TREE_MAXMAPS) return false;
    P.x = x; P.phi = phi; P.jac = jac;
    P.n_evals = n_evals; P.x_sk = x_sk; P.phi_sk = phi_sk; P.j_sr = j_sr; P.j_sc = j_sc;
    for (size_t m = 0; m < plan.maps.size(); ++m)
    {
        const TreeGenPlan::MapShape& sh = plan.maps[m];
        if (sh.phi == 1)
        {
            if (!tma_encode_2d(P.maps[m], phi, (uint64_t)n_evals, (uint64_t)plan.length_phi, (uint64_t)phi_sk * 8, (uint32_t)plan.tile, (uint32_t)sh.nrows)) return false;
        }
        else if (sh.phi == 2)
        {
            if (jac && !tma_encode_2d(P.maps[m], jac, (uint64_t)n_evals, (uint64_t)plan.length_jac * plan.n_ctrl, (uint64_t)j_sc * 8, (uint32_t)plan.tile, (uint32_t)sh.nrows))
                return false;
        }
        else if (jac)
        {
            if (!tma_encode_3d(P.maps[m], jac, (uint64_t)n_evals, (uint64_t)plan.n_ctrl, (uint64_t)plan.length_jac, (uint64_t)j_sc * 8, (uint64_t)j_sr * 8, (uint32_t)plan.tile,
                               (uint32_t)sh.ncols, (uint32_t)sh.nrows))
                return false;
        }
    }
    return true;
}

cudaError_t tree_jit_launch(void* kernel, const TreeGenPlan& plan, const TreeGenParams& P, cudaStream_t stream)
{
    TreeGenParams params = P;
    return launch_kernel(static_cast<Compiled*>(kernel), plan.smem_bytes, plan.warps_per_cta * 32, P.n_evals, &params, stream);
}

}  // namespace exo
